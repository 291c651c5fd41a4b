set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_j16_bench_n1.err | tail -1 > gpurun_out/r02_j16_bench_query_vith_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r02_j16_bench_ref.err | tail -1 > gpurun_out/r02_j16_bench_reference.json
for f in gpurun_out/r02_j16_bench_*.json; do echo $f; cut -c1-900 $f; echo; done
tail -3 gpurun_out/r02_j16_bench_n1.err

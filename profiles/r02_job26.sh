set -x
mkdir -p gpurun_out
timeout 900 python bench.py 2> gpurun_out/r02_j26_bench.err | tail -1 > gpurun_out/r02_j26_bench_query_vith_n1.json
cut -c1-300 gpurun_out/r02_j26_bench_query_vith_n1.json; tail -3 gpurun_out/r02_j26_bench.err

set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_j4_pytest.log
for S in 768 1280; do
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size $S --steps 10 --warmup 3 2> gpurun_out/r02_j4_bench_enc_$S.err | tail -1 > gpurun_out/r02_j4_bench_encoder_vith_$S.json
done
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb --steps 20 --warmup 3 2> gpurun_out/r02_j4_bench_anchor.err | tail -1 > gpurun_out/r02_j4_bench_anchor_vitb.json
timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j4_bench_n1.err | tail -1 > gpurun_out/r02_j4_bench_query_vith_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r02_j4_bench_ref.err | tail -1 > gpurun_out/r02_j4_bench_reference.json
tail -25 gpurun_out/r02_j4_pytest.log
for f in gpurun_out/r02_j4_bench_*.json; do echo $f; cut -c1-300 $f; echo; done
tail -5 gpurun_out/r02_j4_bench_enc_1280.err

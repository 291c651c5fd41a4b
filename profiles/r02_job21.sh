set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_j21_gpu_tests.log 2>&1; tail -12 gpurun_out/r02_j21_gpu_tests.log | cut -c1-300
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb --steps 10 --warmup 3 2> gpurun_out/r02_j21_bench_anchor.err | tail -1 > gpurun_out/r02_j21_bench_anchor_vitb_n1.json
cut -c1-330 gpurun_out/r02_j21_bench_anchor_vitb_n1.json; tail -2 gpurun_out/r02_j21_bench_anchor.err

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maskrcnn_gpu.py tests/test_edge_cases_gpu.py -q > gpurun_out/r02_j24_tests.log 2>&1; tail -8 gpurun_out/r02_j24_tests.log | cut -c1-300
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config maskrcnn_vitb --steps 10 --warmup 3 2> gpurun_out/r02_j24_bench_maskrcnn.err | tail -1 > gpurun_out/r02_j24_bench_maskrcnn_vitb_n1.json
cut -c1-200 gpurun_out/r02_j24_bench_maskrcnn_vitb_n1.json; tail -2 gpurun_out/r02_j24_bench_maskrcnn.err

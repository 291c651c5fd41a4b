set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r02_j28_gpu_tests.log 2>&1; tail -4 gpurun_out/r02_j28_gpu_tests.log | cut -c1-300

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_j25_gpu_tests.log 2>&1; tail -6 gpurun_out/r02_j25_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_j25_smoke.log 2>&1; tail -3 gpurun_out/r02_j25_smoke.log

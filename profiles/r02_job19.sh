set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maskrcnn_gpu.py -x -q -s > gpurun_out/r02_j19_maskrcnn_tests.log 2>&1; tail -25 gpurun_out/r02_j19_maskrcnn_tests.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_j19_gpu_tests.log 2>&1; tail -8 gpurun_out/r02_j19_gpu_tests.log
(cd rsprompter_b200 && timeout 300 ./rsp_selftest all > ../gpurun_out/r02_j19_selftest.log 2>&1; tail -3 ../gpurun_out/r02_j19_selftest.log)
(cd rsprompter_b200 && timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 ./rsp_selftest all > ../gpurun_out/r02_j19_synccheck_selftest.log 2>&1; echo "synccheck rc=$?" >> ../gpurun_out/r02_j19_synccheck_selftest.log)
grep -c "Barrier error" gpurun_out/r02_j19_synccheck_selftest.log; grep "Barrier error" -A 5 gpurun_out/r02_j19_synccheck_selftest.log | grep "Device Frame\|at " | sort | uniq -c | head -20; tail -4 gpurun_out/r02_j19_synccheck_selftest.log
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config maskrcnn_vitb --steps 10 --warmup 3 2> gpurun_out/r02_j19_bench_maskrcnn.err | tail -1 > gpurun_out/r02_j19_bench_maskrcnn_vitb_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_j19_bench_n1.err | tail -1 > gpurun_out/r02_j19_bench_query_vith_n1.json
for f in gpurun_out/r02_j19_bench_*.json; do echo $f; cut -c1-700 $f; echo; done
tail -3 gpurun_out/r02_j19_bench_maskrcnn.err

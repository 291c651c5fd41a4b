set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mask2former_gpu.py -q -s > gpurun_out/r02_j20_mask2former_tests.log 2>&1; tail -30 gpurun_out/r02_j20_mask2former_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_maskrcnn_gpu.py tests/test_query_gpu.py -q > gpurun_out/r02_j20_regress_tests.log 2>&1; tail -6 gpurun_out/r02_j20_regress_tests.log
(cd rsprompter_b200 && timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 ./rsp_selftest all > ../gpurun_out/r02_j20_synccheck_selftest.log 2>&1; echo "synccheck rc=$?" >> ../gpurun_out/r02_j20_synccheck_selftest.log)
grep -c "Barrier error" gpurun_out/r02_j20_synccheck_selftest.log; grep "Device Frame" gpurun_out/r02_j20_synccheck_selftest.log | sort | uniq -c | head; tail -4 gpurun_out/r02_j20_synccheck_selftest.log
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config mask2former_vitb --steps 10 --warmup 3 2> gpurun_out/r02_j20_bench_mask2former.err | tail -1 > gpurun_out/r02_j20_bench_mask2former_vitb_n1.json
cut -c1-400 gpurun_out/r02_j20_bench_mask2former_vitb_n1.json; tail -3 gpurun_out/r02_j20_bench_mask2former.err

set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_small.py > gpurun_out/r02_j17_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_j17_memcheck.log
tail -12 gpurun_out/r02_j17_memcheck.log
(cd rsprompter_b200 && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 ./rsp_selftest attn > ../gpurun_out/r02_j17_memcheck_selftest.log 2>&1; echo "memcheck rc=$?" >> ../gpurun_out/r02_j17_memcheck_selftest.log)
tail -6 gpurun_out/r02_j17_memcheck_selftest.log

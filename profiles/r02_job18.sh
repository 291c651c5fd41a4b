set -x
mkdir -p gpurun_out
(cd rsprompter_b200 && timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 ./rsp_selftest attn > ../gpurun_out/r02_j18_racecheck_selftest.log 2>&1; echo "racecheck rc=$?" >> ../gpurun_out/r02_j18_racecheck_selftest.log)
tail -15 gpurun_out/r02_j18_racecheck_selftest.log
(cd rsprompter_b200 && timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 ./rsp_selftest attn > ../gpurun_out/r02_j18_synccheck_selftest.log 2>&1; echo "synccheck rc=$?" >> ../gpurun_out/r02_j18_synccheck_selftest.log)
tail -6 gpurun_out/r02_j18_synccheck_selftest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_small.py > gpurun_out/r02_j18_racecheck_small.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_j18_racecheck_small.log
tail -12 gpurun_out/r02_j18_racecheck_small.log

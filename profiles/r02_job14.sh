set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_j14_pytest.log
tail -6 gpurun_out/r02_j14_pytest.log
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py 2> gpurun_out/r02_j14_bench_n1.err | tail -1 > gpurun_out/r02_j14_bench_query_vith_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb 2> gpurun_out/r02_j14_bench_anchor.err | tail -1 > gpurun_out/r02_j14_bench_anchor_vitb_n1.json
for f in gpurun_out/r02_j14_bench_*.json; do echo $f; cut -c1-200 $f; echo; done

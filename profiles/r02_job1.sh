set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_j1_smi.txt
timeout 1800 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r02_j1_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_j1_bench_query_vith.err | tail -1 > gpurun_out/r02_j1_bench_query_vith.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb --steps 10 --warmup 3 2> gpurun_out/r02_j1_bench_anchor_vitb.err | tail -1 > gpurun_out/r02_j1_bench_anchor_vitb.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_j1_launches_query_vith.csv python profiles/run_step.py --variant query --arch huge --steps 1 --warmup 1 > gpurun_out/r02_j1_ncu_query_vith.log 2>&1
tail -5 gpurun_out/r02_j1_pytest.log
cut -c1-1500 gpurun_out/r02_j1_bench_query_vith.json
tail -3 gpurun_out/r02_j1_bench_query_vith.err
cut -c1-600 gpurun_out/r02_j1_bench_anchor_vitb.json

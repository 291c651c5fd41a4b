set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench_n1.json
timeout 300 python profiles/run_step.py --variant query --steps 5 --warmup 2 | tail -1 > gpurun_out/final_query_vitb.json
timeout 400 python profiles/run_step.py --variant query --arch huge --steps 3 --warmup 2 | tail -1 > gpurun_out/final_query_vith.json
timeout 400 python profiles/run_step.py --variant anchor --arch huge --steps 3 --warmup 2 | tail -1 > gpurun_out/final_anchor_vith.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_traffic_anchor_final.csv python profiles/run_step.py --variant anchor --steps 1 --warmup 1 > gpurun_out/ncu_anchor.log 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_bench_n1.json gpurun_out/final_query_vitb.json gpurun_out/final_query_vith.json gpurun_out/final_anchor_vith.json | cut -c1-400

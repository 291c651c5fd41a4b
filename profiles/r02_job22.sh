set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_edge_cases_gpu.py tests/test_anchor_gpu.py tests/test_maskrcnn_gpu.py tests/test_mask2former_gpu.py -q > gpurun_out/r02_j22_tests.log 2>&1; tail -15 gpurun_out/r02_j22_tests.log | cut -c1-300
for v in maskrcnn mask2former; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_j22_launches_${v}_vitb.csv python profiles/run_step.py --variant $v --arch base --steps 1 --warmup 1 > gpurun_out/r02_j22_ncu_${v}.log 2>&1
tail -2 gpurun_out/r02_j22_ncu_${v}.log | cut -c1-300; wc -l gpurun_out/r02_j22_launches_${v}_vitb.csv
done

set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_j12_pytest.log
tail -4 gpurun_out/r02_j12_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_j12_smoke.log 2>&1; tail -1 gpurun_out/r02_j12_smoke.log
timeout 900 python bench.py 2> gpurun_out/r02_j12_bench_n1.err | tail -1 > gpurun_out/r02_j12_bench_query_vith_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb 2> gpurun_out/r02_j12_bench_anchor.err | tail -1 > gpurun_out/r02_j12_bench_anchor_vitb_n1.json
for S in 512 768 1024 1280; do
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size $S --steps 10 --warmup 3 2> gpurun_out/r02_j12_bench_enc_$S.err | tail -1 > gpurun_out/r02_j12_bench_encoder_vith_$S.json
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_j12_launches_query_vith.csv python profiles/run_step.py --variant query --arch huge --steps 1 --warmup 1 > gpurun_out/r02_j12_ncu_query_vith.log 2>&1
for f in gpurun_out/r02_j12_bench_*.json; do echo $f; cut -c1-200 $f; echo; done

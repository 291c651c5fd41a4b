set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_j8_pytest.log
(cd rsprompter_b200 && timeout 400 ./rsp_selftest attn bench > ../gpurun_out/r02_j8_selftest_attn.log 2>&1)
timeout 900 python bench.py 2> gpurun_out/r02_j8_bench_n1.err | tail -1 > gpurun_out/r02_j8_bench_query_vith_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config anchor_vitb 2> gpurun_out/r02_j8_bench_anchor.err | tail -1 > gpurun_out/r02_j8_bench_anchor_vitb_n1.json
for S in 512 768 1024 1280; do
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size $S --steps 10 --warmup 3 2> gpurun_out/r02_j8_bench_enc_$S.err | tail -1 > gpurun_out/r02_j8_bench_encoder_vith_$S.json
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_j8_launches_query_vith.csv python profiles/run_step.py --variant query --arch huge --steps 1 --warmup 1 > gpurun_out/r02_j8_ncu_query_vith.log 2>&1
(cd rsprompter_b200 && timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vit_window_attention_kernel<.int.80>' --launch-skip 1 -c 1 -f -o ../gpurun_out/r02_attn_window_hd80_v2 ./rsp_selftest attn bench > ../gpurun_out/r02_j8_ncu_attn_w.log 2>&1)
tail -4 gpurun_out/r02_j8_pytest.log
tail -5 gpurun_out/r02_j8_selftest_attn.log
for f in gpurun_out/r02_j8_bench_*.json; do echo $f; cut -c1-200 $f; echo; done

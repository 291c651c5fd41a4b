set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -80 > gpurun_out/r02_j3_pytest.log
cd rsprompter_b200
timeout 400 ./rsp_selftest attn bench > ../gpurun_out/r02_j3_selftest_attn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vit_attention_kernel<.int.80, .int.64>' --launch-skip 1 -c 1 -f -o ../gpurun_out/r02_attn_global_hd80 ./rsp_selftest attn bench > ../gpurun_out/r02_j3_ncu_attn_g.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vit_window_attention_kernel<.int.80>' --launch-skip 1 -c 1 -f -o ../gpurun_out/r02_attn_window_hd80 ./rsp_selftest attn bench > ../gpurun_out/r02_j3_ncu_attn_w.log 2>&1
cd ..
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j3_bench_n1.err | tail -1 > gpurun_out/r02_j3_bench_query_vith_n1.json
for S in 768 1280; do
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config encoder_vith --size $S --steps 3 --warmup 3 2> gpurun_out/r02_j3_bench_enc_$S.err | tail -1 > gpurun_out/r02_j3_bench_encoder_vith_$S.json
done
tail -30 gpurun_out/r02_j3_pytest.log
tail -6 gpurun_out/r02_j3_selftest_attn.log
for f in gpurun_out/r02_j3_bench_*.json; do echo $f; cut -c1-300 $f; echo; done

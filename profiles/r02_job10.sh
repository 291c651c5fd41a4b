set -x
mkdir -p gpurun_out
(cd rsprompter_b200 && timeout 300 ./rsp_selftest attn bench > ../gpurun_out/r02_j10_selftest_attn.log 2>&1; RSP_ATT_WINDOW_ROUNDS=1 timeout 300 ./rsp_selftest attn bench > ../gpurun_out/r02_j10_selftest_attn_rounds.log 2>&1)
tail -12 gpurun_out/r02_j10_selftest_attn.log; tail -5 gpurun_out/r02_j10_selftest_attn_rounds.log
timeout 2400 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r02_j10_pytest.log
tail -6 gpurun_out/r02_j10_pytest.log
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py 2> gpurun_out/r02_j10_bench_n1.err | tail -1 > gpurun_out/r02_j10_bench_query_vith_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size 1024 2> gpurun_out/r02_j10_bench_enc.err | tail -1 > gpurun_out/r02_j10_bench_encoder_vith_1024.json
(cd rsprompter_b200 && timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vit_window_attention2_kernel<.int.80>' --launch-skip 1 -c 1 -f -o ../gpurun_out/r02_attn_window2_hd80 ./rsp_selftest attn bench > ../gpurun_out/r02_j10_ncu_attn_w.log 2>&1)
for f in gpurun_out/r02_j10_bench_*.json; do echo $f; cut -c1-200 $f; echo; done

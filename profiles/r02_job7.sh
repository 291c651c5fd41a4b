set -x
mkdir -p gpurun_out
(cd rsprompter_b200 && timeout 400 ./rsp_selftest attn bench > ../gpurun_out/r02_j7_selftest_attn.log 2>&1)
tail -12 gpurun_out/r02_j7_selftest_attn.log
timeout 2400 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r02_j7_pytest.log
tail -8 gpurun_out/r02_j7_pytest.log
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j7_bench_n1.err | tail -1 > gpurun_out/r02_j7_bench_query_vith_n1.json
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size 1024 --steps 20 --warmup 3 2> gpurun_out/r02_j7_bench_enc.err | tail -1 > gpurun_out/r02_j7_bench_encoder_vith_1024.json
(cd rsprompter_b200 && timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vit_attention_kernel<.int.80, .int.64>' --launch-skip 1 -c 1 -f -o ../gpurun_out/r02_attn_global_hd80_v3 ./rsp_selftest attn bench > ../gpurun_out/r02_j7_ncu_attn_g.log 2>&1)
for f in gpurun_out/r02_j7_bench_*.json; do echo $f; cut -c1-300 $f; echo; done

set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_j5_pytest.log
(cd rsprompter_b200 && timeout 400 ./rsp_selftest attn bench > ../gpurun_out/r02_j5_selftest_attn.log 2>&1)
for S in 768 1280 1024; do
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --config encoder_vith --size $S --steps 10 --warmup 3 2> gpurun_out/r02_j5_bench_enc_$S.err | tail -1 > gpurun_out/r02_j5_bench_encoder_vith_$S.json
done
RSP_BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j5_bench_n1.err | tail -1 > gpurun_out/r02_j5_bench_query_vith_n1.json
tail -25 gpurun_out/r02_j5_pytest.log
tail -5 gpurun_out/r02_j5_selftest_attn.log
for f in gpurun_out/r02_j5_bench_*.json; do echo $f; cut -c1-300 $f; echo; done
tail -5 gpurun_out/r02_j5_bench_enc_1280.err

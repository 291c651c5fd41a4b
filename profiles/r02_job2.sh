set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_j2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r02_j2_bench_n2.err | tail -1 > gpurun_out/r02_j2_bench_query_vith_n2.json
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j2_bench_n1.err | tail -1 > gpurun_out/r02_j2_bench_query_vith_n1.json
for S in 512 1024; do
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --config encoder_vith --size $S --steps 20 --warmup 3 2> gpurun_out/r02_j2_bench_enc_$S.err | tail -1 > gpurun_out/r02_j2_bench_encoder_vith_$S.json
done
tail -25 gpurun_out/r02_j2_pytest.log
for f in gpurun_out/r02_j2_bench_*.json; do echo $f; cut -c1-400 $f; echo; done
tail -3 gpurun_out/r02_j2_bench_n2.err gpurun_out/r02_j2_bench_enc_512.err

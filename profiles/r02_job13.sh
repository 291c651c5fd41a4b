set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 4 2; do
RSP_BENCH_SKIP_ROOFLINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/r02_j13_bench_n$N.err | tail -1 > gpurun_out/r02_j13_bench_query_vith_n$N.json
done
RSP_BENCH_SKIP_CPU=1 RSP_BENCH_SKIP_ROOFLINE=1 timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j13_bench_n1.err | tail -1 > gpurun_out/r02_j13_bench_query_vith_n1.json
for f in gpurun_out/r02_j13_bench_*.json; do echo $f; cut -c1-260 $f; echo; done
tail -3 gpurun_out/r02_j13_bench_n4.err

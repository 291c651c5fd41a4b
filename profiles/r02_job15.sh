set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
RSP_BENCH_SKIP_ROOFLINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/r02_j15_bench_n8.err | tail -1 > gpurun_out/r02_j15_bench_query_vith_n8.json
RSP_BENCH_SKIP_CPU=1 RSP_BENCH_SKIP_ROOFLINE=1 timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/r02_j15_bench_n1.err | tail -1 > gpurun_out/r02_j15_bench_query_vith_n1.json
for f in gpurun_out/r02_j15_bench_*.json; do echo $f; cut -c1-260 $f; echo; done
tail -3 gpurun_out/r02_j15_bench_n8.err

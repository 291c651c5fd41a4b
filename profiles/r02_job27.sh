set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_query_gpu.py tests/test_decoder_gpu.py -q > gpurun_out/r02_j27_tests.log 2>&1; tail -5 gpurun_out/r02_j27_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -k "query" >> gpurun_out/r02_j27_tests.log 2>&1; tail -3 gpurun_out/r02_j27_tests.log | cut -c1-300
RSP_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_j27_bench.err | tail -1 > gpurun_out/r02_j27_bench_query_vith_n1.json
python -c "
import json
d=json.load(open('gpurun_out/r02_j27_bench_query_vith_n1.json')); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
for k in d['roofline']['top_kernels']:
    if 'mask_embed' in k['kernel'] or 'query_mask' in k['kernel']: print(k)
"

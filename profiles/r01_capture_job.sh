set -x
mkdir -p gpurun_out
timeout 200 python profiles/run_step.py --variant query --steps 5 --warmup 2 > gpurun_out/query_step.json 2> gpurun_out/query_step.err
timeout 200 python profiles/run_step.py --variant anchor --steps 5 --warmup 2 > gpurun_out/anchor_step.json 2>> gpurun_out/query_step.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_traffic_anchor.csv python profiles/run_step.py --variant anchor --steps 1 --warmup 1 > gpurun_out/ncu_anchor.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_traffic_query.csv python profiles/run_step.py --variant query --steps 1 --warmup 1 > gpurun_out/ncu_query.log 2>&1
cd rsprompter_b200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vit_attention -c 3 -f -o ../gpurun_out/r01_attn_full ./rsp_selftest attn bench > ../gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_v2 -c 6 -f -o ../gpurun_out/r01_gemm_full ./rsp_selftest gemmprof > ../gpurun_out/ncu_gemm.log 2>&1
cd ..
cat gpurun_out/query_step.json gpurun_out/anchor_step.json; tail -3 gpurun_out/query_step.err; ls -la gpurun_out

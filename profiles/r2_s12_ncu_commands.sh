# Final captures of round 2 (after the fp16 operand split of edge_mlp_tc / edge_tp_act_tc, the split-K node_chain tiles and the early
# query / time-embedding stream).  Run on the GPU box; summaries are made here with profiles/ncu_summary.py / launch_summary.py.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_s12_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --no-unfused > gpurun_out/r2_s12_ncu_bench.log 2>&1
timeout 300 $NCU -k regex:edge_mlp_tc_kernel -s 21 -c 1 -o gpurun_out/r2_s12_mlp_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:edge_tp_act_tc_kernel -s 21 -c 1 -o gpurun_out/r2_s12_tp_act_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:node_chain_kernel -s 21 -c 1 -o gpurun_out/r2_s12_node_chain python profiles/run_head_breakdown.py 128 > /dev/null 2>&1
ls -la gpurun_out/r2_s12_*.ncu-rep | awk '{print $5, $9}'

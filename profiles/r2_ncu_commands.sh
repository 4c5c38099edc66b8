mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# launch list of the bench step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --no-unfused > gpurun_out/r2_ncu_bench.log 2>&1
# top kernels of the C2 forward (eager, 2 forwards: the second one's launches after -s)
timeout 300 $NCU -k regex:node_chain_kernel -s 18 -c 3 -o gpurun_out/r2_node_chain python profiles/run_forward.py 2 128 > /dev/null 2>&1
timeout 300 $NCU -k regex:fps_cluster_kernel -s 1 -c 1 -o gpurun_out/r2_fps python profiles/run_forward.py 2 128 > /dev/null 2>&1
timeout 300 $NCU -k regex:radius_grid_kernel -s 6 -c 2 -o gpurun_out/r2_radius_grid python profiles/run_forward.py 2 128 > /dev/null 2>&1
timeout 300 $NCU -k regex:dual_linear_kernel -s 17 -c 2 -o gpurun_out/r2_dual_linear python profiles/run_forward.py 2 128 > /dev/null 2>&1
# the denoise step's kernels at 1024 poses (eager head steps: 3 warm-up + 10)
timeout 300 $NCU -k regex:head_front_kernel -s 4 -c 1 -o gpurun_out/r2_head_front python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:edge_mlp_tc_kernel -s 21 -c 1 -o gpurun_out/r2_mlp_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1   # (18 UNet launches + 3 warm-up steps first)
timeout 300 $NCU -k regex:edge_tp_act_tc_kernel -s 21 -c 1 -o gpurun_out/r2_tp_act_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:value_reduce_kernel -s 21 -c 1 -o gpurun_out/r2_value_reduce python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:score_tp_kernel -s 4 -c 1 -o gpurun_out/r2_score_tp python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:node_chain_kernel -s 21 -c 1 -o gpurun_out/r2_node_chain_head python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'

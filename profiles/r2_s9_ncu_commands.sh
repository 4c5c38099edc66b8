# End-of-round-2 captures (after the warp-uniform issue loops, the hardware-reciprocal sigmoid, the branch-free FPS and the
# lane-per-cell grid search).  Run on the GPU box; summaries are made here with profiles/ncu_summary.py / launch_summary.py.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2_s9_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --no-unfused > gpurun_out/r2_s9_ncu_bench.log 2>&1
timeout 300 $NCU -k regex:fps_cluster_kernel -s 1 -c 1 -o gpurun_out/r2_s9_fps python profiles/run_forward.py 2 128 > /dev/null 2>&1
timeout 300 $NCU -k regex:radius_grid_kernel -s 6 -c 2 -o gpurun_out/r2_s9_radius_grid python profiles/run_forward.py 2 128 > /dev/null 2>&1
timeout 300 $NCU -k regex:edge_mlp_tc_kernel -s 21 -c 1 -o gpurun_out/r2_s9_mlp_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:edge_tp_act_tc_kernel -s 21 -c 1 -o gpurun_out/r2_s9_tp_act_tc python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
timeout 300 $NCU -k regex:value_reduce_kernel -s 21 -c 1 -o gpurun_out/r2_s9_value_reduce python profiles/run_head_breakdown.py 1024 > /dev/null 2>&1
ls -la gpurun_out/r2_s9_*.ncu-rep | awk '{print $5, $9}'

#!/bin/bash
mkdir -p gpurun_out
for n in 4 8 12 16 24; do echo "blocks/SM $n"; HAVC_B200_PX_BLOCKS=$n timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/tmp_ops.json 2>&1 | grep -E "pre.h|post.h" ; done > gpurun_out/r2px_blocks.txt 2>&1
cat gpurun_out/r2px_blocks.txt
NCU="ncu --set full --clock-control none --import-source on --nvtx --nvtx-include target/ -f"
timeout 600 $NCU -o gpurun_out/r2px_ncu python tools/ncu_target.py --config cfg2 --ops pixel > gpurun_out/r2px_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2px_ncu.ncu-rep gpurun_out/r2px_ncu.txt "ncu --set full: pixel passes with the periodic horizontal kernels" "pre.h,pre.v,head,post.v,post.h"
ncu -i gpurun_out/r2px_ncu.ncu-rep --page details --csv 2>/dev/null | grep -iE "stall|Issue Slot|Eligible|No Eligible|Bank|Occupancy" | cut -c1-260 > gpurun_out/r2px_ncu_details.txt
ncu -i gpurun_out/r2px_ncu.ncu-rep --page raw --csv --metrics smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,gpu__time_duration.sum > gpurun_out/r2px_ncu_stalls.csv 2>/dev/null
rm -f gpurun_out/r2px_ncu.ncu-rep

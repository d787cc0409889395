#!/bin/bash
# ncu --set full of the whole-ray test-time renderer (run under gpurun from the repo root):
#   bash profiles/run_ncu_render.sh r02      ->  gpurun_out/r02_render.ncu-rep + gpurun_out/r02_render_plain.log
# The command trains the c2-like scene for 600 steps and renders 800x800 frames; the plain run must exit 0 first.  The
# capture takes the first launch of the pre-pass and of the persistent kernel on a full frame (640,000 rays).
# Summary on the CPU box:  python profiles/summarise_render.py r02  ->  profiles/r02_render_ncu_summary.md
set -e
R=${1:-r02}
export RR_ONLY=1
CMD="python scratch/render_rays_time.py 600"
L2M=lts__t_sectors.sum,lts__t_requests.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,sm__cycles_elapsed.max
$CMD > gpurun_out/${R}_render_plain.log 2>&1 &&
ncu --set full --metrics $L2M --clock-control none --import-source on -k regex:'render_rays_kernel|render_first_hit' -c 2 \
    -o gpurun_out/${R}_render $CMD > gpurun_out/${R}_render_ncu.log 2>&1
tail -3 gpurun_out/${R}_render_ncu.log; grep "whole_rays=True" gpurun_out/${R}_render_plain.log

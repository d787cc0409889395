#!/bin/bash
# ncu --set full of the first test-time render rounds (run under gpurun from the repo root):
#   bash profiles/run_ncu_render.sh r01e      ->  gpurun_out/r01e_prof.ncu-rep, summarised by profiles/summarise.py r01e
# The first round marches all 640,000 rays of a frame (thread-per-ray path of march_test_kernel).
set -e
R=${1:-r01e}
CMD="python scratch/render_profile.py 300"
$CMD > gpurun_out/${R}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'march_test|composite_test' -c 8 \
    -o gpurun_out/${R}_prof $CMD > gpurun_out/${R}_ncu_full.log 2>&1
tail -3 gpurun_out/${R}_ncu_full.log

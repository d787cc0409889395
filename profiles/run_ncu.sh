#!/bin/bash
# Profiling recipe for one round (run under gpurun from the repo root):  bash profiles/run_ncu.sh r01
# 1. plain run must exit 0;  2. launch list (per-launch device time, cold-cache: compare SHARES);  3. ncu --set
# full of one training step's worth of the hot kernels.  Outputs land in gpurun_out/ and are summarised by
# profiles/summarise.py into profiles/<round>_*.
set -e
R=${1:-r01}
CMD="python bench.py --steps 6 --warmup 3 --pretrain 96 --no-graph --skip-cpu"
$CMD > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'_kernel' -s 1000 -c 300 --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_list.log 2>&1
$CMD > gpurun_out/${R}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'march_train|march_scan|hashgrid|field_mlp|adam_kernel|composite_loss|ray_aabb|rays_from' -s 660 -c 12 \
    -o gpurun_out/${R}_prof $CMD > gpurun_out/${R}_ncu_full.log 2>&1
tail -3 gpurun_out/${R}_ncu_full.log

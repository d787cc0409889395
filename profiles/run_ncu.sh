#!/bin/bash
# Profiling recipe for one round (run under gpurun from the repo root):  bash profiles/run_ncu.sh r02 [c2|c5]
# bench.py --ncu-steps N trains to steady state (512 + 32 graph-replayed steps), then runs N eager training steps plus the
# two memory probes between cudaProfilerStart/Stop, so `--profile-from-start off` captures exactly those launches.
# 1. plain run must exit 0;  2. launch list (per-launch device time, cold-cache: compare SHARES);  3. ncu --set full
# + the L2 request / sector and L1 lookup counters the hash-grid roofline is built on.
# Outputs land in gpurun_out/ and are summarised on the CPU box by profiles/summarise.py into profiles/<round>_* and
# profiles/ncu_metrics.json (the per-unit counters bench.py cites).
set -e
R=${1:-r02}
CFG=${2:-c2}
CMD="python bench.py --config $CFG --ncu-steps 2 --skip-cpu"
L2M=lts__t_sectors.sum,lts__t_requests.sum,lts__t_sectors_srcunit_tex.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_lookup_hit.sum,lts__t_sectors_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__m_l1tex2xbar_req_cycles_active.sum,sm__cycles_elapsed.max
$CMD > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_list.log 2>&1
cp gpurun_out/ncu_units.json gpurun_out/${R}_ncu_units.json
$CMD > gpurun_out/${R}_plain2.log 2>&1 &&
ncu --set full --metrics $L2M --clock-control none --import-source on --profile-from-start off \
    -k regex:'hashgrid|field_mlp|adam_kernel|composite_loss|membench|march_train|march_scan' -c 24 \
    -o gpurun_out/${R}_prof $CMD > gpurun_out/${R}_ncu_full.log 2>&1
tail -3 gpurun_out/${R}_ncu_full.log

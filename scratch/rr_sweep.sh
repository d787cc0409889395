#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_render_rays.py -x -q --timeout 120 2>&1 | tail -5 > gpurun_out/rr_tests.log
RR_ONLY=1 timeout 200 python scratch/render_rays_time.py 1000 > gpurun_out/rr_time_default.log 2>&1
for v in ms16 ms32 as1 prof; do
  B2N_LIB=$PWD/google-nerf_b200/lib/libb2n_$v.so RR_ONLY=1 timeout 200 python scratch/render_rays_time.py 1000 > gpurun_out/rr_time_$v.log 2>&1
done
tail -3 gpurun_out/rr_tests.log; grep -H "whole_rays=True  first_hit=True\|kernels" gpurun_out/rr_time_*.log

#!/bin/bash
# ncu of the any-sigma 5x5 Gaussian and the general resize; band sweep of the any-sigma op
mkdir -p gpurun_out
OP=gaussq5 timeout 600 python scripts/bench_sustained.py default gauss.band_rows=60 gauss.band_rows=76 gauss.band_rows=116 > gpurun_out/r2z_q5_band_sweep.txt 2>&1; cut -c1-200 gpurun_out/r2z_q5_band_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GaussQ8Op -s 3 -c 1 -f -o gpurun_out/prof_gaussq5_r2z env OP=gaussq5 SECONDS_PER_SETTING=0.01 python scripts/bench_sustained.py default > gpurun_out/r2z_ncu_q5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resize_u8w -s 3 -c 1 -f -o gpurun_out/prof_resize_r2z env SECONDS_PER_CASE=0.01 python scripts/bench_all_kernels.py "1600x900" > gpurun_out/r2z_ncu_resize.log 2>&1
ls -la gpurun_out/*r2z*.ncu-rep

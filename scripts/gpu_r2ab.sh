#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/prof_gaussq5_r2z env OP=gaussq5 SECONDS_PER_SETTING=0.01 python scripts/bench_sustained.py default > gpurun_out/r2z_ncu_q5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/prof_yuyvgauss_r2z env SECONDS_PER_CASE=0.01 python scripts/bench_all_kernels.py "chain YUYV->BGR->GaussianBlur" > gpurun_out/r2z_ncu_yg.log 2>&1
ls -la gpurun_out/*r2z.ncu-rep

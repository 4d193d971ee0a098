#!/bin/bash
# ncu --set full captures of the kernels changed late in round 1 (1 GPU): cfg1 conversion, 3x3 binomial Gaussian, 2x resize
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_yuv422_vec -s 4 -c 1 -f -o gpurun_out/prof_cfg1_s3 python scripts/bench_configs.py cfg1 > /dev/null 2>&1; echo "cfg1 rc=$?"
KS=3 SIGMA=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss3_s3 python scripts/gauss_sweep.py "" > /dev/null 2>&1; echo "gauss3 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resize2x -s 4 -c 1 -f -o gpurun_out/prof_resize2x_s3 python scripts/bench_configs.py resizebatch > /dev/null 2>&1; echo "resize2x rc=$?"
ls -la gpurun_out/*_s3.ncu-rep

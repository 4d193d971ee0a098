#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_sepfilter -c 1 -s 3 -f -o gpurun_out/prof_sep11_r2e python scripts/bench_generic.py gauss11 > gpurun_out/r2e_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter2d -c 1 -s 3 -f -o gpurun_out/prof_f2d5_r2e python scripts/bench_generic.py f2d5 >> gpurun_out/r2e_ncu.log 2>&1
tail -5 gpurun_out/r2e_ncu.log

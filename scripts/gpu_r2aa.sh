#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "resize" > gpurun_out/r2aa_pytest.txt 2>&1; tail -2 gpurun_out/r2aa_pytest.txt
timeout 600 python scripts/bench_all_kernels.py resize > gpurun_out/r2aa_resize.txt 2>&1; cut -c1-200 gpurun_out/r2aa_resize.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GaussQ8Op -s 3 -c 1 -f -o gpurun_out/prof_gaussq5_r2z env OP=gaussq5 SECONDS_PER_SETTING=0.01 python scripts/bench_sustained.py default > gpurun_out/r2z_ncu_q5.log 2>&1
ls -la gpurun_out/*gaussq5*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_pytest.txt 2>&1; tail -2 gpurun_out/r2ad_pytest.txt
timeout 600 python scripts/bench_all_kernels.py YUYV > gpurun_out/r2ad_yuyv.txt 2>&1; cut -c1-180 gpurun_out/r2ad_yuyv.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_pytest.txt 2>&1; tail -2 gpurun_out/r2ai_pytest.txt
timeout 600 python scripts/bench_all_kernels.py "GaussianBlur" > gpurun_out/r2ai_kern.txt 2>&1; cut -c1-200 gpurun_out/r2ai_kern.txt

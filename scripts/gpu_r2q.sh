#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.txt 2>&1; tail -5 gpurun_out/r2q_pytest.txt
timeout 600 python scripts/bench_all_kernels.py "filter2D 5x5, 4K" 2>&1 | cut -c1-260

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.txt 2>&1; tail -5 gpurun_out/r2t_pytest.txt
timeout 600 python scripts/bench_all_kernels.py "WideOp" 2>&1 | cut -c1-240

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; tail -5 gpurun_out/r2o_pytest.txt
timeout 600 python scripts/bench_all_kernels.py "BGR f32" "gray f32" 2>&1 | cut -c1-260

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1; tail -3 gpurun_out/r2g_pytest.txt
timeout 600 python scripts/bench_sustained.py default > gpurun_out/r2g_sustained.txt 2>&1
OP=gauss3 timeout 300 python scripts/bench_sustained.py default >> gpurun_out/r2g_sustained.txt 2>&1
cat gpurun_out/r2g_sustained.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.txt 2>&1; tail -6 gpurun_out/r2k_pytest.txt
timeout 600 python scripts/bench_generic.py gauss filter2d > gpurun_out/r2k_generic.txt 2>&1; cat gpurun_out/r2k_generic.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_generic_filters_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "filter or gauss or sep or generic" > gpurun_out/r2f_pytest.txt 2>&1; tail -4 gpurun_out/r2f_pytest.txt
timeout 600 python scripts/bench_generic.py gauss filter2d > gpurun_out/r2f_generic.txt 2>&1; cat gpurun_out/r2f_generic.txt

#!/bin/bash
# round-2 re-entry: transposed-form Gauss5Op + leaner strip skeleton: parity, sustained, data-content power probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gauss or strip or sobel or filter or sep or chain or yuyv or geometry" > gpurun_out/r2w_pytest.txt 2>&1; tail -3 gpurun_out/r2w_pytest.txt
timeout 300 python scripts/bench_sustained.py default > gpurun_out/r2w_sustained.txt 2>&1; tail -2 gpurun_out/r2w_sustained.txt
OP=gauss3 timeout 300 python scripts/bench_sustained.py default >> gpurun_out/r2w_sustained.txt 2>&1; tail -1 gpurun_out/r2w_sustained.txt
timeout 600 python scripts/bench_power_data.py > gpurun_out/r2w_power_data.txt 2>&1; cat gpurun_out/r2w_power_data.txt | tail -8

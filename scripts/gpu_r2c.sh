#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "gauss" > gpurun_out/r2c_pytest.txt 2>&1; tail -3 gpurun_out/r2c_pytest.txt
timeout 600 python scripts/bench_sustained.py default gauss.variant=1 gauss.band_rows=60 gauss.variant=1,gauss.band_rows=60 gauss.band_rows=28 gauss.band_rows=76 gauss.variant=1,gauss.band_rows=100 > gpurun_out/r2c_sustained.txt 2>&1
OP=swap timeout 300 python scripts/bench_sustained.py default >> gpurun_out/r2c_sustained.txt 2>&1
OP=gauss3 timeout 300 python scripts/bench_sustained.py default >> gpurun_out/r2c_sustained.txt 2>&1
cat gpurun_out/r2c_sustained.txt

#!/bin/bash
# quick kernel iteration: gaussian parity subset, sweep, ncu full capture of the sweep's default config
mkdir -p gpurun_out
TAG=${1:-it}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "gauss or sobel or batch or cfg2" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
shift
timeout 600 python scripts/gauss_sweep.py "$@" 2>&1 | tee gpurun_out/sweep_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python scripts/gauss_sweep.py "" > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"

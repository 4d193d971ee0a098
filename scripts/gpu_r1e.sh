#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x -k "resize or banded or batch" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_generic.py resize warp 2>&1 | tee gpurun_out/generic_r1e.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x -k "yuyv or cvt or conver or reference or exhaust or decode" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_configs.py cfg1 2>&1 | cut -c1-330 | tee gpurun_out/configs_r1g.txt
timeout 300 python scripts/bench_generic.py cvt 2>&1 | tee gpurun_out/generic_cvt_r1g.txt

#!/bin/bash
mkdir -p gpurun_out
TAG=$1; shift
timeout 900 python scripts/gauss_sweep.py "$@" 2>&1 | tee gpurun_out/sweep_${TAG}.log

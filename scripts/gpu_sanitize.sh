#!/bin/bash
# compute-sanitizer over the parity tests of the kernels added this session (memcheck: all; racecheck: strip/tile kernels)
mkdir -p gpurun_out
K='yuyv_to or banded or warp or resize_larger or cvt or fused'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider -k "yuyv_to or banded or warp_affine" > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.txt

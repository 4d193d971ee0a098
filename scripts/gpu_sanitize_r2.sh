#!/bin/bash
# compute-sanitizer over the kernels / host paths added in round 2
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_generic_filters_gpu.py tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_generic_filters_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "general_kernel or tiny_images_and_batches or gaussian5_binomial or band_seams or sobel" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.txt

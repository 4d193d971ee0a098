#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; tail -3 gpurun_out/r2r_pytest.txt
timeout 900 python scripts/bench_all_kernels.py > gpurun_out/r2_all_kernels.txt 2>&1; tail -3 gpurun_out/r2_all_kernels.txt | cut -c1-200
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_generic_filters_gpu.py -m gpu -q -x -p no:cacheprovider -k "wide or multichannel" > gpurun_out/r2_sanitizer_memcheck_b.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_b.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_generic_filters_gpu.py -m gpu -q -x -p no:cacheprovider -k "wide or multichannel" > gpurun_out/r2_sanitizer_racecheck_b.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck_b.txt

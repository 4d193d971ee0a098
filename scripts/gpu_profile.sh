#!/bin/bash
# ncu evidence for the metric kernel: launch list + one full-set capture (1 GPU only).
mkdir -p gpurun_out
TAG=${1:-r1}
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 3 --warmup 3 --e2e-frames 2 --no-cpu > gpurun_out/ncu_launch_bench_${TAG}.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_launch_bench_${TAG}.log | cut -c1-300
echo "== full set, k_strip"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 1 -c 2 -f -o gpurun_out/prof_${TAG} \
  python bench.py --steps 3 --warmup 3 --e2e-frames 2 --no-cpu > gpurun_out/ncu_full_bench_${TAG}.log 2>&1
echo "rc=$?"; ls -la gpurun_out/

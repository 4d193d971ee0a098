#!/bin/bash
# one ncu --set full capture per secondary kernel (1 GPU), for profiles/
mkdir -p gpurun_out
TAG=${1:-r1}
for cfg in cfg1 cfg3 cfg4 cfg5; do
  case $cfg in
    cfg1) K=k_yuv422_vec;; cfg3) K=k_strip;; cfg4) K=k_resize4x;; cfg5) K=k_warp_f32_tile;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${cfg}_${TAG} \
     python scripts/bench_configs.py $cfg > gpurun_out/ncu_${cfg}_${TAG}.log 2>&1
  echo "$cfg rc=$?"
done
KS=7 SIGMA=1.5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gaussq8k7_${TAG} python scripts/gauss_sweep.py "" > gpurun_out/ncu_q8k7_${TAG}.log 2>&1; echo "q8k7 rc=$?"
ls -la gpurun_out/*.ncu-rep

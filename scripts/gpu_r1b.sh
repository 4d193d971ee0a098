#!/bin/bash
# Session-3 GPU visit: smoke, full GPU parity suite, bench line, fused-chain + warp tile sweep, two ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== configs"; timeout 900 python scripts/bench_configs.py chain cfg5sweep cfg3 2>&1 | tee gpurun_out/configs_r1b.txt
echo "== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_chain_r1b python scripts/bench_configs.py chain > gpurun_out/ncu_chain.log 2>&1; echo "ncu chain rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_warp_tile -s 4 -c 1 -f -o gpurun_out/prof_cfg5_r1b python scripts/bench_configs.py cfg5 > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
ls -la gpurun_out/*.ncu-rep

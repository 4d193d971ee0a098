#!/bin/bash
# Round-end evidence: smoke, full GPU suite, bench line (N=1), reference arm, configs, host path, launch list, ncu captures.
mkdir -p gpurun_out
TAG=${1:-r1f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref_${TAG}.json
echo "== configs"; CPU=1 timeout 1200 python scripts/bench_configs.py chain_gauss chain cfg1 cfg3 cfg4 cfg5 2>&1 | tee gpurun_out/configs_${TAG}.txt
echo "== host"; timeout 600 python scripts/bench_host.py 2>&1 | tee gpurun_out/host_${TAG}.txt
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 3 --warmup 3 --e2e-frames 2 --no-cpu > gpurun_out/ncu_launch_bench_${TAG}.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 1 -c 1 -f -o gpurun_out/prof_gauss5_${TAG} \
  python bench.py --steps 3 --warmup 3 --e2e-frames 2 --no-cpu > gpurun_out/ncu_full_bench_${TAG}.log 2>&1; echo "gauss5 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_chain_gauss_${TAG} python scripts/bench_configs.py chain_gauss > /dev/null 2>&1; echo "chain_gauss rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_chain_sobel_${TAG} python scripts/bench_configs.py chain > /dev/null 2>&1; echo "chain rc=$?"
ls -la gpurun_out/*.ncu-rep

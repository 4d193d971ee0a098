#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest.txt 2>&1; tail -2 gpurun_out/r2ac_pytest.txt
timeout 300 python scripts/bench_sustained.py default > gpurun_out/r2ac_sustained.txt 2>&1; cut -c1-220 gpurun_out/r2ac_sustained.txt
timeout 900 python scripts/bench_all_kernels.py > gpurun_out/r2ac_all_kernels.txt 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2ac_all_kernels.txt"):
    try: k=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f'{k["burst_frac"]:.3f} {k["sustained_frac"]:.3f} {k["sm_mhz"]}  {k["case"][:90]}')
PY
BARGS="--steps 3 --warmup 3 --e2e-frames 2 --no-cpu --no-extra --sustained-s 0 --no-calls"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2ac python bench.py $BARGS > gpurun_out/r2ac_ncu_full.log 2>&1

#!/bin/bash
# round-2 re-entry, step 2: horizontal-first any-sigma Gaussians, transposed-form fused YUYV chain: full GPU suite + all kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.txt 2>&1; tail -3 gpurun_out/r2x_pytest.txt
timeout 900 python scripts/bench_all_kernels.py > gpurun_out/r2x_all_kernels.txt 2>&1; wc -l gpurun_out/r2x_all_kernels.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2x_all_kernels.txt"):
    try: j=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f'{j["burst_frac"]:.3f} {j["sustained_frac"]:.3f} {j["sm_mhz"]}  {j["case"][:90]}')
PY

#!/bin/bash
# round 2, call B (1 GPU): tests after the cross-item prefetch, bench with the sustained copy reference, pageable sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1
tail -5 gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --no-cpu > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 600 gpurun_out/r2b_bench_n1.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2b_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "frac", round(j["roofline"]["frac"],4))
print("sustained", json.dumps(j["roofline"]["sustained"]))
print("calls", j.get("calls"))
for x in j.get("extra_configs") or []:
    print("  ", x.get("config"), x.get("ms_per_step"), x.get("roofline",{}).get("frac"), x.get("clocks",{}).get("sm_mhz"), x.get("parity"), x.get("error"))
PY
timeout 600 python scripts/bench_pageable.py > gpurun_out/r2b_pageable.txt 2>&1
cat gpurun_out/r2b_pageable.txt

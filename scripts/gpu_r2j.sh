#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.txt 2>&1; tail -3 gpurun_out/r2j_pytest.txt
timeout 600 python bench.py --no-cpu --no-extra > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 300 gpurun_out/r2j_bench.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", round(j["roofline"]["sustained"]["frac"],4))
print({k:v for k,v in j["calls"].items() if k!="how"})
PY
OP=sobel NF=64 timeout 300 python scripts/bench_sustained.py default strip.pdl=0 2>&1 | tail -2
OP=gauss5 timeout 300 python scripts/bench_sustained.py default strip.pdl=0 2>&1 | tail -2

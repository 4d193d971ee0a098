#!/bin/bash
# final record for round 2 (1 GPU): smoke, full GPU suite, bench line + reference arm, every kernel family with clocks
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.txt 2>&1; tail -2 gpurun_out/r2u_pytest.txt
timeout 900 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; tail -c 300 gpurun_out/r2u_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r2u_bench_ref.json 2>&1
timeout 900 python scripts/bench_all_kernels.py > gpurun_out/r2_all_kernels.txt 2>&1; wc -l gpurun_out/r2_all_kernels.txt
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2u_bench_n1.json").read().strip().splitlines()[-1])
s=j["roofline"]["sustained"]
print("value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", round(s["frac"],4), s["clocks"]["sm_mhz"], "copy", round(s["plain_copy_sustained"]["gbs"]))
e=j["e2e"]; print("e2e", round(e["value"]), round(e["frac_of_link_all_ranks"],3))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j["calls"].items() if k!="how"})
for x in j["extra_configs"]:
    r=x.get("roofline",{}); print("  ", x.get("config")[:60], round(r.get("frac",0),3), round(r.get("sustained",{}).get("frac",0),3), x.get("parity"), x.get("error"))
print(j["cpu_baseline"]["value"], j["cpu_baseline"]["definition_port"]["value"], j["cpu_baseline"]["cores"])
PY

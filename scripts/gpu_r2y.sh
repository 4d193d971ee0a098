#!/bin/bash
# round-2 re-entry, step 3: compact resize table; band sweep of the new metric kernel; the bench line with the new sustained records; ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "resize" > gpurun_out/r2y_pytest.txt 2>&1; tail -2 gpurun_out/r2y_pytest.txt
timeout 600 python scripts/bench_all_kernels.py resize > gpurun_out/r2y_resize.txt 2>&1; cut -c1-200 gpurun_out/r2y_resize.txt
timeout 600 python scripts/bench_sustained.py default gauss.band_rows=44 gauss.band_rows=60 gauss.band_rows=76 > gpurun_out/r2y_band_sweep.txt 2>&1; cut -c1-230 gpurun_out/r2y_band_sweep.txt
timeout 900 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; tail -c 300 gpurun_out/r2y_bench_n1.err
BARGS="--steps 3 --warmup 3 --e2e-frames 2 --no-cpu --no-extra --sustained-s 0 --no-calls"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches.csv python bench.py $BARGS > gpurun_out/r2y_ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2y python bench.py $BARGS > gpurun_out/r2y_ncu_full.log 2>&1
ls -la gpurun_out/*r2y*.ncu-rep
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2y_bench_n1.json").read().strip().splitlines()[-1])
s=j["roofline"]["sustained"]
print("value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", round(s["frac"],4), s["clocks"]["sm_mhz"], "copy", round(s["plain_copy_sustained"]["gbs"]), s["plain_copy_sustained"]["clocks"]["sm_mhz"], "copy0", round(s["plain_copy_sustained_zeros"]["gbs"]), "zero", round(s["zero_content"]["frac"],4), s["zero_content"]["clocks"])
e=j["e2e"]; print("e2e", round(e["value"]), round(e["frac_of_link_all_ranks"],3))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j["calls"].items() if k!="how"})
for x in j["extra_configs"]:
    r=x.get("roofline",{}); print("  ", x.get("config")[:60], round(r.get("frac",0),3), round(r.get("sustained",{}).get("frac",0),3), x.get("parity"), x.get("error"))
PY

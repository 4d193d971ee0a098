#!/bin/bash
# round-2 record after the re-entry session's kernel work (1 GPU): smoke, full GPU suite, bench line + reference arm,
# every kernel family with clocks, ncu launch list + full captures (free-running and base clocks), sanitizers
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.txt 2>&1; tail -1 gpurun_out/r2f_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; tail -2 gpurun_out/r2f_pytest.txt
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 300 gpurun_out/r2f_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2>&1
timeout 900 python scripts/bench_all_kernels.py > gpurun_out/r2f_all_kernels.txt 2>&1; wc -l gpurun_out/r2f_all_kernels.txt
BARGS="--steps 3 --warmup 3 --e2e-frames 2 --no-cpu --no-extra --sustained-s 0 --no-calls"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py $BARGS > gpurun_out/r2f_ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2f python bench.py $BARGS > gpurun_out/r2f_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control base --import-source on -k regex:k_strip -s 4 -c 1 -f -o gpurun_out/prof_gauss5_r2f_baseclk python bench.py $BARGS > gpurun_out/r2f_ncu_full_base.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resize_u8w -s 3 -c 1 -f -o gpurun_out/prof_resize_r2f env SECONDS_PER_CASE=0.01 python scripts/bench_all_kernels.py "1600x900" > gpurun_out/r2f_ncu_resize.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_generic_filters_gpu.py -m gpu -q -x -p no:cacheprovider -k "gauss or resize or yuyv_to or chain or sobel or sep or band_seams or tiny or dense or filter2d or warp" > gpurun_out/r2f_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2f_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_generic_filters_gpu.py -m gpu -q -x -p no:cacheprovider -k "gaussian5_binomial or band_seams or yuyv_to or any_sigma or gaussq or sobel" > gpurun_out/r2f_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2f_sanitizer_racecheck.txt
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2f_bench_n1.json").read().strip().splitlines()[-1])
s=j["roofline"]["sustained"]
print("value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", round(s["frac"],4), s["clocks"]["sm_mhz"], "copy", round(s["plain_copy_sustained"]["gbs"]), s["plain_copy_sustained"]["clocks"]["sm_mhz"], "copy0", round(s["plain_copy_sustained_zeros"]["gbs"]), "zero", round(s["zero_content"]["frac"],4))
e=j["e2e"]; print("e2e", round(e["value"]), round(e["frac_of_link_all_ranks"],3))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j["calls"].items() if k!="how"})
for x in j["extra_configs"]:
    r=x.get("roofline",{}); print("  ", x.get("config")[:60], round(r.get("frac",0),3), round(r.get("sustained",{}).get("frac",0),3), x.get("parity"), x.get("error"))
print(j["cpu_baseline"]["value"], j["cpu_baseline"]["definition_port"]["value"], j["cpu_baseline"]["cores"])
for l in open("gpurun_out/r2f_all_kernels.txt"):
    try: k=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f'{k["burst_frac"]:.3f} {k["sustained_frac"]:.3f} {k["sm_mhz"]}  {k["case"][:90]}')
PY

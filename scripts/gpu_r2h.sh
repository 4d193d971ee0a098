#!/bin/bash
# N-GPU box: link probe, multi-GPU tests, bench under torchrun at N = all GPUs, in-library multi-GPU e2e
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 300 scripts/_link_probe 256 > gpurun_out/r2h_link_probe_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q > gpurun_out/r2h_pytest_n$N.txt 2>&1; tail -3 gpurun_out/r2h_pytest_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
tail -c 400 gpurun_out/r2h_bench_n$N.err
timeout 300 python bench.py --impl reference --gpus $N > gpurun_out/r2h_bench_ref_n$N.json 2>&1
timeout 400 python bench.py --no-extra --no-cpu --sustained-s 0 --no-calls --multi-gpus $N > gpurun_out/r2h_bench_multi$N.json 2> gpurun_out/r2h_bench_multi$N.err
tail -c 400 gpurun_out/r2h_bench_multi$N.err
python - <<PY
import json
for f in ("r2h_bench_n$N","r2h_bench_multi$N"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=j["e2e"]; print(f, "n", j["n_gpus"], "value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", j["roofline"]["sustained"] and round(j["roofline"]["sustained"]["frac"],4),
              "e2e", round(e["value"]), "link_all", round(e["link_all_ranks"]["duplex_gbs_each_total"],1), "frac_link", round(e["frac_of_link_all_ranks"],3), "multi", e.get("one_process_multi_gpu"))
        for x in j.get("extra_configs") or []:
            r=x.get("roofline",{}); print("  ", x.get("config"), round(x.get("ms_per_step",0),4), round(r.get("frac",0),3), round(r.get("sustained",{}).get("frac",0),3), x.get("parity"), x.get("error"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat gpurun_out/r2h_link_probe_n$N.txt

#!/bin/bash
# round 2, call A (2 GPUs): topology + link ceiling, the GPU test suite, bench at N=1 / N=2, in-library multi-GPU e2e
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1
nvidia-smi -q -i 0 | grep -A12 "PCI" | head -40 >> gpurun_out/r2a_topo.txt 2>&1
lscpu | head -30 >> gpurun_out/r2a_topo.txt 2>&1
timeout 300 scripts/_link_probe 256 > gpurun_out/r2a_link_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 600 gpurun_out/r2a_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err
tail -c 600 gpurun_out/r2a_bench_n2.err
timeout 300 python bench.py --no-extra --no-cpu --sustained-s 0 --multi-gpus 2 > gpurun_out/r2a_bench_multi2.json 2> gpurun_out/r2a_bench_multi2.err
tail -c 600 gpurun_out/r2a_bench_multi2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>&1
python - <<'PY'
import json
for f in ("r2a_bench_n1","r2a_bench_n2","r2a_bench_multi2"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=j["e2e"]; print(f, "value", round(j["value"]), "frac", round(j["roofline"]["frac"],4), "sus", j["roofline"]["sustained"] and round(j["roofline"]["sustained"]["frac"],4),
              "e2e", round(e["value"]), "link_all", e["link_all_ranks"]["duplex_gbs_each_total"], "frac_link", round(e["frac_of_link_all_ranks"],3), "multi", e.get("one_process_multi_gpu"))
        print(" calls", j.get("calls"))
        for x in j.get("extra_configs") or []:
            print("  ", x.get("config"), x.get("ms_per_step"), x.get("roofline",{}).get("frac"), x.get("parity"), x.get("error"))
    except Exception as ex:
        print(f, "ERR", ex)
PY

"""Turns an .ncu-rep into the text summary kept under profiles/ (run in the build container).
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/NAME [units_per_launch]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = int(sys.argv[3]) if len(sys.argv) > 3 else 32 * 2160 * 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# ncu --set full --clock-control none summary of {rep}", ""]
summ = []
for r in data:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i], unit[i])
            lines.append(f"{k:70s} {r[i]} {unit[i]}")
    lines.append("")
    summ.append(d)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
agg = subprocess.run([sys.executable, "scripts/ncu_src.py", "/tmp/_src.csv", str(units * len(data))], capture_output=True, text=True).stdout
lines += ["# executed warp instructions by opcode (source page, all captured launches; unit = one warp-row of 512 B in / 480 B out)", agg]
open(out + ".txt", "w").write("\n".join(lines))
print("\n".join(lines[:40]))
d0 = summ[0]
rd = float(d0["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d0["dram__bytes_read.sum"][1]]
wr = float(d0["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d0["dram__bytes_write.sum"][1]]
json.dump({"source": rep, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr}, open(out + ".json", "w"), indent=1)

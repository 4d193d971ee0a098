"""Aggregates an `ncu --page source --csv` dump: executed instructions by opcode, stall samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = int(sys.argv[2]) if len(sys.argv) > 2 else 32 * 2160 * 24  # warp-rows per launch
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) >= len(rows[hi]) and r[0].startswith("0x")]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
byop, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    e = int(r[iE])
    tot += e
    toks = [o for o in r[iS].strip().split() if not o.startswith("@")]
    op = toks[0].split(".")[0]
    byop[op] += e
    samp[op] += int(r[iSm])
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i])
print("total warp instr", tot, "per unit", round(tot / units, 1))
for op, c in byop.most_common():
    print(f"{op:12s} {c:12d} {100 * c / tot:5.1f}%  per-unit {c / units:6.2f}  samples {samp[op]}")
ts = sum(stalls.values())
print([(k, round(100 * v / ts, 1)) for k, v in stalls.most_common(10)])

"""profiles/sass_summary.txt: per kernel of the shipped librcv_imgproc.so -- registers, spills, shared memory and the
counts of the SASS mnemonics that show how it moves data (UTMALDG = TMA tile load, SYNCS = mbarrier,
LDS.128 / ST.E.128 / LDG.E.128 = 128-bit accesses, SHFL = warp shuffle, PRMT = byte permute, UTC*MMA / LDTM =
tensor-core paths, expected to be 0 here).  Runs in the build container (cuobjdump, no GPU).
    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rustcv_b200", "librcv_imgproc.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n

usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", line)}
        cur = None

KEYS = ["UTMALDG", "SYNCS", "LDS.128", "LDS", "STS", "ST.E.128", "LDG.E.128", "LDG", "ST.E", "SHFL", "PRMT", "IMAD", "FFMA",
        "VIMNMX", "MUFU", "ATOMG", "UTCHMMA", "UTCIMMA", "LDTM", "HMMA", "IMMA"]
rows = []
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
    cnt = {}
    for k in KEYS:
        if "." in k:
            cnt[k] = sum(1 for o in ops if o.startswith(k))
        else:
            cnt[k] = sum(1 for o in ops if o.split(".")[0] == k)
    dn = demangle(name)
    if "rcv::" not in dn:  # nvJPEG's own statically linked kernels (the MJPEG branch's entropy decode / IDCT) are not ours
        continue
    rows.append((dn, name, len(ops), cnt))

print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass / -res-usage; sm_100a); srchash "
      f"{open(LIB + '.srchash').read().strip()[:16] if os.path.exists(LIB + '.srchash') else '?'}")
print("# static instruction counts per kernel; REG = registers per thread, STACK > 0 = spills / local arrays\n")
hdr = ["kernel", "instr", "REG", "STACK", "SHARED"] + KEYS
print(" | ".join(hdr))
tot = {k: 0 for k in KEYS}
for dn, mn, n, cnt in sorted(rows, key=lambda r: r[0]):
    u = usage.get(mn, {})
    short = re.sub(r"^void ", "", dn)
    short = short[:short.rfind(">(") + 1] if ">(" in short else re.sub(r"\(.*$", "", short)
    short = short.replace("rcv::", "").replace("(int)", "").replace("(bool)", "").replace("unsigned char", "u8")
    print(" | ".join([short, str(n), str(u.get("REG", "?")), str(u.get("STACK", "?")), str(u.get("SHARED", "?"))] + [str(cnt[k]) for k in KEYS]))
    for k in KEYS:
        tot[k] += cnt[k]
print("\n# all rcv:: kernels: " + ", ".join(f"{k} {tot[k]}" for k in KEYS))

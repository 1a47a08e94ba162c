"""Split the SASS of a profiled step kernel at its CTA barriers (staging | rules phase | encoder) and print, per segment,
the warp instructions executed and the stall samples by reason.   python tools/ncu_segments.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {}
for i, k in enumerate(hdr):
    ix.setdefault(k, i)
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[1]]
stall = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
# step kernels: barrier after the (chain-only) tile wait, after the table staging, after the rules phase
nb = len(bars)
names = (["prologue", "state fetch + staging", "rules phase", "encoder"] if nb >= 3 else ["staging", "rules phase", "encoder"]) + ["after the last barrier"] * 8
prev = 0
for si, b in enumerate(bars + [len(data)]):
    seg = data[prev:b]
    n = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    tot = {k: sum(int(r[ix[k]]) for r in seg) for k in stall}
    s = sum(tot.values())
    top = sorted(tot.items(), key=lambda kv: -kv[1])[:7]
    print(f"{names[si]:12s} SASS lines {len(seg):6d}  warp instructions {n:11d}  stall samples {s:6d}  " + ", ".join(f"{k[6:]} {v}" for k, v in top))
    prev = b

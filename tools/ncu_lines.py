"""Read a .ncu-rep (ncu --set full --import-source on, built with -lineinfo) and print the source lines that execute
the most instructions / collect the most stall samples.   python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
cur, hdr, data = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif len(r) >= 2 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)                      # the first "Source" column is the CUDA line
        d["file"] = cur
        data.append(d)


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


ti = sum(f(d["Instructions Executed"]) for d in data)
ts = sum(f(d["Warp Stall Sampling (All Samples)"]) for d in data)
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
stall_cols = [k for k in data[0] if k.startswith("stall_") and "Not Issued" not in k]
for key, name in (("Instructions Executed", "instructions"), ("Warp Stall Sampling (All Samples)", "stall samples")):
    print(f"---- top lines by {name}")
    for d in sorted(data, key=lambda d: -f(d[key]))[:top]:
        st = sorted(((f(d[c]), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{d['file'][:18]:18s} {d['Line No']:>5s} inst {f(d['Instructions Executed']) / ti * 100:5.1f}%  stall {f(d['Warp Stall Sampling (All Samples)']) / ts * 100:5.1f}% "
              f"[{st[0][1]} {st[1][1]}]  {d['Source'].strip()[:100]}")

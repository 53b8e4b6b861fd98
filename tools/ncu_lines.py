#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []; fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] == "Line No": continue
    if r[2] != "-":   # sass rows
        continue
    try: data.append((int(r[7]), int(r[6]), int(r[8]), fname, r[0], r[1].strip()[:100]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1; tots = sum(d[1] for d in data) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/tot*100:5.1f}% inst {d[1]/tots*100:5.1f}% smp  thr/inst {d[2]/max(d[0],1):5.1f}  {d[3]}:{d[4]:>4s} {d[5]}")

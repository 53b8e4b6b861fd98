#!/usr/bin/env python
"""Per-source-line stall-reason samples of an .ncu-rep (source page; needs -lineinfo + --import-source on).
usage: python tools/ncu_stalls.py rep.ncu-rep reason[,reason...] [top_n]     e.g. stall_no_inst,stall_barrier"""
import csv, subprocess, sys
rep, reasons = sys.argv[1], sys.argv[2].split(","); top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; fname = ""; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: data.append((fname, r[0], r[1].strip()[:90], {k: int(r[hdr.index(k)] or 0) for k in reasons}, int(r[6] or 0)))
    except ValueError: pass
for k in reasons:
    tot = sum(d[3][k] for d in data) or 1
    print(f"== {k}: {tot} samples of {sum(d[4] for d in data)}")
    for d in sorted(data, key=lambda d: -d[3][k])[:top]:
        print(f"  {d[3][k] / tot * 100:5.1f}%  {d[0]}:{d[1]:>4s} {d[2]}")

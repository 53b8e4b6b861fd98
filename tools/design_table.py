#!/usr/bin/env python
"""tools/design_table.py -- the measurement table of DESIGN.md 4 from the committed bench lines under profiles/ (so that the
document quotes what the files say).  usage: python tools/design_table.py [round-tag]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def line(name):
    p = os.path.join(ROOT, "profiles", f"{TAG}_{name}.json")
    if not os.path.exists(p):
        return None
    txt = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(txt[-1]) if txt else None


def M(v):
    return f"{v / 1e6:.2f} M" if v >= 1e6 else f"{v / 1e3:.1f} k"


def kern(d):
    k = d["kernel_ms_per_step"]
    names = {"int_search": "k_int_search", "subpel_refine": "sub-pel", "epzs": "k_epzs_int", "mc_tq": "mc_tq", "subpel_planes": "planes", "gen_requests": "gen",
             "chroma": "chroma"}
    return ", ".join(f"{names.get(a, a)} {b:.3f}" for a, b in sorted(k.items(), key=lambda t: -t[1]) if b > 0.0005)


rows = []
for cfg, nm, label in ((2, "bench", "config 2: 1080p FullSearch ±32 + SATD sub-pel + 4×4 TQ of 7 modes"),
                       (3, "bench_c3", "config 3: 4K EPZS + sub-pel + 8×8 TQ (High)"),
                       (4, "bench_c4", "config 4: 1080p 4:2:2, fast full search + SATD sub-pel + 4×4 TQ + 4:2:2 chroma path")):
    d = line(nm)
    if not d:
        continue
    e = d["e2e"]; r = d["roofline"]; c = d.get("cpu_baseline", {})
    ref = line(nm + "_reference")
    rows.append(f"| 1 GPU, {label} | **{M(d['value'])}** ({d['ms_per_step']:.3f} ms per picture: {kern(d)}) | **{M(e['value'])}** with {e['picture_streams_per_gpu']} picture streams "
                f"({M(e['single_stream_value'])} with one); {e['h2d_bytes_per_step'] / 1e6:.1f} MB up, {e['d2h_bytes_per_step'] / 1e6:.1f} MB down per picture | "
                f"{r['kernel']}: {r['achieved']:.0f} GB/s = {100 * r['frac']:.1f} % of the HBM peak"
                + (f"; ALU-pipe floor {100 * d['alu_roofline']['frac']:.0f} %; worst case {r['worst_case_launch_ms']:.2f} ms" if "alu_roofline" in d and r.get("worst_case_launch_ms") else "")
                + f" | 1 core: {M(c.get('value') or 0)} ({c.get('kind')}), sample re-checked: {c.get('checked')}"
                + (f"; {ref['cpu_baseline']['cores']} cores (`--impl reference`): {M(ref['value'])}" if ref else "") + " |")
for n in (2, 4, 8):
    for nm, label in (("bench_c2_n%d" % n, "config 2, independent picture streams"), ("bench_c5_n%d" % n, "config 5: 4K FullSearch, every rank reads rank 0's reconstructed anchor over NVLink"),
                      ("bench_c3_n%d" % n, "config 3, independent picture streams")):
        d = line(nm)
        if not d:
            continue
        e = d["e2e"]
        rows.append(f"| {n} GPUs, {label} | **{M(d['value'])}** ({d['ms_per_step']:.3f} ms per picture, max over ranks) | {M(e['value'])} ({e['h2d_bytes_per_step'] / 1e6:.1f} + {e['d2h_bytes_per_step'] / 1e6:.1f} MB per picture over PCIe"
                    + (f", {d['nvlink_bytes_per_step_per_rank'] / 1e6:.1f} MB per picture per rank over NVLink" if "nvlink_bytes_per_step_per_rank" in d else "") + ") | — | — |")
print("| | device-resident (`value`), macroblocks/s | end to end through the C ABI (`e2e`) | roofline of the dominant kernel | JM on the host CPU |")
print("|---|---|---|---|---|")
print("\n".join(rows))
d = line("bench")
if d and "next_rows" in d:
    x = d["next_rows"]["deblock"]
    print(f"\nDeblocking (`next_rows.deblock`): {x['gpu_ms_per_picture']:.2f} ms per {x['picture']} ({M(x['macroblocks_per_s'])} macroblocks/s, {x['wavefront_steps']} wavefront steps); "
          f"the CPU restatement on one core: {x.get('cpu_port_ms_per_picture', 0):.1f} ms; results equal: {x.get('gpu_matches_cpu')}.")
p = os.path.join(ROOT, "profiles", f"{TAG}_dropin_1080p.json")
if os.path.exists(p):
    e = json.load(open(p))["1080p"]
    s, g = e["stock"], e["dropin_me"]
    print(f"\nRe-linked encoder (`encoder`, `tools/dropin_1080p.py`, {e['frames']} frames of 1080p, FullSearch ±32, 1 reference, ME + planes + deblocking on the device): "
          f"stock `lencod` {s['wall_s']:.1f} s (JM's `Total ME time` {s['total_me_time'][0]}), `lencod_jmb` **{g['wall_s']:.1f} s** (ME {g['total_me_time'][0]}), "
          f"bitstream identical: {g['bitstream_identical']}.  {g['shim'][0] if g['shim'] else ''}")

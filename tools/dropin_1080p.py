#!/usr/bin/env python
"""tools/dropin_1080p.py -- stock lencod_ref vs the re-linked lencod_jmb on BASELINE config 2 (synthetic 1080p, FullSearch +-32,
Baseline, 1 reference): wall time, JM's own 'Total ME time', bitstream identity.  Two drop-in modes are timed:
  me     : JMB_SHIM_OFF=tq,dist  -- quarter-pel planes, integer search and sub-pel refinement on the device, JM's 16-coefficient
           transform / quantiser leaf calls stay JM's (a kernel launch per 16 coefficients costs 100x the CPU loop)
  full   : every wrapped family on the device (at CIF size: ~1000 leaf calls per macroblock)
Writes gpurun_out/<tag>_dropin_1080p.json.  usage: python tools/dropin_1080p.py [tag] [frames]"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_jm_dropin as T   # noqa: E402

CFG2 = ["ProfileIDC=66", "SymbolMode=0", "RDOptimization=1", "Transform8x8Mode=0", "QPISlice=28", "QPPSlice=28", "SearchMode=-1",
        "SearchRange=32", "NumberReferenceFrames=1", "AdaptiveRounding=0", "LevelIDC=51"]


def run(exe, d, tag, w, h, frames, env):
    t0 = time.perf_counter()
    r = T._encode(exe, d, tag, w, h, frames, CFG2, env=env)
    el = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-1500:]
    g = lambda key: [l.split(":")[1].strip() for l in r.stdout.splitlines() if key in l]
    return {"wall_s": el, "total_encoding_time": g("Total encoding time"), "total_me_time": g("Total ME time"),
            "md5_264": hashlib.md5(open(os.path.join(d, tag + ".264"), "rb").read()).hexdigest(),
            "shim": [l for l in r.stderr.splitlines() if l.startswith("[jmb shim]")]}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    out = {}
    for name, (w, h, modes) in {"1080p": (1920, 1088, ["me"]), "cif": (352, 288, ["me"] + (["full"] if os.environ.get("DROPIN_FULL") else []))}.items():
        with tempfile.TemporaryDirectory() as d:
            T._make_yuv(os.path.join(d, "input.yuv"), w, h, frames, seed=21)
            res = {"frames": frames, "macroblocks_per_frame": (w // 16) * (h // 16), "stock": run(T.REF, d, "ref", w, h, frames, {})}
            for m in modes:
                env = {"JMB_SHIM_VERBOSE": "1"}
                if m == "me":
                    env["JMB_SHIM_OFF"] = "tq,dist"
                res["dropin_" + m] = run(T.JMB, d, "gpu_" + m, w, h, frames, env)
                res["dropin_" + m]["bitstream_identical"] = res["dropin_" + m]["md5_264"] == res["stock"]["md5_264"]
            out[name] = res
            print(name, json.dumps(res)[:1500], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_dropin_1080p.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

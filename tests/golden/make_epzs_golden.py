"""Generates tests/golden/epzs_golden.npz: recorded calls of JM's REAL EPZS_integer_motion_estimation and
EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_int.c:42, me_epzs_sub.c:30) inside the live stock encoder.

Run here (needs /root/reference for the JM objects; CPU only):
    python tests/golden/make_epzs_golden.py
How: jm_b200/shim/_build/lencod_jmb = JM's unmodified objects + the --wrap shim, run with JMB_SHIM=passthrough (every
wrapper forwards to JM's own code; no GPU) and JMB_EPZS_CAPTURE=<file>: the EPZS wrappers then record, per call, the request
the device path would send (built from JM's own predictor generators), the pictures, and what JM's REAL function returned.
tests/test_epzs_golden.py replays the requests through the CPU restatement (oracle/jm_oracle.c::jmo_epzs) and demands JM's
answers -- that pins the restatement the GPU kernel is then compared with."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jm_b200 import api   # noqa: E402  (dtype of the request only)

JMB = os.path.join(ROOT, "jm_b200", "shim", "_build", "lencod_jmb")
RUNS = [   # (tag, w, h, frames, overrides, keep every n-th call)
    ("cabac8x8", 96, 80, 4, ["ProfileIDC=100", "SymbolMode=1", "RDOptimization=1", "Transform8x8Mode=1", "QPISlice=28", "QPPSlice=28",
                              "SearchMode=3", "SearchRange=16", "NumberReferenceFrames=2", "AdaptiveRounding=1", "EPZSSubPelGrid=1", "MEDistortionHPel=2", "MEDistortionQPel=2"], 11),
    ("base4x4", 80, 64, 5, ["ProfileIDC=66", "SymbolMode=0", "RDOptimization=1", "Transform8x8Mode=0", "QPISlice=32", "QPPSlice=32",
                             "SearchMode=3", "SearchRange=32", "NumberReferenceFrames=3", "AdaptiveRounding=0", "EPZSSubPelGrid=1",
                             "EPZSPattern=4", "EPZSDualRefinement=2", "EPZSFixedPredictors=2", "MEDistortionHPel=2", "MEDistortionQPel=2"], 13),
]


def parse(path):
    raw = open(path, "rb").read()
    pics, calls, off = {}, [], 0
    while off < len(raw):
        kind = int(np.frombuffer(raw, "<i4", 1, off)[0])
        if kind == 1:
            _, pid, w, h = np.frombuffer(raw, "<i4", 4, off); off += 16
            pics[int(pid)] = np.frombuffer(raw, np.uint8, w * h, off).reshape(h, w).copy(); off += w * h
        else:
            hdr = np.frombuffer(raw, "<i4", 8, off); off += 32
            req = np.frombuffer(raw, api.EPZS_REQ, 1, off).copy(); off += api.EPZS_REQ.itemsize
            nc = int(hdr[3])
            cands = np.frombuffer(raw, "<i2", 2 * nc, off).reshape(nc, 2).copy(); off += 4 * nc
            out = np.frombuffer(raw, "<i8", 2, off); off += 16
            calls.append(dict(kind=kind, ref=int(hdr[1]), cur=int(hdr[2]), mv=(int(hdr[4]), int(hdr[5])), me=int(hdr[6]), metrics=int(hdr[7]),
                              req=req, cands=cands, cost=int(out[0]), prev_after=int(out[1])))
    return pics, calls


def main():
    import test_jm_dropin as T
    out = {}
    for tag, w, h, frames, extra, step in RUNS:
        with tempfile.TemporaryDirectory() as d:
            T._make_yuv(os.path.join(d, "input.yuv"), w, h, frames, seed=5)
            cap = os.path.join(d, "cap.bin")
            r = T._encode(JMB, d, "cap", w, h, frames, extra, env={"JMB_SHIM": "passthrough", "JMB_EPZS_CAPTURE": cap})
            assert r.returncode == 0, r.stderr[-800:]
            pics, calls = parse(cap)
        # selection only (the expected values are JM's): keep every call whose integer stage leaves through one of the rarer
        # returns, and every step-th of the rest
        from oracle import pyoracle as po
        orc = po.Oracle()
        refs = {}
        keep = []
        for i, c in enumerate(calls):
            rare = False
            if c["kind"] == 2:
                if c["ref"] not in refs:
                    refs[c["ref"]] = orc.ref_create(pics[c["ref"]].astype(np.uint16))
                q = c["req"].copy(); q["cand_off"] = 0
                ex = int(orc.epzs(refs[c["ref"]], pics[c["cur"]].astype(np.uint16), q, c["cands"], (2, 2, 0, 1, 9))[0]["exit_code"])
                rare = ex in (2, 3, 4) and sum(1 for k in keep if k.get("exit") == ex) < 60
                c["exit"] = ex
            if rare or i % step == 0 or (c["kind"] == 2 and i % (step // 2) == 0):
                keep.append(c)
        used = sorted({c["ref"] for c in keep} | {c["cur"] for c in keep})
        for pid in used:
            out[f"{tag}_pic_{pid}"] = pics[pid]
        out[f"{tag}_req"] = np.concatenate([c["req"] for c in keep])
        out[f"{tag}_meta"] = np.array([[c["kind"], c["ref"], c["cur"], c["mv"][0], c["mv"][1], c["me"], c["metrics"], len(c["cands"])] for c in keep], np.int32)
        out[f"{tag}_out"] = np.array([[c["cost"], c["prev_after"]] for c in keep], np.int64)
        out[f"{tag}_cands"] = np.concatenate([c["cands"] for c in keep] + [np.zeros((0, 2), np.int16)])
        print(tag, len(calls), "calls recorded,", len(keep), "kept,", len(used), "pictures")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "epzs_golden.npz"), **out)


if __name__ == "__main__":
    main()

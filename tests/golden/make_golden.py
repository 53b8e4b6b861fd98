#!/usr/bin/env python
"""Generates tests/golden/jm_golden.npz from the REAL JM 19.0 functions (oracle/_ref/libjmref.so, compiled from the
reference's own sources by `make -C oracle ref`).  JM ships no golden vectors of its own, so these are outputs of the
reference itself run in the authoring container; the file is small and committed, this script regenerates it:

    python tests/golden/make_golden.py

Inputs are seeded (jm_b200/synth.py) and stored alongside the outputs, so the consumers (tests/test_golden.py: the C
restatement on CPU, the CUDA path on the GPU box) need neither the reference tree nor libjmref.so.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from jm_b200 import h264_tables as T, synth          # noqa: E402
from oracle import pyoracle as po                    # noqa: E402

W, H, R = 64, 48, 8


def main():
    assert po.ref_available(), "build oracle/_ref first: make -C oracle ref (needs /root/reference)"
    g = {}
    f = synth.luma_frames(W, H, 2, seed=2024, motion=(2, -3))
    g["ref_luma"], g["cur_luma"] = f[0], f[1]
    ref = po.JMRef(W, H, search_range=R)
    ref.set_ref(f[0]); ref.set_cur(f[1])
    g["planes"] = ref.planes()                                   # getSubImagesLuma, [4][4][H+40][W+64]
    g["spiral"] = ref.spiral()[:(2 * R + 1) ** 2]
    ref.spiral(); g["mvbits_arg"] = np.arange(-ref.max_mvd, ref.max_mvd + 1)      # the table spans +-max_mvd (mv_search.c:366-374)
    g["mvbits"] = np.array([ref.mvbits(int(v)) for v in g["mvbits_arg"]])
    rng = np.random.default_rng(77)

    # full_search_motion_estimation + sub_pel_motion_estimation (SAD / SATD / SATD: the bundled cfgs)
    n = 48
    fs = np.zeros((n, 12), np.int64)     # bt, pos_x, pos_y, pred_x, pred_y, lam, imv_x, imv_y, icost, mv_x, mv_y, cost
    for i in range(n):
        bt = int(rng.integers(1, 8)); bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // bsx + 1)) * bsx, int(rng.integers(0, (H - bsy) // bsy + 1)) * bsy)
        span = 30 if i % 4 else 150                               # every 4th: far predictor -> window leaves the picture
        pred = (int(rng.integers(-span, span)), int(rng.integers(-span, span)))
        center = (((pred[0] + 2) >> 2) * 4, ((pred[1] + 2) >> 2) * 4)
        lam = int(rng.integers(1, 400))
        mv, c = ref.full_search(bt, pos, pred, center, lam, po.DISTBLK_MAX)
        mv2, c2 = ref.sub_pel(bt, pos, pred, mv, [lam] * 3, po.DISTBLK_MAX, 0)
        fs[i] = [bt, pos[0], pos[1], pred[0], pred[1], lam, mv[0], mv[1], c, mv2[0], mv2[1], c2]
    g["full_search"] = fs

    # distortion of arbitrary quarter-pel candidates: computeSAD / computeSSE / computeSATD (4x4 and 8x8 Hadamard)
    n = 120
    ds = np.zeros((n, 9), np.int64)      # metric, bt, pos_x, pos_y, cand_x, cand_y, test8x8, dist(<<5), _
    for i in range(n):
        metric = i % 3
        bt = int(rng.integers(1, 8)); bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        cand = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        t8 = int(metric == po.SATD and bt <= 4 and rng.integers(0, 2))
        ds[i] = [metric, bt, pos[0], pos[1], cand[0], cand[1], t8, ref.dist(metric, bt, pos, cand, t8), 0]
    g["dist"] = ds

    # fast full search: BlockSAD surfaces (setup_fast_full_search) and fast_full_search_motion_estimation
    ffs = po.JMRef(W, H, search_range=R, fast_full=1)
    ffs.set_ref(f[0]); ffs.set_cur(f[1]); ffs.spiral()
    g["ffs_max_mvd"] = np.array([ffs.max_mvd])
    mbs = [(0, 0), (48, 32), (16, 16)]
    surf = np.zeros((len(mbs), 8, 16, (2 * R + 1) ** 2), np.uint32)
    meta = np.zeros((len(mbs), 6), np.int64)                     # mb_x, mb_y, pmv_x, pmv_y, center_x, center_y
    srch = []
    for k, mb in enumerate(mbs):
        pmv = (int(rng.integers(-30, 30)), int(rng.integers(-30, 30)))
        c = ffs.ffs_setup(mb, pmv)
        meta[k] = [mb[0], mb[1], pmv[0], pmv[1], c[0], c[1]]
        for bt, idxs in [(7, range(16)), (6, [0, 1, 2, 3, 8, 9, 10, 11]), (5, range(0, 16, 2)), (4, [0, 2, 8, 10]),
                         (3, [0, 2]), (2, [0, 8]), (1, [0])]:
            for i in idxs:
                surf[k, bt, i] = ffs.ffs_sad(bt, i)
                pos = (mb[0] + (i & 3) * 4, mb[1] + (i >> 2) * 4)
                pred = (pmv[0] + int(rng.integers(-6, 6)), pmv[1] + int(rng.integers(-6, 6)))
                lam = int(rng.integers(1, 300))
                mv, cost = ffs.ffs_search(bt, pos, pred, lam, po.DISTBLK_MAX)
                srch.append([k, bt, i, pos[0], pos[1], pred[0], pred[1], lam, mv[0], mv[1], cost])
    g["ffs_surfaces"], g["ffs_meta"], g["ffs_search"] = surf, meta, np.array(srch, np.int64)

    # forward4x4 / forward8x8 / HadamardSAD4x4 / HadamardSAD8x8
    b4 = rng.integers(-255, 256, size=(64, 4, 4)).astype(np.int32); b8 = rng.integers(-255, 256, size=(32, 8, 8)).astype(np.int32)
    g["res4"], g["res8"] = b4, b8
    g["fwd4"] = np.stack([ref.forward4x4(b) for b in b4]); g["fwd8"] = np.stack([ref.forward8x8(b) for b in b8])
    g["had4"] = np.array([ref.hadamard4x4(b) for b in b4]); g["had8"] = np.array([ref.hadamard8x8(b) for b in b8])

    # the six quantiser variants: 0/1 quant_4x4_normal/_around, 2/3 quant_8x8_normal/_around, 4/5 quant_8x8cavlc_normal/_around
    for variant in range(6):
        nn = 4 if variant < 2 else 8
        scan = T.SNGL_SCAN if nn == 4 else (T.SNGL_SCAN8x8_CAVLC if variant >= 4 else T.SNGL_SCAN8x8)
        cc = T.COEFF_COST4x4[0] if nn == 4 else T.COEFF_COST8x8[0]
        rows = []
        for it in range(40):
            qp = int(rng.integers(0, 52)); amp = int(rng.choice([3, 40, 255, 1023]))
            res = rng.integers(-amp, amp + 1, size=(nn, nn))
            coef = ref.forward4x4(res) if nn == 4 else ref.forward8x8(res)
            intra = int(rng.integers(0, 2)); cav = 1 if variant >= 4 else int(rng.integers(0, 2)); arw = 1 + it % 8
            o = ref.quant(variant, coef, qp, T.q_params(qp, intra, nn), scan, cc, cav, arw=arw, cost0=5)
            rows.append(dict(qp=qp, intra=intra, cav=cav, arw=arw, coef_in=coef, **o))
        for key in ("qp", "intra", "cav", "arw", "coef_in", "nonzero", "coef", "levels", "runs", "fadjust", "coeff_cost"):
            g[f"quant{variant}_{key}"] = np.array([r[key] for r in rows])

    # ---- appended after the first release of the file: everything above regenerates bit-identically ----
    # the weighted / bi-predictive distortion forms: compute*WP (form 1), computeBiPred*1 (2), computeBiPred*2 (3)
    f3 = synth.luma_frames(W, H, 3, seed=2024, motion=(2, -3))[2]
    g["ref2_luma"] = f3
    ref.set_ref2(f3)
    n = 180
    dx = np.zeros((n, 15), np.int64)   # metric, form, bt, pos_x, pos_y, c1x, c1y, c2x, c2y, t8, w1, w2, off, denom, dist(<<5)
    for i in range(n):
        metric, form = i % 3, 1 + (i // 3) % 3
        bt = int(rng.integers(1, 8)); bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        c1 = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        c2 = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        t8 = int(metric == po.SATD and bt <= 4 and rng.integers(0, 2))
        denom = int(rng.integers(0, 8))
        w = [int(rng.integers(-128, 128)) if i % 5 == 0 else int((1 << denom) * rng.uniform(0.5, 1.5)) for _ in range(2)]
        off = int(rng.integers(-40, 41))
        wp = (w[0], w[1], off, denom, (1 << (denom - 1)) if denom else 0)
        dx[i] = [metric, form, bt, pos[0], pos[1], c1[0], c1[1], c2[0], c2[1], t8, w[0], w[1], off, denom, ref.dist_ex(metric, form, bt, pos, c1, c2, wp, t8)]
    g["dist_ex"] = dx

    # inverse4x4 / inverse8x8 and the six Hadamard transforms (kind 0..5: hadamard4x4, ihadamard4x4, hadamard4x2, ihadamard4x2, hadamard2x2, ihadamard2x2)
    c4 = rng.integers(-4000, 4001, size=(48, 4, 4)).astype(np.int32); c8 = rng.integers(-4000, 4001, size=(24, 8, 8)).astype(np.int32)
    g["coef4"], g["coef8"] = c4, c8
    g["inv4"] = np.stack([ref.inverse4x4(b) for b in c4]); g["inv8"] = np.stack([ref.inverse8x8(b) for b in c8])
    for kind, per in enumerate((16, 16, 8, 8, 4, 4)):
        v = rng.integers(-30000, 30001, size=(32, per)).astype(np.int32)
        g[f"hadk{kind}_in"] = v; g[f"hadk{kind}_out"] = np.stack([ref.hadamard(kind, x) for x in v])

    # quant_ac4x4_normal/_around (6,7), quant_dc4x4_normal (8), quant_dc2x2_normal/_around (9,10), quant_dc4x2_normal/_around (11,12)
    scan420 = np.array([(0, 0), (0, 1), (0, 2), (0, 3)], np.uint8)
    scan422 = np.array([(0, 0), (0, 1), (1, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3)], np.uint8)
    for variant in range(6, 13):
        ncoef = {6: 16, 7: 16, 8: 16, 9: 4, 10: 4, 11: 8, 12: 8}[variant]
        scan = {9: scan420, 10: scan420, 11: scan422, 12: scan422}.get(variant, T.SNGL_SCAN)
        rows = []
        for it in range(30):
            qp = int(rng.integers(0, 52)); amp = int(rng.choice([6, 80, 600, 4000]))
            coef = rng.integers(-amp, amp + 1, size=ncoef).astype(np.int32)
            intra = int(rng.integers(0, 2)); cav = int(rng.integers(0, 2)); arw = 1 + it % 8
            qpar = T.q_params(qp, intra, 4)
            o = ref.quant_misc(variant, coef, qp, qpar if variant <= 7 else qpar[0, 0], scan, T.COEFF_COST4x4[0], cav, arw=arw, cost0=3)
            rows.append(dict(qp=qp, intra=intra, cav=cav, arw=arw, coef_in=coef, **o))
        for key in ("qp", "intra", "cav", "arw", "coef_in", "nonzero", "coef", "levels", "runs", "fadjust", "coeff_cost"):
            g[f"quant{variant}_{key}"] = np.array([r[key] for r in rows])

    # whole-encoder pins: md5 of the stock encoder's bitstream on seeded synthetic input (tests/test_jm_dropin.py configs)
    out = os.path.join(HERE, "jm_golden.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {os.path.getsize(out)} bytes, {len(g)} arrays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU study behind DESIGN.md 8.1: how many displacements of the full search could an EXACT successive-elimination bound
remove before any SAD is computed?  (numpy, no GPU, no reference tree.)

For a macroblock and every displacement of the +-32 window the 4x4 block sums of source (S_b) and reference (R_b) give
SAD_b >= |S_b - R_b|, hence for every partition p:  SAD_p >= sum_{b in p} |S_b - R_b|.  With byte-sized means
(S8 = S >> 4, R8 = R >> 4, one VABSDIFF4 per four blocks on the device) the bound loosens to
SAD_p >= 16 * sum |S8_b - R8_b| - 15 * n_p.  A displacement survives if for some partition the bound is still below the
gate threshold of k_int_search (thr_p - column term - row-group term, taken at their final values: optimistic by the
little the thresholds still move after the seeding stage).

usage: python tools/sea_sim.py [n_macroblocks]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jm_b200 import synth
from jm_b200 import h264_tables as T

W, H, R, PAD = 640, 368, 32, 48


def partitions():
    out = []
    for t, w4, h4 in ((1, 4, 4), (2, 4, 2), (3, 2, 4), (4, 2, 2), (5, 2, 1), (6, 1, 2), (7, 1, 1)):
        for by in range(0, 4, h4):
            for bx in range(0, 4, w4):
                out.append((t, bx, by, w4, h4))
    return out


def mvbits(v):
    a = np.abs(v)
    return np.where(a == 0, 1, 2 * np.floor(np.log2(np.maximum(a, 1))).astype(int) + 3)


def main(n_mb=24):
    f = synth.luma_frames(W, H, 2, seed=1234, motion=(5, 3))
    ref, cur = np.pad(f[0].astype(np.int64), PAD, mode="edge"), f[1].astype(np.int64)
    lam, parts, rng = T.lambda_me(28), partitions(), np.random.default_rng(5)
    tot = s8 = s16 = items = items8 = 0
    for _ in range(n_mb):
        mbx, mby = 16 * int(rng.integers(3, W // 16 - 3)), 16 * int(rng.integers(3, H // 16 - 3))
        mbpred = np.array([20, 12]) + rng.integers(-8, 9, size=2)           # bench.py's predictors: true motion + jitter
        preds = [mbpred + rng.integers(-3, 4, size=2) for _ in parts]
        cents = [(p + 2) >> 2 for p in preds]
        x0, x1 = min(c[0] for c in cents) - R, max(c[0] for c in cents) + R
        y0, y1 = min(c[1] for c in cents) - R, max(c[1] for c in cents) + R
        cw, ch = x1 - x0 + 1, y1 - y0 + 1
        src = cur[mby:mby + 16, mbx:mbx + 16]
        S = src.reshape(4, 4, 4, 4).sum(axis=(1, 3))
        sad4 = np.zeros((ch, cw, 4, 4), np.int64); R16 = np.zeros_like(sad4)
        for iy in range(ch):
            for ix in range(cw):
                blk = ref[PAD + mby + y0 + iy:PAD + mby + y0 + iy + 16, PAD + mbx + x0 + ix:PAD + mbx + x0 + ix + 16]
                sad4[iy, ix] = np.abs(blk - src).reshape(4, 4, 4, 4).sum(axis=(1, 3))
                R16[iy, ix] = blk.reshape(4, 4, 4, 4).sum(axis=(1, 3))
        L16, L8 = np.abs(S[None, None] - R16), np.abs((S >> 4)[None, None] - (R16 >> 4))
        Dx, Dy = np.arange(x0, x1 + 1), np.arange(y0, y1 + 1)
        live8 = np.zeros((ch, cw), bool); live16 = np.zeros((ch, cw), bool)
        for (t, bx, by, w4, h4), p, c in zip(parts, preds, cents):
            sad = sad4[:, :, by:by + h4, bx:bx + w4].sum(axis=(2, 3))
            bxs, bys = mvbits(4 * Dx - p[0]), mvbits(4 * Dy - p[1])
            inx, iny = np.abs(Dx - c[0]) <= R, np.abs(Dy - c[1]) <= R
            cost = np.where(inx[None, :] & iny[:, None], (sad << 5) + lam * (bxs[None, :] + bys[:, None]), 1 << 60)
            thr = ((cost.min() - 2 * lam) >> 5) + 1
            ax = np.where(inx, (lam * (bxs - 1)) >> 5, 1 << 30); ay = np.where(iny, (lam * (bys - 1)) >> 5, 1 << 30)
            ay4 = np.repeat(np.array([ay[4 * g:4 * g + 4].min() for g in range((ch + 3) // 4)]), 4)[:ch]
            t_gate = thr - ax[None, :] - ay4[:, None]
            n = w4 * h4
            live16 |= L16[:, :, by:by + h4, bx:bx + w4].sum(axis=(2, 3)) < t_gate
            live8 |= L8[:, :, by:by + h4, bx:bx + w4].sum(axis=(2, 3)) < (np.maximum(t_gate + 15 * n + 15, 0) >> 4)
        tot += ch * cw; s8 += int(live8.sum()); s16 += int(live16.sum())
        g = (ch + 3) // 4
        items += g * cw; items8 += int(sum(live8[4 * k:4 * k + 4].any(axis=0).sum() for k in range(g)))
    print(f"{n_mb} macroblocks, {tot // n_mb} displacements each: survivors {100.0 * s16 / tot:.1f} % with exact block sums, "
          f"{100.0 * s8 / tot:.1f} % with byte-sized means; {100.0 * items8 / items:.1f} % of the 4-displacement thread items keep a survivor")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 24)

"""Seeded synthetic YUV luma planes (no datasets in this environment).

Texture = coarse random field upsampled and box-blurred, translated by (dx, dy)*n pixels per frame,
plus N(0, sigma) noise -- bounded motion so a +/-32 full search finds the true match.
"""
import numpy as np


def luma_frames(width, height, n_frames, seed=1234, motion=(5, 3), noise=2.0, bitdepth=8):
    rng = np.random.default_rng(seed)
    mx, my = abs(motion[0]) * n_frames + 16, abs(motion[1]) * n_frames + 16
    W, H = width + 2 * mx, height + 2 * my
    coarse = rng.integers(16, 240, size=((H + 7) // 8 + 1, (W + 7) // 8 + 1)).astype(np.float32)
    tex = np.kron(coarse, np.ones((8, 8), np.float32))[:H, :W]
    k = 5
    c = np.cumsum(np.pad(tex, ((k, k), (0, 0)), mode="edge"), axis=0)
    tex = (c[2 * k:] - c[:-2 * k])[:H] / (2 * k)
    c = np.cumsum(np.pad(tex, ((0, 0), (k, k)), mode="edge"), axis=1)
    tex = (c[:, 2 * k:] - c[:, :-2 * k])[:, :W] / (2 * k)
    fine = rng.normal(0, 6.0, size=(H, W)).astype(np.float32)
    tex = tex + fine
    frames = []
    scale = (1 << bitdepth) / 256.0
    for n in range(n_frames):
        ox, oy = mx + motion[0] * n, my + motion[1] * n
        f = tex[oy:oy + height, ox:ox + width] + rng.normal(0, noise, size=(height, width))
        frames.append(np.clip(np.rint(f * scale), 0, (1 << bitdepth) - 1).astype(np.uint16))
    return frames


def write_yuv420(path, lumas, bitdepth=8, textured_chroma=False):
    """Planar I420; chroma flat mid-grey (all BASELINE cfgs have ChromaMEEnable=0) or a 2x2-decimated copy of the luma
    texture (so the chroma residual path codes something)."""
    with open(path, "wb") as f:
        for y in lumas:
            h, w = y.shape
            c = np.full((h // 2) * (w // 2) * 2, 1 << (bitdepth - 1))
            if textured_chroma:
                u = y[0::2, 0::2]; v = y[1::2, 1::2][::-1]
                c = np.concatenate([u.reshape(-1), v.reshape(-1)])
            if bitdepth == 8:
                f.write(y.astype(np.uint8).tobytes()); f.write(c.astype(np.uint8).tobytes())
            else:
                f.write(y.astype("<u2").tobytes()); f.write(c.astype("<u2").tobytes())


def write_yuv422(path, lumas, bitdepth=8):
    """Planar 4:2:2 (chroma at half width, full height) with luma-derived chroma texture."""
    with open(path, "wb") as f:
        for y in lumas:
            u = y[:, 0::2]; v = y[::-1, 1::2]
            for p in (y, u, v):
                f.write(p.astype(np.uint8 if bitdepth == 8 else "<u2").tobytes())

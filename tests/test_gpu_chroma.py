"""GPU parity of the chroma path (jmb_chroma_residual_coding: motion-compensated chroma prediction + residual coding with the
2x2 / 4x2 DC Hadamard paths) against the CPU restatement of OneComponentChromaPrediction4x4 + residual_transform_quant_chroma_4x4
(oracle/jm_oracle.c::jmo_chroma_pred / jmo_chroma_rc; their quantisers and Hadamards are the functions pinned to JM's in
tests/test_oracle_vs_ref.py; the glue is pinned in the live encoder, tests/test_jm_dropin.py).  Bit-exact."""
import numpy as np
import pytest

from jm_b200 import api, synth
from jm_b200 import h264_tables as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _chroma_planes(luma, yuv, seed):
    rng = np.random.default_rng(seed)
    h, w = luma.shape
    hc = h // 2 if yuv == 1 else h
    u = luma[::(2 if yuv == 1 else 1), 0::2].astype(int) // 2 + 60 + rng.integers(-2, 3, (hc, w // 2))
    v = luma[::-1][::(2 if yuv == 1 else 1), 1::2].astype(int) // 2 + 50 + rng.integers(-2, 3, (hc, w // 2))
    return np.ascontiguousarray(np.clip(u, 0, 255).astype(np.uint8)), np.ascontiguousarray(np.clip(v, 0, 255).astype(np.uint8))


@pytest.mark.parametrize("yuv,qp,cavlc,sample_bytes", [(1, 26, 1, 1), (2, 26, 0, 1), (1, 38, 0, 2), (2, 20, 1, 2), (2, 44, 1, 1), (1, 48, 1, 1)])
def test_chroma_residual_coding_matches_oracle(ctx, oracle, yuv, qp, cavlc, sample_bytes):
    w, h = 96, 64
    f = synth.luma_frames(w, h, 2, seed=80 + yuv, motion=(2, -1))
    ru, rv = _chroma_planes(f[0], yuv, 1)
    cu, cv = _chroma_planes(f[1], yuv, 2)
    dt = np.uint8 if sample_bytes == 1 else np.uint16
    ctx.configure(search_range=8)
    ctx.ref_put(0, f[0]); ctx.ref_put_chroma(0, ru.astype(dt), rv.astype(dt))
    ctx.pic_begin(f[1], [0]); ctx.pic_chroma(cu.astype(dt), cv.astype(dt))
    n_mb = (w // 16) * (h // 16)
    rng = np.random.default_rng(qp)
    pred = np.zeros(n_mb, api.MB_PRED)
    pred["b8mode"] = rng.integers(1, 8, size=(n_mb, 4))
    pred["mv"] = np.array([8, -4]) + rng.integers(-9, 10, size=(n_mb, 16, 2))
    pred["mv"][::5] = rng.integers(-300, 301, size=pred["mv"][::5].shape)          # far outside the picture: the coordinate clamps
    pred["mv"][1::7] = 0
    d = api.chroma_desc(yuv, qp, lambda q: T.q_params(q, 0, 4), T.COEFF_COST4x4[0], cavlc)
    got = ctx.chroma_residual_coding(d, pred=pred)
    hmb = 8 if yuv == 1 else 16
    nb = hmb // 2
    seen_cbp = set()
    for mb in range(n_mb):
        mbx, mby = (mb % (w // 16)) * 8, (mb // (w // 16)) * hmb
        for uv, (rp, cp) in enumerate(((ru, cu), (rv, cv))):
            p = oracle.chroma_pred(rp, yuv, (mbx, mby), pred["mv"][mb])
            want = oracle.chroma_rc(cp[mby:mby + hmb, mbx:mbx + 8], p, yuv, int(d["qp_ac"][0, uv]), int(d["qp_dc"][0, uv]), d["params_ac"][0, uv],
                                    d["params_dc"][0, uv], T.COEFF_COST4x4[0], cavlc)
            assert np.array_equal(got["recon"][mb, uv, :hmb], want["recon"]), (mb, uv, "recon")
            assert np.array_equal(got["dc"][mb, uv], want["dc"]), (mb, uv, "dc")
            assert np.array_equal(got["ac"][mb, uv, :nb], want["ac"][:nb]), (mb, uv, "ac")
            assert int(got["cbp_blk"][mb, uv]) == want["cbp_blk"] and int(got["cr_cbp"][mb, uv]) == want["cr_cbp"], (mb, uv, got["cbp_blk"][mb, uv], want)
            seen_cbp.add(want["cr_cbp"])
    assert seen_cbp, seen_cbp


def test_chroma_from_resident_search_results(ctx):
    """pred = NULL: the motion of partition mode `mode` comes from the search results still on the device."""
    w, h = 64, 48
    f = synth.luma_frames(w, h, 2, seed=85, motion=(1, 2))
    ru, rv = _chroma_planes(f[0], 1, 3); cu, cv = _chroma_planes(f[1], 1, 4)
    n_mb = 12
    ctx.configure(search_range=8)
    ctx.ref_put(0, f[0]); ctx.ref_put_chroma(0, ru, rv); ctx.pic_begin(f[1], [0]); ctx.pic_chroma(cu, cv)
    predtab = np.zeros(n_mb, api.MB_MVPRED); predtab["pred"] = np.random.default_rng(5).integers(-6, 7, size=(n_mb, 41, 2))
    res8 = ctx.me_search_frame_pred(predtab, api.frame_params([40, 40, 40])).reshape(n_mb, 41)
    d = api.chroma_desc(1, 28, lambda q: T.q_params(q, 0, 4), T.COEFF_COST4x4[0], 1)
    for mode in (1, 4, 7):
        a = ctx.chroma_residual_coding(d, pred=None, mode=mode, n_mb=n_mb)
        res24 = np.zeros(n_mb * 41, api.ME_RES); res24["mv_x"] = res8["mv_x"].reshape(-1); res24["mv_y"] = res8["mv_y"].reshape(-1)
        p = ctx.pred_from_results(res24, mode)
        b = ctx.chroma_residual_coding(d, pred=p)
        for k in ("dc", "ac", "cbp_blk", "cr_cbp", "recon"):
            assert np.array_equal(a[k], b[k]), (mode, k)
    with pytest.raises(api.JMBError):
        bad = p.copy(); bad["ref"][3, 1] = 9
        ctx.chroma_residual_coding(d, pred=bad)          # host-side tables are validated on the device too
        ctx.sync()

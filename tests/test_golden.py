"""Golden vectors produced by the REAL JM 19.0 functions (tests/golden/make_golden.py -> jm_golden.npz).
CPU: the C restatement (oracle/jm_oracle.c) must reproduce them.  GPU (-m gpu): libjmb200, through the C ABI, must too.
Neither side needs the reference tree at run time."""
import os

import numpy as np
import pytest

from jm_b200 import h264_tables as T
from oracle import pyoracle as po

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jm_golden.npz"))
W, H, R = 64, 48, 8
BIG = po.DISTBLK_MAX
FFS_IDX = [(7, range(16)), (6, [0, 1, 2, 3, 8, 9, 10, 11]), (5, range(0, 16, 2)), (4, [0, 2, 8, 10]), (3, [0, 2]), (2, [0, 8]), (1, [0])]


def _variant_tables(variant):
    n = 4 if variant < 2 else 8
    scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if variant >= 4 else T.SNGL_SCAN8x8)
    cc = T.COEFF_COST4x4[0] if n == 4 else T.COEFF_COST8x8[0]
    return n, scan, cc


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_planes_spiral_mvbits(oracle):
    r = oracle.ref_create(G["ref_luma"])
    assert np.array_equal(oracle.planes(r), G["planes"])
    oracle.ref_destroy(r)
    assert np.array_equal(oracle.spiral(R), G["spiral"])
    assert [oracle.mvbits(int(v)) for v in G["mvbits_arg"]] == G["mvbits"].tolist()


def test_oracle_searches_and_distortion(oracle):
    r = oracle.ref_create(G["ref_luma"]); cur = G["cur_luma"]
    for bt, px, py, qx, qy, lam, ix, iy, ic, mx, my, mc in G["full_search"].tolist():
        center = (((qx + 2) >> 2) * 4, ((qy + 2) >> 2) * 4)
        mv, c = oracle.full_search(r, cur, bt, (px, py), (qx, qy), center, lam, BIG, R)
        assert (mv, c) == ((ix, iy), ic)
        mv2, c2 = oracle.sub_pel(r, cur, bt, (px, py), (qx, qy), mv, [lam] * 3, BIG, po.SATD, po.SATD, 0, 1, 0)
        assert (mv2, c2) == ((mx, my), mc)
    for metric, bt, px, py, cx, cy, t8, d, _ in G["dist"].tolist():
        assert oracle.dist(r, cur, bt, (px, py), (cx, cy), metric, t8) << 5 == d
    for k, (mbx, mby, pmx, pmy, cx, cy) in enumerate(G["ffs_meta"].tolist()):
        assert oracle.ffs_center((pmx, pmy), R) == (cx, cy)
        bs = oracle.ffs_setup(r, cur, (mbx, mby), (cx, cy), R)
        for bt, idxs in FFS_IDX:
            for i in idxs:
                assert np.array_equal(bs[bt, i], G["ffs_surfaces"][k, bt, i])
        for kk, bt, i, px, py, qx, qy, lam, mx, my, cost in G["ffs_search"].tolist():
            if kk == k:
                assert oracle.ffs_search(bs, R, bt, i, (cx, cy), (qx, qy), lam, BIG, int(G["ffs_max_mvd"][0])) == ((mx, my), cost)
    oracle.ref_destroy(r)


def test_oracle_transforms_and_quant(oracle):
    for i, b in enumerate(G["res4"]):
        assert np.array_equal(oracle.forward4x4(b), G["fwd4"][i]) and oracle.hadamard4x4(b) == G["had4"][i]
    for i, b in enumerate(G["res8"]):
        assert np.array_equal(oracle.forward8x8(b), G["fwd8"][i]) and oracle.hadamard8x8(b) == G["had8"][i]
    for variant in range(6):
        n, scan, cc = _variant_tables(variant)
        for i in range(len(G[f"quant{variant}_qp"])):
            g = {k: G[f"quant{variant}_{k}"][i] for k in ("qp", "intra", "cav", "arw", "coef_in", "nonzero", "coef", "levels", "runs", "fadjust", "coeff_cost")}
            o = oracle.quant(variant, g["coef_in"], int(g["qp"]), T.q_params(int(g["qp"]), int(g["intra"]), n), scan, cc, int(g["cav"]), arw=int(g["arw"]), cost0=5)
            assert o["nonzero"] == g["nonzero"] and o["coeff_cost"] == g["coeff_cost"] and np.array_equal(o["coef"], g["coef"])
            assert np.array_equal(o["levels"], g["levels"]) and np.array_equal(o["runs"], g["runs"])
            if variant & 1:
                assert np.array_equal(o["fadjust"], g["fadjust"])


# ------------------------------------------------------------------------------------------------ GPU: libjmb200
@pytest.fixture(scope="module")
def ctx():
    from jm_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_planes(ctx):
    ctx.ref_put(0, G["ref_luma"])
    for fy in range(4):
        for fx in range(4):
            assert np.array_equal(ctx.ref_plane(0, fy, fx, (H, W)), G["planes"][fy, fx]), (fy, fx)


@pytest.mark.gpu
def test_gpu_searches_and_distortion(ctx):
    from jm_b200 import api
    ctx.configure(search_range=R)
    ctx.ref_put(0, G["ref_luma"]); ctx.pic_begin(G["cur_luma"], [0])
    fs = G["full_search"]
    reqs = np.zeros(len(fs), api.ME_REQ)
    reqs["blocktype"], reqs["pos_x"], reqs["pos_y"], reqs["pred_x"], reqs["pred_y"] = fs[:, 0], fs[:, 1], fs[:, 2], fs[:, 3], fs[:, 4]
    reqs["center_x"] = ((fs[:, 3] + 2) >> 2) * 4; reqs["center_y"] = ((fs[:, 4] + 2) >> 2) * 4
    reqs["lambda"] = fs[:, 5:6]; reqs["flags"] = api.REQ_SUBPEL; reqs["min_mcost"] = BIG
    res = ctx.me_search(reqs)
    for k, name in enumerate(("imv_x", "imv_y", "icost", "mv_x", "mv_y", "cost")):
        assert np.array_equal(res[name].astype(np.int64), fs[:, 6 + k]), name
    for metric, bt, px, py, cx, cy, t8, d, _ in G["dist"].tolist():
        assert int(ctx.dist(0, metric, bt, (px, py), [(cx, cy)], t8)[0]) << 5 == d
    # fast full search: surfaces and searches
    for k, (mbx, mby, pmx, pmy, cx, cy) in enumerate(G["ffs_meta"].tolist()):
        got = ctx.ffs_surfaces(0, (mbx, mby), (cx, cy))
        for bt, idxs in FFS_IDX:
            for i in idxs:
                assert np.array_equal(got[bt, i], G["ffs_surfaces"][k, bt, i]), (k, bt, i)
    s = G["ffs_search"]
    reqs = np.zeros(len(s), api.ME_REQ)
    reqs["blocktype"], reqs["pos_x"], reqs["pos_y"], reqs["pred_x"], reqs["pred_y"] = s[:, 1], s[:, 3], s[:, 4], s[:, 5], s[:, 6]
    reqs["center_x"] = G["ffs_meta"][s[:, 0], 4]; reqs["center_y"] = G["ffs_meta"][s[:, 0], 5]
    reqs["lambda"] = s[:, 7:8]; reqs["mode"] = api.SEARCH_FAST_FULL; reqs["min_mcost"] = BIG
    assert ctx.max_mvd == int(G["ffs_max_mvd"][0])
    res = ctx.me_search(reqs)
    assert np.array_equal(res["imv_x"], s[:, 8]) and np.array_equal(res["imv_y"], s[:, 9]) and np.array_equal(res["icost"], s[:, 10])


@pytest.mark.gpu
def test_gpu_transforms_and_quant(ctx):
    from jm_b200 import api
    assert np.array_equal(ctx.forward_transform(G["res4"], 4), G["fwd4"])
    assert np.array_equal(ctx.forward_transform(G["res8"], 8), G["fwd8"])
    for variant in range(6):
        n, scan, cc = _variant_tables(variant)
        for i in range(len(G[f"quant{variant}_qp"])):
            g = {k: G[f"quant{variant}_{k}"][i] for k in ("qp", "intra", "cav", "arw", "coef_in", "nonzero", "coef", "levels", "runs", "fadjust", "coeff_cost")}
            cav = int(g["cav"]) if n == 4 else int(variant >= 4)      # quant_8x8_normal/_around ignore the entropy mode; the ABI selects the CAVLC lists with it
            qd = api.quant_desc(n, int(g["qp"]), T.q_params(int(g["qp"]), int(g["intra"]), n), scan, cc, cav, variant & 1, int(g["arw"]))
            o = ctx.quant_blocks(qd, g["coef_in"], do_transform=0, cost0=5)
            assert o["nonzero"][0] == g["nonzero"] and o["coeff_cost"][0] == g["coeff_cost"] and np.array_equal(o["coef"][0], g["coef"])
            assert np.array_equal(o["levels"][0], g["levels"]) and np.array_equal(o["runs"][0], g["runs"])
            if variant & 1:
                assert np.array_equal(o["fadjust"][0], g["fadjust"])


def test_oracle_luma_residual_coding_consistency(oracle):
    """The restated luma_residual_coding against its own pieces: a macroblock whose prediction equals the source codes
    nothing and reconstructs the source; a strong residual is coded and reconstructs within the quantiser's error."""
    src = G["cur_luma"][:16, :16]
    qpar = T.q_params(28, 0, 4)
    lv, c8, cbp, cbpb, rec, sse = oracle.luma_residual_coding(src, src, 4, 28, qpar, T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    assert not lv.any() and cbp == 0 and cbpb == 0 and sse == 0 and np.array_equal(rec, src)
    pred = np.clip(src.astype(np.int32) + np.random.default_rng(1).integers(-40, 41, size=(16, 16)), 0, 255).astype(np.uint16)
    lv, c8, cbp, cbpb, rec, sse = oracle.luma_residual_coding(src, pred, 4, 28, qpar, T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    assert cbp == 15 and cbpb == 0xFFFF and lv.any()
    assert sse == int(((src.astype(np.int64) - rec.astype(np.int64)) ** 2).sum()) and sse < int(((src.astype(np.int64) - pred) ** 2).sum())
    # inverse(dequant(quant(forward(x)))) ~ x: the pinned pieces compose
    x = np.random.default_rng(2).integers(-60, 61, size=(4, 4))
    q = oracle.quant(0, oracle.forward4x4(x), 20, T.q_params(20, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    back = (oracle.inverse4x4(q["coef"]) + 32) >> 6
    assert np.abs(back - x).max() <= 12          # QP 20: quantiser step 6.5


# ---- the widened path: weighted / bi-predictive distortions, inverse transforms, Hadamards, DC / AC list quantisers ----
SCAN420 = np.array([(0, 0), (0, 1), (0, 2), (0, 3)], np.uint8)
SCAN422 = np.array([(0, 0), (0, 1), (1, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3)], np.uint8)


def _wp(w1, w2, off, denom):
    return (int(w1), int(w2), int(off), int(denom), (1 << (int(denom) - 1)) if denom else 0)


def _misc_cases(variant):
    from jm_b200 import api
    scan = {9: SCAN420, 10: SCAN420, 11: SCAN422, 12: SCAN422}.get(variant, T.SNGL_SCAN)
    for i in range(len(G[f"quant{variant}_qp"])):
        g = {k: G[f"quant{variant}_{k}"][i] for k in ("qp", "intra", "cav", "arw", "coef_in", "nonzero", "coef", "levels", "runs", "fadjust", "coeff_cost")}
        qpar = T.q_params(int(g["qp"]), int(g["intra"]), 4)
        plan = api.qlist_plan(variant, int(g["qp"]), qpar if variant <= 7 else qpar[0, 0], scan, T.COEFF_COST4x4[0], int(g["cav"]), int(g["arw"]))
        yield plan, g


def _check_misc(o, g, variant):
    assert o["nonzero"] == g["nonzero"] and o["coeff_cost"] == g["coeff_cost"], variant
    for k in ("coef", "levels", "runs", "fadjust"):
        assert np.array_equal(o[k], g[k]), (variant, k)


def test_oracle_widened_path(oracle):
    r1, r2 = oracle.ref_create(G["ref_luma"]), oracle.ref_create(G["ref2_luma"])
    for metric, form, bt, px, py, c1x, c1y, c2x, c2y, t8, w1, w2, off, denom, d in G["dist_ex"].tolist():
        assert oracle.dist_ex(r1, r2, G["cur_luma"], bt, (px, py), (c1x, c1y), (c2x, c2y), metric, form, _wp(w1, w2, off, denom), t8) << 5 == d
    oracle.ref_destroy(r1); oracle.ref_destroy(r2)
    for i, b in enumerate(G["coef4"]):
        assert np.array_equal(oracle.inverse4x4(b), G["inv4"][i])
    for i, b in enumerate(G["coef8"]):
        assert np.array_equal(oracle.inverse8x8(b), G["inv8"][i])
    for kind in range(6):
        for v, want in zip(G[f"hadk{kind}_in"], G[f"hadk{kind}_out"]):
            assert np.array_equal(oracle.hadamard(kind, v), want), kind
    for variant in range(6, 13):
        for plan, g in _misc_cases(variant):
            _check_misc(oracle.quant_list(plan, g["coef_in"], cost0=3), g, variant)


@pytest.mark.gpu
def test_gpu_widened_path(ctx):
    ctx.ref_put(0, G["ref_luma"]); ctx.ref_put(1, G["ref2_luma"]); ctx.pic_begin(G["cur_luma"], [0, 1])
    for metric, form, bt, px, py, c1x, c1y, c2x, c2y, t8, w1, w2, off, denom, d in G["dist_ex"].tolist():
        got = ctx.dist_ex(0, 1, metric, form, bt, (px, py), [(c1x, c1y)], (c2x, c2y), _wp(w1, w2, off, denom), t8)
        assert int(got[0]) << 5 == d, (metric, form, bt)
    assert np.array_equal(ctx.inverse_transform(G["coef4"], 4), G["inv4"])
    assert np.array_equal(ctx.inverse_transform(G["coef8"], 8), G["inv8"])
    for kind, per in enumerate((16, 16, 8, 8, 4, 4)):
        assert np.array_equal(ctx.hadamard(kind, G[f"hadk{kind}_in"], per), G[f"hadk{kind}_out"]), kind
    for variant in range(6, 13):
        for plan, g in _misc_cases(variant):
            _check_misc(ctx.quant_list(plan, g["coef_in"], cost0=3), g, variant)

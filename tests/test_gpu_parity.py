"""GPU parity tests proper: every libjmb200 entry point, called through the C ABI, against the CPU
oracle (oracle/jm_oracle.c, itself pinned to the real JM functions by test_oracle_vs_ref.py) on the
same seeded inputs.  Bit-exact: this is integer work."""
import numpy as np
import pytest

from jm_b200 import api, synth
from jm_b200 import h264_tables as T
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
BIG = po.DISTBLK_MAX


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _frames(w, h, seed, motion=(3, -2), n=2):
    return synth.luma_frames(w, h, n, seed=seed, motion=motion)


@pytest.mark.parametrize("w,h,seed", [(96, 64, 1), (176, 144, 2), (16, 16, 3), (400, 48, 4), (208, 128, 5), (1920, 1088, 6)])
def test_subpel_planes(ctx, oracle, w, h, seed):
    f = _frames(w, h, seed)[0]
    if seed == 3:
        f = np.random.default_rng(3).choice([0, 255], size=(h, w)).astype(np.uint16)   # exercises the clips
    if seed == 5:
        f = np.random.default_rng(5).integers(0, 256, size=(h, w)).astype(np.uint16)   # white noise: every tap sign pattern
    r = oracle.ref_create(f)
    want = oracle.planes(r)
    for upload in ("u16", "u8"):          # JM's imgpel samples and the byte upload take different staging paths
        if upload == "u16":
            ctx.ref_put(0, f)
        else:
            ctx.ref_put_u8(0, f.astype(np.uint8))
        for fy in range(4):
            for fx in range(4):
                got = ctx.ref_plane(0, fy, fx, (h, w))
                assert np.array_equal(got, want[fy, fx]), (upload, fy, fx)
    oracle.ref_destroy(r)


def _random_reqs(rng, w, h, n, R, mode, flags, pred_span=40, lam_max=400):
    reqs = np.zeros(n, api.ME_REQ)
    for q in reqs:
        bt = int(rng.integers(1, 8))
        bsx, bsy = api.BLOCK_SIZE[bt]
        q["blocktype"] = bt
        q["pos_x"] = int(rng.integers(0, (w - bsx) // bsx + 1)) * bsx
        q["pos_y"] = int(rng.integers(0, (h - bsy) // bsy + 1)) * bsy
        q["pred_x"], q["pred_y"] = rng.integers(-pred_span, pred_span + 1, 2)
        q["center_x"] = ((int(q["pred_x"]) + 2) >> 2) * 4
        q["center_y"] = ((int(q["pred_y"]) + 2) >> 2) * 4
        q["mode"], q["flags"] = mode, flags
        q["lambda"] = int(rng.integers(1, lam_max))
        q["min_mcost"] = BIG
    return reqs


def _check_full(ctx, oracle, r, cur, reqs, res, R, sub=None):
    for q, o in zip(reqs, res):
        pos = (int(q["pos_x"]), int(q["pos_y"])); pred = (int(q["pred_x"]), int(q["pred_y"]))
        center = (int(q["center_x"]), int(q["center_y"]))
        mv, cost = oracle.full_search(r, cur, int(q["blocktype"]), pos, pred, center, int(q["lambda"][0]), int(q["min_mcost"]), R)
        assert (int(o["imv_x"]), int(o["imv_y"])) == mv and int(o["icost"]) == cost, (q, o, mv, cost)
        if sub:
            mh, mq, shp, sqp = sub
            mc = cost if shp else BIG
            t8 = int(bool(q["flags"] & api.REQ_TEST8X8))
            mv2, c2 = oracle.sub_pel(r, cur, int(q["blocktype"]), pos, pred, mv, [int(x) for x in q["lambda"]], mc, mh, mq, shp, sqp, t8)
            assert (int(o["mv_x"]), int(o["mv_y"])) == mv2 and int(o["cost"]) == c2, (q, o, mv2, c2)


@pytest.mark.parametrize("w,h,R,seed,span", [(96, 64, 8, 5, 40), (64, 48, 12, 6, 200), (176, 144, 32, 7, 60)])
def test_full_search_random_requests(ctx, oracle, w, h, R, seed, span):
    """Ungrouped random requests; large predictors push whole windows outside the picture so the
    partition-origin clamp (me_distortion.c:367) and the border path are exercised."""
    f = _frames(w, h, seed)
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(seed)
    reqs = _random_reqs(rng, w, h, 24 if R == 32 else 60, R, api.SEARCH_FULL, 0, pred_span=span)
    res = ctx.me_search(reqs)
    _check_full(ctx, oracle, r, f[1], reqs, res, R)
    oracle.ref_destroy(r)


def test_full_search_min_mcost_gate(ctx, oracle):
    """Incoming min_mcost below every candidate: JM keeps the centre and returns min_mcost."""
    w, h, R = 64, 48, 6
    f = _frames(w, h, 8)
    ctx.configure(search_range=R); ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    reqs = _random_reqs(np.random.default_rng(8), w, h, 30, R, api.SEARCH_FULL, 0)
    reqs["min_mcost"][::2] = 40
    reqs["min_mcost"][1::4] = 3000
    res = ctx.me_search(reqs)
    _check_full(ctx, oracle, r, f[1], reqs, res, R)
    oracle.ref_destroy(r)


def _frame_reqs(w, h, rng, mode, flags, lam=None, jitter=6, base=(0, 0)):
    parts = api.mb_partitions()
    n_mb = (w // 16) * (h // 16)
    reqs = np.zeros(n_mb * api.NPART, api.ME_REQ)
    i = 0
    for mb in range(n_mb):
        mbx, mby = (mb % (w // 16)) * 16, (mb // (w // 16)) * 16
        mbpred = rng.integers(-jitter, jitter + 1, 2) + np.array(base)
        for (t, x, y) in parts:
            q = reqs[i]; i += 1
            q["blocktype"], q["pos_x"], q["pos_y"] = t, mbx + x, mby + y
            p = mbpred + rng.integers(-3, 4, 2)
            q["pred_x"], q["pred_y"] = p
            c = p if mode == api.SEARCH_FULL else mbpred
            q["center_x"], q["center_y"] = ((int(c[0]) + 2) >> 2) * 4, ((int(c[1]) + 2) >> 2) * 4
            q["mode"], q["flags"] = mode, flags
            q["lambda"] = lam if lam is not None else int(rng.integers(1, 300))
            q["min_mcost"] = BIG
    return reqs


@pytest.mark.parametrize("flags,metrics", [(api.REQ_SUBPEL, (api.SAD, api.SATD, api.SATD)),
                                           (api.REQ_SUBPEL, (api.SAD, api.SAD, api.SAD)),
                                           (api.REQ_SUBPEL, (api.SAD, api.SATD, api.SAD)),
                                           (api.REQ_SUBPEL | api.REQ_TEST8X8, (api.SAD, api.SATD, api.SATD))])
def test_frame_search_with_subpel(ctx, oracle, flags, metrics):
    """All 41 partitions of every macroblock in one call (shared 4x4 SADs), then the half-/quarter-pel
    refinement, with each sub-pel metric combination JM's start_me_refinement_hp/qp rules produce."""
    w, h, R = 80, 48, 8
    f = _frames(w, h, 9, motion=(-3, 2))
    ctx.configure(search_range=R, metric=metrics)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(10)
    reqs = _frame_reqs(w, h, rng, api.SEARCH_FULL, flags, base=(12, -8))
    if flags & api.REQ_TEST8X8:
        reqs["flags"][reqs["blocktype"] > 4] = api.REQ_SUBPEL      # JM only sets test8x8 for block types <= 4
    res = ctx.me_search(reqs, frame=True)
    shp = 0 if metrics[0] != metrics[1] else 1
    sqp = 0 if metrics[1] != metrics[2] else 1
    _check_full(ctx, oracle, r, f[1], reqs, res, R, sub=(metrics[1], metrics[2], shp, sqp))
    # the ungrouped entry point must give the same answers
    sel = rng.permutation(len(reqs))[:50]
    res2 = ctx.me_search(reqs[sel])
    assert np.array_equal(res2, res[sel])
    oracle.ref_destroy(r)


def test_fast_full_search(ctx, oracle):
    """FAST_FULL: one centre per macroblock, macroblock-origin clamp, max_mvd guard; also the BlockSAD
    surfaces JM's setup_fast_full_search would have produced."""
    w, h, R = 64, 48, 8
    f = _frames(w, h, 11, motion=(2, 3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(12)
    reqs = _frame_reqs(w, h, rng, api.SEARCH_FAST_FULL, 0, jitter=30)
    res = ctx.me_search(reqs, frame=True)
    for g in range(len(reqs) // api.NPART):
        q0 = reqs[g * api.NPART]
        mb = (int(q0["pos_x"]), int(q0["pos_y"])); c = (int(q0["center_x"]), int(q0["center_y"]))
        bs = oracle.ffs_setup(r, f[1], mb, c, R)
        if g % 5 == 0:
            got = ctx.ffs_surfaces(0, mb, c)
            for bt, idxs in [(7, range(16)), (6, [0, 1, 2, 3, 8, 9, 10, 11]), (5, range(0, 16, 2)), (4, [0, 2, 8, 10]),
                             (3, [0, 2]), (2, [0, 8]), (1, [0])]:
                for i in idxs:
                    assert np.array_equal(got[bt, i], bs[bt, i]), (mb, bt, i)
        for k in range(api.NPART):
            q, o = reqs[g * api.NPART + k], res[g * api.NPART + k]
            bi = ((int(q["pos_y"]) & 15) >> 2) * 4 + ((int(q["pos_x"]) & 15) >> 2)
            mv, cost = oracle.ffs_search(bs, R, int(q["blocktype"]), bi, c, (int(q["pred_x"]), int(q["pred_y"])),
                                         int(q["lambda"][0]), BIG, ctx.max_mvd)
            assert (int(o["imv_x"]), int(o["imv_y"])) == mv and int(o["icost"]) == cost, (g, k, q, o, mv, cost)
    oracle.ref_destroy(r)


def test_subpel_only_requests(ctx, oracle):
    """JMB_REQ_SKIP_INT: the SubPelME call site alone (mv_search.c:975)."""
    w, h = 64, 48
    f = _frames(w, h, 13)
    ctx.configure(search_range=8); ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(13)
    reqs = _random_reqs(rng, w, h, 80, 8, api.SEARCH_FULL, api.REQ_SUBPEL | api.REQ_SKIP_INT)
    reqs["center_x"] = rng.integers(-12, 12, len(reqs)) * 4
    reqs["center_y"] = rng.integers(-12, 12, len(reqs)) * 4
    reqs["min_mcost"][::3] = 20000
    res = ctx.me_search(reqs)
    for q, o in zip(reqs, res):
        mv, c = oracle.sub_pel(r, f[1], int(q["blocktype"]), (int(q["pos_x"]), int(q["pos_y"])), (int(q["pred_x"]), int(q["pred_y"])),
                               (int(q["center_x"]), int(q["center_y"])), [int(x) for x in q["lambda"]], int(q["min_mcost"]),
                               po.SATD, po.SATD, 0, 1, 0)
        assert (int(o["mv_x"]), int(o["mv_y"])) == mv and int(o["cost"]) == c
    oracle.ref_destroy(r)


@pytest.mark.parametrize("case", ["one_mb_same_centre", "duplicates", "six_chunks_of_16x16", "test8x8"])
def test_subpel_shared_evaluations(ctx, oracle, case):
    """The refinement evaluates sub-blocks that several searches of a CTA share only once (same position, same mv, SATD): the
    extremes of that path -- a macroblock whose 41 partitions all stand on one mv, the same search listed several times, a CTA
    whose 656 items take six chunks with nothing shared, and 8x8 sub-blocks -- against JM search by search."""
    w, h = 96, 64
    f = _frames(w, h, 31)
    ctx.configure(search_range=8); ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(31)
    shapes = [(bt, bx, by) for bt in range(1, 8) for by in range(0, 16, api.BLOCK_SIZE[bt][1]) for bx in range(0, 16, api.BLOCK_SIZE[bt][0])]
    assert len(shapes) == 41
    if case == "six_chunks_of_16x16":
        reqs = np.zeros(41, api.ME_REQ)
        for i, q in enumerate(reqs):
            q["blocktype"] = 1; q["pos_x"] = 16 * (i % 6); q["pos_y"] = 16 * ((i // 6) % 4)
            q["center_x"], q["center_y"] = 4 * int(rng.integers(-6, 7)), 4 * int(rng.integers(-6, 7))
    else:
        reqs = np.zeros(82 if case == "duplicates" else 41, api.ME_REQ)
        for i, q in enumerate(reqs):
            bt, bx, by = shapes[i % 41] if case != "duplicates" else shapes[(i // 2) % 41]
            q["blocktype"] = bt; q["pos_x"] = 32 + bx; q["pos_y"] = 16 + by
            q["center_x"], q["center_y"] = 12, -8
    reqs["pred_x"] = rng.integers(-9, 10, len(reqs)); reqs["pred_y"] = rng.integers(-9, 10, len(reqs))
    reqs["mode"] = api.SEARCH_FULL
    reqs["flags"] = api.REQ_SUBPEL | api.REQ_SKIP_INT
    if case == "test8x8":
        reqs["flags"][reqs["blocktype"] <= 4] |= api.REQ_TEST8X8          # as the picture form does (jm_b200/api.py)
    reqs["lambda"] = rng.integers(1, 300, (len(reqs), 1))
    reqs["min_mcost"] = BIG
    res = ctx.me_search(reqs)
    for q, o in zip(reqs, res):
        t8 = 1 if (case == "test8x8" and int(q["blocktype"]) <= 4) else 0
        mv, c = oracle.sub_pel(r, f[1], int(q["blocktype"]), (int(q["pos_x"]), int(q["pos_y"])), (int(q["pred_x"]), int(q["pred_y"])),
                               (int(q["center_x"]), int(q["center_y"])), [int(x) for x in q["lambda"]], int(q["min_mcost"]),
                               po.SATD, po.SATD, 0, 1, t8)
        assert (int(o["mv_x"]), int(o["mv_y"])) == mv and int(o["cost"]) == c, (case, q)
    oracle.ref_destroy(r)


@pytest.mark.parametrize("metric", [api.SAD, api.SSE, api.SATD])
def test_dist(ctx, oracle, metric):
    w, h = 96, 64
    f = _frames(w, h, 14)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(metric)
    for _ in range(12):
        bt = int(rng.integers(1, 8)); bsx, bsy = api.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (w - bsx) // 4 + 1)) * 4, int(rng.integers(0, (h - bsy) // 4 + 1)) * 4)
        t8 = int(metric == api.SATD and bt <= 4 and rng.integers(0, 2))
        cands = np.stack([pos[0] * 4 + rng.integers(-220, 220, 40), pos[1] * 4 + rng.integers(-170, 170, 40)], 1).astype(np.int16)
        got = ctx.dist(0, metric, bt, pos, cands, t8)
        want = [oracle.dist(r, f[1], bt, pos, (int(c[0]), int(c[1])), metric, t8) for c in cands]
        assert got.tolist() == want
    oracle.ref_destroy(r)


@pytest.mark.parametrize("form", [api.PRED_WEIGHTED, api.PRED_AVERAGE, api.PRED_WEIGHTED_AVERAGE])
@pytest.mark.parametrize("metric", [api.SAD, api.SSE, api.SATD])
def test_dist_weighted_and_bipred(ctx, oracle, metric, form):
    """jmb_dist_ex standing in for compute*WP, computeBiPred*1, computeBiPred*2 (restatement pinned in test_oracle_vs_ref.py)."""
    w, h = 96, 64
    f = _frames(w, h, 15, n=3)
    ctx.ref_put(0, f[0]); ctx.ref_put(1, f[2]); ctx.pic_begin(f[1], [0, 1])
    r1, r2 = oracle.ref_create(f[0]), oracle.ref_create(f[2])
    rng = np.random.default_rng(10 * form + metric)
    for it in range(16):
        bt = int(rng.integers(1, 8)); bsx, bsy = api.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (w - bsx) // 4 + 1)) * 4, int(rng.integers(0, (h - bsy) // 4 + 1)) * 4)
        t8 = int(metric == api.SATD and bt <= 4 and rng.integers(0, 2))
        cands = np.stack([pos[0] * 4 + rng.integers(-220, 220, 30), pos[1] * 4 + rng.integers(-170, 170, 30)], 1).astype(np.int16)
        c2 = (pos[0] * 4 + int(rng.integers(-220, 220)), pos[1] * 4 + int(rng.integers(-170, 170)))
        denom = int(rng.integers(0, 8))
        wt = [int(rng.integers(-128, 128)) if it % 4 == 0 else int((1 << denom) * rng.uniform(0.5, 1.5)) for _ in range(2)]
        wp = (wt[0], wt[1], int(rng.integers(-40, 41)), denom, (1 << (denom - 1)) if denom else 0)
        got = ctx.dist_ex(0, 1, metric, form, bt, pos, cands, c2, wp, t8)
        want = [oracle.dist_ex(r1, r2, f[1], bt, pos, (int(c[0]), int(c[1])), c2, metric, form, wp, t8) for c in cands]
        assert got.tolist() == want, (it, bt, t8, wp)
    oracle.ref_destroy(r1); oracle.ref_destroy(r2)
    with pytest.raises(api.JMBError):
        ctx.dist_ex(0, 5, metric, api.PRED_AVERAGE, 1, (0, 0), [[0, 0]], (0, 0))          # ref2 not in the list
    with pytest.raises(api.JMBError):
        ctx.dist_ex(0, 1, metric, api.PRED_WEIGHTED, 1, (0, 0), [[0, 0]], (0, 0), wp=(32, 32, 0, 9, 16))   # denom out of range


def test_forward_transforms(ctx, oracle):
    rng = np.random.default_rng(20)
    b4 = rng.integers(-255, 256, size=(500, 4, 4)); b8 = rng.integers(-255, 256, size=(300, 8, 8))
    g4 = ctx.forward_transform(b4, 4); g8 = ctx.forward_transform(b8, 8)
    for i in range(len(b4)):
        assert np.array_equal(g4[i], oracle.forward4x4(b4[i]))
    for i in range(len(b8)):
        assert np.array_equal(g8[i], oracle.forward8x8(b8[i]))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("do_transform", [0, 1])
def test_quant_blocks(ctx, oracle, variant, do_transform):
    n = 4 if variant < 2 else 8
    around, cavlc8 = variant & 1, variant >= 4
    scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if cavlc8 else T.SNGL_SCAN8x8)
    cc = T.COEFF_COST4x4[0] if n == 4 else T.COEFF_COST8x8[0]
    rng = np.random.default_rng(100 + variant)
    for qp in [0, 7, 22, 28, 37, 51]:
        for cav in ([1] if cavlc8 else [0, 1]):
            if variant in (2, 3) and cav:
                continue        # n=8 with is_cavlc selects the cavlc variants
            intra = int(rng.integers(0, 2)); arw = int(rng.integers(1, 9))
            qpar = T.q_params(qp, intra, n)
            qd = api.quant_desc(n, qp, qpar, scan, cc, cav, around, arw)
            res = np.concatenate([rng.integers(-a, a + 1, size=(40, n, n)) for a in (2, 30, 255, 1023)])
            res[::9, 0, 0] = 0
            coef = res if do_transform else np.stack([(oracle.forward4x4 if n == 4 else oracle.forward8x8)(b) for b in res])
            got = ctx.quant_blocks(qd, coef, do_transform=do_transform, cost0=3)
            for i in range(len(res)):
                tc = (oracle.forward4x4 if n == 4 else oracle.forward8x8)(res[i])
                want = oracle.quant(variant, tc, qp, qpar, scan, cc, cav, arw=arw, cost0=3)
                assert got["nonzero"][i] == want["nonzero"] and got["coeff_cost"][i] == want["coeff_cost"], (qp, i)
                assert np.array_equal(got["coef"][i], want["coef"])
                assert np.array_equal(got["levels"][i], want["levels"]) and np.array_equal(got["runs"][i], want["runs"])
                if around:
                    assert np.array_equal(got["fadjust"][i], want["fadjust"])


def _mc_tq_reference(oracle, planes, cur, pred, w, h, n, qp, qpar, scan, cc, cav):
    """luma_prediction + compute_residue + forward + quant, composed from oracle pieces."""
    n_mb = len(pred); mb_w = w // 16
    levels = np.zeros((n_mb, 256), np.int16); cost = np.zeros((n_mb, 4), np.int32); cbp = np.zeros(n_mb, np.uint32)
    variant = 0 if n == 4 else (4 if cav else 2)
    for mb in range(n_mb):
        mbx, mby = (mb % mb_w) * 16, (mb // mb_w) * 16
        step = n // 4
        for by4 in range(0, 4, step):
            for bx4 in range(0, 4, step):
                b8 = (by4 >> 1) * 2 + (bx4 >> 1)
                mode = int(pred[mb]["b8mode"][b8])
                ux4, uy4 = (bx4 & ~1, by4 & ~1) if (mode < 5 or n == 8) else (bx4, by4)
                if mode == 1:
                    ux4, uy4 = 0, 0         # P16x16: one 16x16 prediction unit (macroblock.c:1225)
                mvx, mvy = [int(v) for v in pred[mb]["mv"][uy4 * 4 + ux4]]
                qx, qy = ((mbx + ux4 * 4) << 2) + mvx, ((mby + uy4 * 4) << 2) + mvy
                iy = min(max(qy >> 2, -20), h + 20 - 1 - 16); ix = min(max(qx >> 2, -32), w + 32 - 1 - 16)
                y0 = iy + 20 + (by4 - uy4) * 4; x0 = ix + 32 + (bx4 - ux4) * 4
                p = planes[qy & 3, qx & 3, y0:y0 + n, x0:x0 + n].astype(np.int32)
                s = cur[mby + by4 * 4: mby + by4 * 4 + n, mbx + bx4 * 4: mbx + bx4 * 4 + n].astype(np.int32)
                tc = (oracle.forward4x4 if n == 4 else oracle.forward8x8)(s - p)
                o = oracle.quant(variant, tc, qp, qpar, scan, cc, cav)
                # (level, run) lists -> level per scan position
                lv = np.zeros(n * n, np.int16)
                for sgrp in range(4 if variant == 4 else 1):
                    k = sgrp * 16 if variant == 4 else 0
                    for L, Rn in zip(o["levels"][17 * sgrp:], o["runs"][17 * sgrp:]):
                        if L == 0:
                            break
                        k += Rn; lv[k] = L; k += 1
                b = by4 * 4 + bx4 if n == 4 else b8
                levels[mb, b * n * n:(b + 1) * n * n] = lv
                cost[mb, b8] += o["coeff_cost"]
                if o["nonzero"]:
                    cbp[mb] |= (1 << (by4 * 4 + bx4)) if n == 4 else (51 << (4 * b8 - 2 * (b8 & 1)))
    return levels, cost, cbp


@pytest.mark.parametrize("n,cav", [(4, 1), (4, 0), (8, 0), (8, 1)])
def test_mc_tq(ctx, oracle, n, cav):
    w, h = 64, 48
    f = _frames(w, h, 15)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0]); planes = oracle.planes(r)
    rng = np.random.default_rng(15 + n)
    n_mb = (w // 16) * (h // 16)
    pred = np.zeros(n_mb, api.MB_PRED)
    pred["mv"] = rng.integers(-60, 60, size=(n_mb, 16, 2))
    pred["mv"][0] = rng.integers(-400, 400, size=(16, 2))       # far outside: origin clamps
    pred["b8mode"] = rng.integers(1, 5 if n == 8 else 8, size=(n_mb, 4))
    for qp in (20, 28, 40):
        qpar = T.q_params(qp, 0, n)
        scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if cav else T.SNGL_SCAN8x8)
        cc = T.COEFF_COST4x4[0] if n == 4 else T.COEFF_COST8x8[0]
        qd = api.quant_desc(n, qp, qpar, scan, cc, cav)
        got = ctx.mc_tq(pred, qd)
        want = _mc_tq_reference(oracle, planes, f[1], pred, w, h, n, qp, qpar, scan, cc, cav)
        for a, b, name in zip(got, want, ("levels", "coeff_cost", "cbp_blk")):
            assert np.array_equal(a, b), (name, qp)
    oracle.ref_destroy(r)


def test_errors_are_loud(ctx):
    with pytest.raises(api.JMBError):
        ctx.ref_put(0, np.zeros((30, 30), np.uint16))           # not a multiple of 16
    with pytest.raises(api.JMBError):
        ctx.ref_put(99, np.zeros((32, 32), np.uint16))
    bad = np.zeros(1, api.ME_REQ); bad["blocktype"] = 9
    ctx.ref_put(0, np.zeros((32, 32), np.uint16)); ctx.pic_begin(np.zeros((32, 32), np.uint16), [0])
    with pytest.raises(api.JMBError):
        ctx.me_search(bad)
    # the frame layout is validated on the device: a wrong partition order / a bad field is reported, not computed
    ctx.configure(search_range=4)
    good = _frame_reqs(32, 32, np.random.default_rng(0), api.SEARCH_FULL, api.REQ_SUBPEL, lam=10)
    ctx.me_search(good, frame=True)
    swapped = good.copy(); swapped[[3, 4]] = swapped[[4, 3]]
    with pytest.raises(api.JMBError, match="frame-layout"):
        ctx.me_search(swapped, frame=True)
    badlam = good.copy(); badlam["lambda"][50] = 1 << 20
    with pytest.raises(api.JMBError, match="lambda"):
        ctx.me_search(badlam, frame=True)
    ctx.me_search(good, frame=True)        # the error word is cleared once reported
    # the integer full search is a SAD search: under MEDistortionFPel = SSE it refuses, sub-pel-only requests still run
    ctx.configure(search_range=4, metric=(api.SSE, api.SSE, api.SSE))
    with pytest.raises(api.JMBError, match="MEDistortionFPel"):
        ctx.me_search(good, frame=True)
    sub = good.copy(); sub["flags"] = api.REQ_SUBPEL | api.REQ_SKIP_INT
    ctx.me_search(sub, frame=True)
    ctx.configure(search_range=4)


def test_pred_from_results_and_resident_chain(ctx, oracle):
    """all_mv fill (mv_search.c:1005-1014) on the device, and the host-buffer-free chain
    me_search_frame -> pred_from_results(NULL, NULL) -> mc_tq(NULL) gives what the explicit chain gives."""
    w, h, R = 64, 48, 8
    f = _frames(w, h, 21, motion=(2, 1))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    rng = np.random.default_rng(21)
    reqs = _frame_reqs(w, h, rng, api.SEARCH_FULL, api.REQ_SUBPEL, lam=40)
    res = ctx.me_search(reqs, frame=True)
    n_mb = len(reqs) // api.NPART
    qd = api.quant_desc(4, 30, T.q_params(30, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    r = res.reshape(n_mb, api.NPART)
    for mode in range(1, 8):
        pred = ctx.pred_from_results(res, mode)
        for b in range(16):
            slot = api.part_slot(mode, b & 3, b >> 2)
            assert np.array_equal(pred["mv"][:, b, 0], r["mv_x"][:, slot]) and np.array_equal(pred["mv"][:, b, 1], r["mv_y"][:, slot])
        assert (pred["b8mode"] == mode).all() and (pred["ref"] == 0).all()
        want = ctx.mc_tq(pred, qd)
        ctx._ck(ctx.L.jmb_pred_from_results(ctx.h, None, n_mb, mode, None, api.HOST))
        lv = np.zeros((n_mb, 256), np.int16); cc = np.zeros((n_mb, 4), np.int32); cbp = np.zeros(n_mb, np.uint32)
        ctx._ck(ctx.L.jmb_mc_tq(ctx.h, None, n_mb, qd.ctypes.data, lv.ctypes.data, cc.ctypes.data, cbp.ctypes.data, api.HOST))
        assert np.array_equal(lv, want[0]) and np.array_equal(cc, want[1]) and np.array_equal(cbp, want[2])


@pytest.mark.parametrize("n", [4, 8])
def test_mc_tq_modes_equals_per_mode_chain(ctx, n):
    """jmb_mc_tq_modes (all partition modes, one launch, prediction straight from the search results) gives exactly what
    pred_from_results + mc_tq give mode by mode -- and those are checked against the oracle above."""
    w, h, R = 64, 48, 8
    f = _frames(w, h, 31, motion=(1, 2))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    reqs = _frame_reqs(w, h, np.random.default_rng(31), api.SEARCH_FULL, api.REQ_SUBPEL, lam=30)
    res = ctx.me_search(reqs, frame=True)
    n_mb = len(reqs) // api.NPART
    scan = T.SNGL_SCAN if n == 4 else T.SNGL_SCAN8x8
    qd = api.quant_desc(n, 26, T.q_params(26, 0, n), scan, (T.COEFF_COST4x4 if n == 4 else T.COEFF_COST8x8)[0], 0)
    mask = 0x7F if n == 4 else 0x0F
    lv, cc, cbp = ctx.mc_tq_modes(res, qd, mask)
    lv2, cc2, cbp2 = ctx.mc_tq_modes(None, qd, mask, n_mb=n_mb)          # resident results
    assert np.array_equal(lv, lv2) and np.array_equal(cc, cc2) and np.array_equal(cbp, cbp2)
    for mode in range(1, 8):
        if not (mask >> (mode - 1)) & 1:
            assert not lv[mode - 1].any() or True
            continue
        want = ctx.mc_tq(ctx.pred_from_results(res, mode), qd)
        assert np.array_equal(lv[mode - 1], want[0]) and np.array_equal(cc[mode - 1], want[1]) and np.array_equal(cbp[mode - 1], want[2]), mode
    if n == 8:
        with pytest.raises(api.JMBError):
            ctx.mc_tq_modes(res, qd, 0x7F)


def test_inverse_transforms(ctx, oracle):
    rng = np.random.default_rng(40)
    b4 = rng.integers(-40000, 40001, size=(400, 4, 4)); b8 = rng.integers(-40000, 40001, size=(200, 8, 8))
    g4 = ctx.inverse_transform(b4, 4); g8 = ctx.inverse_transform(b8, 8)
    for i in range(len(b4)):
        assert np.array_equal(g4[i], oracle.inverse4x4(b4[i]))
    for i in range(len(b8)):
        assert np.array_equal(g8[i], oracle.inverse8x8(b8[i]))


@pytest.mark.parametrize("n,cav", [(4, 1), (4, 0), (8, 0), (8, 1)])
@pytest.mark.parametrize("qp", [22, 30, 38])
def test_luma_residual_coding_modes(ctx, oracle, n, cav, qp):
    """Device luma_residual_coding (prediction .. thresholding .. reconstruction .. SSE) for every mode of every macroblock
    against the oracle's restatement of macroblock.c:806-1257, fed with the same prediction (oracle planes + the mvs found)."""
    w, h, R = 64, 48, 8
    f = _frames(w, h, 41, motion=(2, -1))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    reqs = _frame_reqs(w, h, np.random.default_rng(41), api.SEARCH_FULL, api.REQ_SUBPEL, lam=T.lambda_me(qp))
    res = ctx.me_search(reqs, frame=True)
    n_mb = len(reqs) // api.NPART
    scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if cav else T.SNGL_SCAN8x8)
    cc = (T.COEFF_COST4x4 if n == 4 else T.COEFF_COST8x8)[0]
    qpar = T.q_params(qp, 0, n)
    qd = api.quant_desc(n, qp, qpar, scan, cc, cav)
    mask = 0x7F if n == 4 else 0x0F
    got = ctx.luma_residual_coding_modes(res, qd, mask)
    r = oracle.ref_create(f[0]); planes = oracle.planes(r); oracle.ref_destroy(r)
    mbw = w // 16
    seen_reset8 = seen_reset_mb = seen_coded = 0
    for mode in range(1, 8):
        if not (mask >> (mode - 1)) & 1:
            continue
        pred_tab = ctx.pred_from_results(res, mode)
        for mb in range(n_mb):
            mbx, mby = (mb % mbw) * 16, (mb // mbw) * 16
            pred = np.zeros((16, 16), np.uint16)
            for by4 in range(4):
                for bx4 in range(4):
                    ux4, uy4 = (bx4 & ~1, by4 & ~1) if (mode < 5 or n == 8) else (bx4, by4)
                    if mode == 1:
                        ux4, uy4 = 0, 0         # P16x16: one 16x16 prediction unit (macroblock.c:1225)
                    mvx, mvy = [int(v) for v in pred_tab[mb]["mv"][uy4 * 4 + ux4]]
                    qx, qy = ((mbx + ux4 * 4) << 2) + mvx, ((mby + uy4 * 4) << 2) + mvy
                    iy = min(max(qy >> 2, -20), h + 20 - 1 - 16); ix = min(max(qx >> 2, -32), w + 32 - 1 - 16)
                    y0 = iy + 20 + (by4 - uy4) * 4; x0 = ix + 32 + (bx4 - ux4) * 4
                    pred[by4 * 4:by4 * 4 + 4, bx4 * 4:bx4 * 4 + 4] = planes[qy & 3, qx & 3, y0:y0 + 4, x0:x0 + 4]
            src = f[1][mby:mby + 16, mbx:mbx + 16]
            lv, c8, cbp, cbpb, rec, sse = oracle.luma_residual_coding(src, pred, n, qp, qpar, scan, cc, cav)
            m = mode - 1
            assert np.array_equal(got["levels"][m, mb], lv), (mode, mb)
            assert np.array_equal(got["cost8"][m, mb], c8) and got["cbp"][m, mb] == cbp and got["cbp_blk"][m, mb] == cbpb, (mode, mb)
            assert np.array_equal(got["recon"][m, mb], rec) and got["sse"][m, mb] == sse, (mode, mb)
            seen_coded += cbp != 0; seen_reset_mb += (cbp == 0 and lv.any()); seen_reset8 += bool((c8 == 0).any() and cbp != 0)
    assert seen_coded > 0 or qp > 30                        # coarse quantisers may threshold every macroblock away
    no_recon = ctx.luma_residual_coding_modes(None, qd, mask, n_mb=n_mb, want_recon=False)        # resident results, no recon
    assert np.array_equal(no_recon["sse"], got["sse"]) and np.array_equal(no_recon["levels"], got["levels"])


def test_luma_residual_coding_explicit_prediction(ctx):
    """The explicit-prediction form (what the shim's verify hook feeds from JM's own state) equals the all-modes form,
    also for a sub-range of macroblocks (first_mb)."""
    w, h, R = 64, 48, 8
    f = _frames(w, h, 43, motion=(-2, 2))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    reqs = _frame_reqs(w, h, np.random.default_rng(43), api.SEARCH_FULL, api.REQ_SUBPEL, lam=60)
    res = ctx.me_search(reqs, frame=True)
    qd = api.quant_desc(4, 27, T.q_params(27, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    allm = ctx.luma_residual_coding_modes(res, qd)
    for mode in (1, 2, 4, 7):
        pred = ctx.pred_from_results(res, mode)
        for first, cnt in ((0, len(pred)), (5, 4)):
            o = ctx.luma_residual_coding(pred[first:first + cnt], qd, first_mb=first)
            for k in ("levels", "cost8", "cbp_blk", "cbp", "recon", "sse"):
                assert np.array_equal(o[k], allm[k][mode - 1][first:first + cnt]), (mode, first, k)


def test_full_size_1080p_properties(ctx):
    """BASELINE size (1080p, +-32): size-independent properties of the whole-picture search instead of a CPU replay.
    For every sampled request: the integer cost returned is exactly lambda*bits + (SAD << 5) at the returned mv (SAD
    re-measured through jmb_dist); no candidate of a random sample of the window is cheaper (and none ties with a smaller
    spiral index); searching again with min_mcost = the found cost finds nothing better (idempotence)."""
    w, h, R = 1920, 1088, 32
    f = synth.luma_frames(w, h, 2, seed=1234, motion=(5, 3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    rng = np.random.default_rng(77)
    reqs = _frame_reqs(w, h, rng, api.SEARCH_FULL, 0, lam=187, jitter=8, base=(20, 12))
    res = ctx.me_search(reqs, frame=True)
    assert len(res) == 8160 * 41

    def mvbits(v):
        return 1 if v == 0 else 2 * (abs(int(v)).bit_length() - 1) + 3

    def spiral_index(dx, dy):
        l = max(abs(dx), abs(dy))
        if l == 0:
            return 0
        base = (2 * l - 1) ** 2
        if abs(dy) == l and abs(dx) < l:
            return base + 2 * (dx + l - 1) + (dy > 0)
        return base + 2 * (2 * l - 1) + 2 * (dy + l) + (dx > 0)

    for i in rng.choice(len(reqs), 60, replace=False):
        q, o = reqs[i], res[i]
        pos = (int(q["pos_x"]), int(q["pos_y"])); bt = int(q["blocktype"]); lam = int(q["lambda"][0])
        px, py, cx, cy = int(q["pred_x"]), int(q["pred_y"]), int(q["center_x"]), int(q["center_y"])
        mvx, mvy = int(o["imv_x"]), int(o["imv_y"])
        assert abs(mvx - cx) <= 4 * R and abs(mvy - cy) <= 4 * R and mvx % 4 == 0 and mvy % 4 == 0
        cands = [(mvx, mvy)] + [(cx + 4 * int(rng.integers(-R, R + 1)), cy + 4 * int(rng.integers(-R, R + 1))) for _ in range(40)]
        sads = ctx.dist(0, api.SAD, bt, pos, [(pos[0] * 4 + a, pos[1] * 4 + b) for a, b in cands])
        costs = [lam * (mvbits(a - px) + mvbits(b - py)) + (int(s) << 5) for (a, b), s in zip(cands, sads)]
        assert costs[0] == int(o["icost"]), (i, costs[0], int(o["icost"]))
        best_idx = spiral_index((mvx - cx) // 4, (mvy - cy) // 4)
        for (a, b), c in zip(cands[1:], costs[1:]):
            assert c > costs[0] or (c == costs[0] and spiral_index((a - cx) // 4, (b - cy) // 4) >= best_idx), (i, (a, b), c, costs[0])
    # idempotence on one macroblock group: min_mcost = found cost -> nothing is strictly better, the centre is kept
    g = reqs[:41].copy(); g["min_mcost"] = res["icost"][:41]
    again = ctx.me_search(g, frame=True)
    assert np.array_equal(again["icost"], res["icost"][:41])
    assert np.array_equal(again["imv_x"], g["center_x"]) and np.array_equal(again["imv_y"], g["center_y"])


@pytest.mark.parametrize("R", [16, 32])
def test_full_search_group_with_scattered_centres(ctx, oracle, R):
    """One macroblock's 41 requests with centres tens of pels apart: the union of their windows exceeds one staged chunk
    in both directions, so the chunk loops, the repeated TMA loads and the 'a displacement belongs to some partitions only'
    logic are all exercised -- each request must still equal its own stand-alone search."""
    w, h = 208, 160
    f = _frames(w, h, 51, motion=(4, -3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(51)
    parts = api.mb_partitions()
    for mbx, mby in [(96, 64), (0, 0), (192, 144)]:
        reqs = np.zeros(api.NPART, api.ME_REQ)
        for k, (t, x, y) in enumerate(parts):
            q = reqs[k]
            q["blocktype"], q["pos_x"], q["pos_y"] = t, mbx + x, mby + y
            p = rng.integers(-160, 161, 2)                    # +-40 pels
            q["pred_x"], q["pred_y"] = p
            q["center_x"], q["center_y"] = ((int(p[0]) + 2) >> 2) * 4, ((int(p[1]) + 2) >> 2) * 4
            q["mode"], q["flags"], q["lambda"], q["min_mcost"] = api.SEARCH_FULL, 0, int(rng.integers(1, 300)), BIG
        res = ctx.me_search(reqs, frame=True)
        _check_full(ctx, oracle, r, f[1], reqs, res, R)
    oracle.ref_destroy(r)


def test_luma_residual_coding_with_a_non_standard_scan(ctx, oracle):
    """A scan other than the frame zig-zag (here JM's 4x4 field scan, block.c:178) takes the table-driven quantiser instead
    of the compile-time one: same oracle, same answers."""
    field_scan = np.array([(0, 0), (0, 1), (1, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 0), (2, 1), (2, 2), (2, 3),
                           (3, 0), (3, 1), (3, 2), (3, 3)], np.uint8)
    w, h, R, qp = 64, 48, 8, 26
    f = _frames(w, h, 61, motion=(1, -2))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    reqs = _frame_reqs(w, h, np.random.default_rng(61), api.SEARCH_FULL, api.REQ_SUBPEL, lam=T.lambda_me(qp))
    res = ctx.me_search(reqs, frame=True)
    qpar = T.q_params(qp, 0, 4)
    qd = api.quant_desc(4, qp, qpar, field_scan, T.COEFF_COST4x4[0], 1)
    got = ctx.luma_residual_coding_modes(res, qd, 0x01)
    raw = ctx.mc_tq_modes(res, qd, 0x01)
    r = oracle.ref_create(f[0]); planes = oracle.planes(r); oracle.ref_destroy(r)
    pred_tab = ctx.pred_from_results(res, 1)
    for mb in range(len(pred_tab)):
        mbx, mby = (mb % (w // 16)) * 16, (mb // (w // 16)) * 16
        mvx, mvy = [int(v) for v in pred_tab[mb]["mv"][0]]
        qx, qy = (mbx << 2) + mvx, (mby << 2) + mvy
        iy = min(max(qy >> 2, -20), h + 20 - 1 - 16); ix = min(max(qx >> 2, -32), w + 32 - 1 - 16)
        pred = planes[qy & 3, qx & 3, iy + 20:iy + 36, ix + 32:ix + 48]
        lv, c8, cbp, cbpb, rec, sse = oracle.luma_residual_coding(f[1][mby:mby + 16, mbx:mbx + 16], pred, 4, qp, qpar, field_scan, T.COEFF_COST4x4[0], 1)
        assert np.array_equal(got["levels"][0, mb], lv) and got["cbp"][0, mb] == cbp and got["sse"][0, mb] == sse and np.array_equal(got["recon"][0, mb], rec)
        if cbp == 15:                                       # nothing thresholded away: the raw transform/quant levels are the same
            assert np.array_equal(raw[0][0, mb], lv)


@pytest.mark.parametrize("variant", [6, 7, 8, 9, 10, 11, 12])
def test_quant_dc_ac_family(ctx, oracle, variant):
    """jmb_quant_list standing in for quant_ac4x4_*, quant_dc4x4_normal, quant_dc2x2_*, quant_dc4x2_* (mapping: api.qlist_plan,
    pinned against the real functions in test_oracle_vs_ref.py)."""
    scan420 = np.array([(0, 0), (0, 1), (0, 2), (0, 3)], np.uint8)
    scan422 = np.array([(0, 0), (0, 1), (1, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3)], np.uint8)
    ncoef = {6: 16, 7: 16, 8: 16, 9: 4, 10: 4, 11: 8, 12: 8}[variant]
    scan = {9: scan420, 10: scan420, 11: scan422, 12: scan422}.get(variant, T.SNGL_SCAN)
    rng = np.random.default_rng(80 + variant)
    for it in range(60):
        qp = int(rng.integers(0, 52)); amp = int(rng.choice([6, 80, 600, 4000]))
        coef = rng.integers(-amp, amp + 1, size=ncoef)
        qpar = T.q_params(qp, intra=int(rng.integers(0, 2)), n=4)
        plan = api.qlist_plan(variant, qp, qpar if variant <= 7 else qpar[0, 0], scan, T.COEFF_COST4x4[0], int(rng.integers(0, 2)), 1 + it % 8)
        got, want = ctx.quant_list(plan, coef, cost0=3), oracle.quant_list(plan, coef, cost0=3)
        for k in ("nonzero", "coeff_cost"):
            assert got[k] == want[k], (variant, it, k)
        for k in ("coef", "levels", "runs", "fadjust"):
            assert np.array_equal(got[k], want[k]), (variant, it, k)


def test_hadamards(ctx, oracle):
    rng = np.random.default_rng(90)
    for kind, per in [(api.HAD_4X4, 16), (api.IHAD_4X4, 16), (api.HAD_4X2, 8), (api.IHAD_4X2, 8), (api.HAD_2X2, 4), (api.IHAD_2X2, 4)]:
        v = rng.integers(-30000, 30001, size=(100, per))
        got = ctx.hadamard(kind, v, per)
        for i in range(len(v)):
            assert np.array_equal(got[i], oracle.hadamard(kind, v[i])), (kind, i)

"""Pins the oracle (oracle/jm_oracle.c) to the REAL JM 19.0 leaf functions (oracle/_ref/libjmref.so,
compiled from the reference's own sources).  CPU only.  Skipped when the reference build is absent."""
import numpy as np
import pytest

from jm_b200 import h264_tables as T
from jm_b200 import synth
from oracle import pyoracle as po

W, H, R = 96, 64, 8


@pytest.fixture(scope="module")
def pair(oracle, have_ref):
    f = synth.luma_frames(W, H, 2, seed=7, motion=(3, -2))
    ref = po.JMRef(W, H, search_range=R)
    ref.set_ref(f[0]); ref.set_cur(f[1])
    r = oracle.ref_create(f[0])
    yield oracle, ref, r, f
    oracle.ref_destroy(r)


def test_subpel_planes(pair):
    o, ref, r, f = pair
    a, b = o.planes(r), ref.planes()
    for fy in range(4):
        for fx in range(4):
            assert np.array_equal(a[fy, fx], b[fy, fx]), (fy, fx)


def test_subpel_planes_extreme_values(oracle, have_ref):
    rng = np.random.default_rng(3)
    img = rng.choice([0, 255], size=(H, W)).astype(np.uint16)      # forces the iClip1 clamps
    ref = po.JMRef(W, H, search_range=R); ref.set_ref(img)
    r = oracle.ref_create(img)
    assert np.array_equal(oracle.planes(r), ref.planes())
    oracle.ref_destroy(r)


def test_spiral_and_mvbits(pair):
    o, ref, r, f = pair
    assert np.array_equal(o.spiral(R), ref.spiral()[:(2 * R + 1) ** 2])
    for v in list(range(-130, 131)) + [255, -255, ref.max_mvd, -ref.max_mvd]:   # table spans +-max_mvd
        assert o.mvbits(v) == ref.mvbits(v), v
    big = po.JMRef(64, 64, search_range=32); big.spiral()
    assert big.max_mvd == 1023
    for v in [256, 511, 512, 1023, -1023]:
        assert o.mvbits(v) == big.mvbits(v), v


@pytest.mark.parametrize("metric", [po.SAD, po.SSE, po.SATD])
def test_distortion(pair, metric):
    o, ref, r, f = pair
    rng = np.random.default_rng(metric)
    for _ in range(300):
        bt = int(rng.integers(1, 8))
        bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        # include candidates far outside the picture so every clamp rule is hit
        cand = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        t8 = int(metric == po.SATD and bt <= 4 and rng.integers(0, 2))
        assert o.dist(r, f[1], bt, pos, cand, metric, t8) << 5 == ref.dist(metric, bt, pos, cand, t8)


@pytest.mark.parametrize("form", [1, 2, 3])
@pytest.mark.parametrize("metric", [po.SAD, po.SSE, po.SATD])
def test_weighted_and_bipred_distortion(pair, metric, form):
    """computeSADWP/SSEWP/SATDWP, computeBiPredSAD1/SSE1/SATD1, computeBiPredSAD2/SSE2/SATD2 (me_distortion.c:434-1520)."""
    o, ref, r, f = pair
    f2 = synth.luma_frames(W, H, 3, seed=7, motion=(3, -2))[2]
    ref.set_ref2(f2)
    r2 = o.ref_create(f2)
    rng = np.random.default_rng(10 * form + metric)
    for it in range(300):
        bt = int(rng.integers(1, 8))
        bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        c1 = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        c2 = (pos[0] * 4 + int(rng.integers(-200, 200)), pos[1] * 4 + int(rng.integers(-160, 160)))
        t8 = int(metric == po.SATD and bt <= 4 and rng.integers(0, 2))
        denom = int(rng.integers(0, 8))
        # weights around 1.0 and wild ones (negative, large: the iClip1 of the weighted sample is hit), offsets of both signs
        w = [int(rng.integers(-128, 128)) if it % 4 == 0 else int((1 << denom) * rng.uniform(0.5, 1.5)) for _ in range(2)]
        wp = (w[0], w[1], int(rng.integers(-40, 41)), denom, (1 << (denom - 1)) if denom else 0)
        assert o.dist_ex(r, r2, f[1], bt, pos, c1, c2, metric, form, wp, t8) << 5 == ref.dist_ex(metric, form, bt, pos, c1, c2, wp, t8), (it, bt, wp)
    o.ref_destroy(r2)


def test_distortion_early_exit_is_min_mcost(pair):
    """JM's row / sub-block early exit returns the threshold itself (dist_scale_f, mv_search.h:19-23): a losing candidate."""
    o, ref, r, f = pair
    ref.set_ref2(f[0])
    for form in range(4):
        for metric in (po.SAD, po.SSE, po.SATD):
            full = ref.dist_ex(metric, form, 1, (16, 16), (70, 60), (50, 70))
            assert ref.dist_ex(metric, form, 1, (16, 16), (70, 60), (50, 70), min_mcost=full - 32) == full - 32
            assert ref.dist_ex(metric, form, 1, (16, 16), (70, 60), (50, 70), min_mcost=full) == full


def _bipred_replay(o, r1, r2, cur, bt, pos, pred1, pred2, mv1, mv2, lam, min_mcost, metric, form, wp, t8, offsets, start=0):
    """What jm_b200/shim/jm_wrap.c does for the bi-predictive searches: the COMPLETE distortion of every candidate (one
    jmb_dist_ex call on the device; here the restatement), then JM's sequential selection -- candidates whose mv cost alone
    reaches min_mcost skipped, a distortion above the remaining budget returned as the budget itself (dist_scale_f,
    mv_search.h:19), strict '<'.  offsets = the candidates' quarter-pel offsets from mv1 in JM's order."""
    px, py = pos[0] * 4, pos[1] * 4
    c2 = (px + mv2[0], py + mv2[1])
    mc2 = lam * (o.mvbits(mv2[0] - pred2[0]) + o.mvbits(mv2[1] - pred2[1]))
    best = 0
    for k, (ox, oy) in enumerate(offsets):
        if k < start:
            continue
        cx, cy = mv1[0] + ox, mv1[1] + oy
        mcost = lam * (o.mvbits(cx - pred1[0]) + o.mvbits(cy - pred1[1])) + mc2
        if mcost >= min_mcost:
            continue
        d = o.dist_ex(r1, r2, cur, bt, pos, (px + cx, py + cy), c2, metric, form, wp, t8)
        thr = min_mcost - mcost
        mcost += thr if d > (thr >> 5) else (d << 5)
        if mcost < min_mcost:
            best, min_mcost = k, mcost
    return (mv1[0] + offsets[best][0], mv1[1] + offsets[best][1]), min_mcost


@pytest.mark.parametrize("form", [2, 3])
def test_bipred_searches_decide_like_batched_distortions_plus_replay(pair, form):
    """full_search_bipred_motion_estimation (me_fullsearch.c:112) and sub_pel_bipred_motion_estimation (:299), the REAL
    functions with their early exits, against complete distortions + the replay the shim performs."""
    o, ref, r, f = pair
    f2 = synth.luma_frames(W, H, 3, seed=7, motion=(3, -2))[2]
    ref.set_ref2(f2)
    r2 = o.ref_create(f2)
    sp = o.spiral(R)
    rng = np.random.default_rng(40 + form)
    for it in range(24):
        bt = int(rng.integers(1, 5)); bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // bsx + 1)) * bsx, int(rng.integers(0, (H - bsy) // bsy + 1)) * bsy)
        pred1 = (int(rng.integers(-24, 25)), int(rng.integers(-24, 25))); pred2 = (int(rng.integers(-24, 25)), int(rng.integers(-24, 25)))
        mv1 = (int(rng.integers(-6, 7)) * 4, int(rng.integers(-6, 7)) * 4); mv2 = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
        lam = int(rng.integers(1, 300)); t8 = int(bt <= 4 and rng.integers(0, 2))
        wp = (int(rng.integers(20, 45)), int(rng.integers(20, 45)), int(rng.integers(-5, 6)), 5, 16)
        big = po.DISTBLK_MAX if it % 3 else int(rng.integers(2000, 60000))          # an incoming bound as well
        # integer stage: (2*2+1)^2 candidates in spiral order, SAD
        sr = 2
        offs = [(4 * int(x), 4 * int(y)) for x, y in sp[:(2 * sr + 1) ** 2]]
        got = _bipred_replay(o, r, r2, f[1], bt, pos, pred1, pred2, mv1, mv2, lam, big, po.SAD, form, wp, 0, offs)
        assert got == ref.bipred_search(0, form, bt, pos, pred1, pred2, mv1, mv2, sr << 2, [lam] * 3, big, 0, wp), (it, "full")
        # sub-pel stages (me_fullsearch.c:330-392): nine half-pel candidates from position 0 (the full-pel metric differs from the
        # half-pel one: start_me_refinement_hp = 0), then quarter-pel positions 1..8 around the winner (start_me_refinement_qp = 1)
        m1, c = _bipred_replay(o, r, r2, f[1], bt, pos, pred1, pred2, mv1, mv2, lam, big, po.SATD, form, wp, t8,
                               [(2 * int(x), 2 * int(y)) for x, y in sp[:9]])
        m1, c = _bipred_replay(o, r, r2, f[1], bt, pos, pred1, pred2, m1, mv2, lam, c, po.SATD, form, wp, t8,
                               [(int(x), int(y)) for x, y in sp[:9]], start=1)
        assert (m1, c) == ref.bipred_search(1, form, bt, pos, pred1, pred2, mv1, mv2, 0, [lam] * 3, big, t8, wp), (it, "sub-pel")
    o.ref_destroy(r2)


def test_full_search(pair):
    o, ref, r, f = pair
    rng = np.random.default_rng(11)
    for _ in range(40):
        bt = int(rng.integers(1, 8))
        bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        pred = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))
        center = (((pred[0] + 2) >> 2) * 4, ((pred[1] + 2) >> 2) * 4)
        lam = int(rng.integers(1, 400))
        big = po.DISTBLK_MAX
        assert o.full_search(r, f[1], bt, pos, pred, center, lam, big, R) == ref.full_search(bt, pos, pred, center, lam, big)


def test_sub_pel(pair):
    o, ref, r, f = pair
    rng = np.random.default_rng(12)
    for _ in range(60):
        bt = int(rng.integers(1, 8))
        bsx, bsy = po.BLOCK_SIZE[bt]
        pos = (int(rng.integers(0, (W - bsx) // 4 + 1)) * 4, int(rng.integers(0, (H - bsy) // 4 + 1)) * 4)
        pred = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))
        mv = (int(rng.integers(-12, 12)) * 4, int(rng.integers(-12, 12)) * 4)
        lam = [int(rng.integers(1, 400))] * 3
        t8 = int(bt <= 4 and rng.integers(0, 2))
        big = po.DISTBLK_MAX
        got = o.sub_pel(r, f[1], bt, pos, pred, mv, lam, big, po.SATD, po.SATD, 0, 1, t8)
        assert got == ref.sub_pel(bt, pos, pred, mv, lam, big, t8)


def test_fast_full_search(oracle, have_ref):
    f = synth.luma_frames(W, H, 2, seed=9, motion=(-4, 3))
    ref = po.JMRef(W, H, search_range=R, fast_full=1)
    ref.set_ref(f[0]); ref.set_cur(f[1]); ref.spiral()
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(5)
    for mb in [(0, 0), (80, 48), (32, 16), (80, 0), (0, 48)]:
        pmv = (int(rng.integers(-30, 30)), int(rng.integers(-30, 30)))
        c_ref = ref.ffs_setup(mb, pmv)
        c = oracle.ffs_center(pmv, R)
        assert c == c_ref
        bs = oracle.ffs_setup(r, f[1], mb, c, R)
        for bt, idxs in [(7, range(16)), (6, [0, 1, 2, 3, 8, 9, 10, 11]), (5, range(0, 16, 2)), (4, [0, 2, 8, 10]),
                         (3, [0, 2]), (2, [0, 8]), (1, [0])]:
            for i in idxs:
                assert np.array_equal(bs[bt, i], ref.ffs_sad(bt, i)), (mb, bt, i)
                pos = (mb[0] + (i & 3) * 4, mb[1] + (i >> 2) * 4)
                pred = (pmv[0] + int(rng.integers(-6, 6)), pmv[1] + int(rng.integers(-6, 6)))
                lam = int(rng.integers(1, 300))
                assert oracle.ffs_search(bs, R, bt, i, c, pred, lam, po.DISTBLK_MAX, ref.max_mvd) == \
                    ref.ffs_search(bt, pos, pred, lam, po.DISTBLK_MAX)
    oracle.ref_destroy(r)


def test_transforms_and_hadamard(oracle, have_ref):
    ref = po.JMRef(32, 32, search_range=4)
    rng = np.random.default_rng(1)
    for _ in range(200):
        b4 = rng.integers(-255, 256, size=(4, 4)); b8 = rng.integers(-255, 256, size=(8, 8))
        assert np.array_equal(oracle.forward4x4(b4), ref.forward4x4(b4))
        assert np.array_equal(oracle.forward8x8(b8), ref.forward8x8(b8))
        assert oracle.hadamard4x4(b4) == ref.hadamard4x4(b4)
        assert oracle.hadamard8x8(b8) == ref.hadamard8x8(b8)


def test_inverse_transforms(oracle, have_ref):
    """inverse4x4 / inverse8x8 (lcommon/src/transform.c:70, :450) on dequantised-coefficient-like input."""
    ref = po.JMRef(32, 32, search_range=4)
    rng = np.random.default_rng(2)
    for _ in range(300):
        amp = int(rng.choice([64, 2000, 40000]))
        b4 = rng.integers(-amp, amp + 1, size=(4, 4)); b8 = rng.integers(-amp, amp + 1, size=(8, 8))
        assert np.array_equal(oracle.inverse4x4(b4), ref.inverse4x4(b4))
        assert np.array_equal(oracle.inverse8x8(b8), ref.inverse8x8(b8))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_quant(oracle, have_ref, variant):
    ref = po.JMRef(32, 32, search_range=4)
    rng = np.random.default_rng(variant)
    n = 4 if variant < 2 else 8
    scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if variant >= 4 else T.SNGL_SCAN8x8)
    cc = T.COEFF_COST4x4[0] if n == 4 else T.COEFF_COST8x8[0]
    for it in range(300):
        qp = int(rng.integers(0, 52))
        amp = int(rng.choice([3, 40, 255, 1023]))   # residual range of 8- and 10-bit video (JM itself overflows int32 beyond)
        res = rng.integers(-amp, amp + 1, size=(n, n))
        coef = oracle.forward4x4(res) if n == 4 else oracle.forward8x8(res)
        if it % 7 == 0:
            coef[rng.integers(0, n), rng.integers(0, n)] = 0
        qpar = T.q_params(qp, intra=int(rng.integers(0, 2)), n=n)
        cav = int(rng.integers(0, 2))
        arw = 1 + (it % 8)
        a = oracle.quant(variant, coef, qp, qpar, scan, cc, cav, arw=arw, cost0=5)
        b = ref.quant(variant, coef, qp, qpar, scan, cc, cav, arw=arw, cost0=5)
        assert a["nonzero"] == b["nonzero"] and a["coeff_cost"] == b["coeff_cost"]
        assert np.array_equal(a["coef"], b["coef"])
        assert np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["runs"], b["runs"])
        if variant % 2 == 1:
            assert np.array_equal(a["fadjust"], b["fadjust"])


SCAN_YUV420 = np.array([(0, 0), (0, 1), (0, 2), (0, 3)], np.uint8)                                  # block.c:79
SCAN_YUV422 = np.array([(0, 0), (0, 1), (1, 0), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3)], np.uint8)  # block.c:88


@pytest.mark.parametrize("variant", [6, 7, 8, 9, 10, 11, 12])
def test_quant_dc_ac_family(oracle, have_ref, variant):
    """quant_ac4x4_normal/_around, quant_dc4x4_normal, quant_dc2x2_*, quant_dc4x2_* against the list quantiser fed through
    the mapping of jm_b200.api.qlist_plan (the same mapping the shim applies)."""
    from jm_b200 import api
    ref = po.JMRef(32, 32, search_range=4)
    rng = np.random.default_rng(60 + variant)
    ncoef = {6: 16, 7: 16, 8: 16, 9: 4, 10: 4, 11: 8, 12: 8}[variant]
    scan = {9: SCAN_YUV420, 10: SCAN_YUV420, 11: SCAN_YUV422, 12: SCAN_YUV422}.get(variant, T.SNGL_SCAN)
    for it in range(300):
        qp = int(rng.integers(0, 52)); amp = int(rng.choice([6, 80, 600, 4000]))
        coef = rng.integers(-amp, amp + 1, size=ncoef)
        if it % 5 == 0:
            coef[rng.integers(0, ncoef)] = 0
        qpar = T.q_params(qp, intra=int(rng.integers(0, 2)), n=4)
        qp_arg = qpar if variant <= 7 else qpar[0, 0]
        cav = int(rng.integers(0, 2)); arw = 1 + it % 8
        want = ref.quant_misc(variant, coef, qp, qp_arg, scan, T.COEFF_COST4x4[0], cav, arw=arw, cost0=7)
        got = oracle.quant_list(api.qlist_plan(variant, qp, qp_arg, scan, T.COEFF_COST4x4[0], cav, arw), coef, cost0=7)
        assert got["nonzero"] == want["nonzero"] and np.array_equal(got["coef"], want["coef"]), (variant, it)
        assert np.array_equal(got["levels"], want["levels"]) and np.array_equal(got["runs"], want["runs"])
        if variant <= 7:
            assert got["coeff_cost"] == want["coeff_cost"]
        if variant == 7:
            assert np.array_equal(got["fadjust"], want["fadjust"])


def test_hadamards(oracle, have_ref):
    ref = po.JMRef(32, 32, search_range=4)
    rng = np.random.default_rng(70)
    for kind, per in [(0, 16), (1, 16), (2, 8), (3, 8), (4, 4), (5, 4)]:
        for _ in range(200):
            v = rng.integers(-30000, 30001, size=per)
            assert np.array_equal(oracle.hadamard(kind, v), ref.hadamard(kind, v)), kind


def test_mv_predictor_matches_jm():
    """oracle.pyoracle.mv_predictor (what the chain kernel k_mb_chain computes on the device, and what tests/test_gpu_frame.py
    checks it against) vs JM's own GetMotionVectorPredictorNormal: every availability pattern, reference match pattern and
    block shape / position that selects a directional rule."""
    rng = np.random.default_rng(7)
    shapes = [(16, 16, 0, 0), (16, 8, 0, 0), (16, 8, 0, 8), (8, 16, 0, 0), (8, 16, 8, 0), (8, 8, 8, 8), (8, 4, 0, 4), (4, 8, 4, 0), (4, 4, 12, 12)]
    n = 0
    for avail in range(8):
        for refs in range(27):
            for (bsx, bsy, mbx, mby) in shapes:
                nb = []
                for k in range(3):
                    r = (refs // 3 ** k) % 3 - 1          # -1 (intra), 0, 1
                    nb.append(((avail >> k) & 1, r, int(rng.integers(-300, 300)), int(rng.integers(-300, 300))))
                for ref_frame in (0, 1):
                    assert po.mv_predictor(nb, ref_frame, mbx, mby, bsx, bsy) == po.jmref_mv_predictor(nb, ref_frame, mbx, mby, bsx, bsy), (nb, ref_frame, bsx, bsy, mbx, mby)
                    n += 1
    assert n == 8 * 27 * 9 * 2


@pytest.mark.parametrize("yuv,slice_type,idc", [(1, 0, 0), (1, 1, 0), (2, 0, 0), (2, 1, 2), (0, 0, 0), (1, 2, 0), (1, 0, 2)])
def test_deblock_matches_jm(yuv, slice_type, idc):
    """jmo_deblock (the CPU restatement the GPU deblocking kernel is checked against) vs JM's own DeblockFrame on synthetic coded
    pictures: P / B / I slices, 4:0:0 / 4:2:0 / 4:2:2, 8x8-transform macroblocks, skipped and coefficient-free macroblocks,
    differing references and vectors on both lists, QPs around the filter threshold, filter offsets, DFDisableIdc 0 / 1 / 2."""
    for seed in range(3):
        rng = np.random.default_rng(100 * yuv + 10 * slice_type + seed)
        w, h = 96, 64
        luma, cb, cr, mbs = po.random_deblock_picture(rng, w, h, yuv, slice_type, idc=idc)
        want = po.jmref_deblock(luma, cb, cr, yuv, slice_type, mbs)
        got = po.deblock(luma, cb, cr, yuv, slice_type, mbs)
        assert np.array_equal(got[0], want[0]), np.argwhere(got[0] != want[0])[:5]
        assert (luma != want[0]).mean() > 0.05, "the picture must actually be filtered"
        if yuv:
            assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
            assert (cb != want[1]).mean() > 0.02

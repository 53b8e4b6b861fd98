"""GPU parity of the EPZS kernel (jmb_epzs_search / jmb_epzs_search_frame) against the CPU restatement of
EPZS_integer_motion_estimation + EPZS_sub_pel_motion_estimation (oracle/jm_oracle.c::jmo_epzs, itself pinned to the real JM
functions by tests/test_epzs_golden.py).  Bit-exact; every return path of the integer stage is exercised."""
import numpy as np
import pytest

from jm_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
BIG = po.DISTBLK_MAX
FIELDS = ("mv_x", "mv_y", "imv_x", "imv_y", "cost", "icost", "prev_sad", "exit_code")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _random_requests(rng, w, h, n, motion_q, range_q=128, subpel=True, ref_gt0_share=0.3):
    reqs = np.zeros(n, api.EPZS_REQ)
    cands = []
    for q in reqs:
        bt = int(rng.integers(1, 8)); bsx, bsy = api.BLOCK_SIZE[bt]
        q["blocktype"] = bt
        q["pos_x"] = int(rng.integers(0, (w - bsx) // bsx + 1)) * bsx; q["pos_y"] = int(rng.integers(0, (h - bsy) // bsy + 1)) * bsy
        base = np.array(motion_q) + rng.integers(-10, 11, 2)
        q["pred_x"], q["pred_y"] = base
        q["start_x"], q["start_y"] = base if rng.random() < 0.7 else base + rng.integers(-6, 7, 2)
        lam = int(rng.integers(4, 200))
        q["lambda"] = [lam, lam + 3, lam + 5]
        gt0 = rng.random() < ref_gt0_share
        fl = (api.EPZS_REF_GT0_FRAME if gt0 else 0) | (api.EPZS_ADAPT_PATTERN if rng.random() < 0.7 else 0)
        fl |= (api.EPZS_SQUARE_HINT if gt0 and bt != 1 else 0) | (api.EPZS_DUAL if rng.random() < 0.7 else 0)
        fl |= api.EPZS_SUBPEL if subpel else 0
        if bt <= 4 and rng.random() < 0.4:
            fl |= api.EPZS_TEST8X8
        q["ref"] = q["jm_ref"] = 1 if gt0 else 0
        q["pattern"] = int(rng.integers(0, 6)); q["pattern_dual"] = int(rng.integers(0, 6))
        q["range_x"] = q["range_y"] = range_q
        med = (int(rng.choice([12, 48, 192])) << 5) * int(rng.integers(0, 3))
        q["medthres"] = med; q["subthres"] = med
        q["stop"] = int(rng.integers(0, 3000)) << int(rng.integers(0, 6))
        q["prev_sad"] = BIG if rng.random() < 0.3 else int(rng.integers(0, 40000))
        q["min_mcost"] = BIG
        # predictor segments: near the true motion, far away, duplicates, out of range
        q["cand_off"] = len(cands)
        for s in range(4):
            k = int(rng.integers(0, 14 if s == 0 else 8))
            gen = s == 2 and rng.random() < 0.5
            if gen:
                fl |= api.EPZS_WINDOW_GEN
                k = 8 * int(rng.integers(1, 5)) - 1
            q["n_cand"][s] = k
            q["gate"][s] = 0 if s == 0 else int(rng.integers(0, 4))
            if not gen:
                for _ in range(k):
                    r = rng.random()
                    if r < 0.55:
                        c = np.array(motion_q) + rng.integers(-12, 13, 2)
                    elif r < 0.75:
                        c = base + rng.integers(-range_q - 20, range_q + 21, 2)
                    elif r < 0.9 and len(cands) > q["cand_off"]:
                        c = np.array(cands[int(rng.integers(q["cand_off"], len(cands)))])
                    else:
                        c = np.array([0, 0])
                    cands.append((int(c[0]), int(c[1])))
        q["flags"] = fl
    return reqs, np.array(cands, np.int16).reshape(-1, 2)


def _assert_same(got, want, reqs):
    for f in FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, (f, int(bad[0]), got[bad[0]], want[bad[0]], reqs[bad[0]])


@pytest.mark.parametrize("seed,range_q,metrics", [(1, 128, (api.SAD, api.SATD, api.SATD)), (2, 24, (api.SAD, api.SATD, api.SATD)),
                                                  (3, 128, (api.SAD, api.SAD, api.SAD)), (4, 64, (api.SAD, api.SSE, api.SSE))])
def test_epzs_search_matches_oracle(ctx, oracle, seed, range_q, metrics):
    w, h = 176, 144
    f = synth.luma_frames(w, h, 2, seed=50 + seed, motion=(3, -2))
    ctx.configure(search_range=32, metric=metrics)
    ctx.ref_put(0, f[0]); ctx.ref_put(1, f[0]); ctx.pic_begin(f[1], [0, 1])      # ref 1 = the same picture: one oracle reference serves both
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(seed)
    reqs, cands = _random_requests(rng, w, h, 700, motion_q=(-12, 8), range_q=range_q)
    got = ctx.epzs_search(reqs, cands)
    shp = 0 if metrics[0] != metrics[1] else 1
    want = oracle.epzs(r, f[1], reqs, cands, (metrics[1], metrics[2], shp, 1, 9))
    _assert_same(got, want, reqs)
    # (return 4, me_epzs_int.c:362, cannot be reached: its condition implies the one of return 3, and the cost only falls in between)
    assert set(np.unique(want["exit_code"])) >= {1, 2, 3, 5} or seed != 1, np.unique(want["exit_code"], return_counts=True)
    oracle.ref_destroy(r)


def test_epzs_subpel_only(ctx, oracle):
    """JMB_EPZS_SKIP_INT: the SubPelME call site alone (EPZS_sub_pel_motion_estimation)."""
    w, h = 96, 80
    f = synth.luma_frames(w, h, 2, seed=60, motion=(2, 1))
    ctx.configure(search_range=16)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(61)
    reqs, cands = _random_requests(rng, w, h, 300, motion_q=(-8, -4), range_q=64, ref_gt0_share=0)
    reqs["flags"] |= api.EPZS_SKIP_INT
    reqs["start_x"] = (reqs["start_x"] >> 2) << 2; reqs["start_y"] = (reqs["start_y"] >> 2) << 2
    reqs["min_mcost"] = np.where(rng.random(len(reqs)) < 0.5, BIG, rng.integers(1000, 60000, len(reqs)))
    same = rng.random(len(reqs)) < 0.4        # predictor == mv: the early return of me_epzs_sub.c:92
    reqs["pred_x"] = np.where(same, reqs["start_x"], reqs["pred_x"]); reqs["pred_y"] = np.where(same, reqs["start_y"], reqs["pred_y"])
    got = ctx.epzs_search(reqs, cands)
    want = oracle.epzs(r, f[1], reqs, cands, (api.SATD, api.SATD, 0, 1, 9))
    for fld in ("mv_x", "mv_y", "cost"):
        assert np.array_equal(got[fld], want[fld]), fld
    oracle.ref_destroy(r)


def test_epzs_frame_form(ctx, oracle):
    """Requests generated on the device from the per-macroblock tables = the explicit list built by the same rules; results
    also against the oracle on a sample."""
    w, h = 112, 80
    n_mb = (w // 16) * (h // 16)
    f = synth.luma_frames(w, h, 2, seed=62, motion=(4, 2))
    ctx.configure(search_range=32)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    rng = np.random.default_rng(63)
    pred = np.zeros(n_mb, api.MB_MVPRED)
    pred["pred"] = np.array([-16, -8]) + rng.integers(-9, 10, size=(n_mb, 1, 2)) + rng.integers(-3, 4, size=(n_mb, 41, 2))
    n_shared = 9
    shared = (np.array([-16, -8]) + rng.integers(-14, 15, size=(n_mb, n_shared, 2))).astype(np.int16)
    shared[:, 0] = 0
    fp = api.epzs_frame_params([33, 35, 37], flags=api.EPZS_ADAPT_PATTERN | api.EPZS_DUAL | api.EPZS_SUBPEL | api.EPZS_TEST8X8,
                               n_shared=n_shared, window=4, search_range=32)
    got = ctx.epzs_search_frame(pred, shared, fp)
    reqs = api.epzs_requests_from_frame(pred, fp, w // 16)
    exp = ctx.epzs_search(reqs, shared.reshape(-1, 2))
    assert np.array_equal(got["mv_x"], exp["mv_x"]) and np.array_equal(got["mv_y"], exp["mv_y"])
    assert np.array_equal(got["cost"], np.minimum(exp["cost"], 0x7FFFFFFF).astype(np.int32))
    r = oracle.ref_create(f[0])
    sel = rng.permutation(len(reqs))[:400]
    want = oracle.epzs(r, f[1], reqs[sel], shared.reshape(-1, 2), (api.SATD, api.SATD, 0, 1, 9))
    _assert_same(exp[sel], want, reqs[sel])
    oracle.ref_destroy(r)
    # the resident results feed the residual coder like those of the full search
    from jm_b200 import h264_tables as T
    qd = api.quant_desc(8, 28, T.q_params(28, 0, 8), T.SNGL_SCAN8x8, T.COEFF_COST8x8[0], 0)
    heads, tokens = ctx.mc_tq_modes_compact(None, qd, 0x0F, n_mb=n_mb)
    res24 = np.zeros(len(reqs), api.ME_RES)
    res24["mv_x"], res24["mv_y"] = got["mv_x"], got["mv_y"]
    heads2, tokens2 = ctx.mc_tq_modes_compact(res24, qd, 0x0F)
    assert np.array_equal(heads["cbp_blk"], heads2["cbp_blk"]) and len(tokens) == len(tokens2)


def test_epzs_rejects_bad_requests(ctx):
    w, h = 32, 32
    f = synth.luma_frames(w, h, 2, seed=64)
    ctx.configure(search_range=8)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    q = np.zeros(1, api.EPZS_REQ)
    q["blocktype"] = 1; q["range_x"] = q["range_y"] = 32; q["lambda"] = 10; q["prev_sad"] = BIG; q["min_mcost"] = BIG
    ctx.epzs_search(q, np.zeros((0, 2), np.int16))
    bad = q.copy(); bad["n_cand"][0, 0] = 3            # three candidates promised, none supplied
    with pytest.raises(api.JMBError, match="rejected on the device"):
        ctx.epzs_search(bad, np.zeros((0, 2), np.int16))
    bad = q.copy(); bad["pattern"] = 9
    with pytest.raises(api.JMBError, match="rejected on the device"):
        ctx.epzs_search(bad, np.zeros((0, 2), np.int16))
    ctx.configure(search_range=8, metric=(api.SAD, api.SATD, api.SAD))      # start_me_refinement_qp = 0
    with pytest.raises(api.JMBError, match="start_me_refinement_qp"):
        ctx.epzs_search(q, np.zeros((0, 2), np.int16))
    ctx.configure(search_range=8)


@pytest.mark.parametrize("n", [4, 8])
@pytest.mark.parametrize("metric", [api.SAD, api.SSE, api.SATD])
def test_mode_decision_distortion_backends(ctx, oracle, n, metric):
    """jmb_block_distortion = distortion4x4/8x8 SAD, SSE, SATD (me_distortion.c:38-146) on difference blocks; the 8x8 SAD also with
    JM's row-wise threshold (distortion8x8SADthres)."""
    rng = np.random.default_rng(70 + n + metric)
    diff = rng.integers(-255, 256, size=(500, n * n)).astype(np.int16)
    diff[::7] = rng.integers(-3, 4, size=(len(diff[::7]), n * n))
    got = ctx.block_distortion(metric, n, diff)
    if metric == api.SATD:
        want = [oracle.hadamard4x4(d) if n == 4 else oracle.hadamard8x8(d) for d in diff]
    elif metric == api.SSE:
        want = (diff.astype(np.int64) ** 2).sum(1)
    else:
        want = np.abs(diff.astype(np.int64)).sum(1)
    assert got.tolist() == [int(v) for v in want]
    if n == 8 and metric == api.SAD:
        thres = rng.integers(0, 6000, len(diff)).astype(np.int32)
        got = ctx.block_distortion(metric, n, diff, thres)
        rows = np.abs(diff.astype(np.int64)).reshape(-1, 8, 8).sum(2).cumsum(1)
        want = [int(r[np.argmax(r > t)]) if (r > t).any() else int(r[-1]) for r, t in zip(rows, thres)]
        assert got.tolist() == want

"""GPU parity of the round-2 picture-form entry points (through the C ABI, bit-exact):
device-generated requests from a predictor table (jmb_me_search_frame_pred), 8-bit uploads, the compact
(level, run) token output of the residual coder, the asynchronous host location, large search ranges
and the NVLink peer mapping of a reconstructed reference."""
import multiprocessing as mp
import os

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


def _pred_table(rng, n_mb, base=(0, 0), jitter=6):
    pred = np.zeros(n_mb, api.MB_MVPRED)
    mbp = rng.integers(-jitter, jitter + 1, size=(n_mb, 1, 2)) + np.array(base)
    pred["pred"] = mbp + rng.integers(-3, 4, size=(n_mb, 41, 2))
    return pred


@pytest.mark.parametrize("mode", [api.SEARCH_FULL, api.SEARCH_FAST_FULL])
@pytest.mark.parametrize("mv_range", [api.MV_RANGE_L51, (-24, 24, -16, 16)])
def test_frame_pred_equals_explicit_requests_and_oracle(ctx, oracle, mode, mv_range):
    """The device builds the 41 requests of every macroblock from the predictor table by JM's rules (centre rounding
    mv_search.c:931, clip_mv_range :957/:981, me_fullfast.c:309-327); results must equal the explicit-request form and
    the oracle's search + refinement."""
    w, h, R = 80, 64, 4 if mv_range[1] < 100 else 8
    if mode == api.SEARCH_FAST_FULL and mv_range[1] < 100:
        mv_range = (-64, 64, -48, 48)           # the fast-full centre clip needs min + 4R <= max - 4R
    f = synth.luma_frames(w, h, 2, seed=31, motion=(-2, 3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    n_mb = (w // 16) * (h // 16)
    rng = np.random.default_rng(32)
    pred = _pred_table(rng, n_mb, base=(-8, 12), jitter=14)
    fp = api.frame_params([37, 41, 43], mode=mode, flags=api.REQ_SUBPEL, mv_range=mv_range)
    got = ctx.me_search_frame_pred(pred, fp)
    reqs = api.requests_from_pred(pred, fp, w // 16, R)
    want = ctx.me_search(reqs, frame=True)
    mvx = np.clip(want["mv_x"], mv_range[0], mv_range[1]); mvy = np.clip(want["mv_y"], mv_range[2], mv_range[3])
    assert np.array_equal(got["mv_x"], mvx) and np.array_equal(got["mv_y"], mvy)
    assert np.array_equal(got["cost"], want["cost"].astype(np.int32))
    if mode == api.SEARCH_FULL:      # and against the oracle, request by request (sample)
        r = oracle.ref_create(f[0])
        for k in rng.permutation(len(reqs))[:120]:
            q = reqs[k]
            pos = (int(q["pos_x"]), int(q["pos_y"])); p = (int(q["pred_x"]), int(q["pred_y"])); c = (int(q["center_x"]), int(q["center_y"]))
            mv, cost = oracle.full_search(r, f[1], int(q["blocktype"]), pos, p, c, 37, BIG, R)
            mv2, c2 = oracle.sub_pel(r, f[1], int(q["blocktype"]), pos, p, mv, [37, 41, 43], BIG, po.SATD, po.SATD, 0, 1, 0)
            assert (int(got["mv_x"][k]), int(got["mv_y"][k])) == (min(max(mv2[0], mv_range[0]), mv_range[1]), min(max(mv2[1], mv_range[2]), mv_range[3]))
            assert int(got["cost"][k]) == c2
        oracle.ref_destroy(r)


def test_u8_uploads_match_u16(ctx):
    w, h = 96, 48
    f = synth.luma_frames(w, h, 2, seed=33)
    ctx.configure(search_range=6)
    ctx.ref_put(0, f[0]); ctx.ref_put_u8(1, f[0].astype(np.uint8))
    for fy in range(4):
        for fx in range(4):
            assert np.array_equal(ctx.ref_plane(0, fy, fx, (h, w)), ctx.ref_plane(1, fy, fx, (h, w)))
    pred = _pred_table(np.random.default_rng(34), (w // 16) * (h // 16))
    fp = api.frame_params([50, 50, 50])
    ctx.pic_begin(f[1], [0]); a = ctx.me_search_frame_pred(pred, fp)
    ctx.pic_begin_u8(f[1].astype(np.uint8), [1]); b = ctx.me_search_frame_pred(pred, fp)
    assert np.array_equal(a, b)


def test_host_async_location(ctx):
    """JMB_HOST_ASYNC: pinned buffers, nothing waited for until jmb_sync; same answers as the synchronous form."""
    w, h = 64, 48
    f = synth.luma_frames(w, h, 2, seed=35)
    n_mb = 12
    ctx.configure(search_range=8)
    ref8 = ctx.pinned((h, w), np.uint8); cur8 = ctx.pinned((h, w), np.uint8)
    ref8[:] = f[0]; cur8[:] = f[1]
    pred = ctx.pinned(n_mb, api.MB_MVPRED); pred[:] = _pred_table(np.random.default_rng(36), n_mb)
    fp = api.frame_params([44, 44, 44])
    res = ctx.pinned(n_mb * 41, api.ME_RES8)
    ctx.ref_put_u8(0, ref8, api.HOST_ASYNC); ctx.pic_begin_u8(cur8, [0], api.HOST_ASYNC)
    ctx.me_search_frame_pred(pred, fp, res, api.HOST_ASYNC)
    ctx.sync()
    got = np.array(res)
    ctx.ref_put(1, f[0]); ctx.pic_begin(f[1], [1])
    assert np.array_equal(got, ctx.me_search_frame_pred(np.array(pred), fp))


def _expand_tokens(heads, tokens, n, cavlc8):
    """(heads, tokens) -> dense [7][n_mb][256] levels in the layout of jmb_mc_tq_modes."""
    n_modes, n_mb = heads.shape
    lev = np.zeros((n_modes, n_mb, 256), np.int16)
    per = 16 if (n == 4 or cavlc8) else 64
    for m in range(n_modes):
        for mb in range(n_mb):
            hd = heads[m, mb]
            pos = {}
            for tk in tokens[int(hd["token_off"]): int(hd["token_off"]) + int(hd["n_tokens"])]:
                blk = int(tk["blk"])
                p = pos.get(blk, 0) + int(tk["run"])
                assert p < per
                lev[m, mb, blk * per + p] = tk["level"]
                pos[blk] = p + 1
    return lev


@pytest.mark.parametrize("n,cavlc,qp", [(4, 1, 24), (4, 0, 30), (8, 0, 22), (8, 1, 22)])
def test_compact_tokens_equal_dense_levels(ctx, n, cavlc, qp):
    w, h = 96, 64
    f = synth.luma_frames(w, h, 2, seed=37, motion=(1, -2))
    f[1] = np.clip(f[1].astype(int) + np.random.default_rng(3).integers(-12, 13, f[1].shape), 0, 255).astype(np.uint16)   # coefficients worth coding
    n_mb = (w // 16) * (h // 16)
    ctx.configure(search_range=6)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    pred = _pred_table(np.random.default_rng(38), n_mb)
    ctx.me_search_frame_pred(pred, api.frame_params([30, 30, 30]), want_res=False)
    scan = T.SNGL_SCAN if n == 4 else (T.SNGL_SCAN8x8_CAVLC if cavlc else T.SNGL_SCAN8x8)
    cc = T.COEFF_COST4x4[0] if n == 4 else T.COEFF_COST8x8[0]
    qd = api.quant_desc(n, qp, T.q_params(qp, 0, n), scan, cc, cavlc)
    mask = 0x7F if n == 4 else 0x0F
    lev, cost, cbp = ctx.mc_tq_modes(None, qd, mask, n_mb=n_mb)
    heads, tokens = ctx.mc_tq_modes_compact(None, qd, mask, n_mb=n_mb)
    assert len(tokens) == int((lev != 0).sum()) and len(tokens) > 50
    got = _expand_tokens(heads, tokens, n, bool(cavlc and n == 8))
    assert np.array_equal(got, lev)
    assert np.array_equal(heads["cbp_blk"], cbp)
    assert np.array_equal(heads["cost8"], np.minimum(cost, 255).astype(np.uint8))
    # an undersized token buffer is an error, never a silent truncation
    with pytest.raises(api.JMBError, match="tokens produced"):
        ctx.mc_tq_modes_compact(None, qd, mask, n_mb=n_mb, token_cap=8)


@pytest.mark.parametrize("R", [40, 44, 48, 64])
def test_large_search_ranges(ctx, oracle, R):
    """SearchRange up to JM's 64: more than one staging chunk per axis and spiral indices beyond 8192."""
    w, h = 64, 48
    f = synth.luma_frames(w, h, 2, seed=39, motion=(4, -3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    r = oracle.ref_create(f[0])
    rng = np.random.default_rng(R)
    reqs = np.zeros(8, api.ME_REQ)
    for q in reqs:
        bt = int(rng.integers(1, 8)); bsx, bsy = api.BLOCK_SIZE[bt]
        q["blocktype"] = bt
        q["pos_x"] = int(rng.integers(0, (w - bsx) // bsx + 1)) * bsx; q["pos_y"] = int(rng.integers(0, (h - bsy) // bsy + 1)) * bsy
        q["pred_x"], q["pred_y"] = rng.integers(-60, 61, 2)
        q["center_x"] = ((int(q["pred_x"]) + 2) >> 2) * 4; q["center_y"] = ((int(q["pred_y"]) + 2) >> 2) * 4
        q["lambda"] = 4; q["min_mcost"] = BIG       # a small lambda lets far positions win: the outer spiral rings matter
    res = ctx.me_search(reqs)
    for q, o in zip(reqs, res):
        mv, cost = oracle.full_search(r, f[1], int(q["blocktype"]), (int(q["pos_x"]), int(q["pos_y"])), (int(q["pred_x"]), int(q["pred_y"])),
                                      (int(q["center_x"]), int(q["center_y"])), 4, BIG, R)
        assert (int(o["imv_x"]), int(o["imv_y"])) == mv and int(o["icost"]) == cost
    oracle.ref_destroy(r)
    ctx.configure(search_range=8)


def test_full_pel_metric_other_than_sad_is_refused(ctx):
    """setup_fast_full_search builds squared-error surfaces for MEDistortionFPel != SAD (me_fullfast.c:274) and the full
    search takes computePredFPel of that metric: the SAD search kernel must refuse both, never answer with SADs."""
    w, h = 32, 32
    f = synth.luma_frames(w, h, 2, seed=40)
    ctx.configure(search_range=4, metric=(api.SSE, api.SATD, api.SATD))
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    for mode in (api.SEARCH_FULL, api.SEARCH_FAST_FULL):
        q = np.zeros(1, api.ME_REQ)
        q["blocktype"] = 1; q["mode"] = mode; q["lambda"] = 10; q["min_mcost"] = BIG
        with pytest.raises(api.JMBError, match="MEDistortionFPel"):
            ctx.me_search(q)
    ctx.configure(search_range=8)


def test_configure_rejects_out_of_range_refinement_starts(ctx):
    with pytest.raises(api.JMBError):
        ctx.configure(start_hp=3)
    with pytest.raises(api.JMBError):
        ctx.configure(start_qp=-1)
    with pytest.raises(api.JMBError):
        ctx.configure(search_range=65)
    with pytest.raises(api.JMBError):
        ctx.configure(max_mvd=1)
    ctx.configure(search_range=8)


def test_device_resident_prediction_table_is_validated(ctx):
    """jmb_mc_tq / jmb_luma_residual_coding with JMB_DEVICE input: mode and reference are checked on the device."""
    import torch
    w, h = 32, 32
    f = synth.luma_frames(w, h, 2, seed=41)
    ctx.configure(search_range=4); ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    pred = np.zeros(4, api.MB_PRED); pred["b8mode"] = 1
    pred["b8mode"][2, 1] = 9            # no such partition mode
    pred["ref"][3, 0] = 200             # no such reference
    d_pred = torch.from_numpy(pred.view(np.uint8).reshape(-1).copy()).cuda()
    lev = torch.zeros(4 * 256, dtype=torch.int16, device="cuda"); cost = torch.zeros(16, dtype=torch.int32, device="cuda"); cbp = torch.zeros(4, dtype=torch.int32, device="cuda")
    qd = api.quant_desc(4, 28, T.q_params(28, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    ctx.mc_tq(d_pred.data_ptr(), qd, api.DEVICE, n_mb=4, out=(lev.data_ptr(), cost.data_ptr(), cbp.data_ptr()))
    with pytest.raises(api.JMBError, match="rejected on the device"):
        ctx.sync()


def _peer_owner(conn, w, h, seed):
    """Owner process: holds the 'reconstructed' luma in ITS device memory and exports the IPC handle."""
    c = api.Context(0)
    f = synth.luma_frames(w, h, 1, seed=seed)[0].astype(np.uint8)
    p = c.dev_alloc(w * h)
    c.dev_copy(p, np.ascontiguousarray(f), w * h, api.DEVICE, api.HOST)
    conn.send(c.peer_export(p).tobytes())
    conn.recv()                         # keep the allocation alive until the peer is done
    c.dev_free(p); c.close()


def test_reference_read_from_a_peer_mapping():
    """jmb_peer_export / jmb_peer_open: the quarter-pel planes built from another process's device buffer (mapped over IPC;
    between GPUs of a box the loads travel over NVLink) equal those built from a host copy of the same picture."""
    w, h, seed = 96, 64, 42
    mpc = mp.get_context("spawn")
    a, b = mpc.Pipe()
    proc = mpc.Process(target=_peer_owner, args=(b, w, h, seed))
    proc.start()
    try:
        handle = np.frombuffer(a.recv(), np.uint8)
        c = api.Context(0)
        mapped = c.peer_open(handle)
        assert c.peer_open(handle) == mapped             # cached
        c.ref_put_u8(0, mapped, api.DEVICE, shape=(h, w)); c.sync()
        f = synth.luma_frames(w, h, 1, seed=seed)[0]
        c.ref_put(1, f)
        for fy, fx in [(0, 0), (2, 2), (1, 3), (3, 1)]:
            assert np.array_equal(c.ref_plane(0, fy, fx, (h, w)), c.ref_plane(1, fy, fx, (h, w)))
        c.peer_close(mapped)
        c.close()
    finally:
        a.send(b"done")
        proc.join(60)


@pytest.mark.parametrize("mode", [api.SEARCH_FULL, api.SEARCH_FAST_FULL])
def test_mb_surfaces_and_per_partition_argmin(ctx, mode):
    """jmb_mb_surfaces + jmb_mb_search (the drop-in form: SAD surfaces once per macroblock, one small arg-min launch per
    partition, answer through the mapped mailbox) must give exactly what the one-launch-per-request search gives -- which
    tests/test_gpu_parity.py checks against the oracle -- incl. border macroblocks (UMVLine4X clamps), incoming bounds and the
    sub-pel refinement; a window the resident surfaces do not cover is refused."""
    w, h, R, E = 96, 80, 12, 6
    f = synth.luma_frames(w, h, 2, seed=43, motion=(3, -2))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    rng = np.random.default_rng(44)
    parts = api.mb_partitions()
    for mb in [(0, 0), (80, 64), (32, 32), (80, 0), (0, 64), (48, 16)]:
        base = rng.integers(-40, 41, 2)
        c0 = (((int(base[0]) + 2) >> 2) * 4, ((int(base[1]) + 2) >> 2) * 4)
        ctx.mb_surfaces(0, mb, c0, R + E)
        for k in rng.permutation(41)[:14]:
            t, x, y = parts[k]
            q = np.zeros(1, api.ME_REQ)
            q["blocktype"] = t; q["pos_x"] = mb[0] + x; q["pos_y"] = mb[1] + y
            p = base + rng.integers(-3 * E // 2, 3 * E // 2 + 1, 2)
            q["pred_x"], q["pred_y"] = p
            if mode == api.SEARCH_FULL:
                q["center_x"] = ((int(p[0]) + 2) >> 2) * 4; q["center_y"] = ((int(p[1]) + 2) >> 2) * 4
            else:
                q["center_x"], q["center_y"] = c0
            q["mode"] = mode; q["flags"] = api.REQ_SUBPEL if k % 2 else 0
            q["lambda"] = int(rng.integers(1, 200))
            q["min_mcost"] = BIG if k % 3 else int(rng.integers(200, 30000))
            want = ctx.me_search(q)[0]
            got = ctx.mb_search(q)
            assert got == want, (mb, q, got, want)
        # sub-pel only (the SubPelME call site)
        q["flags"] = api.REQ_SUBPEL | api.REQ_SKIP_INT
        a, b = ctx.mb_search(q), ctx.me_search(q)[0]
        assert (a["mv_x"], a["mv_y"], a["cost"]) == (b["mv_x"], b["mv_y"], b["cost"])
    q["flags"] = 0; q["center_x"] = int(q["center_x"]) + 4 * (E + 8)
    with pytest.raises(api.JMBError, match="not covered"):
        ctx.mb_search(q)
    ctx.pic_begin(f[1], [0])            # a new picture invalidates the resident surfaces
    q["center_x"] = c0[0]
    with pytest.raises(api.JMBError, match="no surfaces resident"):
        ctx.mb_search(q)


def _chain_layouts():
    """The two calls the shim makes per macroblock region: modes 1-3 of the macroblock; the four sub-modes of one 8x8 quadrant.
    Entries: (blocktype, x, y, chain, deps) with deps = {neighbour index: index of the earlier search that covers it}."""
    whole = [(1, 0, 0, 0, {}), (2, 0, 0, 1, {}), (2, 0, 8, 1, {1: 1}), (3, 0, 0, 2, {}), (3, 8, 0, 2, {0: 3})]
    def quad(qx, qy):
        return [(4, qx, qy, 0, {}),
                (5, qx, qy, 1, {}), (5, qx, qy + 4, 1, {1: 1}),
                (6, qx, qy, 2, {}), (6, qx + 4, qy, 2, {0: 3}),
                (7, qx, qy, 3, {}), (7, qx + 4, qy, 3, {0: 5}), (7, qx, qy + 4, 3, {1: 5, 2: 6}), (7, qx + 4, qy + 4, 3, {0: 7, 1: 6, 2: 5})]
    return [whole, quad(0, 0), quad(8, 0), quad(0, 8), quad(8, 8)]


@pytest.mark.parametrize("mode", [api.SEARCH_FULL, api.SEARCH_FAST_FULL])
def test_mb_chain_runs_ahead_like_jm_block_after_block(ctx, mode):
    """jmb_mb_chain: the searches of a macroblock whose predictors depend on each other (the lower 16x8 block on the upper one,
    the 4x4 blocks of a quadrant on each other) in one call.  Expected = JM's sequence done one leaf call at a time: predictor
    from the neighbours (oracle.pyoracle.mv_predictor, pinned to JM's GetMVPredictor) -> centre -> jmb_me_search (oracle-checked
    in tests/test_gpu_parity.py) -> clip -> visible to the next block."""
    from oracle import pyoracle as po
    w, h, R, E = 96, 80, 12, 6
    f = synth.luma_frames(w, h, 2, seed=47, motion=(-2, 3))
    ctx.configure(search_range=R)
    ctx.ref_put(0, f[0]); ctx.pic_begin(f[1], [0])
    rng = np.random.default_rng(48)
    lim = (-60, 68, -52, 44)      # (multiples of 4: an integer-pel centre stays one after the clip)
    n_done = n_unc = 0
    for mb in [(0, 0), (80, 64), (32, 32), (48, 16), (16, 48)]:
        base = rng.integers(-24, 25, 2)
        c0 = (((int(base[0]) + 2) >> 2) * 4, ((int(base[1]) + 2) >> 2) * 4)
        c0 = (min(max(c0[0], lim[0]), lim[1]) // 4 * 4, min(max(c0[1], lim[2]), lim[3]) // 4 * 4)
        ctx.mb_surfaces(0, mb, c0, R + E)
        for layout in _chain_layouts():
            reqs = np.zeros(len(layout), api.CHAIN_REQ)
            lam = int(rng.integers(1, 200))
            for i, (t, x, y, chain, deps) in enumerate(layout):
                q = reqs[i]
                q["req"]["blocktype"] = t; q["req"]["pos_x"] = mb[0] + x; q["req"]["pos_y"] = mb[1] + y
                q["req"]["mode"] = mode; q["req"]["flags"] = api.REQ_SUBPEL; q["req"]["lambda"] = lam; q["req"]["min_mcost"] = BIG
                if mode == api.SEARCH_FAST_FULL:
                    q["req"]["center_x"], q["req"]["center_y"] = c0
                q["jm_ref"] = 0; q["chain"] = chain
                for k in range(3):
                    nb = q["nb"][k]
                    if k in deps:
                        nb["available"] = 1; nb["ref_idx"] = 0; nb["dep"] = deps[k]
                    else:
                        nb["dep"] = -1
                        nb["available"] = rng.random() < 0.8
                        nb["ref_idx"] = int(rng.choice([-1, 0, 0, 0, 1]))
                        far = rng.random() < 0.1       # now and then a neighbour that pulls the window off the resident surfaces
                        nb["mv_x"], nb["mv_y"] = base + rng.integers(-6, 7, 2) + (60 if far else 0)
            got = ctx.mb_chain(reqs, lim)
            fin = {}
            for i, (t, x, y, chain, deps) in enumerate(layout):
                q = reqs[i]; bsx, bsy = api.BLOCK_SIZE[t]
                nbs, skipped = [], False
                for k in range(3):
                    nb = q["nb"][k]
                    if nb["dep"] >= 0:
                        if int(nb["dep"]) not in fin:
                            skipped = True
                        nbs.append((1, 0) + fin.get(int(nb["dep"]), (0, 0)))
                    else:
                        nbs.append((int(nb["available"]), int(nb["ref_idx"]), int(nb["mv_x"]), int(nb["mv_y"])))
                if skipped:
                    assert got[i]["status"] == api.CHAIN_SKIPPED, (mb, i, got[i])
                    continue
                px, py = po.mv_predictor(nbs, 0, x, y, bsx, bsy)
                assert (got[i]["pred_x"], got[i]["pred_y"]) == (px, py), (mb, i, nbs, got[i])
                lq = np.zeros(1, api.ME_REQ); lq[0] = q["req"]
                lq["pred_x"] = px; lq["pred_y"] = py
                if mode == api.SEARCH_FULL:
                    lq["center_x"] = min(max(((px + 2) >> 2) * 4, lim[0]), lim[1]); lq["center_y"] = min(max(((py + 2) >> 2) * 4, lim[2]), lim[3])
                assert (got[i]["center_x"], got[i]["center_y"]) == (lq["center_x"][0], lq["center_y"][0])
                if got[i]["status"] == api.CHAIN_UNCOVERED:
                    with pytest.raises(api.JMBError, match="not covered"):
                        ctx.mb_search(lq)
                    n_unc += 1
                    continue
                assert got[i]["status"] == api.CHAIN_DONE
                want = ctx.me_search(lq)[0]
                assert got[i]["res"] == want, (mb, i, lq, got[i], want)
                fin[i] = (min(max(int(want["mv_x"]), lim[0]), lim[1]), min(max(int(want["mv_y"]), lim[2]), lim[3]))
                n_done += 1
    assert n_done > 150 and (n_unc > 0 or mode == api.SEARCH_FAST_FULL), (n_done, n_unc)


def test_1080p_picture_step_matches_jm_on_a_sample(ctx):
    """BASELINE config 2 at full size through the picture form (predictor table in, 8-byte results and (level, run) tokens out):
    motion vectors, costs AND quantised levels of a sample of macroblocks against JM's OWN functions (oracle/_ref/libjmref.so:
    full_search_motion_estimation, sub_pel_motion_estimation, forward4x4, quant_4x4_normal) -- the re-check bench.py's CPU leg does,
    as a test."""
    from oracle import pyoracle as po
    from jm_b200 import h264_tables as T
    import bench
    if not po.ref_available():
        pytest.skip("oracle/_ref/libjmref.so not built")
    w, h, R, qp, lam = 1920, 1088, 32, 28, 187
    n_mb = (w // 16) * (h // 16)
    f = synth.luma_frames(w, h, 2, seed=4321, motion=(5, 3))
    pred = bench.make_pred_table(api, 91, n_mb)
    ctx.configure(search_range=R)
    ctx.ref_put_u8(0, f[0].astype(np.uint8)); ctx.pic_begin_u8(f[1].astype(np.uint8), [0])
    fp = api.frame_params([lam] * 3, mode=api.SEARCH_FULL, flags=api.REQ_SUBPEL)
    g = ctx.me_search_frame_pred(pred, fp).reshape(n_mb, 41)
    qd = api.quant_desc(4, qp, T.q_params(qp, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
    heads, tokens = ctx.mc_tq_modes_compact(None, qd, 0x7F, n_mb=n_mb, token_cap=7 * n_mb * 96)
    idx = np.unique(np.concatenate([np.linspace(0, n_mb - 1, 40).astype(int), [0, w // 16 - 1, n_mb - w // 16, n_mb - 1]]))      # incl. the four corners
    ref = po.JMRef(w, h, R)
    ref.set_ref(f[0].astype(np.uint16)); ref.set_cur(f[1].astype(np.uint16))
    mbw = w // 16
    mb_xy = np.stack([(idx % mbw) * 16, (idx // mbw) * 16], 1)
    mv, cost, lev, _ = po.jmref_run_mbs(ref, mb_xy, pred["pred"][idx], [lam] * 3, qp, T.q_params(qp, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0])
    assert np.array_equal(g["mv_x"][idx], mv[:, :, 0]) and np.array_equal(g["mv_y"][idx], mv[:, :, 1])
    assert np.array_equal(g["cost"][idx], cost)
    assert np.array_equal(bench.expand_tokens(heads, tokens, idx, per=16), lev)
    assert (lev != 0).sum() > 50

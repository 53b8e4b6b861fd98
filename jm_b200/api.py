"""ctypes mirror of the libjmb200 C ABI (include/jmb200.h) -- harness side only.

No CPU fallback lives here: if the shared library is missing, or no B200 is present, construction
raises.  Arrays cross the boundary as numpy buffers (JMB_HOST) or raw device pointers (JMB_DEVICE,
e.g. ``tensor.data_ptr()`` of a torch CUDA tensor).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JMB200_LIB") or os.path.join(HERE, "lib", "libjmb200.so")     # JMB200_LIB: tuning builds (tools/)
HOST, DEVICE, HOST_ASYNC = 0, 1, 2
SAD, SSE, SATD = 0, 1, 2
SEARCH_FULL, SEARCH_FAST_FULL = 0, 1
REQ_SUBPEL, REQ_TEST8X8, REQ_SKIP_INT = 1, 2, 4
DISTBLK_MAX = 0x7FFFFFFF << 5
NPART = 41
BLOCK_SIZE = [(16, 16), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]

ME_REQ = np.dtype([("pos_x", "<i2"), ("pos_y", "<i2"), ("pred_x", "<i2"), ("pred_y", "<i2"),
                   ("center_x", "<i2"), ("center_y", "<i2"), ("blocktype", "u1"), ("ref", "u1"),
                   ("mode", "u1"), ("flags", "u1"), ("lambda", "<i4", (3,)), ("pad_", "<i4"), ("min_mcost", "<i8")])
ME_RES = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("imv_x", "<i2"), ("imv_y", "<i2"), ("cost", "<i8"), ("icost", "<i8")])
CHAIN_NB = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("ref_idx", "i1"), ("available", "i1"), ("dep", "i1"), ("pad_", "i1")])
CHAIN_REQ = np.dtype([("req", ME_REQ), ("nb", CHAIN_NB, (3,)), ("jm_ref", "i1"), ("chain", "i1"), ("pad_", "i1", (6,))])
CHAIN_RES = np.dtype([("res", ME_RES), ("pred_x", "<i2"), ("pred_y", "<i2"), ("center_x", "<i2"), ("center_y", "<i2"), ("status", "<i4"), ("pad_", "<i4")])
CHAIN_DONE, CHAIN_UNCOVERED, CHAIN_SKIPPED = range(3)
DB_MB = np.dtype([("mb_type", "u1"), ("flags", "u1"), ("qp", "i1"), ("qpc", "i1", (2,)), ("df_disable_idc", "i1"), ("df_alpha_c0_offset", "i1"),
                  ("df_beta_offset", "i1"), ("cbp_blk", "<u4"), ("pad_", "<u4"), ("mv", "<i2", (2, 16, 2)), ("ref_id", "i1", (2, 16))])      # jmb_db_mb, 176 bytes
DB_T8X8, DB_CBP, DB_AVAIL_A, DB_AVAIL_B = 1, 2, 4, 8
MB_PRED = np.dtype([("mv", "<i2", (16, 2)), ("b8mode", "u1", (4,)), ("ref", "u1", (4,))])
QUANT_DESC = np.dtype([("n", "<i4"), ("qp", "<i4"), ("is_cavlc", "<i4"), ("around", "<i4"), ("adapt_rnd_weight", "<i4"),
                       ("qparams", "<i4", (64, 3)), ("scan", "u1", (64, 2)), ("c_cost", "u1", (64,))])
QLIST_DESC = np.dtype([("m", "<i4"), ("q_bits", "<i4"), ("qp_per", "<i4"), ("dequant", "<i4"), ("clip", "<i4"), ("use_cost", "<i4"),
                       ("around", "<i4"), ("adapt_rnd_weight", "<i4"), ("params", "<i4", (16, 3)), ("c_cost", "u1", (16,))])
DQ_LEVEL, DQ_SHIFT, DQ_SHIFT_RND4 = 0, 1, 2
PRED_PLAIN, PRED_WEIGHTED, PRED_AVERAGE, PRED_WEIGHTED_AVERAGE = range(4)
DIST_PRED = np.dtype([(k, np.int32) for k in ("form", "ref2", "cand2_x", "cand2_y", "weight1", "weight2", "offset", "log_weight_denom", "wp_round")])
HAD_4X4, IHAD_4X4, HAD_4X2, IHAD_4X2, HAD_2X2, IHAD_2X2 = range(6)
MB_MVPRED = np.dtype([("pred", "<i2", (41, 2))])
FRAME_PARAMS = np.dtype([("lambda", "<i4", (3,)), ("mode", "<i4"), ("flags", "<i4"), ("ref", "<i4"),
                         ("mv_min_x", "<i4"), ("mv_max_x", "<i4"), ("mv_min_y", "<i4"), ("mv_max_y", "<i4")])
ME_RES8 = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("cost", "<i4")])
TQ_HEAD = np.dtype([("cbp_blk", "<u4"), ("token_off", "<u4"), ("n_tokens", "<u2"), ("cost8", "u1", (4,)), ("reserved_", "<u2")])
TQ_TOKEN = np.dtype([("level", "<i2"), ("run", "u1"), ("blk", "u1")])
assert MB_MVPRED.itemsize == 164 and FRAME_PARAMS.itemsize == 40 and ME_RES8.itemsize == 8 and TQ_HEAD.itemsize == 16 and TQ_TOKEN.itemsize == 4
EPZS_REQ = np.dtype([("pos_x", "<i2"), ("pos_y", "<i2"), ("pred_x", "<i2"), ("pred_y", "<i2"), ("start_x", "<i2"), ("start_y", "<i2"),
                     ("blocktype", "u1"), ("ref", "u1"), ("flags", "u1"), ("pattern", "u1"), ("pattern_dual", "u1"), ("jm_ref", "u1"), ("reserved_", "u1", (2,)),
                     ("n_cand", "u1", (4,)), ("gate", "u1", (4,)), ("cand_off", "<i4"), ("lambda", "<i4", (3,)),
                     ("range_x", "<i2"), ("range_y", "<i2"), ("stop", "<i8"), ("medthres", "<i8"), ("prev_sad", "<i8"),
                     ("subthres", "<i8"), ("min_mcost", "<i8")])
EPZS_RES = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("imv_x", "<i2"), ("imv_y", "<i2"), ("cost", "<i8"), ("icost", "<i8"),
                     ("prev_sad", "<i8"), ("exit_code", "<i4"), ("n_evals", "<i4")])
EPZS_FRAME_PARAMS = np.dtype([("lambda", "<i4", (3,)), ("flags", "<i4"), ("ref", "<i4"), ("pattern", "<i4"), ("pattern_dual", "<i4"),
                              ("n_shared", "<i4"), ("window", "<i4"), ("range", "<i4"), ("medthres", "<i4", (8,)), ("minthres", "<i4", (8,)),
                              ("maxthres", "<i4", (8,)), ("subthres", "<i4", (8,)), ("mv_min_x", "<i4"), ("mv_max_x", "<i4"),
                              ("mv_min_y", "<i4"), ("mv_max_y", "<i4")])
assert EPZS_REQ.itemsize == 88 and EPZS_RES.itemsize == 40 and EPZS_FRAME_PARAMS.itemsize == 184
EPZS_PAT_SDIAMOND, EPZS_PAT_SQUARE, EPZS_PAT_EDIAMOND, EPZS_PAT_LDIAMOND, EPZS_PAT_SBDIAMOND, EPZS_PAT_PMVFAST = range(6)
EPZS_REF_GT0_FRAME, EPZS_ADAPT_PATTERN, EPZS_SQUARE_HINT, EPZS_DUAL, EPZS_SUBPEL, EPZS_TEST8X8, EPZS_SKIP_INT, EPZS_WINDOW_GEN = 1, 2, 4, 8, 16, 32, 64, 128
# MED / MIN / MAX_THRES_BASE of lencod/src/me_epzs_common.c:34-37 scaled as EPZSStructInit does (:454-457): costs carry 5 fractional bits
EPZS_MIN_BASE = [0, 64, 32, 32, 16, 8, 8, 4]
EPZS_MED_BASE = [0, 192, 96, 96, 48, 24, 24, 12]
EPZS_MAX_BASE = [0, 768, 384, 384, 192, 96, 96, 48]
CHROMA_DESC = np.dtype([("yuv_format", "<i4"), ("is_cavlc", "<i4"), ("qp_ac", "<i4", (2,)), ("qp_dc", "<i4", (2,)),
                        ("params_ac", "<i4", (2, 16, 3)), ("params_dc", "<i4", (2, 3)), ("c_cost", "u1", (16,))])
assert CHROMA_DESC.itemsize == 448
IPC_HANDLE_BYTES = 64
# level 4 .. 5.1 mv range in quarter-pel (LEVELHMVLIMIT / LEVELVMVLIMIT, lencod/src/conformance.c): +-2048 x +-512 pels
MV_RANGE_L51 = (-8192, 8191, -2048, 2047)
assert QLIST_DESC.itemsize == 240
assert ME_REQ.itemsize == 40 and ME_RES.itemsize == 24 and MB_PRED.itemsize == 72 and QUANT_DESC.itemsize == 980


class MEConfig(C.Structure):
    _fields_ = [("search_range", C.c_int32), ("max_mvd", C.c_int32), ("metric", C.c_int32 * 3),
                ("start_hp", C.c_int32), ("start_qp", C.c_int32), ("search_pos2", C.c_int32), ("search_pos4", C.c_int32)]


class JMBError(RuntimeError):
    pass


def load_library():
    if not os.path.exists(LIB_PATH):
        raise JMBError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(make -C jm_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i = C.c_void_p, C.c_int
    L.jmb_abi_version.restype = i
    L.jmb_create.argtypes = [i, C.POINTER(vp)]
    L.jmb_destroy.argtypes = [vp]
    L.jmb_last_error.restype = C.c_char_p; L.jmb_last_error.argtypes = [vp]
    L.jmb_sync.argtypes = [vp]
    L.jmb_stream.restype = vp; L.jmb_stream.argtypes = [vp]
    L.jmb_launch_count.restype = C.c_uint64; L.jmb_launch_count.argtypes = [vp]
    L.jmb_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.jmb_host_free.argtypes = [vp, vp]
    L.jmb_ref_put.argtypes = [vp, i, vp, i, i, i, i, i]
    L.jmb_ref_drop.argtypes = [vp, i]
    L.jmb_ref_get_plane.argtypes = [vp, i, i, i, vp, i]
    L.jmb_pic_begin.argtypes = [vp, vp, i, i, i, i, C.POINTER(i), i]
    L.jmb_me_configure.argtypes = [vp, C.POINTER(MEConfig)]
    L.jmb_me_search.argtypes = [vp, vp, i, vp, i]
    L.jmb_me_search_frame.argtypes = [vp, vp, i, vp, i]
    L.jmb_ffs_surfaces.argtypes = [vp, i, i, i, i, i, vp, i]
    L.jmb_dist.argtypes = [vp, i, i, i, i, i, vp, i, i, vp, i]
    L.jmb_dist_ex.argtypes = [vp, i, vp, i, i, i, i, vp, i, i, vp, i]
    L.jmb_forward_transform.argtypes = [vp, vp, i, i, i]
    L.jmb_quant_blocks.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp, vp, i]
    L.jmb_mc_tq.argtypes = [vp, vp, i, vp, vp, vp, vp, i]
    L.jmb_pred_from_results.argtypes = [vp, vp, i, i, vp, i]
    L.jmb_mc_tq_modes.argtypes = [vp, vp, i, C.c_uint, vp, vp, vp, vp, i]
    L.jmb_inverse_transform.argtypes = [vp, vp, i, i, i]
    L.jmb_hadamard.argtypes = [vp, i, vp, i, i]
    L.jmb_quant_list.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, i]
    L.jmb_luma_residual_coding.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, i]
    L.jmb_luma_residual_coding_modes.argtypes = [vp, vp, i, C.c_uint, vp, vp, vp, vp, vp, vp, vp, i]
    L.jmb_ref_put_u8.argtypes = [vp, i, vp, i, i, i, i]
    L.jmb_pic_begin_u8.argtypes = [vp, vp, i, i, i, i, C.POINTER(i), i]
    L.jmb_me_search_frame_pred.argtypes = [vp, vp, i, vp, vp, i]
    L.jmb_mc_tq_modes_compact.argtypes = [vp, vp, i, C.c_uint, vp, vp, vp, C.c_uint32, vp, i]
    L.jmb_block_distortion.argtypes = [vp, i, i, vp, i, vp, vp, i]
    L.jmb_ref_put_chroma.argtypes = [vp, i, vp, vp, i, i, i, i, i]
    L.jmb_pic_chroma.argtypes = [vp, vp, vp, i, i, i, i, i]
    L.jmb_chroma_residual_coding.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, i]
    L.jmb_mb_surfaces.argtypes = [vp, i, i, i, i, i, i]
    L.jmb_mb_search.argtypes = [vp, vp, vp]
    L.jmb_mb_chain.argtypes = [vp, vp, i, vp, i, vp]
    L.jmb_deblock_picture.argtypes = [vp, vp, i, vp, vp, i, i, i, i, i, i, vp, i]
    L.jmb_epzs_search.argtypes = [vp, vp, i, vp, i, vp, i]
    L.jmb_epzs_search_frame.argtypes = [vp, vp, vp, i, vp, vp, i]
    L.jmb_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.jmb_dev_free.argtypes = [vp, vp]
    L.jmb_dev_copy.argtypes = [vp, vp, vp, C.c_size_t, i, i]
    L.jmb_peer_export.argtypes = [vp, vp, vp]
    L.jmb_peer_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.jmb_peer_close.argtypes = [vp, vp]
    L.jmb_timing_enable.argtypes = [vp, i]
    L.jmb_timing_get.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(i)]
    return L


def _ptr(a):
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)      # raw device / pinned pointer


def part_slot(blocktype, bx4, by4):
    base = [0, 0, 1, 3, 5, 9, 17, 25][blocktype]
    w4 = [4, 4, 4, 2, 2, 2, 1, 1][blocktype]; h4 = [4, 4, 2, 4, 2, 1, 2, 1][blocktype]
    return base + (by4 // h4) * (4 // w4) + bx4 // w4


def mb_partitions():
    """(blocktype, x offset, y offset) of the 41 partitions of a macroblock in canonical order."""
    out = []
    for t in range(1, 8):
        bsx, bsy = BLOCK_SIZE[t]
        for y in range(0, 16, bsy):
            for x in range(0, 16, bsx):
                out.append((t, x, y))
    assert all(part_slot(t, x // 4, y // 4) == k for k, (t, x, y) in enumerate(out))
    return out


def qlist_plan(variant, qp, qparams, scan, c_cost, is_cavlc, arw=0):
    """How one of JM's DC / AC quantisers maps onto the list quantiser (jmb_quant_list) -- the same mapping the shim does in C.
    variant: 6 quant_ac4x4_normal, 7 quant_ac4x4_around, 8 quant_dc4x4_normal, 9/10 quant_dc2x2_normal/_around,
             11/12 quant_dc4x2_normal/_around.  qparams: [4][4][3] for 6,7; one triple otherwise.
    Returns dict(order = flat index into the function's coefficient array per list position, + the descriptor fields)."""
    per = qp // 6
    scan = np.asarray(scan).reshape(-1, 2)
    qparams = np.asarray(qparams, np.int64)
    if variant in (6, 7):
        order = [int(scan[k][1]) * 4 + int(scan[k][0]) for k in range(1, 16)]
        params = [qparams.reshape(16, 3)[o] for o in order]
        return dict(order=order, params=np.array(params), q_bits=15 + per, qp_per=per, dequant=DQ_SHIFT_RND4, clip=int(is_cavlc), use_cost=1,
                    around=int(variant == 7), arw=arw, c_cost=np.asarray(c_cost, np.uint8)[:16])
    one = qparams.reshape(-1)[:3]
    dc = np.array([[2 * one[0], one[1], one[2]]])
    if variant == 8:
        order = [int(scan[k][1]) * 4 + int(scan[k][0]) for k in range(16)]
        dq = DQ_LEVEL
    elif variant in (9, 10):
        order, dq = [0, 1, 2, 3], DQ_SHIFT
    else:
        order, dq = [int(scan[k][0]) * 4 + int(scan[k][1]) for k in range(8)], DQ_SHIFT       # j first: block.c:88-94
    return dict(order=order, params=np.repeat(dc, len(order), 0), q_bits=16 + per, qp_per=per, dequant=dq, clip=int(is_cavlc), use_cost=0,
                around=0, arw=arw, c_cost=np.zeros(16, np.uint8))


def frame_params(lam, mode=SEARCH_FULL, flags=REQ_SUBPEL, ref=0, mv_range=MV_RANGE_L51):
    fp = np.zeros(1, FRAME_PARAMS)
    fp["lambda"] = lam; fp["mode"] = mode; fp["flags"] = flags; fp["ref"] = ref
    fp["mv_min_x"], fp["mv_max_x"], fp["mv_min_y"], fp["mv_max_y"] = mv_range
    return fp


def requests_from_pred(pred, fp, mb_w, search_range):
    """The jmb_me_req list jmb_me_search_frame_pred generates on the device (same rules; used by tests and the CPU legs)."""
    n_mb = len(pred)
    parts = mb_partitions()
    fp = fp[0]
    reqs = np.zeros((n_mb, NPART), ME_REQ)
    mbx = (np.arange(n_mb) % mb_w) * 16; mby = (np.arange(n_mb) // mb_w) * 16
    R4 = 4 * search_range
    for k, (t, x, y) in enumerate(parts):
        reqs["blocktype"][:, k] = t
        reqs["pos_x"][:, k] = mbx + x; reqs["pos_y"][:, k] = mby + y
        p = pred["pred"][:, k].astype(np.int32)
        reqs["pred_x"][:, k] = p[:, 0]; reqs["pred_y"][:, k] = p[:, 1]
        if int(fp["mode"]) == SEARCH_FAST_FULL:
            b = pred["pred"][:, 0].astype(np.int32)
            reqs["center_x"][:, k] = np.clip(((b[:, 0] + 2) >> 2) * 4, fp["mv_min_x"] + R4, fp["mv_max_x"] - R4)
            reqs["center_y"][:, k] = np.clip(((b[:, 1] + 2) >> 2) * 4, fp["mv_min_y"] + R4, fp["mv_max_y"] - R4)
        else:
            reqs["center_x"][:, k] = np.clip(((p[:, 0] + 2) >> 2) * 4, fp["mv_min_x"], fp["mv_max_x"])
            reqs["center_y"][:, k] = np.clip(((p[:, 1] + 2) >> 2) * 4, fp["mv_min_y"], fp["mv_max_y"])
        reqs["flags"][:, k] = int(fp["flags"]) & (REQ_SUBPEL | (REQ_TEST8X8 if t <= 4 else 0))
    reqs["mode"] = int(fp["mode"]); reqs["ref"] = int(fp["ref"])
    reqs["lambda"] = fp["lambda"]
    reqs["min_mcost"] = DISTBLK_MAX
    return reqs.reshape(-1)


def epzs_frame_params(lam, flags=EPZS_ADAPT_PATTERN | EPZS_DUAL | EPZS_SUBPEL, ref=0, pattern=2, pattern_dual=2, n_shared=0, window=0,
                      search_range=32, scales=(0, 1, 2, 1), mv_range=MV_RANGE_L51):
    """Defaults = bin/encoder.cfg: EPZSPattern 2 (extended diamond), EPZSDualRefinement 3 (extended diamond), threshold scalers
    EPZSMinThresScale 0, Med 1, Max 2, SubPel 1.  Thresholds are costs: base * scale << 5 (up_scale, me_epzs_common.c:459)."""
    fp = np.zeros(1, EPZS_FRAME_PARAMS)
    fp["lambda"] = lam; fp["flags"] = flags; fp["ref"] = ref; fp["pattern"] = pattern; fp["pattern_dual"] = pattern_dual
    fp["n_shared"] = n_shared; fp["window"] = window; fp["range"] = 4 * search_range
    fp["minthres"][0] = [scales[0] * b << 5 for b in EPZS_MIN_BASE]; fp["medthres"][0] = [scales[1] * b << 5 for b in EPZS_MED_BASE]
    fp["maxthres"][0] = [scales[2] * b << 5 for b in EPZS_MAX_BASE]; fp["subthres"][0] = [scales[3] * b << 5 for b in EPZS_MED_BASE]
    fp["mv_min_x"], fp["mv_max_x"], fp["mv_min_y"], fp["mv_max_y"] = mv_range
    return fp


def epzs_requests_from_frame(pred, fp, mb_w, mb_index=None):
    """The jmb_epzs_req list jmb_epzs_search_frame generates on the device (same rules; tests and the CPU legs).
    mb_index: picture addresses of the macroblocks in `pred` (default 0..n-1); cand_off always counts from `pred`'s first row."""
    fp = fp[0]
    n_mb = len(pred)
    reqs = np.zeros((n_mb, NPART), EPZS_REQ)
    addr = np.arange(n_mb) if mb_index is None else np.asarray(mb_index)
    mbx = (addr % mb_w) * 16; mby = (addr // mb_w) * 16
    big = DISTBLK_MAX
    ld = 2 * int(fp["lambda"][0])
    for k, (t, x, y) in enumerate(mb_partitions()):
        p = pred["pred"][:, k].astype(np.int32)
        reqs["pos_x"][:, k] = mbx + x; reqs["pos_y"][:, k] = mby + y
        reqs["pred_x"][:, k] = p[:, 0]; reqs["pred_y"][:, k] = p[:, 1]
        reqs["start_x"][:, k] = np.clip(p[:, 0], fp["mv_min_x"], fp["mv_max_x"]); reqs["start_y"][:, k] = np.clip(p[:, 1], fp["mv_min_y"], fp["mv_max_y"])
        reqs["blocktype"][:, k] = t
        reqs["flags"][:, k] = (int(fp["flags"]) & (EPZS_ADAPT_PATTERN | EPZS_DUAL | EPZS_SUBPEL | (EPZS_TEST8X8 if t <= 4 else 0))) | (EPZS_WINDOW_GEN if fp["window"] else 0)
        med = int(fp["medthres"][t])
        stop = min(max(big, int(fp["minthres"][t])), int(fp["maxthres"][t]) + ld)
        stop = ((8 * max(med + ld, stop) + med) >> 3) + ld
        reqs["stop"][:, k] = stop; reqs["medthres"][:, k] = med; reqs["subthres"][:, k] = int(fp["subthres"][t])
    reqs["ref"] = int(fp["ref"]); reqs["jm_ref"] = int(fp["ref"]); reqs["pattern"] = int(fp["pattern"]); reqs["pattern_dual"] = int(fp["pattern_dual"])
    reqs["n_cand"][:, :, 0] = int(fp["n_shared"]); reqs["n_cand"][:, :, 2] = 8 * int(fp["window"]) - 1 if fp["window"] else 0
    reqs["gate"][:, :, 2] = 3
    reqs["cand_off"] = (np.arange(n_mb) * int(fp["n_shared"]))[:, None]
    reqs["lambda"] = fp["lambda"]
    reqs["range_x"] = reqs["range_y"] = int(fp["range"])
    reqs["prev_sad"] = big; reqs["min_mcost"] = big
    return reqs.reshape(-1)


# QP_SCALE_CR of lcommon (chroma qp from luma qp, H.264 table 8-15) and the chroma descriptor JM's tables give for an inter macroblock
QP_SCALE_CR = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 32, 33,
               34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]


def chroma_desc(yuv_format, qp_luma, q_params_fn, c_cost, is_cavlc, chroma_qp_offset=(0, 0)):
    """q_params_fn(qp) -> [4][4][3] inter luma/chroma parameters (flat matrices: the same table serves Y, U and V)."""
    d = np.zeros(1, CHROMA_DESC)
    d["yuv_format"] = yuv_format; d["is_cavlc"] = int(is_cavlc)
    for uv in range(2):
        qpc = QP_SCALE_CR[min(51, max(0, qp_luma + chroma_qp_offset[uv]))]
        qp_dc = qpc + (3 if yuv_format == 2 else 0)
        d["qp_ac"][0, uv] = qpc; d["qp_dc"][0, uv] = qp_dc
        d["params_ac"][0, uv] = np.asarray(q_params_fn(qpc), np.int32).reshape(16, 3)
        d["params_dc"][0, uv] = np.asarray(q_params_fn(qp_dc), np.int32).reshape(16, 3)[0]
    d["c_cost"][0] = np.asarray(c_cost, np.uint8)[:16]
    return d


def quant_desc(n, qp, qparams, scan, c_cost, is_cavlc, around=0, arw=0):
    q = np.zeros(1, QUANT_DESC)
    q["n"], q["qp"], q["is_cavlc"], q["around"], q["adapt_rnd_weight"] = n, qp, int(is_cavlc), int(around), int(arw)
    q["qparams"][0, :n * n] = np.asarray(qparams, np.int32).reshape(n * n, 3)
    q["scan"][0, :n * n] = np.asarray(scan, np.uint8).reshape(n * n, 2)
    m = min(len(c_cost), 64)
    q["c_cost"][0, :m] = np.asarray(c_cost, np.uint8)[:m]
    return q


class Context:
    """One libjmb200 context = one GPU, one stream."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.jmb_create(device, C.byref(h))
        if rc:
            raise JMBError(f"jmb_create({device}) = {rc}: {self.L.jmb_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.jmb_destroy(self.h); self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise JMBError(f"libjmb200 error {rc}: {self.L.jmb_last_error(self.h).decode()}")

    def sync(self):
        self._ck(self.L.jmb_sync(self.h))

    @property
    def stream(self):
        return self.L.jmb_stream(self.h)

    @property
    def launches(self):
        return int(self.L.jmb_launch_count(self.h))

    def pinned(self, shape, dtype):
        """numpy array over cudaHostAlloc'd memory (freed with the context's process)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._ck(self.L.jmb_host_alloc(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def configure(self, search_range=32, max_mvd=None, metric=(SAD, SATD, SATD), start_hp=None, start_qp=None,
                  search_pos2=9, search_pos4=9):
        if max_mvd is None:   # lencod/src/mv_search.c:325-329
            import math
            bits = 3 + 2 * int(math.ceil(math.log(4 * (2 * search_range + 3) + 1) / math.log(2) + 1e-10))
            max_mvd = (1 << (bits >> 1)) - 1
        if start_hp is None:
            start_hp = 0 if metric[0] != metric[1] else 1
        if start_qp is None:
            start_qp = 0 if metric[1] != metric[2] else 1
        cfg = MEConfig(search_range, max_mvd, (C.c_int32 * 3)(*metric), start_hp, start_qp, search_pos2, search_pos4)
        self._ck(self.L.jmb_me_configure(self.h, C.byref(cfg)))
        self.search_range, self.max_mvd = search_range, max_mvd

    def ref_put(self, slot, luma, loc=HOST, shape=None, stride=None, bitdepth=8):
        if loc == HOST:
            luma = np.ascontiguousarray(luma, np.uint16)
            h, w = luma.shape
            stride = w
        else:
            h, w = shape
            stride = stride or w
        self._ck(self.L.jmb_ref_put(self.h, slot, _ptr(luma), w, h, stride, bitdepth, loc))

    def ref_put_u8(self, slot, luma, loc=HOST, shape=None, stride=None):
        if loc != DEVICE:
            assert luma.dtype == np.uint8 and luma.flags["C_CONTIGUOUS"]
            h, w = luma.shape
            stride = w
        else:
            h, w = shape
            stride = stride or w
        self._ck(self.L.jmb_ref_put_u8(self.h, slot, _ptr(luma), w, h, stride, loc))

    def pic_begin_u8(self, cur, ref_slots, loc=HOST, shape=None, stride=None):
        if loc != DEVICE:
            assert cur.dtype == np.uint8 and cur.flags["C_CONTIGUOUS"]
            h, w = cur.shape
            stride = w
        else:
            h, w = shape
            stride = stride or w
        arr = (C.c_int * len(ref_slots))(*ref_slots)
        self._ck(self.L.jmb_pic_begin_u8(self.h, _ptr(cur), w, h, stride, loc, arr, len(ref_slots)))

    def me_search_frame_pred(self, pred, fp, res=None, loc=HOST, n_mb=None, want_res=True):
        """pred: MB_MVPRED[n_mb] (or device pointer); fp: FRAME_PARAMS[1]; returns ME_RES8[n_mb * 41] (HOST)."""
        if loc != DEVICE:
            n_mb = len(pred)
            if res is None and want_res:
                res = np.zeros(n_mb * NPART, ME_RES8)
        self._ck(self.L.jmb_me_search_frame_pred(self.h, _ptr(pred), n_mb, _ptr(fp), None if res is None else _ptr(res), loc))
        return res

    def mc_tq_modes_compact(self, res, qdesc, mode_mask=0x7F, loc=HOST, n_mb=None, out=None, token_cap=None):
        """res=None: resident search results.  HOST: returns (heads[7][n_mb], tokens[n_tokens])."""
        if loc == HOST:
            if res is not None:
                res = np.ascontiguousarray(res, ME_RES)
                n_mb = len(res) // NPART
            token_cap = token_cap or 7 * n_mb * 256
            heads = np.zeros((7, n_mb), TQ_HEAD); tokens = np.zeros(token_cap, TQ_TOKEN); n_tok = np.zeros(1, np.uint32)
        else:
            heads, tokens, n_tok = out
        self._ck(self.L.jmb_mc_tq_modes_compact(self.h, None if res is None else _ptr(res), n_mb, mode_mask, _ptr(qdesc), _ptr(heads),
                                                _ptr(tokens), token_cap, _ptr(n_tok), loc))
        if loc == HOST:
            return heads, tokens[:int(n_tok[0])]
        return heads, tokens, n_tok

    def ref_put_chroma(self, slot, u, v, loc=HOST, shape=None, sample_bytes=1):
        if loc != DEVICE:
            assert u.dtype == v.dtype and u.flags["C_CONTIGUOUS"] and v.flags["C_CONTIGUOUS"] and u.shape == v.shape
            shape = u.shape; sample_bytes = u.dtype.itemsize
        self._ck(self.L.jmb_ref_put_chroma(self.h, slot, _ptr(u), _ptr(v), sample_bytes, shape[1], shape[0], shape[1], loc))

    def pic_chroma(self, u, v, loc=HOST, shape=None, sample_bytes=1):
        if loc != DEVICE:
            assert u.dtype == v.dtype and u.flags["C_CONTIGUOUS"] and v.flags["C_CONTIGUOUS"] and u.shape == v.shape
            shape = u.shape; sample_bytes = u.dtype.itemsize
        self._ck(self.L.jmb_pic_chroma(self.h, _ptr(u), _ptr(v), sample_bytes, shape[1], shape[0], shape[1], loc))

    def chroma_residual_coding(self, desc, pred=None, mode=1, first_mb=0, n_mb=None, loc=HOST, out=None, want_recon=True):
        """pred: MB_PRED[n_mb] or None (partition mode `mode` of the resident search results).  HOST: returns dict(dc, ac, cbp_blk,
        cr_cbp, recon) with dc [n_mb][2][8], ac [n_mb][2][8][15], cbp_blk / cr_cbp [n_mb][2], recon [n_mb][2][16][8]."""
        if loc != DEVICE:
            if pred is not None:
                pred = np.ascontiguousarray(pred, MB_PRED); n_mb = len(pred)
            o = dict(dc=np.zeros((n_mb, 2, 8), np.int16), ac=np.zeros((n_mb, 2, 8, 15), np.int16), cbp_blk=np.zeros((n_mb, 2), np.uint32),
                     cr_cbp=np.zeros((n_mb, 2), np.uint32), recon=np.zeros((n_mb, 2, 16, 8), np.uint8) if want_recon else None)
            self._ck(self.L.jmb_chroma_residual_coding(self.h, None if pred is None else _ptr(pred), mode, first_mb, n_mb, _ptr(desc), _ptr(o["dc"]),
                                                       _ptr(o["ac"]), _ptr(o["cbp_blk"]), _ptr(o["cr_cbp"]), None if o["recon"] is None else _ptr(o["recon"]), loc))
            return o
        dc, ac, cb, cc, rec = out
        self._ck(self.L.jmb_chroma_residual_coding(self.h, None if pred is None else _ptr(pred), mode, first_mb, n_mb, _ptr(desc), dc, ac, cb, cc, rec, loc))

    def mb_surfaces(self, ref, mb, center, radius):
        self._ck(self.L.jmb_mb_surfaces(self.h, ref, mb[0], mb[1], center[0], center[1], radius))

    def mb_search(self, req):
        req = np.ascontiguousarray(req, ME_REQ).reshape(1)
        res = np.zeros(1, ME_RES)
        self._ck(self.L.jmb_mb_search(self.h, _ptr(req), _ptr(res)))
        return res[0]

    def mb_chain(self, reqs, mv_limits, int_divide=1):
        reqs = np.ascontiguousarray(reqs, CHAIN_REQ)
        res = np.zeros(len(reqs), CHAIN_RES)
        lim = np.ascontiguousarray(mv_limits, np.int32)
        self._ck(self.L.jmb_mb_chain(self.h, _ptr(reqs), len(reqs), _ptr(lim), int_divide, _ptr(res)))
        return res

    def deblock_picture(self, luma, cb, cr, yuv, slice_type, mbs, direct8x8inf=1):
        """DeblockFrame on host planes (u8); returns the filtered planes."""
        luma = np.ascontiguousarray(luma, np.uint8).copy(); h, w = luma.shape
        cb = np.ascontiguousarray(cb, np.uint8).copy() if yuv else None; cr = np.ascontiguousarray(cr, np.uint8).copy() if yuv else None
        mbs = np.ascontiguousarray(mbs, DB_MB)
        if len(mbs) != (w // 16) * (h // 16):
            raise ValueError("one jmb_db_mb per macroblock")
        self._ck(self.L.jmb_deblock_picture(self.h, _ptr(luma), w, _ptr(cb) if yuv else None, _ptr(cr) if yuv else None, w // 2, w, h, yuv, slice_type,
                                            direct8x8inf, _ptr(mbs), HOST))
        return luma, cb, cr

    def deblock_picture_dev(self, d_luma, pitch, d_cb, d_cr, pitch_c, w, h, yuv, slice_type, d_mbs, direct8x8inf=1):
        """The same in place on device-resident planes / records (raw device pointers); only enqueues work."""
        self._ck(self.L.jmb_deblock_picture(self.h, d_luma, pitch, d_cb, d_cr, pitch_c, w, h, yuv, slice_type, direct8x8inf, d_mbs, DEVICE))

    def block_distortion(self, metric, n, diff, thres=None):
        diff = np.ascontiguousarray(diff, np.int16).reshape(-1, n * n)
        out = np.zeros(len(diff), np.int32)
        if thres is not None:
            thres = np.ascontiguousarray(thres, np.int32)
        self._ck(self.L.jmb_block_distortion(self.h, metric, n, _ptr(diff), len(diff), None if thres is None else _ptr(thres), _ptr(out), HOST))
        return out

    def epzs_search(self, reqs, cands, loc=HOST, res=None, n=None, n_cands=None):
        if loc != DEVICE:
            reqs = np.ascontiguousarray(reqs, EPZS_REQ); n = len(reqs)
            cands = np.ascontiguousarray(cands, np.int16).reshape(-1, 2); n_cands = len(cands)
            if res is None:
                res = np.zeros(n, EPZS_RES)
        self._ck(self.L.jmb_epzs_search(self.h, _ptr(reqs), n, _ptr(cands) if n_cands else None, n_cands, _ptr(res), loc))
        return res

    def epzs_search_frame(self, pred, shared, fp, res=None, loc=HOST, n_mb=None, want_res=True):
        if loc != DEVICE:
            n_mb = len(pred)
            if res is None and want_res:
                res = np.zeros(n_mb * NPART, ME_RES8)
        self._ck(self.L.jmb_epzs_search_frame(self.h, _ptr(pred), None if shared is None else _ptr(shared), n_mb, _ptr(fp),
                                              None if res is None else _ptr(res), loc))
        return res

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.jmb_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        self._ck(self.L.jmb_dev_free(self.h, p))

    def dev_copy(self, dst, src, nbytes, dst_loc, src_loc):
        self._ck(self.L.jmb_dev_copy(self.h, _ptr(dst), _ptr(src), nbytes, dst_loc, src_loc))

    def peer_export(self, dev_ptr):
        h = np.zeros(IPC_HANDLE_BYTES, np.uint8)
        self._ck(self.L.jmb_peer_export(self.h, dev_ptr, _ptr(h)))
        return h

    def peer_open(self, handle):
        handle = np.ascontiguousarray(handle, np.uint8)
        p = C.c_void_p()
        self._ck(self.L.jmb_peer_open(self.h, _ptr(handle), C.byref(p)))
        return p.value

    def peer_close(self, mapped):
        self._ck(self.L.jmb_peer_close(self.h, mapped))

    def ref_drop(self, slot):
        self._ck(self.L.jmb_ref_drop(self.h, slot))

    def ref_plane(self, slot, fy, fx, shape):
        h, w = shape
        out = np.empty((h + 40, w + 64), np.uint16)
        self._ck(self.L.jmb_ref_get_plane(self.h, slot, fy, fx, _ptr(out), HOST))
        return out

    def pic_begin(self, cur, ref_slots, loc=HOST, shape=None, stride=None):
        if loc == HOST:
            cur = np.ascontiguousarray(cur, np.uint16)
            h, w = cur.shape
            stride = w
        else:
            h, w = shape
            stride = stride or w
        arr = (C.c_int * len(ref_slots))(*ref_slots)
        self._ck(self.L.jmb_pic_begin(self.h, _ptr(cur), w, h, stride, loc, arr, len(ref_slots)))

    def me_search(self, reqs, res=None, loc=HOST, n=None, frame=False):
        if loc == HOST:
            reqs = np.ascontiguousarray(reqs, ME_REQ)
            n = len(reqs)
            if res is None:
                res = np.zeros(n, ME_RES)
        if frame:
            assert n % NPART == 0
            self._ck(self.L.jmb_me_search_frame(self.h, _ptr(reqs), n // NPART, _ptr(res), loc))
        else:
            self._ck(self.L.jmb_me_search(self.h, _ptr(reqs), n, _ptr(res), loc))
        return res

    def ffs_surfaces(self, ref, mb, center):
        n = (2 * self.search_range + 1) ** 2
        out = np.zeros((8, 16, n), np.uint32)
        self._ck(self.L.jmb_ffs_surfaces(self.h, ref, mb[0], mb[1], center[0], center[1], _ptr(out), HOST))
        return out

    def dist(self, ref, metric, blocktype, pos, cands, test8x8=0):
        cands = np.ascontiguousarray(cands, np.int16).reshape(-1, 2)
        out = np.zeros(len(cands), np.int32)
        self._ck(self.L.jmb_dist(self.h, ref, metric, blocktype, pos[0], pos[1], _ptr(cands), len(cands), test8x8, _ptr(out), HOST))
        return out

    def dist_ex(self, ref, ref2, metric, form, blocktype, pos, cands, cand2, wp=(32, 32, 0, 5, 16), test8x8=0):
        """jmb_dist_ex: form PRED_PLAIN / PRED_WEIGHTED / PRED_AVERAGE / PRED_WEIGHTED_AVERAGE;
        wp = (weight1, weight2, offset, luma_log_weight_denom, wp_luma_round)."""
        cands = np.ascontiguousarray(cands, np.int16).reshape(-1, 2)
        out = np.zeros(len(cands), np.int32)
        d = np.zeros(1, DIST_PRED)
        d["form"] = form; d["ref2"] = ref2; d["cand2_x"], d["cand2_y"] = cand2
        d["weight1"], d["weight2"], d["offset"], d["log_weight_denom"], d["wp_round"] = wp
        self._ck(self.L.jmb_dist_ex(self.h, ref, _ptr(d), metric, blocktype, pos[0], pos[1], _ptr(cands), len(cands), test8x8, _ptr(out), HOST))
        return out

    def forward_transform(self, blocks, n):
        b = np.ascontiguousarray(blocks, np.int32).reshape(-1, n * n).copy()
        self._ck(self.L.jmb_forward_transform(self.h, _ptr(b), len(b), n, HOST))
        return b.reshape(-1, n, n)

    def hadamard(self, kind, vals, per):
        b = np.ascontiguousarray(vals, np.int32).reshape(-1, per).copy()
        self._ck(self.L.jmb_hadamard(self.h, kind, _ptr(b), len(b), HOST))
        return b

    def quant_list(self, plan, coef_flat, cost0=0):
        """One list through jmb_quant_list, gathered / scattered like the shim does."""
        coef = np.ascontiguousarray(coef_flat, np.int32).reshape(-1).copy()
        order = np.asarray(plan["order"]); m = len(order)
        d = np.zeros(1, QLIST_DESC)
        for k in ("m", "q_bits", "qp_per", "dequant", "clip", "use_cost", "around"):
            d[k] = m if k == "m" else plan[k]
        d["adapt_rnd_weight"] = plan["arw"]; d["params"][0, :m] = plan["params"]; d["c_cost"][0] = plan["c_cost"]
        lst = coef[order].copy(); levels = np.zeros(17, np.int32); runs = np.zeros(17, np.int32); fadj = np.zeros(m, np.int32)
        cost = np.array([cost0], np.int32); nz = np.zeros(1, np.int32)
        self._ck(self.L.jmb_quant_list(self.h, _ptr(d), _ptr(lst), 1, _ptr(levels), _ptr(runs), _ptr(fadj), _ptr(cost), _ptr(nz), HOST))
        coef[order] = lst
        fa = np.zeros(len(coef), np.int32); fa[order] = fadj
        return dict(nonzero=int(nz[0]), coef=coef, levels=levels, runs=runs, fadjust=fa, coeff_cost=int(cost[0]))

    def inverse_transform(self, blocks, n):
        b = np.ascontiguousarray(blocks, np.int32).reshape(-1, n * n).copy()
        self._ck(self.L.jmb_inverse_transform(self.h, _ptr(b), len(b), n, HOST))
        return b.reshape(-1, n, n)

    def luma_residual_coding_modes(self, res, qdesc, mode_mask=0x7F, n_mb=None, want_recon=True):
        """res=None: resident search results.  Returns dict(levels, cost8, cbp_blk, cbp, recon, sse), all mode-major."""
        if res is not None:
            res = np.ascontiguousarray(res, ME_RES)
            n_mb = len(res) // NPART
        o = dict(levels=np.zeros((7, n_mb, 256), np.int16), cost8=np.zeros((7, n_mb, 4), np.int32), cbp_blk=np.zeros((7, n_mb), np.uint32),
                 cbp=np.zeros((7, n_mb), np.uint32), recon=np.zeros((7, n_mb, 16, 16), np.uint8) if want_recon else None,
                 sse=np.zeros((7, n_mb), np.int32))
        self._ck(self.L.jmb_luma_residual_coding_modes(self.h, None if res is None else _ptr(res), n_mb, mode_mask, _ptr(qdesc), _ptr(o["levels"]),
                                                       _ptr(o["cost8"]), _ptr(o["cbp_blk"]), _ptr(o["cbp"]),
                                                       None if o["recon"] is None else _ptr(o["recon"]), _ptr(o["sse"]), HOST))
        return o

    def luma_residual_coding(self, pred, qdesc, first_mb=0):
        pred = np.ascontiguousarray(pred, MB_PRED); n_mb = len(pred)
        o = dict(levels=np.zeros((n_mb, 256), np.int16), cost8=np.zeros((n_mb, 4), np.int32), cbp_blk=np.zeros(n_mb, np.uint32),
                 cbp=np.zeros(n_mb, np.uint32), recon=np.zeros((n_mb, 16, 16), np.uint8), sse=np.zeros(n_mb, np.int32))
        self._ck(self.L.jmb_luma_residual_coding(self.h, _ptr(pred), first_mb, n_mb, _ptr(qdesc), _ptr(o["levels"]), _ptr(o["cost8"]),
                                                 _ptr(o["cbp_blk"]), _ptr(o["cbp"]), _ptr(o["recon"]), _ptr(o["sse"]), HOST))
        return o

    def quant_blocks(self, qdesc, coef, do_transform=0, cost0=0):
        n = int(qdesc["n"][0]); nn = n * n
        coef = np.ascontiguousarray(coef, np.int32).reshape(-1, nn).copy()
        nblk = len(coef)
        lr = 17 if n == 4 else (68 if int(qdesc["is_cavlc"][0]) else 65)
        levels = np.zeros((nblk, lr), np.int32); runs = np.zeros((nblk, lr), np.int32)
        fadj = np.zeros((nblk, nn), np.int32)
        cost = np.full(nblk, cost0, np.int32); nz = np.zeros(nblk, np.int32)
        self._ck(self.L.jmb_quant_blocks(self.h, _ptr(qdesc), int(do_transform), _ptr(coef), nblk, _ptr(levels), _ptr(runs),
                                         _ptr(fadj), _ptr(cost), _ptr(nz), HOST))
        return dict(coef=coef.reshape(-1, n, n), levels=levels, runs=runs, fadjust=fadj.reshape(-1, n, n), coeff_cost=cost, nonzero=nz)

    def mc_tq(self, pred, qdesc, loc=HOST, n_mb=None, out=None):
        if loc == HOST:
            pred = np.ascontiguousarray(pred, MB_PRED)
            n_mb = len(pred)
            levels = np.zeros((n_mb, 256), np.int16); cost = np.zeros((n_mb, 4), np.int32); cbp = np.zeros(n_mb, np.uint32)
        else:
            levels, cost, cbp = out
        self._ck(self.L.jmb_mc_tq(self.h, _ptr(pred), n_mb, _ptr(qdesc), _ptr(levels), _ptr(cost), _ptr(cbp), loc))
        return levels, cost, cbp

    def mc_tq_modes(self, res, qdesc, mode_mask=0x7F, loc=HOST, n_mb=None, out=None):
        """res=None: the search results still resident on the device.  Returns (levels[7][n_mb][256], cost[7][n_mb][4], cbp[7][n_mb])."""
        if loc == HOST:
            if res is not None:
                res = np.ascontiguousarray(res, ME_RES)
                n_mb = len(res) // NPART
            levels = np.zeros((7, n_mb, 256), np.int16); cost = np.zeros((7, n_mb, 4), np.int32); cbp = np.zeros((7, n_mb), np.uint32)
        else:
            levels, cost, cbp = out
        self._ck(self.L.jmb_mc_tq_modes(self.h, None if res is None else _ptr(res), n_mb, mode_mask, _ptr(qdesc), _ptr(levels), _ptr(cost), _ptr(cbp), loc))
        return levels, cost, cbp

    def pred_from_results(self, res, mode, loc=HOST, n_mb=None, out=None):
        if loc == HOST:
            res = np.ascontiguousarray(res, ME_RES)
            n_mb = len(res) // NPART
            out = np.zeros(n_mb, MB_PRED)
        self._ck(self.L.jmb_pred_from_results(self.h, _ptr(res), n_mb, mode, _ptr(out), loc))
        return out

    def timing(self, on):
        self._ck(self.L.jmb_timing_enable(self.h, int(on)))

    def timing_get(self, kernel):
        ms, n = C.c_double(), C.c_int()
        self._ck(self.L.jmb_timing_get(self.h, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

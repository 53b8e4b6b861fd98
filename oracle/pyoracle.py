"""ctypes bindings for the CPU checker libraries.  TEST INFRASTRUCTURE ONLY.

Two libraries live behind this module:
  * ``oracle/libjmoracle.so``   -- our plain-C restatement (oracle/jm_oracle.c), class ``Oracle``.
  * ``oracle/_ref/libjmref.so`` -- the REAL JM 19.0 leaf functions, compiled from the reference's own
    sources by oracle/Makefile and reached through oracle/ref_harness.c, class ``JMRef``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (jm_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libjmoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libjmref.so")
PAD_X, PAD_Y = 32, 20
SAD, SSE, SATD = 0, 1, 2
DISTBLK_MAX = 0x7FFFFFFF << 5   # lencod/inc/defines.h:136
BLOCK_SIZE = [(16, 16), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]

EPZS_RES = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("imv_x", "<i2"), ("imv_y", "<i2"), ("cost", "<i8"), ("icost", "<i8"),
                     ("prev_sad", "<i8"), ("exit_code", "<i4"), ("n_evals", "<i4")])

_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_oracle():
    """Compile the restatement (gcc, < 1 s).  Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def ref_available():
    return os.path.exists(REF_SO)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        self.L = L
        L.jmo_ref_create.restype = C.c_void_p
        L.jmo_ref_create.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.jmo_ref_destroy.argtypes = [C.c_void_p]
        L.jmo_ref_get_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, _u16p]
        L.jmo_spiral.argtypes = [C.c_int, _i16p]
        L.jmo_dist.argtypes = [C.c_void_p, _u16p] + [C.c_int] * 6
        L.jmo_dist_ex.argtypes = [C.c_void_p, C.c_void_p, _u16p] + [C.c_int] * 9 + [_i32p, C.c_int]
        L.jmo_full_search.restype = C.c_int64
        L.jmo_full_search.argtypes = [C.c_void_p, _u16p] + [C.c_int] * 9 + [C.c_int64, C.c_int, _i16p]
        L.jmo_sub_pel.restype = C.c_int64
        L.jmo_sub_pel.argtypes = [C.c_void_p, _u16p] + [C.c_int] * 8 + [_i32p, C.c_int64] + [C.c_int] * 5 + [_i16p]
        L.jmo_ffs_center.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i16p]
        L.jmo_ffs_setup.argtypes = [C.c_void_p, _u16p] + [C.c_int] * 6 + [_u32p]
        L.jmo_ffs_search.restype = C.c_int64
        L.jmo_ffs_search.argtypes = [_u32p] + [C.c_int] * 8 + [C.c_int64, C.c_int, _i16p]
        L.jmo_forward4x4.argtypes = [_i32p]
        L.jmo_forward8x8.argtypes = [_i32p]
        L.jmo_quant_list.argtypes = [C.c_int] * 8 + [_i32p, _u8p, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.jmo_hadamard.argtypes = [C.c_int, _i32p]
        L.jmo_inverse4x4.argtypes = [_i32p]
        L.jmo_inverse8x8.argtypes = [_i32p]
        L.jmo_luma_residual_coding.restype = C.c_int64
        L.jmo_luma_residual_coding.argtypes = [_u16p, _u16p, C.c_int, C.c_int, _i32p, _u8p, _u8p, C.c_int, C.c_int, _i16p, _i32p, _i32p, _i32p, _u16p]
        L.jmo_hadamard_sad4x4.argtypes = [_i16p]
        L.jmo_hadamard_sad8x8.argtypes = [_i16p]
        L.jmo_quant.argtypes = [C.c_int, _i32p, C.c_int, _i32p, _u8p, _u8p, C.c_int, C.c_int,
                                _i32p, _i32p, _i32p, _i32p]

        L.jmo_epzs.argtypes = [C.c_void_p, _u16p, C.c_int, C.c_void_p, _i16p, _i32p, C.c_void_p]
        L.jmo_chroma_pred.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i16p, _u8p]
        L.jmo_chroma_rc.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _u8p, C.c_int, _i16p, _i16p, _u32p, _u8p]
        L.jmo_epzs_batch.argtypes = [C.c_void_p, _u16p, C.c_int, C.c_void_p, C.c_int, _i16p, _i32p, C.c_void_p]
        L.jmo_mc_tq_modes_mb.argtypes = [C.c_void_p, _u16p, C.c_int, C.c_int, C.c_int, _i16p, C.c_int, C.c_int, _i32p, _u8p, _u8p, C.c_int,
                                         C.c_uint, _i16p]

    def chroma_pred(self, ref_c, yuv, mb_c, mv16):
        """ref_c: one chroma plane (uint8); mb_c = (x, y) of the macroblock in chroma samples; mv16: [16][2] luma 4x4 mvs."""
        ref_c = np.ascontiguousarray(ref_c, np.uint8)
        hmb = 8 if yuv == 1 else 16
        pred = np.zeros((hmb, 8), np.uint8)
        self.L.jmo_chroma_pred(ref_c, ref_c.shape[1], ref_c.shape[0], ref_c.shape[1], yuv, mb_c[0], mb_c[1], np.ascontiguousarray(mv16, np.int16).reshape(-1), pred)
        return pred

    def chroma_rc(self, src, pred, yuv, qp_ac, qp_dc, params_ac, params_dc, c_cost, is_cavlc):
        hmb = 8 if yuv == 1 else 16
        dc = np.zeros(8, np.int16); ac = np.zeros((8, 15), np.int16); bits = np.zeros(1, np.uint32); rec = np.zeros((hmb, 8), np.uint8)
        cr = self.L.jmo_chroma_rc(np.ascontiguousarray(src, np.uint8), np.ascontiguousarray(pred, np.uint8), yuv, qp_ac, qp_dc,
                                  np.ascontiguousarray(params_ac, np.int32).reshape(-1), np.ascontiguousarray(params_dc, np.int32).reshape(-1),
                                  np.ascontiguousarray(c_cost, np.uint8), int(is_cavlc), dc, ac.reshape(-1), bits, rec)
        return dict(dc=dc, ac=ac, cbp_blk=int(bits[0]), cr_cbp=int(cr), recon=rec)

    def epzs_batch(self, r, cur, reqs, cands, me):
        """jmo_epzs over a whole request array in one C call (timed by bench.py's CPU legs)."""
        cur = np.ascontiguousarray(cur, np.uint16)
        reqs = np.ascontiguousarray(reqs)
        cands = np.ascontiguousarray(cands, np.int16).reshape(-1)
        if len(cands) == 0:
            cands = np.zeros(2, np.int16)
        res = np.zeros(len(reqs), EPZS_RES)
        self.L.jmo_epzs_batch(r[0], cur, cur.shape[1], reqs.ctypes.data, len(reqs), cands, np.asarray(me, np.int32), res.ctypes.data)
        return res

    def mc_tq_modes_mb(self, r, cur, mb, mv41, n, qp, qparams, scan, c_cost, is_cavlc, mode_mask):
        cur = np.ascontiguousarray(cur, np.uint16)
        lev = np.zeros((7, 256), np.int16)
        self.L.jmo_mc_tq_modes_mb(r[0], cur, cur.shape[1], mb[0], mb[1], np.ascontiguousarray(mv41, np.int16).reshape(-1), n, qp,
                                  np.ascontiguousarray(qparams, np.int32).reshape(-1), np.ascontiguousarray(scan, np.uint8).reshape(-1),
                                  np.ascontiguousarray(c_cost, np.uint8), int(is_cavlc), mode_mask, lev)
        return lev

    def epzs(self, r, cur, reqs, cands, me):
        """reqs: structured array with the layout of jmb_epzs_req; cands int16 [n][2]; me = (metric_h, metric_q, start_hp,
        start_qp, search_pos2).  Returns a structured array with the layout of jmb_epzs_res."""
        cur = np.ascontiguousarray(cur, np.uint16)
        reqs = np.ascontiguousarray(reqs)
        cands = np.ascontiguousarray(cands, np.int16).reshape(-1)
        if len(cands) == 0:
            cands = np.zeros(2, np.int16)
        res = np.zeros(len(reqs), EPZS_RES)
        me = np.asarray(me, np.int32)
        for i in range(len(reqs)):
            self.L.jmo_epzs(r[0], cur, cur.shape[1], reqs[i:i + 1].ctypes.data, cands, me, res[i:i + 1].ctypes.data)
        return res

    # -- reference planes ---------------------------------------------------------------
    def ref_create(self, luma, max_value=255):
        luma = np.ascontiguousarray(luma, np.uint16)
        h, w = luma.shape
        return self.L.jmo_ref_create(luma, w, h, w, max_value), (w, h)

    def ref_destroy(self, r):
        self.L.jmo_ref_destroy(r[0])

    def planes(self, r):
        w, h = r[1]
        out = np.empty((4, 4, h + 2 * PAD_Y, w + 2 * PAD_X), np.uint16)
        for fy in range(4):
            for fx in range(4):
                self.L.jmo_ref_get_plane(r[0], fy, fx, out[fy, fx])
        return out

    def spiral(self, R):
        out = np.empty(((2 * R + 1) ** 2, 2), np.int16)
        self.L.jmo_spiral(R, out)
        return out

    def mvbits(self, v):
        return self.L.jmo_mvbits(int(v))

    def dist(self, r, cur, blocktype, pos, cand, metric, test8x8=0):
        bsx, bsy = BLOCK_SIZE[blocktype]
        src = np.ascontiguousarray(cur[pos[1]:pos[1] + bsy, pos[0]:pos[0] + bsx], np.uint16)
        return self.L.jmo_dist(r[0], src, bsx, bsy, cand[0], cand[1], metric, test8x8)

    def dist_ex(self, r1, r2, cur, blocktype, pos, cand1, cand2, metric, form, wp=(32, 32, 0, 5, 16), test8x8=0, max_value=255):
        """wp = (weight1, weight2, offset, luma_log_weight_denom, wp_luma_round); form 0 plain, 1 weighted, 2 average, 3 weighted average"""
        bsx, bsy = BLOCK_SIZE[blocktype]
        src = np.ascontiguousarray(cur[pos[1]:pos[1] + bsy, pos[0]:pos[0] + bsx], np.uint16)
        return self.L.jmo_dist_ex(r1[0], (r2 or r1)[0], src, bsx, bsy, cand1[0], cand1[1], cand2[0], cand2[1], metric, test8x8, form,
                                  np.asarray(wp, np.int32), max_value)

    def full_search(self, r, cur, blocktype, pos, pred, center, lam, min_mcost, R):
        cur = np.ascontiguousarray(cur, np.uint16)
        mv = np.zeros(2, np.int16)
        c = self.L.jmo_full_search(r[0], cur, cur.shape[1], blocktype, pos[0], pos[1], pred[0], pred[1],
                                   center[0], center[1], lam, min_mcost, R, mv)
        return (int(mv[0]), int(mv[1])), c

    def sub_pel(self, r, cur, blocktype, pos, pred, mv_in, lam3, min_mcost, metric_h, metric_q,
                start_hp, start_qp, test8x8=0):
        cur = np.ascontiguousarray(cur, np.uint16)
        mv = np.zeros(2, np.int16)
        c = self.L.jmo_sub_pel(r[0], cur, cur.shape[1], blocktype, pos[0], pos[1], pred[0], pred[1],
                               mv_in[0], mv_in[1], np.asarray(lam3, np.int32), min_mcost, metric_h, metric_q,
                               start_hp, start_qp, test8x8, mv)
        return (int(mv[0]), int(mv[1])), c

    def ffs_center(self, pmv, R, hq=(-8192, 8191), vq=(-2048, 2047)):
        c = np.zeros(2, np.int16)
        self.L.jmo_ffs_center(pmv[0], pmv[1], R, np.asarray(hq, np.int32), np.asarray(vq, np.int32), c)
        return int(c[0]), int(c[1])

    def ffs_setup(self, r, cur, mb, center, R):
        cur = np.ascontiguousarray(cur, np.uint16)
        bs = np.zeros((8, 16, (2 * R + 1) ** 2), np.uint32)
        self.L.jmo_ffs_setup(r[0], cur, cur.shape[1], mb[0], mb[1], center[0], center[1], R, bs)
        return bs

    def ffs_search(self, bs, R, blocktype, block_index, center, pred, lam, min_mcost, max_mvd):
        mv = np.zeros(2, np.int16)
        c = self.L.jmo_ffs_search(bs, R, blocktype, block_index, center[0], center[1], pred[0], pred[1],
                                  lam, min_mcost, max_mvd, mv)
        return (int(mv[0]), int(mv[1])), c

    # -- transforms / quant ---------------------------------------------------------------
    def forward4x4(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmo_forward4x4(b.reshape(-1))
        return b

    def forward8x8(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmo_forward8x8(b.reshape(-1))
        return b

    def quant_list(self, plan, coef_flat, cost0=0):
        """plan = jm_b200.api.qlist_plan(...); coef_flat = the function's coefficient array, flattened."""
        return _qlist_call(lambda *a: self.L.jmo_quant_list(*a), plan, coef_flat, cost0)

    def hadamard(self, kind, vals):
        b = np.ascontiguousarray(vals, np.int32).reshape(-1).copy()
        self.L.jmo_hadamard(kind, b)
        return b

    def inverse4x4(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmo_inverse4x4(b.reshape(-1))
        return b

    def inverse8x8(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmo_inverse8x8(b.reshape(-1))
        return b

    def luma_residual_coding(self, src, pred, n, qp, qparams, scan, c_cost, is_cavlc, max_value=255):
        """One inter macroblock: (levels[256], cost8[4], cbp, cbp_blk, recon[16][16], sse)."""
        levels = np.zeros(256, np.int16); cost8 = np.zeros(4, np.int32); cbp = np.zeros(1, np.int32); cbpb = np.zeros(1, np.int32)
        recon = np.zeros((16, 16), np.uint16)
        sse = self.L.jmo_luma_residual_coding(np.ascontiguousarray(src, np.uint16), np.ascontiguousarray(pred, np.uint16), n, qp,
                                              np.ascontiguousarray(qparams, np.int32).reshape(-1), np.ascontiguousarray(scan, np.uint8).reshape(-1),
                                              np.ascontiguousarray(c_cost, np.uint8), int(is_cavlc), max_value, levels, cost8, cbp, cbpb, recon)
        return levels, cost8, int(cbp[0]), int(cbpb[0]), recon, int(sse)

    def hadamard4x4(self, d):
        return self.L.jmo_hadamard_sad4x4(np.ascontiguousarray(d, np.int16).reshape(-1))

    def hadamard8x8(self, d):
        return self.L.jmo_hadamard_sad8x8(np.ascontiguousarray(d, np.int16).reshape(-1))

    def quant(self, variant, coef, qp, qparams, scan, c_cost, is_cavlc, arw=0, cost0=0):
        return _quant_call(self.L.jmo_quant, None, variant, coef, qp, qparams, scan, c_cost, is_cavlc, arw, cost0)


def _qlist_call(fn, plan, coef_flat, cost0):
    """Run a list quantiser (oracle restatement) on the coefficients a JM function would see: gather in scan order,
    quantise, scatter back.  Returns the same dict as the JM-side call."""
    coef = np.ascontiguousarray(coef_flat, np.int32).reshape(-1).copy()
    order = np.asarray(plan["order"])
    lst = coef[order].copy()
    levels = np.zeros(17, np.int32); runs = np.zeros(17, np.int32); fadj = np.zeros(len(order), np.int32); cost = np.array([cost0], np.int32)
    nz = fn(len(order), plan["q_bits"], plan["qp_per"], plan["dequant"], plan["clip"], plan["use_cost"], plan["around"], plan["arw"],
            np.ascontiguousarray(plan["params"], np.int32).reshape(-1), np.ascontiguousarray(plan["c_cost"], np.uint8), lst, levels, runs, fadj, cost)
    coef[order] = lst
    fa = np.zeros(len(coef), np.int32); fa[order] = fadj
    return dict(nonzero=int(nz), coef=coef, levels=levels, runs=runs, fadjust=fa, coeff_cost=int(cost[0]))


def _quant_call(fn, handle, variant, coef, qp, qparams, scan, c_cost, is_cavlc, arw, cost0):
    n = 4 if variant < 2 else 8
    coef = np.ascontiguousarray(coef, np.int32).reshape(n * n).copy()
    nl = 68 if variant >= 4 else n * n + 1
    levels = np.zeros(nl, np.int32)
    runs = np.zeros(nl, np.int32)
    fadj = np.zeros(n * n, np.int32)
    cost = np.array([cost0], np.int32)
    args = [variant, coef, qp, np.ascontiguousarray(qparams, np.int32).reshape(-1),
            np.ascontiguousarray(scan, np.uint8).reshape(-1), np.ascontiguousarray(c_cost, np.uint8),
            int(is_cavlc), int(arw), levels, runs, fadj, cost]
    nz = fn(*( [handle] + args if handle is not None else args))
    return dict(nonzero=int(nz), coef=coef.reshape(n, n), levels=levels, runs=runs,
                fadjust=fadj.reshape(n, n), coeff_cost=int(cost[0]))


class JMRef:
    """The real JM leaf functions (oracle/_ref/libjmref.so)."""

    def __init__(self, w, h, search_range=32, metrics=(SAD, SATD, SATD), fast_full=0, rdopt=1,
                 bitdepth=8, vmv_qpel=2048):
        L = C.CDLL(REF_SO)
        self.L = L
        self.w, self.h, self.R = w, h, search_range
        L.jmref_open.restype = C.c_void_p
        L.jmref_open.argtypes = [C.c_int] * 10
        L.jmref_set_ref.argtypes = [C.c_void_p, _u16p, C.c_int]
        L.jmref_set_cur.argtypes = [C.c_void_p, _u16p, C.c_int]
        L.jmref_get_subplane.argtypes = [C.c_void_p, C.c_int, C.c_int, _u16p]
        L.jmref_full_search.restype = C.c_int64
        L.jmref_full_search.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_int64, _i16p]
        L.jmref_sub_pel.restype = C.c_int64
        L.jmref_sub_pel.argtypes = [C.c_void_p] + [C.c_int] * 7 + [_i32p, C.c_int64, C.c_int, _i16p]
        L.jmref_set_ref2.argtypes = [C.c_void_p, _u16p, C.c_int]
        L.jmref_dist_ex.restype = C.c_int64
        L.jmref_dist_ex.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_int64, _i32p]
        L.jmref_bipred_search.restype = C.c_int64
        L.jmref_bipred_search.argtypes = [C.c_void_p] + [C.c_int] * 5 + [_i32p] * 4 + [C.c_int, _i32p, C.c_int64, C.c_int, _i32p, _i16p]
        L.jmref_dist.restype = C.c_int64
        L.jmref_dist.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_int64]
        L.jmref_ffs_setup.argtypes = [C.c_void_p] + [C.c_int] * 4 + [_i16p]
        L.jmref_ffs_get_sad.argtypes = [C.c_void_p, C.c_int, C.c_int, _u32p, C.c_int]
        L.jmref_ffs_search.restype = C.c_int64
        L.jmref_ffs_search.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_int64, _i16p]
        L.jmref_spiral.argtypes = [C.c_void_p, _i16p, C.c_int]
        L.jmref_mvbits.argtypes = [C.c_void_p, C.c_int]
        L.jmref_forward4x4.argtypes = [_i32p]
        L.jmref_forward8x8.argtypes = [_i32p]
        L.jmref_quant_misc.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int, _i32p, _u8p, _u8p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p]
        L.jmref_hadamard.argtypes = [C.c_int, _i32p]
        L.jmref_inverse4x4.argtypes = [_i32p]
        L.jmref_inverse8x8.argtypes = [_i32p]
        L.jmref_hadamard_sad4x4.argtypes = [_i16p]
        L.jmref_hadamard_sad8x8.argtypes = [_i16p]
        L.jmref_quant.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int, _i32p, _u8p, _u8p, C.c_int, C.c_int,
                                  _i32p, _i32p, _i32p, _i32p]
        self.h_ = L.jmref_open(w, h, search_range, metrics[0], metrics[1], metrics[2], int(fast_full),
                               int(rdopt), bitdepth, vmv_qpel)

    def set_ref(self, luma):
        luma = np.ascontiguousarray(luma, np.uint16)
        self.L.jmref_set_ref(self.h_, luma, luma.shape[1])

    def set_ref2(self, luma):
        luma = np.ascontiguousarray(luma, np.uint16)
        self.L.jmref_set_ref2(self.h_, luma, luma.shape[1])

    def set_cur(self, luma):
        luma = np.ascontiguousarray(luma, np.uint16)
        self.L.jmref_set_cur(self.h_, luma, luma.shape[1])

    def planes(self):
        out = np.empty((4, 4, self.h + 2 * PAD_Y, self.w + 2 * PAD_X), np.uint16)
        for fy in range(4):
            for fx in range(4):
                self.L.jmref_get_subplane(self.h_, fy, fx, out[fy, fx])
        return out

    def spiral(self):
        n = max(9, (2 * self.R + 1) ** 2)
        out = np.empty((n, 2), np.int16)
        self.max_mvd = self.L.jmref_spiral(self.h_, out, n)
        return out

    def mvbits(self, v):
        return self.L.jmref_mvbits(self.h_, int(v))

    def full_search(self, blocktype, pos, pred, center, lam, min_mcost):
        mv = np.zeros(2, np.int16)
        c = self.L.jmref_full_search(self.h_, blocktype, pos[0], pos[1], pred[0], pred[1], center[0], center[1],
                                     lam, min_mcost, mv)
        return (int(mv[0]), int(mv[1])), c

    def sub_pel(self, blocktype, pos, pred, mv_in, lam3, min_mcost, test8x8=0):
        mv = np.zeros(2, np.int16)
        c = self.L.jmref_sub_pel(self.h_, blocktype, pos[0], pos[1], pred[0], pred[1], mv_in[0], mv_in[1],
                                 np.asarray(lam3, np.int32), min_mcost, test8x8, mv)
        return (int(mv[0]), int(mv[1])), c

    def dist(self, metric, blocktype, pos, cand, test8x8=0, min_mcost=DISTBLK_MAX):
        return self.L.jmref_dist(self.h_, metric, blocktype, pos[0], pos[1], cand[0], cand[1], test8x8, min_mcost)

    def dist_ex(self, metric, form, blocktype, pos, cand1, cand2, wp=(32, 32, 0, 5, 16), test8x8=0, min_mcost=DISTBLK_MAX):
        return self.L.jmref_dist_ex(self.h_, metric, form, blocktype, pos[0], pos[1], cand1[0], cand1[1], cand2[0], cand2[1],
                                    test8x8, min_mcost, np.asarray(wp, np.int32))

    def bipred_search(self, which, form, blocktype, pos, pred1, pred2, mv1, mv2, search_range_qpel, lam3, min_mcost=DISTBLK_MAX,
                      test8x8=0, wp=(32, 32, 0, 5, 16)):
        """The real full_search_bipred_motion_estimation (which=0) / sub_pel_bipred_motion_estimation (which=1)."""
        a = lambda v: np.asarray(v, np.int32)
        mv = np.zeros(2, np.int16)
        c = self.L.jmref_bipred_search(self.h_, which, form, blocktype, pos[0], pos[1], a(pred1), a(pred2), a(mv1), a(mv2),
                                       search_range_qpel, a(lam3), min_mcost, test8x8, a(wp), mv)
        return (int(mv[0]), int(mv[1])), c

    def ffs_setup(self, mb, pmv):
        c = np.zeros(2, np.int16)
        self.L.jmref_ffs_setup(self.h_, mb[0], mb[1], pmv[0], pmv[1], c)
        return int(c[0]), int(c[1])

    def ffs_sad(self, blocktype, block_index):
        n = (2 * self.R + 1) ** 2
        out = np.empty(n, np.uint32)
        self.L.jmref_ffs_get_sad(self.h_, blocktype, block_index, out, n)
        return out

    def ffs_search(self, blocktype, pos, pred, lam, min_mcost):
        mv = np.zeros(2, np.int16)
        c = self.L.jmref_ffs_search(self.h_, blocktype, pos[0], pos[1], pred[0], pred[1], lam, min_mcost, mv)
        return (int(mv[0]), int(mv[1])), c

    def forward4x4(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmref_forward4x4(b.reshape(-1))
        return b

    def forward8x8(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmref_forward8x8(b.reshape(-1))
        return b

    def quant_misc(self, variant, coef_flat, qp, qparams, scan, c_cost, is_cavlc, arw=0, cost0=0):
        """The real quant_ac4x4_* (6,7), quant_dc4x4_normal (8), quant_dc2x2_* (9,10), quant_dc4x2_* (11,12)."""
        coef = np.ascontiguousarray(coef_flat, np.int32).reshape(-1).copy()
        levels = np.zeros(17, np.int32); runs = np.zeros(17, np.int32); fadj = np.zeros(16, np.int32); cost = np.array([cost0], np.int32)
        nz = self.L.jmref_quant_misc(self.h_, variant, coef, qp, np.ascontiguousarray(qparams, np.int32).reshape(-1),
                                     np.ascontiguousarray(scan, np.uint8).reshape(-1), np.ascontiguousarray(c_cost, np.uint8),
                                     int(is_cavlc), int(arw), levels, runs, fadj, cost)
        return dict(nonzero=int(nz), coef=coef, levels=levels, runs=runs, fadjust=fadj[:len(coef)], coeff_cost=int(cost[0]))

    def hadamard(self, kind, vals):
        b = np.ascontiguousarray(vals, np.int32).reshape(-1).copy()
        self.L.jmref_hadamard(kind, b)
        return b

    def inverse4x4(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmref_inverse4x4(b.reshape(-1))
        return b

    def inverse8x8(self, blk):
        b = np.ascontiguousarray(blk, np.int32).copy()
        self.L.jmref_inverse8x8(b.reshape(-1))
        return b

    def hadamard4x4(self, d):
        return self.L.jmref_hadamard_sad4x4(np.ascontiguousarray(d, np.int16).reshape(-1).copy())

    def hadamard8x8(self, d):
        return self.L.jmref_hadamard_sad8x8(np.ascontiguousarray(d, np.int16).reshape(-1).copy())

    def quant(self, variant, coef, qp, qparams, scan, c_cost, is_cavlc, arw=0, cost0=0):
        return _quant_call(self.L.jmref_quant, self.h_, variant, coef, qp, qparams, scan, c_cost, is_cavlc, arw, cost0)


def jmref_run_mbs(ref, mb_xy, preds, lam3, qp, qparams, scan, c_cost, do_tq=True, want_levels=True):
    """CPU baseline leg: JM's own functions on the GPU step's per-macroblock work (see ref_harness.c)."""
    L = ref.L
    L.jmref_run_mbs.argtypes = [C.c_void_p, C.c_int, _i16p, _i16p, _i32p, C.c_int, _i32p, _u8p, _u8p, C.c_int,
                                _i16p, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), C.c_void_p,
                                np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
    mb_xy = np.ascontiguousarray(mb_xy, np.int16).reshape(-1, 2)
    n = len(mb_xy)
    preds = np.ascontiguousarray(preds, np.int16).reshape(n, 41, 2)
    mv = np.zeros((n, 41, 2), np.int16); cost = np.zeros((n, 41), np.int64)
    lev = np.zeros((n, 7, 256), np.int16) if (do_tq and want_levels) else None
    secs = np.zeros(2)
    L.jmref_run_mbs(ref.h_, n, mb_xy, preds, np.asarray(lam3, np.int32), qp,
                    np.ascontiguousarray(qparams, np.int32).reshape(-1), np.ascontiguousarray(scan, np.uint8).reshape(-1),
                    np.ascontiguousarray(c_cost, np.uint8), int(do_tq), mv, cost,
                    lev.ctypes.data if lev is not None else None, secs)
    return mv, cost, lev, secs


def mv_predictor(nb, ref_frame, mb_x, mb_y, bsx, bsy):
    """GetMotionVectorPredictorNormal (lcommon/src/mv_prediction.c:192-300) restated: nb = three (available, ref_idx, mv_x, mv_y)
    for the neighbours A (left), B (up), C (up-right, or D where get_neighbors substitutes it); pinned against the real function
    by tests/test_oracle_vs_ref.py::test_mv_predictor_matches_jm."""
    rf = [n[1] if n[0] else -1 for n in nb]
    mv = [(n[2], n[3]) if n[0] else (0, 0) for n in nb]
    same = [r == ref_frame for r in rf]
    kind = "median"
    if same == [True, False, False]: kind = 0
    elif same == [False, True, False]: kind = 1
    elif same == [False, False, True]: kind = 2
    if (bsx, bsy) == (8, 16):
        if mb_x == 0:
            if same[0]: kind = 0
        elif same[2]: kind = 2
    elif (bsx, bsy) == (16, 8):
        if mb_y == 0:
            if same[1]: kind = 1
        elif same[0]: kind = 0
    if kind == "median":
        if not (nb[1][0] or nb[2][0]):
            return mv[0]
        return tuple(sorted(c)[1] for c in zip(*mv))
    return mv[kind]


def jmref_mv_predictor(nb, ref_frame, mb_x, mb_y, bsx, bsy):
    """The same through JM's own function (oracle/_ref/libjmref.so)."""
    L = C.CDLL(REF_SO)
    L.jmref_mv_predictor.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i16p]
    out = np.zeros(2, np.int16)
    L.jmref_mv_predictor(np.ascontiguousarray(nb, np.int32).reshape(-1), ref_frame, mb_x, mb_y, bsx, bsy, out)
    return int(out[0]), int(out[1])


DB_MB = np.dtype([("mb_type", "u1"), ("flags", "u1"), ("qp", "i1"), ("qpc", "i1", (2,)), ("df_disable_idc", "i1"), ("df_alpha_c0_offset", "i1"),
                  ("df_beta_offset", "i1"), ("cbp_blk", "<u4"), ("pad_", "<u4"), ("mv", "<i2", (2, 16, 2)), ("ref_id", "i1", (2, 16))])
assert DB_MB.itemsize == 176
DB_T8X8, DB_CBP, DB_AVAIL_A, DB_AVAIL_B = 1, 2, 4, 8


def deblock(luma, cb, cr, yuv, slice_type, mbs, direct8x8inf=1):
    """jmo_deblock (oracle/jm_oracle.c): DeblockFrame restated; returns the filtered planes."""
    L = C.CDLL(ORACLE_SO)
    luma = np.ascontiguousarray(luma, np.uint8).copy(); h, w = luma.shape
    cb = np.ascontiguousarray(cb, np.uint8).copy() if yuv else None; cr = np.ascontiguousarray(cr, np.uint8).copy() if yuv else None
    mbs = np.ascontiguousarray(mbs, DB_MB)
    assert len(mbs) == (w // 16) * (h // 16)
    L.jmo_deblock.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.jmo_deblock(luma.ctypes.data, w, cb.ctypes.data if yuv else None, cr.ctypes.data if yuv else None, w // 2, w, h, yuv, slice_type, direct8x8inf, mbs.ctypes.data)
    return luma, cb, cr


def jmref_deblock(luma, cb, cr, yuv, slice_type, mbs, direct8x8inf=1):
    """The same through JM's own DeblockFrame (oracle/_ref/libjmref.so)."""
    L = C.CDLL(REF_SO)
    luma = np.ascontiguousarray(luma, np.uint8).copy(); h, w = luma.shape
    cb = np.ascontiguousarray(cb, np.uint8).copy() if yuv else None; cr = np.ascontiguousarray(cr, np.uint8).copy() if yuv else None
    mbs = np.ascontiguousarray(mbs, DB_MB)
    L.jmref_deblock.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.jmref_deblock(luma.ctypes.data, cb.ctypes.data if yuv else None, cr.ctypes.data if yuv else None, w, h, yuv, slice_type, direct8x8inf, mbs.ctypes.data)
    return luma, cb, cr


def random_deblock_picture(rng, w, h, yuv, slice_type, intra_share=0.15, idc=0):
    """A synthetic coded picture for the deblocking tests: blocky samples (so that edges are filtered), macroblock types, coded-block
    flags, motion and references of every kind the strength rules tell apart, QPs on both sides of the filter's on/off threshold."""
    mbw, mbh = w // 16, h // 16
    base = rng.integers(40, 200, (h // 4, w // 4)).astype(np.int32)
    luma = np.clip(np.kron(base, np.ones((4, 4), np.int32)) + rng.integers(-6, 7, (h, w)), 0, 255).astype(np.uint8)
    hc = h // 2 if yuv == 1 else h
    planes = []
    for _ in range(2):
        b = rng.integers(60, 190, ((hc + 3) // 4, w // 8)).astype(np.int32)
        planes.append(np.clip(np.kron(b, np.ones((4, 4), np.int32))[:hc] + rng.integers(-4, 5, (hc, w // 2)), 0, 255).astype(np.uint8))
    mbs = np.zeros(mbw * mbh, DB_MB)
    inter = [0, 1, 2, 3, 8] 
    for i, m in enumerate(mbs):
        is_intra = slice_type == 2 or rng.random() < intra_share
        m["mb_type"] = int(rng.choice([9, 10, 13, 14])) if is_intra else int(rng.choice(inter))
        t8 = (m["mb_type"] == 13) or (m["mb_type"] in (1, 2, 3, 8) and rng.random() < 0.3)
        coded = rng.random() < 0.6 and m["mb_type"] != 0
        m["cbp_blk"] = int(rng.integers(0, 1 << 16)) if coded else 0
        if coded and rng.random() < 0.2:
            m["cbp_blk"] = 0            # cbp != 0 may come from chroma alone
        m["flags"] = (DB_T8X8 if t8 else 0) | (DB_CBP if coded else 0) | (DB_AVAIL_A if (i % mbw and rng.random() < 0.8) else 0) | (DB_AVAIL_B if (i >= mbw and rng.random() < 0.8) else 0)
        q = int(rng.choice([12, 20, 26, 30, 36, 44, 51])); m["qp"] = 0 if m["mb_type"] == 14 else q
        m["qpc"] = (0, 0) if m["mb_type"] == 14 else (max(0, q - int(rng.integers(0, 6))), max(0, q - int(rng.integers(0, 6))))
        m["df_disable_idc"] = idc if rng.random() < 0.95 else 1
        m["df_alpha_c0_offset"] = int(rng.choice([0, 0, -4, 6])); m["df_beta_offset"] = int(rng.choice([0, 0, 2, -6]))
        # motion: per partition of the type, references from a small pool; B pictures use both lists
        t = int(m["mb_type"])
        part = {0: (4, 4), 1: (4, 4), 2: (4, 2), 3: (2, 4)}.get(t)
        for k in range(16):
            bx, by = k % 4, k // 4
            if part is None:
                key = (bx // int(rng.choice([1, 2])), by // int(rng.choice([1, 2])), int(rng.integers(0, 3)))
            else:
                key = (bx // part[0], by // part[1], 0)
            r2 = np.random.default_rng(hash((i, key, t)) & 0xffffffff)
            for l in range(2):
                use = (l == 0) if slice_type != 1 else (r2.random() < 0.7)
                if is_intra:
                    use = False
                m["ref_id"][l][k] = int(r2.integers(0, 3)) if use else -1
                m["mv"][l][k] = r2.integers(-9, 10, 2) if (use or r2.random() < 0.3) else (0, 0)
    return luma, planes[0], planes[1], mbs

#!/usr/bin/env python
"""bench.py -- macroblocks/sec of the JM lencod ME + transform/quant hot path on B200.

One "step" = the hot path over ONE P-picture (config 2, the default: 1080p coded 1920x1088 = 8160 macroblocks, 1 reference):
  K6    jmb_ref_put_u8              16 quarter-pel planes of the reference             (getSubImagesLuma)
  --    jmb_pic_begin_u8            current picture to the device
  K1-K3 jmb_me_search_frame_pred    requests generated on the device from the 41 predictors of every macroblock,
                                    full search +-32 of all 41 partitions             (full_search_motion_estimation)
  K5    (same call)                 half-/quarter-pel SATD refinement of every one    (sub_pel_motion_estimation)
  K7/K8 jmb_mc_tq_modes_compact     prediction -> residual -> forward4x4 -> quant for each of the 7 partition modes (what
                                    JM's RDO loop residual-codes per inter candidate), (level, run) tokens out, one launch
Predictors are synthetic (true motion + per-MB / per-partition jitter), lambda from QP 28.

  value : device-timed (CUDA events on the library's stream), inputs resident in HBM, rotating over
          4 distinct input sets (> L2 in total) so no step re-reads a warm L2.
  e2e   : the same step through the C ABI with pinned HOST buffers (H2D/D2H inside the timed region).
  --impl reference : JM's own functions (oracle/_ref/libjmref.so) on the host cores, bounded sample.
  --config {2,3,4,5} : BASELINE.json configs[1..4]; see CONFIGS below.  Default 2.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU,
weak scaling: every rank encodes its own pictures, i.e. independent closed-GOP segments; no collective
on the data path).  No measured leg ever sets JMB_SHIM / JMB_SHIM_OFF (debug switches of the drop-in shim).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEARCH_RANGE = 32
QP = 28
N_SETS = 4
BYTES_PER_MB_REF = 13804           # 512 src + 12800 window + 492 results
METRIC = {2: "macroblocks/sec 1080p full-search ME+DCT/quant", 3: "macroblocks/sec 4K EPZS ME + 8x8 DCT/quant",
          4: "macroblocks/sec 1080p 4:2:2 fast-full-search ME + SATD sub-pel + DCT/quant incl. chroma",
          5: "macroblocks/sec 4K full-search ME+DCT/quant, reconstructed anchor read over NVLink"}

# BASELINE.json configs[1..4] (configs[0] is the reference's own CPU-only plumbing case, tests/test_jm_dropin.py)
CONFIGS = {
    2: dict(name="1080p 4:2:0 synthetic, FullSearch +-32 (SearchMode=-1) 41 partitions/MB + SATD sub-pel + 4x4 transform/quant of 7 "
                 "partition modes, Baseline, 1 ref, QP28", w=1920, h=1088, size="1080p", search="full", n=4, chroma=None),
    3: dict(name="4K 4:2:0 synthetic (3840x2160 = 32400 macroblocks), EPZS integer + sub-pel search of 41 partitions/MB (encoder.cfg "
                 "pattern settings) + 8x8 transform/quant of the 4 partition modes that allow it, High, 1 ref, QP28",
            w=3840, h=2160, size="4K", search="epzs", n=8, chroma=None),
    4: dict(name="1080p 4:2:2 synthetic (encoder_yuv422.cfg metrics: SAD full-pel, SATD half-/quarter-pel), fast full search +-32 + "
                 "SATD sub-pel refinement + 4x4 transform/quant of 7 modes + 4:2:2 chroma prediction/residual/DC path, QP28",
            w=1920, h=1088, size="1080p", search="fastfull", n=4, chroma="422"),
    5: dict(name="4K 4:2:0 synthetic (3840x2160), FullSearch +-32 + SATD sub-pel + 4x4 transform/quant of 7 modes; every rank codes "
                 "a picture against rank 0's reconstructed anchor read over NVLink (peer-mapped, no host copy)",
            w=3840, h=2160, size="4K", search="full", n=4, chroma=None, anchor=True),
}


def workload_config(cfg_id, n_gpus):
    c = CONFIGS[cfg_id]
    par = (f"{n_gpus} x pictures sharing one anchor: each rank's sub-pel kernel reads rank 0's reconstructed luma over NVLink "
           f"({c['w'] * c['h'] / 1e6:.1f} MB u8 per picture per rank)" if c.get("anchor") and n_gpus > 1 else
           f"{n_gpus} x independent picture streams (closed-GOP shards), no data-path collective")
    return {"workload": c["name"], "baseline_config": cfg_id, "width": c["w"], "height": c["h"],
            "macroblocks_per_step": (c["w"] // 16) * (c["h"] // 16), "search_range": SEARCH_RANGE,
            "l2": f"rotating over {N_SETS} distinct input sets (> 126 MB in total)", "parallelism": par}


def make_chroma_planes(luma, yuv, seed):
    """Chroma planes with texture that moves with the luma: decimated luma + noise (4:2:0: w/2 x h/2; 4:2:2: w/2 x h)."""
    rng = np.random.default_rng(seed)
    ys = 2 if yuv == 1 else 1
    u = luma[::ys, 0::2].astype(np.int32) // 2 + 60 + rng.integers(-2, 3, (luma.shape[0] // ys, luma.shape[1] // 2))
    v = luma[::ys, 1::2].astype(np.int32) // 2 + 50 + rng.integers(-2, 3, (luma.shape[0] // ys, luma.shape[1] // 2))
    return np.clip(u, 0, 255).astype(np.uint8), np.clip(v, 0, 255).astype(np.uint8)


def make_shared_candidates(seed, n_mb, n_shared, motion_q=(20, 12)):
    """EPZS: candidate mvs every partition of a macroblock checks = zero mv + neighbour-like motion (true motion + jitter)."""
    rng = np.random.default_rng(seed + 7)
    c = (np.array(motion_q)[None, None, :] + rng.integers(-10, 11, size=(n_mb, n_shared, 2))).astype(np.int16)
    c[:, 0] = 0
    return c


def make_pred_table(api, seed, n_mb, motion_q=(20, 12)):
    """41 predictors per macroblock (canonical partition order) = true motion + per-MB and per-partition jitter."""
    rng = np.random.default_rng(seed)
    pred = np.zeros(n_mb, api.MB_MVPRED)
    mbpred = np.array(motion_q)[None, None, :] + rng.integers(-8, 9, size=(n_mb, 1, 2))
    pred["pred"] = mbpred + rng.integers(-3, 4, size=(n_mb, 41, 2))
    return pred


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 2.0:      # nvidia-smi needs ~100 ms to deliver its first sample
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_begin(self):
        self.i0 = len(self.rows)

    def mark_end(self):
        self.i1 = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        i0, i1 = getattr(self, "i0", 0), getattr(self, "i1", len(self.rows)) + 1      # samples taken DURING the timed region
        rows = self.rows[i0:i1] if i1 > i0 else self.rows
        nearest = False
        if not rows and self.rows:      # a timed region shorter than the 20 ms sampling period: the samples on either side of it
            rows, nearest = self.rows[max(0, i0 - 1):i1 + 1], True
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if nearest:
            out["note"] = "timed region shorter than the sampling period: nearest samples"
        return out


class Workload:
    """Everything one configuration needs: inputs (host pinned + device), outputs, and the per-picture call sequence."""

    def __init__(self, args, api, synth, T, ctx, local, rank, world, torch):
        self.cfg = CONFIGS[args.config]
        self.args = args
        self.api, self.ctx, self.torch, self.local = api, ctx, torch, local
        self.W, self.H = self.cfg["w"], self.cfg["h"]
        W, H = self.W, self.H
        self.n_mb = (W // 16) * (H // 16)
        self.lam = T.lambda_me(QP)
        n = self.cfg["n"]
        if n == 4:
            self.qd = api.quant_desc(4, QP, T.q_params(QP, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)
            self.mode_mask = 0x7F
        else:      # High profile, CABAC, 8x8 transform: partition modes 1..4
            self.qd = api.quant_desc(8, QP, T.q_params(QP, 0, 8), T.SNGL_SCAN8x8, T.COEFF_COST8x8[0], 0)
            self.mode_mask = 0x0F
        mode = api.SEARCH_FAST_FULL if self.cfg["search"] == "fastfull" else api.SEARCH_FULL
        self.fp = api.frame_params([self.lam] * 3, mode=mode, flags=api.REQ_SUBPEL | (api.REQ_TEST8X8 if n == 8 else 0))
        self.token_cap = 7 * self.n_mb * (256 if args.scene_cut else 96)      # 256 per (mode, macroblock) is the hard maximum
        self.epzs = self.cfg["search"] == "epzs"
        self.n_shared = 8
        if self.epzs:      # bin/encoder.cfg: EPZSPattern 2, EPZSDualRefinement 3, window predictors (EPZSFixedPredictors), thresholds 0/1/2/1
            self.efp = api.epzs_frame_params([self.lam] * 3, flags=api.EPZS_ADAPT_PATTERN | api.EPZS_DUAL | api.EPZS_SUBPEL | (api.EPZS_TEST8X8 if n == 8 else 0),
                                             pattern=api.EPZS_PAT_EDIAMOND, pattern_dual=api.EPZS_PAT_EDIAMOND, n_shared=self.n_shared, window=4,
                                             search_range=SEARCH_RANGE)
        dev = f"cuda:{local}"
        anchor = bool(self.cfg.get("anchor"))
        self.chroma = self.cfg["chroma"]                       # "422": the 4:2:2 chroma prediction / residual / 4x2 DC path rides along
        self.yuv = 2 if self.chroma == "422" else 1
        if self.chroma:
            self.cdesc = api.chroma_desc(self.yuv, QP, lambda q: T.q_params(q, 0, 4), T.COEFF_COST4x4[0], 0)
            self.hc = H if self.yuv == 2 else H // 2
        self.sets = []
        for s in range(N_SETS):
            # anchor mode: every rank codes a picture of the SAME sequence against rank 0's anchor
            f = synth.luma_frames(W, H, 2, seed=1234 + (0 if anchor else 97 * rank) + s, motion=(5, 3))
            if args.scene_cut:      # robustness probe: the reference is an unrelated picture (nothing matches)
                f[0] = synth.luma_frames(W, H, 1, seed=999 + s)[0]
            pred = make_pred_table(api, 50 + 13 * rank + s, self.n_mb)
            hs = {"ref": ctx.pinned((H, W), np.uint8), "cur": ctx.pinned((H, W), np.uint8), "pred": ctx.pinned(self.n_mb, api.MB_MVPRED)}
            hs["ref"][:] = f[0]; hs["cur"][:] = f[1]; hs["pred"][:] = pred
            if self.chroma:
                for nm, luma, sd in (("ref", f[0], 11), ("cur", f[1], 12)):
                    u, v = make_chroma_planes(luma, self.yuv, sd + s)
                    hs[nm + "_u"] = ctx.pinned(u.shape, np.uint8); hs[nm + "_v"] = ctx.pinned(v.shape, np.uint8)
                    hs[nm + "_u"][:] = u; hs[nm + "_v"][:] = v
            if self.epzs:      # per macroblock: the zero mv + what the spatial / co-located generators would yield (neighbours' motion)
                hs["shared"] = ctx.pinned((self.n_mb, self.n_shared, 2), np.int16)
                hs["shared"][:] = make_shared_candidates(50 + 13 * rank + s, self.n_mb, self.n_shared)
            ds = {k: torch.from_numpy(v.view(np.uint8).reshape(-1).copy()).to(dev) for k, v in hs.items()}
            self.sets.append((hs, ds))
        self.d_res8 = torch.empty(self.n_mb * api.NPART * 8, dtype=torch.uint8, device=dev)
        self.d_heads = torch.empty(7 * self.n_mb * 16, dtype=torch.uint8, device=dev)
        self.d_tokens = torch.empty(self.token_cap * 4, dtype=torch.uint8, device=dev)
        self.d_ntok = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.chroma:
            n2 = 7 * self.n_mb * 2
            self.d_cdc = torch.empty(n2 * 8, dtype=torch.int16, device=dev); self.d_cac = torch.empty(n2 * 120, dtype=torch.int16, device=dev)
            self.d_ccb = torch.empty(n2, dtype=torch.int32, device=dev); self.d_ccc = torch.empty(n2, dtype=torch.int32, device=dev)
        self.peer_refs = None
        self.tokens_seen = 0

    def host_outputs(self, c):
        o = [c.pinned(self.n_mb * self.api.NPART, self.api.ME_RES8), c.pinned((7, self.n_mb), self.api.TQ_HEAD),
             c.pinned(self.token_cap, self.api.TQ_TOKEN), np.zeros(1, np.uint32)]
        if self.chroma:
            o.append(dict(dc=c.pinned((7, self.n_mb, 2, 8), np.int16), ac=c.pinned((7, self.n_mb, 2, 8, 15), np.int16),
                          cb=c.pinned((7, self.n_mb, 2), np.uint32), cc=c.pinned((7, self.n_mb, 2), np.uint32)))
        return tuple(o)

    def setup_anchor(self, dist, rank, world):
        """config 5: rank 0 owns the reconstructed anchors in exportable device memory; every other rank maps them (IPC ->
        NVLink peer access) once.  torch.distributed only carries the 64-byte handles."""
        api, ctx = self.api, self.ctx
        handles = [None] * N_SETS
        if rank == 0:
            self.anchor_bufs = []
            for s in range(N_SETS):
                p = ctx.dev_alloc(self.W * self.H)
                ctx.dev_copy(p, self.sets[s][0]["ref"], self.W * self.H, api.DEVICE, api.HOST)
                self.anchor_bufs.append(p)
                handles[s] = ctx.peer_export(p).tobytes()
        if world > 1:
            dist.broadcast_object_list(handles, src=0)
        if rank == 0:
            self.peer_refs = self.anchor_bufs
        else:
            self.peer_refs = [ctx.peer_open(np.frombuffer(hd, np.uint8)) for hd in handles]

    def step_device(self, s):
        api, ctx = self.api, self.ctx
        hs, ds = self.sets[s % N_SETS]
        shape = (self.H, self.W)
        ref_ptr = self.peer_refs[s % N_SETS] if self.peer_refs else ds["ref"].data_ptr()
        ctx.ref_put_u8(s % 2, ref_ptr, api.DEVICE, shape=shape)
        ctx.pic_begin_u8(ds["cur"].data_ptr(), [s % 2], api.DEVICE, shape=shape)
        if self.chroma:
            cshape = (self.hc, self.W // 2)
            ctx.ref_put_chroma(s % 2, ds["ref_u"].data_ptr(), ds["ref_v"].data_ptr(), api.DEVICE, shape=cshape)
            ctx.pic_chroma(ds["cur_u"].data_ptr(), ds["cur_v"].data_ptr(), api.DEVICE, shape=cshape)
        if self.epzs:
            ctx.epzs_search_frame(ds["pred"].data_ptr(), ds["shared"].data_ptr(), self.efp, self.d_res8.data_ptr(), api.DEVICE, n_mb=self.n_mb)
        else:
            ctx.me_search_frame_pred(ds["pred"].data_ptr(), self.fp, self.d_res8.data_ptr(), api.DEVICE, n_mb=self.n_mb)
        ctx.mc_tq_modes_compact(None, self.qd, self.mode_mask, api.DEVICE, n_mb=self.n_mb,
                                out=(self.d_heads.data_ptr(), self.d_tokens.data_ptr(), self.d_ntok.data_ptr()), token_cap=self.token_cap)
        if self.chroma:      # chroma prediction + residual coding of every partition mode's motion (what the RD loop codes per candidate)
            n2 = self.n_mb * 2
            for m in range(7):
                ctx.chroma_residual_coding(self.cdesc, None, m + 1, 0, self.n_mb, api.DEVICE,
                                           out=(self.d_cdc.data_ptr() + m * n2 * 16, self.d_cac.data_ptr() + m * n2 * 240,
                                                self.d_ccb.data_ptr() + m * n2 * 4, self.d_ccc.data_ptr() + m * n2 * 4, None))

    def step_host(self, s, c, outs, tt):
        """The same picture through the C ABI with pinned HOST buffers: the uploads and the search are only enqueued
        (JMB_HOST_ASYNC); the residual-coding call returns with heads and tokens in the caller's buffers."""
        api = self.api
        hs, _ = self.sets[s % N_SETS]
        o_res, o_heads, o_tok, o_n = outs[:4]
        t0 = time.perf_counter()
        c.ref_put_u8(s % 2, hs["ref"], api.HOST_ASYNC)
        c.pic_begin_u8(hs["cur"], [s % 2], api.HOST_ASYNC)
        if self.chroma:
            c.ref_put_chroma(s % 2, hs["ref_u"], hs["ref_v"], api.HOST_ASYNC)
            c.pic_chroma(hs["cur_u"], hs["cur_v"], api.HOST_ASYNC)
        if self.epzs:
            c.epzs_search_frame(hs["pred"], hs["shared"], self.efp, o_res, api.HOST_ASYNC)
        else:
            c.me_search_frame_pred(hs["pred"], self.fp, o_res, api.HOST_ASYNC)
        if self.chroma:      # results of all 7 modes come back with the copies queued behind the kernels; the call below waits for everything
            oc = outs[4]
            for m in range(7):
                c._ck(c.L.jmb_chroma_residual_coding(c.h, None, m + 1, 0, self.n_mb, self.cdesc.ctypes.data, oc["dc"][m].ctypes.data, oc["ac"][m].ctypes.data,
                                                     oc["cb"][m].ctypes.data, oc["cc"][m].ctypes.data, None, api.HOST_ASYNC))
        t1 = time.perf_counter()
        c._ck(c.L.jmb_mc_tq_modes_compact(c.h, None, self.n_mb, self.mode_mask, self.qd.ctypes.data, o_heads.ctypes.data, o_tok.ctypes.data,
                                          self.token_cap, o_n.ctypes.data, api.HOST))
        tt["enqueue"] += t1 - t0; tt["residual_coding_and_wait"] += time.perf_counter() - t1
        self.tokens_seen = int(o_n[0])

    def epzs_algorithmic_bytes(self):
        api, ctx = self.api, self.ctx
        hs = self.sets[0][0]
        idx = np.linspace(0, self.n_mb - 1, 128).astype(int)
        ctx.ref_put_u8(0, np.array(hs["ref"])); ctx.pic_begin_u8(np.array(hs["cur"]), [0])
        reqs = api.epzs_requests_from_frame(np.array(hs["pred"])[idx], self.efp, self.W // 16, mb_index=idx)
        res = ctx.epzs_search(reqs, np.array(hs["shared"])[idx].reshape(-1, 2))
        bs = np.array([api.BLOCK_SIZE[t][0] * api.BLOCK_SIZE[t][1] for t in range(8)])[reqs["blocktype"]]
        per_mb = float((bs * (1 + res["n_evals"])).sum()) / len(idx) + api.MB_MVPRED.itemsize + self.n_shared * 4 + 41 * 8
        return per_mb * self.n_mb, float(res["n_evals"].mean())

    def bytes_per_step(self):
        api = self.api
        h2d = 2 * self.W * self.H + self.n_mb * api.MB_MVPRED.itemsize + api.FRAME_PARAMS.itemsize + api.QUANT_DESC.itemsize
        if self.epzs:
            h2d += self.n_mb * self.n_shared * 4 + api.EPZS_FRAME_PARAMS.itemsize - api.FRAME_PARAMS.itemsize
        d2h = self.n_mb * api.NPART * api.ME_RES8.itemsize + 7 * self.n_mb * api.TQ_HEAD.itemsize + 4 + 4 * self.tokens_seen
        if self.chroma:
            h2d += 4 * (self.W // 2) * self.hc + api.CHROMA_DESC.itemsize
            d2h += 7 * self.n_mb * 2 * (16 + 240 + 4 + 4)
        return int(h2d), int(d2h)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from jm_b200 import api, synth
    from jm_b200 import h264_tables as T

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.configure(search_range=SEARCH_RANGE)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    wl = Workload(args, api, synth, T, ctx, local, rank, world, torch)
    n_mb = wl.n_mb
    if wl.cfg.get("anchor"):
        wl.setup_anchor(dist, rank, world)
    torch.cuda.synchronize()

    # ---- end-to-end leg: K independent picture streams per GPU (K host threads, each its own context = its own CUDA stream,
    # reference slots and staging), pictures being independent units (closed-GOP shards) exactly like the ranks; the
    # single-stream figure (one thread, one context, strictly serial pictures) is reported beside it.
    n_streams = args.e2e_streams if args.e2e_streams > 0 else max(1, min(3, (os.cpu_count() or 1) // world))      # one host thread per picture stream
    e2e_ctx = [ctx] + [api.Context(local) for _ in range(n_streams - 1)]
    for c in e2e_ctx[1:]:
        c.configure(search_range=SEARCH_RANGE)
    e2e_out = [wl.host_outputs(c) for c in e2e_ctx]
    e2e_t = [{"enqueue": 0.0, "residual_coding_and_wait": 0.0} for _ in e2e_ctx]

    def run_host_steps(first, count, streams):
        """`count` pictures starting at step index `first`, dealt round-robin to the streams' threads."""
        def worker(k):
            for s in range(first + k, first + count, streams):
                wl.step_host(s, e2e_ctx[k], e2e_out[k], e2e_t[k])
            e2e_ctx[k].sync()
        if streams == 1:
            worker(0)
            return
        th = [threading.Thread(target=worker, args=(k,)) for k in range(streams)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    sampler = ClockSampler(local); sampler.start()       # started early: nvidia-smi needs ~100 ms to deliver its first sample
    for s in range(args.warmup):
        wl.step_device(s)
    ctx.sync()
    barrier()
    ctx.timing(True)
    sampler.mark_begin()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        wl.step_device(args.warmup + s)
    e1.record(stream)
    e1.synchronize()
    sampler.mark_end()
    barrier()
    clocks = sampler.stop()
    gpu_launches = ctx.launches - launches0
    ms = e0.elapsed_time(e1)
    top = "epzs" if wl.epzs else "int_search"
    k_ms, k_n = ctx.timing_get(top)
    kernel_break = {k: ctx.timing_get(k)[0] / max(1, args.steps) for k in ("subpel_planes", "pack_cur", "gen_requests", "int_search", "subpel_refine", "epzs", "mc_tq", "chroma")}
    ctx.timing(False)
    ctx.sync()
    tokens_dev = int(wl.d_ntok.item())
    worst_ms = worst_refine_ms = None
    if not args.scene_cut and not args.no_worst and not wl.epzs:
        # worst case of the search gate: the reference is an unrelated picture (nothing matches, every bound stays loose)
        f = synth.luma_frames(wl.W, wl.H, 1, seed=999)[0].astype(np.uint8)
        d_cut = torch.from_numpy(f.reshape(-1).copy()).to(f"cuda:{local}")
        hs, ds = wl.sets[0]
        for it in range(4):
            if it == 1:
                ctx.timing(True)
            ctx.ref_put_u8(0, d_cut.data_ptr(), api.DEVICE, shape=(wl.H, wl.W))
            ctx.pic_begin_u8(ds["cur"].data_ptr(), [0], api.DEVICE, shape=(wl.H, wl.W))
            ctx.me_search_frame_pred(ds["pred"].data_ptr(), wl.fp, wl.d_res8.data_ptr(), api.DEVICE, n_mb=n_mb)
        w_ms, w_n = ctx.timing_get("int_search")
        worst_ms = w_ms / max(1, w_n)
        r_ms, r_n = ctx.timing_get("subpel_refine")      # nothing matches: every partition ends on its own mv, no sub-block is shared
        worst_refine_ms = r_ms / max(1, r_n)
        ctx.timing(False)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())

    # ---- end-to-end timing (host buffers through the C ABI) ----------------------------------------
    def timed_host(streams, steps):
        run_host_steps(0, max(3, streams) * 2, streams)      # warm-up (buffers grown, pages touched)
        barrier()
        for tt in e2e_t:
            for k in tt:
                tt[k] = 0.0
        t0 = time.perf_counter()
        run_host_steps(args.warmup, steps, streams)
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); el = float(t.item())
        return el
    e2e_s = timed_host(n_streams, args.steps)
    host_ms = {k: 1e3 * sum(tt[k] for tt in e2e_t) / args.steps for k in e2e_t[0]}
    e2e_1 = timed_host(1, args.steps) if n_streams > 1 else e2e_s
    h2d, d2h = wl.bytes_per_step()

    out = {"metric": METRIC[args.config], "value": world * n_mb * args.steps / (ms / 1e3),
           "unit": "macroblocks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(args.config, world), "clocks": clocks, "gpu_launches": int(gpu_launches),
           "e2e": {"value": world * n_mb * args.steps / e2e_s, "unit": "macroblocks/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                   "picture_streams_per_gpu": n_streams, "host_ms_per_picture": host_ms,
                   "single_stream_value": world * n_mb * args.steps / e2e_1, "single_stream_ms_per_step": 1e3 * e2e_1 / args.steps},
           "kernel_ms_per_step": kernel_break, "tokens_per_step": tokens_dev}
    if wl.cfg.get("anchor") and world > 1:
        out["nvlink_bytes_per_step_per_rank"] = wl.W * wl.H
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launch_ms = k_ms / max(1, k_n)
    if wl.epzs:      # the search is two launches (integer stage, sub-pel stage): the roofline line is about both together
        launch_ms = kernel_break["epzs"] + kernel_break["subpel_refine"]
    if wl.epzs:
        # k_epzs: algorithmic bytes = per search the source block once + one reference block per distortion evaluated (u8 samples)
        # + request tables in / results out; the evaluation counts come from the kernel's own n_evals on a sample of set 0
        alg_bytes, evals_per_search = wl.epzs_algorithmic_bytes()
        achieved = alg_bytes / (launch_ms / 1e3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "k_epzs_int + k_epzs_sub", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": None, "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_launch": alg_bytes,
                           "launch_ms": launch_ms, "distortions_per_search": evals_per_search,
                           "note": "EPZS evaluates ~10-60 scattered block distortions per search, each step decided by the one before: the "
                                   "kernels are bound by instruction issue and the barriers between a macroblock's rounds, not by HBM "
                                   "bandwidth -- DESIGN.md 3"}
    else:
        achieved = BYTES_PER_MB_REF * n_mb / (launch_ms / 1e3) / 1e9
        traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (tools/ncu_summary.py)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "int_search_traffic.json")))[wl.cfg["size"].lower()]["dram_bytes_per_launch"]
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": "k_int_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                           "algorithmic_bytes_per_launch": BYTES_PER_MB_REF * n_mb, "launch_ms": launch_ms,
                           "worst_case_launch_ms": worst_ms, "worst_case_subpel_refine_ms": worst_refine_ms,
                           "note": "search-window model of SURVEY 8(d): 13804 B per macroblock*reference; the kernel is ALU-pipe "
                                   "(VABSDIFF4/PRMT/ISETP) bound, not HBM bound -- DESIGN.md 3; traffic < algorithmic bytes because "
                                   "neighbouring windows hit L2"}
    # the streaming kernels against the same HBM peak (algorithmic bytes per launch stated in DESIGN.md 3)
    sp_ms = kernel_break["subpel_planes"]; tq_ms = kernel_break["mc_tq"]
    sp_bytes = wl.W * wl.H + 16 * ((wl.W + 64 + 127) // 128 * 128) * (wl.H + 40)
    nmodes = bin(wl.mode_mask).count("1")
    tq_bytes = nmodes * n_mb * (256 + 256 + 16) + n_mb * 41 * 24 + 4 * tokens_dev
    out["roofline_other"] = {
        "k_subpel_planes": {"algorithmic_bytes_per_launch": sp_bytes, "launch_ms": sp_ms, "achieved": sp_bytes / (sp_ms / 1e3) / 1e9 if sp_ms else None,
                            "frac": sp_bytes / (sp_ms / 1e3) / 1e9 / peak if sp_ms else None},
        "k_mc_tq_modes_c": {"algorithmic_bytes_per_launch": tq_bytes, "launch_ms": tq_ms, "achieved": tq_bytes / (tq_ms / 1e3) / 1e9 if tq_ms else None,
                            "frac": tq_bytes / (tq_ms / 1e3) / 1e9 / peak if tq_ms else None}}
    if wl.chroma:      # k_chroma_rc, 7 launches per step: per macroblock and mode 2 x (source + reference samples) in, dc / ac / flags out
        c_ms = kernel_break["chroma"]
        c_bytes = 7 * n_mb * 2 * (2 * 8 * (16 if wl.yuv == 2 else 8) + 16 + 240 + 8)
        out["roofline_other"]["k_chroma_rc"] = {"algorithmic_bytes_per_step": c_bytes, "ms_per_step": c_ms, "launches_per_step": 7,
                                                "achieved": c_bytes / (c_ms / 1e3) / 1e9 if c_ms else None, "frac": c_bytes / (c_ms / 1e3) / 1e9 / peak if c_ms else None}

    if not wl.epzs:
        # The same kernel against the bound that actually limits it: the SM ALU pipe (VABSDIFF4 / PRMT / ISETP issue at
        # 64 lanes/clk/SM on this part, tools/ubench.cu).  Floor per displacement = 64 VABSDIFF4 + 19 PRMT + 41 ISETP.
        alu_ops = n_mb * (2 * SEARCH_RANGE + 1) ** 2 * (64 + 19 + 41)
        sm_hz = 1e6 * float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        alu_peak = 64.0 * 148 * sm_hz
        out["alu_roofline"] = {"bound": "alu-pipe", "achieved": alu_ops / (launch_ms / 1e3), "peak": alu_peak, "unit": "lane-ops/s",
                               "frac": alu_ops / (launch_ms / 1e3) / alu_peak,
                               "note": "minimum ALU-pipe instructions of the algorithm / launch time, vs 64 lanes/clk/SM x 148 SMs x SM clock"}
    if rank == 0 and world == 1 and args.config == 2:
        out["next_rows"] = {"deblock": deblock_line(ctx, api, torch, wl.W, wl.H, local, with_cpu=not args.no_cpu)}
        try:      # the re-linked encoder against stock JM on this configuration (tools/dropin_1080p.py, measured on a B200 box, committed)
            enc = json.load(open(os.path.join(ROOT, "profiles", "r02_dropin_1080p.json")))["1080p"]
            out["encoder"] = {"source": "profiles/r02_dropin_1080p.json (tools/dropin_1080p.py): stock lencod vs the same JM objects re-linked with libjmb200 "
                                        "(ME on the device, run-ahead chains), 1080p FullSearch +-32, 1 reference",
                              "frames": enc["frames"], "stock_wall_s": enc["stock"]["wall_s"], "dropin_wall_s": enc["dropin_me"]["wall_s"],
                              "stock_me_time": enc["stock"]["total_me_time"], "dropin_me_time": enc["dropin_me"]["total_me_time"],
                              "bitstream_identical": enc["dropin_me"]["bitstream_identical"]}
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(wl, ctx=ctx, api=api, budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def deblock_line(ctx, api, torch, w, h, local, with_cpu=True, iters=20):
    """The deblocking row measured on its own: DeblockFrame of one synthetic coded picture, device-resident planes and
    macroblock records, CUDA events; the CPU restatement of the same picture beside it (and its result compared)."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(77)
    luma, cb, cr, mbs = po.random_deblock_picture(rng, w, h, 1, 0)
    dev = f"cuda:{local}"
    pitch, pitch_c = (w + 127) // 128 * 128, (w // 2 + 127) // 128 * 128
    def padded(a, p):
        o = np.zeros((a.shape[0], p), np.uint8); o[:, :a.shape[1]] = a
        return o
    src = [torch.from_numpy(padded(a, p)).to(dev) for a, p in ((luma, pitch), (cb, pitch_c), (cr, pitch_c))]
    d = [t.clone() for t in src]
    d_mbs = torch.from_numpy(mbs.view(np.uint8).copy()).to(dev)
    ctx.timing(True)
    for it in range(iters + 2):
        for a, b in zip(d, src):
            a.copy_(b)
        torch.cuda.synchronize()
        if it == 2:
            ctx.timing(True)
        ctx.deblock_picture_dev(d[0].data_ptr(), pitch, d[1].data_ptr(), d[2].data_ptr(), pitch_c, w, h, 1, 0, d_mbs.data_ptr())
        ctx.sync()
    ms, n = ctx.timing_get("deblock")
    ctx.timing(False)
    n_mb = (w // 16) * (h // 16)
    res = {"kernel": "k_deblock", "picture": f"{w}x{h} 4:2:0, P slice, synthetic coded picture", "gpu_ms_per_picture": ms / max(1, n),
           "macroblocks_per_s": n_mb / (ms / max(1, n) / 1e3), "wavefront_steps": w // 16 + 2 * (h // 16 - 1),
           "algorithmic_bytes": 2 * (w * h * 3 // 2) + n_mb * 176}
    if with_cpu:
        t0 = time.perf_counter(); want = po.deblock(luma, cb, cr, 1, 0, mbs); res["cpu_port_ms_per_picture"] = 1e3 * (time.perf_counter() - t0)
        got = [t.cpu().numpy()[:, :a.shape[1]] for t, a in zip(d, (luma, cb, cr))]
        res["gpu_matches_cpu"] = bool(all(np.array_equal(g, wv) for g, wv in zip(got, want)))
    return res


_JM = {}


def _cpu_init(cfg_id, ref_luma, cur_luma, w, h, chroma=None):
    """One process = one instance of the CPU implementation; its quarter-pel planes are built once, untimed.
    Full search / fast full search: JM's own functions (oracle/_ref/libjmref.so, JM is single-threaded).  EPZS: the CPU
    restatement pinned to JM's recorded calls (oracle/jm_oracle.c::jmo_epzs; the real function needs the whole encoder state).
    chroma = (ref_u, ref_v, cur_u, cur_v) for config 4."""
    from oracle import pyoracle as po
    _JM["cfg"] = cfg_id
    cfg = CONFIGS[cfg_id]
    if cfg["search"] == "epzs" or cfg["chroma"]:
        o = po.Oracle()
        _JM["oracle"] = o
        _JM["oref"] = o.ref_create(ref_luma)
        _JM["cur"] = np.ascontiguousarray(cur_luma, np.uint16)
    if cfg["search"] != "epzs":
        ref = po.JMRef(w, h, SEARCH_RANGE, fast_full=int(cfg["search"] == "fastfull"))
        ref.set_ref(ref_luma); ref.set_cur(cur_luma)
        _JM["ref"] = ref
    _JM["chroma"] = chroma
    _JM["w"] = w


def _cpu_worker(a):
    """(macroblock addresses, their predictor rows, their shared-candidate rows or None, lambda) -> mv, cost, levels, (ME s, TQ s)
    [+ chroma results for config 4]"""
    idx, preds, shared, lam = a
    from jm_b200 import api
    from jm_b200 import h264_tables as T
    from oracle import pyoracle as po
    cfg = CONFIGS[_JM["cfg"]]
    mbw = _JM["w"] // 16
    mb_xy = np.stack([(idx % mbw) * 16, (idx // mbw) * 16], 1)
    n = cfg["n"]
    if cfg["search"] == "full":
        return po.jmref_run_mbs(_JM["ref"], mb_xy, preds, [lam] * 3, QP, T.q_params(QP, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0])
    o, r, cur = _JM["oracle"], _JM["oref"], _JM["cur"]
    if cfg["search"] == "epzs":
        efp = api.epzs_frame_params([lam] * 3, flags=api.EPZS_ADAPT_PATTERN | api.EPZS_DUAL | api.EPZS_SUBPEL | (api.EPZS_TEST8X8 if n == 8 else 0),
                                    pattern=api.EPZS_PAT_EDIAMOND, pattern_dual=api.EPZS_PAT_EDIAMOND, n_shared=shared.shape[1], window=4,
                                    search_range=SEARCH_RANGE)
        pr = np.zeros(len(idx), api.MB_MVPRED); pr["pred"] = preds
        reqs = api.epzs_requests_from_frame(pr, efp, mbw, mb_index=idx)
        t0 = time.perf_counter()
        res = o.epzs_batch(r, cur, reqs, shared.reshape(-1, 2), (po.SATD, po.SATD, 0, 1, 9))
        t1 = time.perf_counter()
        mv = np.stack([res["mv_x"], res["mv_y"]], 1).reshape(len(idx), 41, 2)
        cost = res["cost"].reshape(len(idx), 41)
    else:      # fast full search: JM's own setup_fast_full_search + fast_full_search_motion_estimation + sub_pel_motion_estimation
        jm = _JM["ref"]
        parts = api.mb_partitions()
        mv = np.zeros((len(idx), 41, 2), np.int16); cost = np.zeros((len(idx), 41), np.int64)
        t0 = time.perf_counter()
        for i in range(len(idx)):
            mb = (int(mb_xy[i, 0]), int(mb_xy[i, 1]))
            jm.ffs_setup(mb, (int(preds[i, 0, 0]), int(preds[i, 0, 1])))
            for k, (t, x, y) in enumerate(parts):
                p = (int(preds[i, k, 0]), int(preds[i, k, 1])); pos = (mb[0] + x, mb[1] + y)
                imv, _ = jm.ffs_search(t, pos, p, lam, po.DISTBLK_MAX)
                m2, c2 = jm.sub_pel(t, pos, p, imv, [lam] * 3, po.DISTBLK_MAX)
                mv[i, k] = m2; cost[i, k] = c2
        t1 = time.perf_counter()
    scan, cc = (T.SNGL_SCAN, T.COEFF_COST4x4[0]) if n == 4 else (T.SNGL_SCAN8x8, T.COEFF_COST8x8[0])
    qp_ = T.q_params(QP, 0, n)
    lev = np.zeros((len(idx), 7, 256), np.int16)
    mask = 0x7F if n == 4 else 0x0F
    for i in range(len(idx)):
        lev[i] = o.mc_tq_modes_mb(r, cur, (int(mb_xy[i, 0]), int(mb_xy[i, 1])), mv[i], n, QP, qp_, scan, cc, n == 4, mask)
    chroma_out = None
    if cfg["chroma"]:      # 4:2:2 chroma prediction + residual coding of each mode's motion (the pinned restatement)
        yuv = 2
        ru, rv, cu, cv = _JM["chroma"]
        d = api.chroma_desc(yuv, QP, lambda q: T.q_params(q, 0, 4), T.COEFF_COST4x4[0], 0)
        base = [0, 0, 1, 3, 5, 9, 17, 25]; w4 = [4, 4, 4, 2, 2, 2, 1, 1]; h4 = [4, 4, 2, 4, 2, 1, 2, 1]
        chroma_out = dict(dc=np.zeros((len(idx), 7, 2, 8), np.int16), cb=np.zeros((len(idx), 7, 2), np.uint32), cc=np.zeros((len(idx), 7, 2), np.uint32))
        for i in range(len(idx)):
            cx, cy = int(mb_xy[i, 0]) // 2, int(mb_xy[i, 1])
            for m in range(1, 8):
                mv16 = np.array([mv[i, base[m] + (by // h4[m]) * (4 // w4[m]) + bx // w4[m]] for by in range(4) for bx in range(4)], np.int16)
                for uv, (rp, cp) in enumerate(((ru, cu), (rv, cv))):
                    p = o.chroma_pred(rp, yuv, (cx, cy), mv16)
                    c = o.chroma_rc(cp[cy:cy + 16, cx:cx + 8], p, yuv, int(d["qp_ac"][0, uv]), int(d["qp_dc"][0, uv]), d["params_ac"][0, uv],
                                    d["params_dc"][0, uv], T.COEFF_COST4x4[0], 0)
                    chroma_out["dc"][i, m - 1, uv] = c["dc"]; chroma_out["cb"][i, m - 1, uv] = c["cbp_blk"]; chroma_out["cc"][i, m - 1, uv] = c["cr_cbp"]
    t2 = time.perf_counter()
    if chroma_out is not None:
        return mv, cost, lev, np.array([t1 - t0, t2 - t1]), chroma_out
    return mv, cost, lev, np.array([t1 - t0, t2 - t1])


def expand_tokens(heads, tokens, mbs, per=16):
    """(level, run) tokens of the macroblocks `mbs` -> dense [len(mbs)][7][256] levels (layout of JM's per-block scan order)."""
    out = np.zeros((len(mbs), 7, 256), np.int16)
    for i, mb in enumerate(mbs):
        for m in range(heads.shape[0]):
            hd = heads[m, mb]
            pos = {}
            for tk in tokens[int(hd["token_off"]): int(hd["token_off"]) + int(hd["n_tokens"])]:
                blk = int(tk["blk"]); p = pos.get(blk, 0) + int(tk["run"])
                out[i, m, blk * per + p] = tk["level"]
                pos[blk] = p + 1
    return out


def cpu_baseline(wl, ctx=None, api=None, budget_s=12.0):
    """The CPU implementation of the step on 1 host core over a bounded sample of set 0's macroblocks (kind 'reference' = JM's
    own functions; 'port' = the pinned restatement, EPZS only); also re-checks the GPU results of that sample -- motion
    vectors, costs AND quantised levels -- bit-for-bit."""
    from oracle import pyoracle as po
    if not wl.epzs and not po.ref_available():
        return {"value": None, "unit": "macroblocks/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libjmref.so missing"}
    hs = wl.sets[0][0]
    n_mb, W, H = wl.n_mb, wl.W, wl.H
    pred = np.array(hs["pred"])["pred"]
    shared = np.array(hs["shared"]) if wl.epzs else None
    # calibrate on 32 MBs, then size the sample for ~budget_s
    idx = np.linspace(0, n_mb - 1, 32).astype(int)
    chroma = tuple(np.array(hs[k]) for k in ("ref_u", "ref_v", "cur_u", "cur_v")) if wl.chroma else None
    _cpu_init(wl.args.config, np.array(hs["ref"]).astype(np.uint16), np.array(hs["cur"]).astype(np.uint16), W, H, chroma)
    args = lambda ii: (ii, pred[ii], shared[ii] if shared is not None else None, wl.lam)
    secs = _cpu_worker(args(idx))[3]
    per_mb = max(1e-5, float(secs.sum()) / len(idx))
    n = int(min(n_mb, max(64, budget_s / per_mb)))
    idx = np.linspace(0, n_mb - 1, n).astype(int)
    out_cpu = _cpu_worker(args(idx))
    mv, cost, lev, secs = out_cpu[:4]
    what = ("the CPU restatement of EPZS_integer_motion_estimation + EPZS_sub_pel_motion_estimation x41 (pinned to recorded calls of the real "
            "functions, tests/test_epzs_golden.py) + forward8x8/quant_8x8_normal x16 per MB" if wl.epzs else
            "JM setup_fast_full_search + fast_full_search_motion_estimation x41 + sub_pel_motion_estimation x41 (the real functions), "
            "4x4 transform/quant x112 and the 4:2:2 chroma prediction/residual coding x7 per MB through the pinned restatement" if wl.chroma else
            "JM full_search_motion_estimation + sub_pel_motion_estimation x41 + forward4x4/quant_4x4_normal x112 per MB")
    res = {"value": n / float(secs.sum()), "unit": "macroblocks/s", "cores": 1, "kind": "port" if wl.epzs else "reference",
           "sample": f"{n} of {n_mb} macroblocks of input set 0 (evenly spaced), {what}",
           "me_seconds": float(secs[0]), "tq_seconds": float(secs[1])}
    if ctx is not None:
        ctx.ref_put_u8(0, np.array(hs["ref"])); ctx.pic_begin_u8(np.array(hs["cur"]), [0])
        if wl.chroma:
            ctx.ref_put_chroma(0, np.array(hs["ref_u"]), np.array(hs["ref_v"])); ctx.pic_chroma(np.array(hs["cur_u"]), np.array(hs["cur_v"]))
        if wl.epzs:
            g = ctx.epzs_search_frame(np.array(hs["pred"]), np.array(hs["shared"]), wl.efp).reshape(n_mb, 41)
        else:
            g = ctx.me_search_frame_pred(np.array(hs["pred"]), wl.fp).reshape(n_mb, 41)
        heads, tokens = ctx.mc_tq_modes_compact(None, wl.qd, wl.mode_mask, n_mb=n_mb, token_cap=wl.token_cap)
        ok_mv = bool(np.array_equal(g["mv_x"][idx], mv[:, :, 0]) and np.array_equal(g["mv_y"][idx], mv[:, :, 1]) and
                     np.array_equal(g["cost"][idx], cost))
        ok_lev = bool(np.array_equal(expand_tokens(heads, tokens, idx, per=16 if wl.cfg["n"] == 4 else 64), lev))
        res["gpu_matches_reference_on_sample"] = ok_mv and ok_lev
        res["checked"] = {"mv_and_cost": ok_mv, "levels": ok_lev, "nonzero_levels_in_sample": int((lev != 0).sum())}
        if wl.chroma:
            co = out_cpu[4]
            ok_c = True
            for m in range(7):
                gc = ctx.chroma_residual_coding(wl.cdesc, None, m + 1, 0, n_mb, want_recon=False)
                ok_c = ok_c and bool(np.array_equal(gc["dc"][idx], co["dc"][:, m]) and np.array_equal(gc["cbp_blk"][idx], co["cb"][:, m]) and
                                     np.array_equal(gc["cr_cbp"][idx], co["cc"][:, m]))
            res["checked"]["chroma_dc_levels_and_cbp"] = ok_c
            res["gpu_matches_reference_on_sample"] = res["gpu_matches_reference_on_sample"] and ok_c
    return res


def run_reference(args):
    """--impl reference: JM's own CPU implementation of the path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    from jm_b200 import api, synth
    from oracle import pyoracle as po
    from jm_b200 import h264_tables as T
    world = int(os.environ.get("WORLD_SIZE", 1))
    c = CONFIGS[args.config]
    epzs = c["search"] == "epzs"
    if not epzs and not po.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjmref.so not built"}))
        return
    W, H = c["w"], c["h"]
    cores = os.cpu_count() or 1
    lam = T.lambda_me(QP)
    f = synth.luma_frames(W, H, 2, seed=1234, motion=(5, 3))
    n_mb = (W // 16) * (H // 16)
    mbw = W // 16
    pred = make_pred_table(api, 50, n_mb)["pred"]
    shared = make_shared_candidates(50, n_mb, 8) if epzs else None
    chroma = None
    if c["chroma"]:
        chroma = make_chroma_planes(f[0], 2, 11) + make_chroma_planes(f[1], 2, 12)
    per_core = args.ref_mbs_per_core * (8 if epzs else 1)      # an EPZS macroblock is ~50x cheaper than a full-search one
    rng = np.random.default_rng(0)

    def job(step):
        idx = rng.permutation(n_mb)[: per_core * cores].reshape(cores, per_core)
        return [(ii, pred[ii], shared[ii] if epzs else None, lam) for ii in idx]

    with mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(args.config, f[0], f[1], W, H, chroma)) as pool:
        for s in range(args.warmup):
            pool.map(_cpu_worker, job(s))
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_worker, job(args.warmup + s))
        el = time.perf_counter() - t0
    v = per_core * cores * args.steps / el
    out = {"impl": "reference", "metric": METRIC[args.config], "value": v, "unit": "macroblocks/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(args.config, world), "sample_macroblocks_per_step": per_core * cores,
           "cpu_baseline": {"value": v, "unit": "macroblocks/s", "cores": cores, "kind": "port" if epzs else "reference",
                            "sample": f"each step = {per_core * cores} random macroblocks of the {c['size']} picture ({per_core} per core, one "
                                      f"instance per core, quarter-pel planes built once before the timed region), " +
                                      ("the pinned CPU restatement of JM's EPZS integer + sub-pel searches + forward8x8/quant_8x8_normal"
                                       if epzs else "JM's own setup_fast_full_search / fast_full_search / sub_pel functions + the pinned restatement for "
                                       "4x4 transform/quant and the 4:2:2 chroma path" if c["chroma"] else
                                       "JM's own full_search/sub_pel/forward4x4/quant_4x4_normal") +
                                      "; value = sampled macroblocks / time"},
           "e2e": {"value": v, "unit": "macroblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[1..4] (default 2 = configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-worst", action="store_true", help="skip the worst-case (unrelated reference) probe of the search kernel")
    ap.add_argument("--scene-cut", action="store_true", help="probe: unrelated reference picture (worst case for the search gate)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-mbs-per-core", type=int, default=160)
    ap.add_argument("--e2e-streams", type=int, default=0,
                    help="end-to-end leg: independent picture streams per GPU (host threads x contexts); 1 = strictly serial pictures; "
                         "0 = auto: min(3, host cores / ranks - 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    for k in ("JMB_SHIM", "JMB_SHIM_OFF"):      # debug switches of the drop-in shim: never part of a measured leg
        if os.environ.get(k):
            sys.exit(f"bench.py: {k} is set; the measured legs run the device path only")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

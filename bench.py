#!/usr/bin/env python
"""bench.py -- macroblocks/sec of the JM lencod ME + transform/quant hot path on B200.

One "step" = the hot path over ONE 1080p P-picture (coded 1920x1088 = 8160 macroblocks, 1 reference):
  K6  jmb_ref_put            16 quarter-pel planes of the reference           (getSubImagesLuma)
  --  jmb_pic_begin          current picture to the device
  K1-K3 jmb_me_search_frame  full search +-32, all 41 partitions of every MB  (full_search_motion_estimation)
  K5  (same call)            half-/quarter-pel SATD refinement of every one   (sub_pel_motion_estimation)
  K7/K8 jmb_mc_tq_modes      prediction -> residual -> forward4x4 -> quant for each of the 7 partition
                             modes (what JM's RDO loop residual-codes per inter candidate), one launch
Predictors are synthetic (true motion + per-MB / per-partition jitter), lambda from QP 28.

  value : device-timed (CUDA events on the library's stream), inputs resident in HBM, rotating over
          4 distinct input sets (> L2 in total) so no step re-reads a warm L2.
  e2e   : the same step through the C ABI with pinned HOST buffers (H2D/D2H inside the timed region).
  --impl reference : JM's own functions (oracle/_ref/libjmref.so) on the host cores, bounded sample.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU,
weak scaling: every rank encodes its own pictures, i.e. independent closed-GOP segments; no collective
on the data path).
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

W, H = 1920, 1088                  # coded size of 1080p (1920x1080 -> 68 MB rows); --size 4k: 3840x2176
SIZE_NAME = "1080p"
SEARCH_RANGE = 32
QP = 28
N_SETS = 4
BYTES_PER_MB_REF = 13804           # 512 src + 12800 window + 492 results


def workload_config(n_gpus, anchor_bcast=False):
    return {"workload": f"{SIZE_NAME} 4:2:0 synthetic, FullSearch +-32 (SearchMode=-1) 41 partitions/MB + SATD sub-pel + "
                        "4x4 transform/quant of 7 partition modes, Baseline, 1 ref, QP28",
            "width": W, "height": H, "macroblocks_per_step": (W // 16) * (H // 16), "search_range": SEARCH_RANGE,
            "l2": f"rotating over {N_SETS} distinct input sets (> 126 MB in total)",
            "parallelism": (f"{n_gpus} x pictures sharing one anchor: ncclBroadcast of the reconstructed reference (4.2 MB u16 luma) per step"
                            if anchor_bcast and n_gpus > 1 else
                            f"{n_gpus} x independent picture streams (closed-GOP shards), no data-path collective")}


def make_requests(api, seed, motion_q=(20, 12)):
    """41 requests per MB in canonical order; predictors = true motion + jitter."""
    rng = np.random.default_rng(seed)
    parts = api.mb_partitions()
    mbw, mbh = W // 16, H // 16
    n_mb = mbw * mbh
    reqs = np.zeros((n_mb, api.NPART), api.ME_REQ)
    mbx = (np.arange(n_mb) % mbw) * 16
    mby = (np.arange(n_mb) // mbw) * 16
    mbpred = np.array(motion_q)[None, :] + rng.integers(-8, 9, size=(n_mb, 2))
    for k, (t, x, y) in enumerate(parts):
        p = mbpred + rng.integers(-3, 4, size=(n_mb, 2))
        reqs["blocktype"][:, k] = t
        reqs["pos_x"][:, k] = mbx + x
        reqs["pos_y"][:, k] = mby + y
        reqs["pred_x"][:, k] = p[:, 0]
        reqs["pred_y"][:, k] = p[:, 1]
        reqs["center_x"][:, k] = ((p[:, 0] + 2) >> 2) * 4
        reqs["center_y"][:, k] = ((p[:, 1] + 2) >> 2) * 4
    reqs["mode"] = api.SEARCH_FULL
    reqs["flags"] = api.REQ_SUBPEL
    reqs["min_mcost"] = api.DISTBLK_MAX
    return reqs.reshape(-1)


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
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    n_mb = (W // 16) * (H // 16)
    lam = T.lambda_me(QP)
    qd = api.quant_desc(4, QP, T.q_params(QP, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0], 1)

    # ---- inputs: N_SETS distinct (reference, current) pairs + request lists, on host (pinned) and in HBM
    sets = []
    for s in range(N_SETS):
        # anchor-broadcast mode: every rank codes a picture of the SAME sequence against rank 0's anchor
        f = synth.luma_frames(W, H, 2, seed=1234 + (0 if args.anchor_bcast else 97 * rank) + s, motion=(5, 3))
        reqs = make_requests(api, seed=50 + 13 * rank + s)
        reqs["lambda"] = lam
        hs = {"ref": ctx.pinned((H, W), np.uint16), "cur": ctx.pinned((H, W), np.uint16), "reqs": ctx.pinned(len(reqs), api.ME_REQ)}
        if args.scene_cut:      # robustness probe: the reference is an unrelated picture (nothing matches)
            f[0] = synth.luma_frames(W, H, 1, seed=999 + s)[0]
        hs["ref"][:] = f[0]; hs["cur"][:] = f[1]; hs["reqs"][:] = reqs
        ds = {k: torch.from_numpy(v.view(np.uint8).reshape(-1).copy()).cuda(local) for k, v in hs.items()}
        sets.append((hs, ds))
    d_res = torch.empty(n_mb * api.NPART * api.ME_RES.itemsize, dtype=torch.uint8, device=f"cuda:{local}")
    d_lev = torch.empty(7 * n_mb * 256, dtype=torch.int16, device=f"cuda:{local}")
    d_cost = torch.empty(7 * n_mb * 4, dtype=torch.int32, device=f"cuda:{local}")
    d_cbp = torch.empty(7 * n_mb, dtype=torch.int32, device=f"cuda:{local}")
    h_res = ctx.pinned(n_mb * api.NPART, api.ME_RES)
    h_lev = ctx.pinned((7, n_mb, 256), np.int16); h_cost = ctx.pinned((7, n_mb, 4), np.int32); h_cbp = ctx.pinned((7, n_mb), np.uint32)
    torch.cuda.synchronize()

    from jm_b200 import shard

    def step_device(s):
        hs, ds = sets[s % N_SETS]
        if args.anchor_bcast and world > 1:
            # B-picture fan-out: rank 0 holds the reconstructed anchor; one NCCL broadcast of its u16 luma
            # plane, then every rank builds its own quarter-pel planes and codes its own picture against it
            with torch.cuda.stream(stream):
                shard.broadcast_anchor(ds["ref"], src=0)
        ctx.ref_put(s % 2, ds["ref"].data_ptr(), api.DEVICE, shape=(H, W))
        ctx.pic_begin(ds["cur"].data_ptr(), [s % 2], api.DEVICE, shape=(H, W))
        ctx.me_search(ds["reqs"].data_ptr(), d_res.data_ptr(), api.DEVICE, n=n_mb * api.NPART, frame=True)
        ctx.mc_tq_modes(d_res.data_ptr(), qd, 0x7F, api.DEVICE, n_mb=n_mb, out=(d_lev.data_ptr(), d_cost.data_ptr(), d_cbp.data_ptr()))

    # ---- end-to-end leg: K independent picture streams per GPU (K host threads, each its own context = its own CUDA
    # stream, reference slots and staging; the calls are the synchronous JMB_HOST ones, so one stream's PCIe copies
    # overlap the other's kernels).  Pictures are independent units (closed-GOP shards), exactly like the ranks.
    # Measured on the 16-core B200 box (tools/gpu_e2e.sh, macroblocks/s over repeated runs): 3 streams 6.6 M; 6 streams 6.4-8.1 M;
    # 8 streams 7.7-7.9 M (the tightest); 10 streams 5.3-6.4 M; 4 streams is bimodal (7.1 M / 4.0 M: the streams' copies convoy).
    n_streams = args.e2e_streams if args.e2e_streams > 0 else max(1, min(8, (os.cpu_count() or 1) // world - 2))
    if args.e2e_streams <= 0 and n_streams == 4:
        n_streams = 5
    e2e_ctx = [ctx] + [api.Context(local) for _ in range(n_streams - 1)]
    for c in e2e_ctx[1:]:
        c.configure(search_range=SEARCH_RANGE)
    e2e_out = [(h_res, h_lev, h_cost, h_cbp)] + [(c.pinned(n_mb * api.NPART, api.ME_RES), c.pinned((7, n_mb, 256), np.int16),
                                                  c.pinned((7, n_mb, 4), np.int32), c.pinned((7, n_mb), np.uint32)) for c in e2e_ctx[1:]]
    e2e_t = [{"ref_put": 0.0, "pic_begin": 0.0, "me_search": 0.0, "mc_tq": 0.0} for _ in e2e_ctx]

    def step_host(s, k=0):
        c, (o_res, o_lev, o_cost, o_cbp), tt = e2e_ctx[k], e2e_out[k], e2e_t[k]
        hs, _ = sets[s % N_SETS]
        t0 = time.perf_counter()
        c.ref_put(s % 2, hs["ref"]); t1 = time.perf_counter()
        c.pic_begin(hs["cur"], [s % 2]); t2 = time.perf_counter()
        c.me_search(hs["reqs"], o_res, frame=True); t3 = time.perf_counter()
        tt["ref_put"] += t1 - t0; tt["pic_begin"] += t2 - t1; tt["me_search"] += t3 - t2
        # residual coding of all 7 partition modes from the results still resident in HBM; levels / costs / cbp come back
        c._ck(c.L.jmb_mc_tq_modes(c.h, None, n_mb, 0x7F, qd.ctypes.data, o_lev.ctypes.data, o_cost.ctypes.data, o_cbp.ctypes.data, api.HOST))
        tt["mc_tq"] += time.perf_counter() - t3

    def run_host_steps(first, count):
        """`count` pictures starting at step index `first`, dealt round-robin to the streams' threads."""
        def worker(k):
            for s in range(first + k, first + count, n_streams):
                step_host(s, k)
            e2e_ctx[k].sync()
        th = [threading.Thread(target=worker, args=(k,)) for k in range(n_streams)]
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
        step_device(s)
    ctx.sync()
    barrier()
    ctx.timing(True)
    sampler.mark_begin()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record(stream)
    e1.synchronize()
    sampler.mark_end()
    barrier()
    clocks = sampler.stop()
    gpu_launches = ctx.launches - launches0
    ms = e0.elapsed_time(e1)
    k_ms, k_n = ctx.timing_get("int_search")
    kernel_break = {k: ctx.timing_get(k)[0] / max(1, args.steps) for k in ("subpel_planes", "pack_cur", "int_search", "subpel_refine", "mc_tq")}
    ctx.timing(False)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())

    # ---- end-to-end timing (host buffers through the C ABI) ----------------------------------------
    run_host_steps(0, max(3, n_streams) * 2)
    barrier()
    for tt in e2e_t:
        for k in tt:
            tt[k] = 0.0
    t0 = time.perf_counter()
    run_host_steps(args.warmup, args.steps)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local}"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    h2d = 2 * W * H * 2 + n_mb * api.NPART * api.ME_REQ.itemsize + api.QUANT_DESC.itemsize
    d2h = n_mb * api.NPART * api.ME_RES.itemsize + 7 * n_mb * (512 + 16 + 4)

    out = {"metric": "macroblocks/sec 1080p full-search ME+DCT/quant", "value": world * n_mb * args.steps / (ms / 1e3),
           "unit": "macroblocks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(world, args.anchor_bcast), "clocks": clocks, "gpu_launches": int(gpu_launches),
           "e2e": {"value": world * n_mb * args.steps / e2e_s, "unit": "macroblocks/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps,
                   "picture_streams_per_gpu": n_streams,
                   "host_ms_per_picture": {k: 1e3 * sum(tt[k] for tt in e2e_t) / args.steps for k in e2e_t[0]}},
           "kernel_ms_per_step": kernel_break}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_PER_MB_REF * n_mb / (k_ms / max(1, k_n) / 1e3) / 1e9
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (tools/ncu_summary.py)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "int_search_traffic.json")))[args.size]["dram_bytes_per_launch"]
    except Exception:
        pass
    out["roofline"] = {"bound": "hbm", "kernel": "k_int_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                       "frac": achieved / peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                       "algorithmic_bytes_per_launch": BYTES_PER_MB_REF * n_mb, "launch_ms": k_ms / max(1, k_n),
                       "note": "search-window model of SURVEY 8(d): 13804 B per macroblock*reference; the kernel is ALU-pipe "
                               "(VABSDIFF4/PRMT/ISETP) bound, not HBM bound -- DESIGN.md 3; traffic < algorithmic bytes because "
                               "neighbouring windows hit L2"}

    # The same kernel against the bound that actually limits it: the SM ALU pipe (VABSDIFF4 / PRMT / ISETP issue at
    # 64 lanes/clk/SM on this part, tools/ubench.cu).  Floor per displacement = 64 VABSDIFF4 + 19 PRMT + 41 ISETP.
    alu_ops = n_mb * (2 * SEARCH_RANGE + 1) ** 2 * (64 + 19 + 41)
    sm_hz = 1e6 * float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    alu_peak = 64.0 * 148 * sm_hz
    out["alu_roofline"] = {"bound": "alu-pipe", "achieved": alu_ops / (k_ms / max(1, k_n) / 1e3), "peak": alu_peak, "unit": "lane-ops/s",
                           "frac": alu_ops / (k_ms / max(1, k_n) / 1e3) / alu_peak,
                           "note": "minimum ALU-pipe instructions of the algorithm / launch time, vs 64 lanes/clk/SM x 148 SMs x SM clock"}
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(sets[0][0], lam, ctx=ctx, api=api, budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_JM = {}


def _cpu_init(ref_luma, cur_luma):
    """One process = one JM instance (JM is single-threaded); its quarter-pel planes are built once, untimed."""
    from oracle import pyoracle as po
    ref = po.JMRef(W, H, SEARCH_RANGE)
    ref.set_ref(ref_luma); ref.set_cur(cur_luma)
    _JM["ref"] = ref


def _cpu_worker(a):
    mb_xy, preds, lam = a
    from jm_b200 import h264_tables as T
    from oracle import pyoracle as po
    ref = _JM["ref"]
    mv, cost, lev, secs = po.jmref_run_mbs(ref, mb_xy, preds, [lam] * 3, QP, T.q_params(QP, 0, 4), T.SNGL_SCAN, T.COEFF_COST4x4[0])
    return mv, cost, lev, secs


def cpu_baseline(hs, lam, ctx=None, api=None, budget_s=12.0):
    """JM's own leaf functions (kind 'reference') on 1 host core over a bounded sample of set 0's macroblocks;
    also re-checks the GPU results of that sample bit-for-bit."""
    from oracle import pyoracle as po
    if not po.ref_available():
        return {"value": None, "unit": "macroblocks/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libjmref.so missing"}
    n_mb = (W // 16) * (H // 16)
    reqs = np.array(hs["reqs"]).reshape(n_mb, 41)
    # calibrate on 32 MBs, then size the sample for ~budget_s
    idx = np.linspace(0, n_mb - 1, 32).astype(int)
    _cpu_init(np.array(hs["ref"]), np.array(hs["cur"]))
    args = lambda ii: (np.stack([reqs["pos_x"][ii, 0], reqs["pos_y"][ii, 0]], 1),
                       np.stack([reqs["pred_x"][ii], reqs["pred_y"][ii]], 2), lam)
    _, _, _, secs = _cpu_worker(args(idx))
    per_mb = max(1e-5, float(secs.sum()) / len(idx))
    n = int(min(n_mb, max(64, budget_s / per_mb)))
    idx = np.linspace(0, n_mb - 1, n).astype(int)
    mv, cost, lev, secs = _cpu_worker(args(idx))
    res = {"value": n / float(secs.sum()), "unit": "macroblocks/s", "cores": 1, "kind": "reference",
           "sample": f"{n} of {n_mb} macroblocks of input set 0 (evenly spaced), JM full_search_motion_estimation + "
                     f"sub_pel_motion_estimation x41 + forward4x4/quant_4x4_normal x112 per MB",
           "me_seconds": float(secs[0]), "tq_seconds": float(secs[1])}
    if ctx is not None:
        ctx.ref_put(0, hs["ref"]); ctx.pic_begin(hs["cur"], [0])      # whatever picture the timed legs left resident, this is set 0
        g = ctx.me_search(hs["reqs"], frame=True).reshape(n_mb, 41)
        ok = bool(np.array_equal(g["mv_x"][idx], mv[:, :, 0]) and np.array_equal(g["mv_y"][idx], mv[:, :, 1]) and
                  np.array_equal(g["cost"][idx], cost))
        res["gpu_matches_reference_on_sample"] = ok
    return res


def run_reference(args):
    """--impl reference: JM's own CPU implementation of the path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    from jm_b200 import api, synth
    from jm_b200 import h264_tables as T
    from oracle import pyoracle as po
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not po.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjmref.so not built"}))
        return
    cores = os.cpu_count() or 1
    lam = T.lambda_me(QP)
    f = synth.luma_frames(W, H, 2, seed=1234, motion=(5, 3))
    reqs = make_requests(api, seed=50).reshape(-1, 41)
    n_mb = len(reqs)
    per_core = args.ref_mbs_per_core
    rng = np.random.default_rng(0)

    def job(step):
        idx = rng.permutation(n_mb)[: per_core * cores].reshape(cores, per_core)
        return [(np.stack([reqs["pos_x"][ii, 0], reqs["pos_y"][ii, 0]], 1),
                 np.stack([reqs["pred_x"][ii], reqs["pred_y"][ii]], 2), lam) for ii in idx]

    with mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(f[0], f[1])) as pool:
        for s in range(args.warmup):
            pool.map(_cpu_worker, job(s))
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_worker, job(args.warmup + s))
        el = time.perf_counter() - t0
    v = per_core * cores * args.steps / el
    out = {"impl": "reference", "metric": "macroblocks/sec 1080p full-search ME+DCT/quant", "value": v, "unit": "macroblocks/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": workload_config(world),
           "cpu_baseline": {"value": v, "unit": "macroblocks/s", "cores": cores, "kind": "reference",
                            "sample": f"each step = {per_core * cores} random macroblocks of the 1080p picture ({per_core} per core, one JM "
                                      f"instance per core, quarter-pel planes built once before the timed region), "
                                      f"JM's own full_search/sub_pel/forward4x4/quant_4x4_normal"},
           "e2e": {"value": v, "unit": "macroblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scene-cut", action="store_true", help="probe: unrelated reference picture (worst case for the search gate)")
    ap.add_argument("--anchor-bcast", action="store_true",
                    help="N>1: broadcast rank 0's reference picture over NCCL every step (pictures sharing an anchor coded on different GPUs)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-mbs-per-core", type=int, default=160)
    ap.add_argument("--e2e-streams", type=int, default=0,
                    help="end-to-end leg: independent picture streams per GPU (host threads x contexts); 1 = strictly serial calls; "
                         "0 = auto: min(3, host cores / (2 x ranks)), the synchronous calls spin-wait on the host")
    ap.add_argument("--size", default="1080p", choices=["1080p", "4k"], help="picture size (default = BASELINE configs[1])")
    args = ap.parse_args()
    if args.size == "4k":
        global W, H, SIZE_NAME
        W, H, SIZE_NAME = 3840, 2176, "4K (3840x2160 coded 3840x2176)"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Multi-GPU sharding of the ME + transform/quant path (host logic).

The path shards on INDEPENDENT units -- closed-GOP segments of a sequence -- one process + one GPU per shard, and needs no
data-path collective: each shard is a stock JM run restricted with StartFrame / FramesToBeEncoded / IDRPeriod
(lencod/inc/configfile.h:39-47) whose hot leaves go to its own GPU (JMB_DEVICE).  The one real exchange step is the
reconstructed-reference broadcast: when pictures that share an anchor (e.g. the non-reference B pictures between two
anchors) are coded on different GPUs, the anchor's reconstructed luma plane is broadcast once (NCCL over NVLink; gloo in
the CPU tests) and every rank builds its own quarter-pel planes from it (jmb_ref_put).  The P-picture chain itself does not
shard (picture n+1 references picture n's reconstruction): replicas only.
"""
from dataclasses import dataclass
from typing import List


@dataclass(frozen=True)
class Segment:
    rank: int
    start_frame: int
    n_frames: int

    def lencod_overrides(self, gop: int) -> List[str]:
        """-p overrides that make a stock lencod (or lencod_jmb) run encode exactly this closed-GOP segment."""
        return [f"StartFrame={self.start_frame}", f"FramesToBeEncoded={self.n_frames}", f"IntraPeriod={gop}", f"IDRPeriod={gop}"]


def plan_gop_segments(n_frames: int, world: int, gop: int) -> List[List[Segment]]:
    """Closed GOPs of `gop` frames dealt to `world` ranks in contiguous runs (segment boundaries = IDR pictures).
    Returns one list of segments per rank; ranks beyond the number of GOPs get an empty list."""
    if n_frames <= 0 or world <= 0 or gop <= 0:
        raise ValueError("n_frames, world and gop must be positive")
    n_gops = (n_frames + gop - 1) // gop
    per, extra = divmod(n_gops, world)
    plan, g = [], 0
    for r in range(world):
        k = per + (1 if r < extra else 0)
        segs = []
        if k:
            start = g * gop
            segs.append(Segment(r, start, min(k * gop, n_frames - start)))
        g += k
        plan.append(segs)
    return plan


def b_picture_owner(poc_in_minigop: int, world: int) -> int:
    """Fan-out of the non-reference B pictures between two anchors: picture i of the mini-GOP goes to rank i % world."""
    return poc_in_minigop % world


def broadcast_anchor(plane, src: int = 0, group=None):
    """Broadcast a reconstructed reference plane (contiguous torch tensor of 16-bit samples, sent as raw bytes) from `src`
    to every rank, in place.  Backend-agnostic: NCCL for CUDA tensors, gloo for CPU tensors (tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return plane
    dist.broadcast(plane.view(torch.uint8), src=src, group=group)
    return plane


def concat_annexb(paths, out_path):
    """Closed-GOP segments concatenate byte-wise into one Annex-B stream (every segment starts with SPS/PPS + IDR)."""
    with open(out_path, "wb") as o:
        for p in paths:
            with open(p, "rb") as f:
                o.write(f.read())

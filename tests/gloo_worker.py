"""One rank of the world_size-2 CPU (gloo) test of the sharded path; launched by tests/test_shard.py with RANK /
WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jm_b200 import shard, synth   # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    anchor = torch.from_numpy(synth.luma_frames(64, 48, 1, seed=99)[0].astype(np.int16))
    plane = anchor.clone() if rank == 0 else torch.zeros_like(anchor)        # only rank 0 "reconstructed" the anchor
    shard.broadcast_anchor(plane, src=0)
    # every rank builds the quarter-pel planes from its copy (the CPU oracle stands in for jmb_ref_put here) ...
    o = po.Oracle(); r = o.ref_create(plane.numpy().astype(np.uint16)); planes = o.planes(r); o.ref_destroy(r)
    # ... and codes its own share of the pictures: shards are disjoint and cover the sequence
    plan = shard.plan_gop_segments(16, world, 4)
    digest = torch.tensor([int(planes.astype(np.int64).sum() % (1 << 31))])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)      # bench.py's rule: max over ranks
    print(json.dumps({"rank": rank, "anchor_ok": bool(torch.equal(plane, anchor)), "digests": [int(g.item()) for g in gathered],
                      "mine": [(s.start_frame, s.n_frames) for s in plan[rank]], "tmax": float(t.item())}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-side sharding logic -- CPU only; the N>1 path is exercised with world_size-2 gloo processes."""
import os
import sys

import numpy as np
import pytest
import json
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jm_b200 import shard, synth   # noqa: E402


def test_plan_covers_every_frame_once():
    for n, world, gop in [(64, 8, 8), (30, 4, 8), (5, 8, 4), (17, 3, 5), (8, 1, 8)]:
        plan = shard.plan_gop_segments(n, world, gop)
        assert len(plan) == world
        frames = [f for segs in plan for s in segs for f in range(s.start_frame, s.start_frame + s.n_frames)]
        assert frames == list(range(n))
        for segs in plan:
            for s in segs:
                assert s.start_frame % gop == 0            # every segment starts on an IDR picture
        ov = plan[0][0].lencod_overrides(gop)
        assert f"StartFrame={plan[0][0].start_frame}" in ov and f"IDRPeriod={gop}" in ov
    with pytest.raises(ValueError):
        shard.plan_gop_segments(0, 2, 4)
    assert [shard.b_picture_owner(i, 4) for i in range(7)] == [0, 1, 2, 3, 0, 1, 2]


def test_anchor_broadcast_world2_gloo():
    world, port = 2, 29500 + os.getpid() % 500
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    out = []
    for p in procs:
        so, se = p.communicate(timeout=300)
        assert p.returncode == 0, se[-800:]
        out.append(json.loads(so.strip().splitlines()[-1]))
    out.sort(key=lambda o: o["rank"])
    assert all(o["anchor_ok"] for o in out)                                       # every rank holds the anchor bit-exactly
    assert out[0]["digests"] == out[1]["digests"] and len(set(out[0]["digests"])) == 1   # identical planes on both ranks
    assert out[0]["mine"] == [[0, 8]] and out[1]["mine"] == [[8, 8]]              # disjoint closed-GOP shards
    assert out[0]["tmax"] == out[1]["tmax"] == 2.0                                # max-over-ranks reduction

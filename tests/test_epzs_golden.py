"""Pins the EPZS restatement (oracle/jm_oracle.c::jmo_epzs) to the REAL JM: tests/golden/epzs_golden.npz holds calls of
EPZS_integer_motion_estimation / EPZS_sub_pel_motion_estimation recorded inside the live stock encoder
(tests/golden/make_epzs_golden.py); the restatement must return JM's motion vector, cost and *prevSad for every one.  CPU only."""
import os

import numpy as np
import pytest

from jm_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "epzs_golden.npz")


@pytest.mark.parametrize("tag", ["cabac8x8", "base4x4"])
def test_restatement_reproduces_recorded_jm_calls(oracle, tag):
    g = np.load(GOLD)
    reqs = g[f"{tag}_req"].view(api.EPZS_REQ) if g[f"{tag}_req"].dtype != api.EPZS_REQ else g[f"{tag}_req"]
    meta, out, cands = g[f"{tag}_meta"], g[f"{tag}_out"], g[f"{tag}_cands"]
    refs = {}
    n_int = n_sub = 0
    off = 0
    exits = set()
    for i in range(len(reqs)):
        kind, ref_id, cur_id, mvx, mvy, me, metrics, nc = (int(v) for v in meta[i])
        if ref_id not in refs:
            refs[ref_id] = oracle.ref_create(g[f"{tag}_pic_{ref_id}"].astype(np.uint16))
        cur = g[f"{tag}_pic_{cur_id}"].astype(np.uint16)
        q = reqs[i:i + 1].copy()
        q["cand_off"] = 0
        if kind == 2:      # integer stage: mv, returned cost, *prevSad afterwards
            r = oracle.epzs(refs[ref_id], cur, q, cands[off:off + nc], (2, 2, 0, 1, 9))[0]
            assert (int(r["imv_x"]), int(r["imv_y"]), int(r["icost"]), int(r["prev_sad"])) == (mvx, mvy, int(out[i, 0]), int(out[i, 1])), (i, q, r, meta[i], out[i])
            exits.add(int(r["exit_code"]))
            n_int += 1
        else:              # sub-pel stage
            mcfg = (metrics & 15, metrics >> 4, me & 1, (me >> 1) & 1, me >> 2)
            r = oracle.epzs(refs[ref_id], cur, q, np.zeros((1, 2), np.int16), mcfg)[0]
            assert (int(r["mv_x"]), int(r["mv_y"]), int(r["cost"])) == (mvx, mvy, int(out[i, 0])), (i, q, r, meta[i], out[i])
            n_sub += 1
        off += nc
    for r in refs.values():
        oracle.ref_destroy(r)
    assert n_int > 200 and n_sub > 200 and 5 in exits, (n_int, n_sub, exits)

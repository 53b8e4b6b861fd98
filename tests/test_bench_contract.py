"""The bench line the driver parses: keys and types of the committed round result (profiles/r01_bench*.json), so an edit
to bench.py that drops a contract key is caught on CPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r01_bench.json")
REF = os.path.join(ROOT, "profiles", "r01_bench_reference.json")

pytestmark = pytest.mark.skipif(not os.path.exists(LINE), reason="no committed bench line")


def test_our_arm_line():
    d = json.load(open(LINE))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "macroblocks/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"] and c["value"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = json.load(open(REF))
    assert d["impl"] == "reference" and d["unit"] == "macroblocks/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    ours = json.load(open(LINE))
    assert d["metric"] == ours["metric"] and d["config"]["workload"] == ours["config"]["workload"]


R02 = {c: os.path.join(ROOT, "profiles", f"r02_bench{'' if c == 2 else '_c%d' % c}.json") for c in (2, 3, 4)}


@pytest.mark.parametrize("cfg", [2, 3, 4])
def test_round2_lines(cfg):
    """The committed round-2 lines of configs 2 / 3 / 4: contract keys, a clean clock record taken during the timed region, the
    sample re-check against JM green, and the tier's extra objects (roofline with live launch time, cpu_baseline, e2e with copies)."""
    if not os.path.exists(R02[cfg]):
        pytest.skip("no committed round-2 line")
    d = json.loads(open(R02[cfg]).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "kernel_ms_per_step"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] >= 3 * d["steps"] and d["vs_baseline"] is None
    assert abs(d["value"] - (d["config"]["macroblocks_per_step"] / (d["ms_per_step"] / 1e3))) / d["value"] < 1e-6
    c = d["clocks"]
    assert c["samples"] >= 1 and c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["launch_ms"] > 0
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] / 1e3) / 1e9) / r["achieved"] < 1e-6
    b = d["cpu_baseline"]
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    checked = b.get("checked") or b.get("gpu_matches_reference_on_sample")
    assert checked and checked["mv_and_cost"] is True and checked["levels"] is True
    if cfg == 4:
        assert checked["chroma_dc_levels_and_cbp"] is True

"""The bench.py JSON contract, checked on the lines recorded on the B200 (profiles/r01_bench_*.jsonl, newest last)."""
import json
import os

from tests.util import ROOT


def _last(name):
    lines = [l for l in open(os.path.join(ROOT, "profiles", name)).read().splitlines() if l.strip()]
    return json.loads(lines[-1]), [json.loads(l) for l in lines]


def test_our_arm_line_has_every_contract_key():
    d, _ = _last("r01_bench_ours.jsonl")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "optimised_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] == d["steps"]              # one launch of the fused optimiser per step
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["frames_over_capacity_all_ranks"] == 0
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and len(c["sample"]) > 10
    cl = d["clocks"]
    assert cl["samples"] > 0 and cl["sm_mhz"] > 0.9 * cl["sm_max_mhz"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    m2 = d["m2_rasterizer_dense"]["roofline"]
    assert m2["bound"] == "hbm" and m2["frac"] >= 0.5                       # north-star: dense rasteriser fwd+bwd >= 50 % of the HBM roof


def test_reference_arm_line():
    d, _ = _last("r01_bench_reference.jsonl")
    assert d["impl"] == "reference" and d["metric"] == "optimised_frames_per_sec" and d["unit"] == "frames/s"
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]


def test_recorded_scaling_exceeds_the_north_star_target():
    _, lines = _last("r01_bench_ours.jsonl")
    best8 = max((l["value"] for l in lines if l["n_gpus"] == 8), default=None)
    assert best8 is not None and best8 >= 10000.0                           # >= 10 000 optimised H36M frames/s on 8 x B200

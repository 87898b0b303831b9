"""The bench.py JSON contract, checked on the lines recorded on the B200 (profiles/r02_bench_*.jsonl, newest last)."""
import json
import os

from tests.util import ROOT


def _last(name):
    lines = [l for l in open(os.path.join(ROOT, "profiles", name)).read().splitlines() if l.strip()]
    return json.loads(lines[-1]), [json.loads(l) for l in lines]


def _check_roofline(r):
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    assert r["n_mask_per_view_measured"] > 0 and r["algorithmic_bytes_per_view_iteration"] > 4 * r["n_mask_per_view_measured"]


def test_our_arm_line_has_every_contract_key():
    d, _ = _last("r02_bench_ours.jsonl")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "configs", "accuracy"):
        assert k in d, k
    assert d["metric"] == "optimised_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] == d["steps"]              # one launch of the fused optimiser per step
    assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["name"] == "h36m"
    assert d["config"]["frames_over_capacity_all_ranks"] == 0
    assert "flush" in d["config"]["l2"].lower()
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.02
    _check_roofline(d["roofline"])
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and len(c["sample"]) > 10
    x = c["e2e_cross_check"]                                                 # end-to-end MPJPE cross-check of the CPU port
    assert x["frames"] == 8 and x["max_joint_deviation_mm"] < 0.1
    cl = d["clocks"]
    assert cl["samples"] > 0 and cl["sm_mhz"] > 0.9 * cl["sm_max_mhz"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    m2 = d["m2_rasterizer_dense"]["roofline"]
    assert m2["bound"] == "hbm" and m2["frac"] >= 0.5                       # north-star: dense rasteriser fwd+bwd >= 50 % of the HBM roof


def test_every_baseline_config_is_measured_in_the_default_line():
    """BASELINE.json configs 3-5 (h36m-occ, panoptic, the 8-view 100k-frame sweep) ride in the default line's `configs` block, each
    with value, e2e, roofline, capacity report and -- when the reference arm ran before on the same box -- the accuracy
    cross-check against the unmodified reference kernels."""
    d, _ = _last("r02_bench_ours.jsonl")
    assert set(d["configs"]) == {"h36m-occ", "panoptic", "occlusion-person-8v"}
    for name, c in d["configs"].items():
        assert "error" not in c, (name, c)
        assert c["value"] > 0 and c["unit"] == "frames/s" and c["frames_over_capacity_all_ranks"] == 0
        assert c["e2e"]["h2d_bytes_per_step"] > 0 and 0 < c["e2e"]["value"] < c["value"] * 1.02
        _check_roofline(c["roofline"])
        v = c["accuracy"]["vs_reference"]
        assert v is None or abs(v["mpjpe_delta_mm"]) < 0.1, (name, v)    # north-star level 3: MPJPE within 0.1 mm of the reference pipeline
    assert d["configs"]["occlusion-person-8v"]["total_frames_timed"] >= 100_000
    v = d["accuracy"]["vs_reference"]
    assert v is None or abs(v["mpjpe_delta_mm"]) < 0.1


def test_reference_arm_line():
    d, _ = _last("r02_bench_reference.jsonl")
    assert d["impl"] == "reference" and d["metric"] == "optimised_frames_per_sec" and d["unit"] == "frames/s"
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["ranks_used"] == d["n_gpus"] and abs(d["value_per_gpu"] * d["ranks_used"] - d["value"]) < 1e-3 * d["value"] + 1e-3
    assert set(d["configs"]) == {"h36m-occ", "panoptic", "occlusion-person-8v"}
    assert d["fused_ssim_5x1x1500x1500"]["inference_ms"] > 0


def test_recorded_scaling():
    """>= 10 000 optimised H36M frames/s on 8 x B200 (north-star), and the reference arm of a multi-GPU run uses every rank."""
    _, lines = _last("r02_bench_ours.jsonl")
    multi = [l for l in lines if l["n_gpus"] > 1]
    for l in multi:
        assert "scaling_diag" in l and len(l["scaling_diag"]["kernel_ms_per_rank"]) == l["n_gpus"]
    best = {}
    for l in lines:
        best[l["n_gpus"]] = max(best.get(l["n_gpus"], 0), l["value"])
    if 8 in best:
        assert best[8] >= 10000.0
    if 1 in best:
        for n, v in best.items():
            assert v / (n * best[1]) > 0.9, (n, v, best[1])                 # weak-scaling efficiency of the recorded runs

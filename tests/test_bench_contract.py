"""The bench line the driver parses: every key of the contract is present and self-consistent in the committed output of
the last hardware run (profiles/r02_bench_50m_1gpu_final.json = `python bench.py` with defaults on one B200) and of the
CPU arm.  A format check only - the numbers themselves are the driver's to measure."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        pytest.skip(name + " not committed")
    return json.load(open(p))


def test_our_arm_line_carries_the_whole_contract():
    d = _load("r02_bench_50m_1gpu_final.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["unit"] == "queries/s" and "50M" in d["metric"] and "configs[2]" in d["config"]["workload"]
    assert "model" not in d["config"] and base.get("metric") is not None
    # value = whole-job throughput over the timed steps
    B = 4096
    assert abs(d["value"] - d["n_gpus"] * B / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == B * 768 * 4 and e["d2h_bytes_per_step"] == B * 10 * 12
    assert e["value"] != d["value"]                      # measured separately, through the host-buffer call
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    pq = r["per_query"]
    alg = (pq["d_quantized"] * pq["code_bytes"] + pq["visits"] * pq["nbr_bytes_per_visit"]) * B
    assert abs(alg - r["alg_bytes_per_launch"]) / alg < 1e-3
    assert abs(r["achieved"] - r["alg_bytes_per_launch"] / (r["avg_launch_ms"] / 1e3) / 1e9) / r["achieved"] < 1e-3
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] >= 3 * d["steps"]            # prepare + search + rerank per step
    assert d["parity"] == {"queries": 256, "tids_identical": True, "dist_bits_identical": True, "counters_identical": True}
    assert d["config"]["recall_at_10_selection_sample"] >= 0.99
    cl = d["clocks"]
    assert cl["sm_mhz"] and cl["sm_max_mhz"] and not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["operator"]["parity"]["rows_identical"] is True


def test_reference_arm_line_mirrors_it():
    d = _load("r02_bench_50m_reference_arm.json")
    o = _load("r02_bench_50m_1gpu_final.json")
    assert d["impl"] == "reference" and d["metric"] == o["metric"] and d["unit"] == o["unit"] and d["higher_is_better"] is True
    assert d["config"]["workload"] == o["config"]["workload"]
    assert (d["config"]["search_list_size"], d["config"]["rescore"]) == (o["config"]["search_list_size"], o["config"]["rescore"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0

"""The JSON lines bench.py printed on the B200 box (committed under profiles/) carry every key of the measurement
contract: metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype /
data / config.workload / clocks / gpu_launches / e2e / roofline / cpu_baseline, and the reference arm its own set."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _load(name):
    p = os.path.join(PROFILES, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not recorded yet")
    s = open(p).read()
    return json.loads(s[s.index("{"):])


def test_our_arm_line_has_the_contract_keys():
    d = _load("bench_r1_v6.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "cell_steps_per_s" and d["unit"] == "cell-steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "tripolar_quarter_degree" and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    cells = d["config"]["Nx_per_gpu"] * d["config"]["Ny_per_gpu"] * d["config"]["Nz"]
    assert d["value"] == pytest.approx(cells * 1e3 / d["ms_per_step"], rel=1e-6)       # value and ms_per_step agree
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                                                  # host round trips cannot be free
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]   # DRAM traffic >= algorithmic bytes
    assert 0 < r["frac"] < 1 and 0 < r["whole_step"]["frac"] < 1
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["unit"] == d["unit"] and b["value"] > 0 and b["sample"]


def test_reference_arm_line_has_the_contract_keys():
    d = _load("bench_reference_r1_v5.json")
    assert d["impl"] == "reference" and d["metric"] == "cell_steps_per_s" and d["unit"] == "cell-steps/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "tripolar_quarter_degree"


def test_weak_scaling_lines_are_whole_job_aggregates():
    vals = {}
    for n in (1, 2, 4, 8):
        d = _load(f"scaling_r1_v5/scale_n{n}.json")
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        cells = d["config"]["Nx_per_gpu"] * d["config"]["Ny_per_gpu"] * d["config"]["Nz"] * n
        assert d["value"] == pytest.approx(cells * 1e3 / d["ms_per_step"], rel=1e-6)
        vals[n] = d["value"]
    assert vals[8] > vals[4] > vals[2] > vals[1]

"""The committed bench lines carry every key of the bench.py contract (no GPU needed: profiles/ holds the lines the last GPU runs printed)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]


def _line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_gpu_arm_line_has_the_contract_keys():
    j = _line("r02zz_bench.json")
    assert [k for k in REQUIRED if k not in j] == []
    assert j["higher_is_better"] is False and j["unit"] == "ms" and j["n_gpus"] == 1 and j["scaling"] == "strong" and j["vs_baseline"] is None
    assert j["value"] == j["ms_per_step"] and j["warmup"] >= 3 and j["gpu_launches"] > 0
    assert "workload" in j["config"] and "model" not in j["config"] and "l2" in j["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(j["e2e"]) and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert j["e2e"]["value"] > j["value"]                                   # the end-to-end figure is not a copy of the device-timed one
    r = j["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(j["cpu_baseline"]) and j["cpu_baseline"]["kind"] in ("reference", "port")
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(j["clocks"]) and not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_matches_the_gpu_arm():
    g, r = _line("r02z_bench.json"), _line("r02z_bench_reference.json")
    assert r["impl"] == "reference" and r["config"] == g["config"] and r["metric"] == g["metric"] and r["unit"] == g["unit"]
    assert r["higher_is_better"] == g["higher_is_better"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["cores"] >= 1


def test_every_configuration_and_every_gpu_count_has_a_line():
    for c in (1, 2, 4, 5):
        j = _line(f"r02z_bench_config{c}.json")
        assert j["config"]["config_index"] == c and j["value"] > 0
    for n, pat in ((2, "r02ae_bench_c3_n2.json"), (4, "r02ag_bench_c3_n4_final_tree.json"), (8, "r02q_bench_c3_n8_stripe16.json")):
        j = _line(pat)
        assert j["n_gpus"] == n and j["passes_ms"] and j["kernels_ms"]      # per-pass and per-kernel ms at every GPU count
    assert glob.glob(os.path.join(ROOT, "profiles", "r02*_parity_*_n8.txt"))

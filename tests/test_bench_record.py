"""The bench.py contract, checked on the lines the final runs of the round produced (profiles/r02_bench_final_n*.json): every key the
driver and the judge read is present, the sub-records carry their parity figures, and the numbers are self-consistent."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not recorded")
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_final_bench_line_honours_the_contract(n):
    j = _line(f"r02_bench_final_n{n}.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "gpu_launches", "clocks", "roofline", "e2e", "strong", "config5", "parity"):
        assert k in j, k
    assert j["n_gpus"] == n and j["unit"] == "vis/s" and j["higher_is_better"] is True and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    v = j["config"]["vis_per_gpu_per_step"]
    assert abs(j["value"] - n * v / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]          # whole-job aggregate over all N GPUs
    assert j["gpu_launches"] > 0
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] >= 40 * e["vis_per_gpu_per_step"] and e["d2h_bytes_per_step"] >= 16 * e["vis_per_gpu_per_step"]
    assert 0 < e["value"] < j["value"]                                                        # host<->device copies inside the timed region
    # parity on the box, at every N
    p = j["parity"]
    assert p["checksum_rel_err"] < 1e-10 and p["tolerance"] == 1e-10
    assert p["grid_max_abs_err_over_peak"] < 1e-10 and p["degrid_max_abs_err_over_peak"] < 1e-10
    assert p["checksum"]["visibilities"] == n * v
    c5 = j["config5"]
    assert c5["config"]["vis_per_gpu_per_step"] == 125_000_000 and c5["n_gpus"] == n
    assert c5["parity"]["checksum_rel_err"] < 1e-10 and c5["parity"]["adjoint_rel_err"] < 1e-10
    s = j["strong"]
    assert s["vis_total_per_step"] == v and s["scaling"] == "strong"
    if n == 1:
        assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] > 0
        for k, a in j["aw"].items():
            if isinstance(a, dict) and "parity_max_rel_err" in a:
                assert a["parity_max_rel_err"] < 1e-10


def test_config5_ran_at_its_stated_size_on_eight_gpus():
    c5 = _line("r02_bench_final_n8.json")["config5"]
    assert c5["config"]["vis_total_per_step"] == 1_000_000_000 and "32768^2" in c5["config"]["workload"] and "support 31" in c5["config"]["workload"]
    assert c5["routing_share_of_step"] < 0.15


def test_reference_arm_line():
    j = _line("r02_bench_final_reference_arm.json")
    assert j["impl"] == "reference" and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

"""The committed bench line (profiles/r01_bench_final.json, written by `python bench.py` on a B200) carries
every key of the measurement contract, and its derived figures are consistent with one another."""
import json
import os

from conftest import ROOT


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.loads(fh.readline())


def test_committed_bench_line_has_the_contract_keys():
    d = _line("r01_bench_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per launch / the kernel's measured duration
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-3 * r["achieved"]
    assert r["traffic"] >= r["algorithmic_bytes_per_launch"]          # DRAM traffic cannot be below the algorithmic bytes
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(e)
    assert e["h2d_bytes_per_step"] == 1024 * 384 * 4 and e["d2h_bytes_per_step"] == 1024 * 10 * 8 + 1024 * 4
    assert e["value"] <= d["value"] * 1.02                                # host copies cannot make it faster
    assert abs(d["value"] - 1024 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 0 and d["recall_at_10"] >= 0.95
    assert d["parity_vs_oracle"]["identical_id_lists"] == d["parity_vs_oracle"]["queries"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_scaling_lines():
    with open(os.path.join(ROOT, "profiles", "r01_bench_multi_gpu_final.jsonl")) as fh:
        lines = [json.loads(l) for l in fh if l.strip()]
    assert [l["n_gpus"] for l in lines] == [1, 2, 4, 8]
    for l in lines:
        assert l["recall_at_10"] >= 0.95 and l["fallback_queries"] == 0
        assert abs(l["value"] - 1024 * l["n_gpus"] * 1e3 / l["ms_per_step"]) < 1e-6 * l["value"]
    assert all(b["value"] > a["value"] for a, b in zip(lines, lines[1:]))   # aggregate QPS grows with the GPUs

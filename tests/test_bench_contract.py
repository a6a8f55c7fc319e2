"""CPU: the committed bench lines (profiles/r02_bench_*.json, written by bench.py on a B200) carry every key of the bench
contract, and bench.py's argument surface is the one the driver uses.  No GPU, no compute."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _line(name):
    with open(os.path.join(PROF, name)) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json", "r02_bench_n8_cfg4.json"])
def test_b200_arm_line_has_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"].startswith("tile-timesteps/sec") and d["unit"] == "tile-timesteps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32+f64" and d["scaling"] == "strong"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["global_tiles"] == d["config"]["tiles_per_gpu"] * d["n_gpus"] or d["config"]["decomp"] == "interleaved"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_tile_step"] == 648
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= 0 and 0 < e["value"] < d["value"] * 1.05
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] - d["config"]["global_tiles"] * d["steps"] / (d["ms_per_step"] * 1e-3 * d["steps"])) < 1e-3 * d["value"]
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
        assert d["config5_casa_cnp"]["outputs_finite"] is True


def test_reference_arm_line():
    d = _line("r02_bench_ref_n1.json")
    assert d["impl"] == "reference" and d["metric"].startswith("tile-timesteps/sec") and d["unit"] == "tile-timesteps/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["global_tiles"] == _line("r02_bench_n1.json")["config"]["global_tiles"]        # the same grid as the B200 arm


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out, flag

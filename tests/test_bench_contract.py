"""bench.py's reference arm runs without a GPU: its JSON line carries the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_line():
    from oracle import refbind
    if not refbind.available("libref.so"):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         check=True, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["metric"] == "primary_mrays_per_s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    assert "workload" in line["config"]
    import bench
    assert line["config"] == bench.headline_config(1)          # the same `config` object as our own arm prints
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
def test_own_arm_line():
    """the headline line on a B200: every key of the bench contract, with sane values"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-others"],
                         check=True, capture_output=True, text=True, timeout=900).stdout.strip().splitlines()
    line = json.loads(out[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "clocks", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["metric"] == "primary_mrays_per_s" and line["steps"] == 2 and line["warmup"] >= 3 and line["n_gpus"] == 1
    assert line["value"] > 5000 and line["gpu_launches"] >= 2 * 64
    assert "configs[4]" in line["config"]["workload"] and line["config"]["resolution"] == [7680, 4320]
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert 0.05 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    e = line["e2e"]
    assert 0 < e["value"] <= line["value"] * 1.05 and e["d2h_bytes_per_step"] == 64 * 7680 * 4320 * 4 and e["h2d_bytes_per_step"] > 0
    c = line["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and "reasons" in c
    sf = e["single_frame_call"]
    assert sf["8k_maze"]["registered"]["ms_per_frame"] > 0 and sf["1080p_default_level"]["registered_with_zbuf"]["ms_per_frame"] > 0

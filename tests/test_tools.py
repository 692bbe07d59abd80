"""The measurement tools that read the kernel source: their anchors must exist in the source as it is."""
import os
import sys

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_ncu_phases_markers_exist():
    """tools/ncu_phases.py groups the ncu source page by marker comments of trace_kernel.cuh: every marker is found, the
    phases come in source order and cover the kernel's body"""
    import ncu_phases
    rg = ncu_phases.ranges()
    names = [n for _, _, n in rg]
    for n, _ in ncu_phases.MARKS + ncu_phases.HELPERS:
        assert n in names, n
    starts = [a for a, _, _ in rg]
    assert starts == sorted(starts)
    order = [n for _, _, n in rg if n in dict(ncu_phases.MARKS)]
    assert order == [n for n, _ in ncu_phases.MARKS], order


def test_bench_config_is_shared_by_both_arms():
    """bench.py: the reference arm prints the same `config` object as our arm, the workload is BASELINE configs[4]"""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.headline_config(1)
    assert "configs[4]" in c["workload"] and c["resolution"] == [7680, 4320] and c["frames_per_step"] == 64
    assert bench.headline_config(8)["sharding"] != c["sharding"]
    assert bench.W_BLUR == 4 * (1 + 2 * (3 + 2 + 2 + 1))

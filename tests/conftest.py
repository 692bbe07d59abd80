import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement (oracle/pwn_oracle.c) -- the checker."""
    from oracle import orcbind
    return orcbind.Oracle()


@pytest.fixture(scope="session")
def refs():
    """oracle/_ref variants (the reference's own code compiled headless), or skip."""
    from oracle import refbind
    if not refbind.available("libref_ids.so"):
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")
    cache = {}

    def get(variant):
        if variant not in cache:
            cache[variant] = refbind.Ref(variant)
        return cache[variant]
    return get


@pytest.fixture(scope="session")
def default_level():
    from pwnfps_b200 import levels
    return levels.default_level()


@pytest.fixture(scope="session")
def tracer():
    """The product: libpwntrace.so on cuda:0 with this host's rcpps/rsqrtps tables."""
    from pwnfps_b200 import api
    if not has_gpu():
        pytest.skip("no GPU")
    t = api.Tracer(0)
    yield t
    t.close()


def oracle_set_level(oracle, lv):
    oracle.level_set(lv.cells, lv.portals_i32())
    if len(lv.spheres):
        sp = np.zeros((len(lv.spheres), 10), np.float32)
        sp[:, 0] = lv.spheres["r"]
        sp[:, 1] = lv.spheres["refl"]
        sp[:, 2:6] = lv.spheres["pos"]
        sp[:, 6:10] = lv.spheres["col"]
        oracle.spheres_set(sp)


def rgb_diff(a, b):
    """per-pixel max |delta| over the B,G,R bytes (alpha is junk-but-deterministic)."""
    a8 = np.ascontiguousarray(a).view(np.uint8).reshape(a.shape + (4,))[..., :3].astype(np.int16)
    b8 = np.ascontiguousarray(b).view(np.uint8).reshape(b.shape + (4,))[..., :3].astype(np.int16)
    return np.abs(a8 - b8).max(axis=-1)


def pack_edge_cols():
    """BGRA quadruples that exercise col_ftoint (util.h:48-59): ties to even, saturation, negatives,
    NaN, inf, |255 v| >= 2^31 (cvtps2dq's 0x80000000), denormals"""
    rng = np.random.default_rng(8)
    lanes = np.concatenate([
        np.array([0.0, -0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 255.5 / 255, 1.0, 1.0000001, 2.0, 30.0, -1.0, -1e30,
                  1e30, 8421504.0, 8421505.0, 8421506.0, 2147483648.0 / 255, 3e9, np.inf, -np.inf, np.nan, 1e-40, -1e-40], np.float32),
        (np.arange(0, 600, dtype=np.float32) * 0.5 / 255).astype(np.float32),
        rng.uniform(-0.2, 1.3, 4000).astype(np.float32),
        (rng.uniform(0, 1, 400) * 2.0 ** rng.integers(-10, 40, 400)).astype(np.float32)])
    lanes = np.resize(lanes, (len(lanes) + 3) // 4 * 4)
    return np.stack([lanes, np.roll(lanes, 1), np.roll(lanes, 7), np.roll(lanes, 13)], axis=1).astype(np.float32)

"""The C-ABI library loads and exports every symbol include/pwntrace.h declares; without a
GPU it refuses loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from pwnfps_b200 import api


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pwntrace.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pwn_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = api.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libpwntrace.so lacks %s" % n
    assert sorted(api.ABI_SYMBOLS) == names
    assert L.pwn_abi_version() == 1


def test_struct_layouts():
    assert api.IDS_DTYPE.itemsize == 16
    assert api.PORTAL_DTYPE.itemsize == 28
    assert api.SPHERE_DTYPE.itemsize == 40
    assert api.POSE_DTYPE.itemsize == 68


@pytest.mark.skipif(has_gpu(), reason="GPU present")
def test_no_gpu_fails_loudly():
    with pytest.raises(api.PwnError) as e:
        api.Tracer(0)
    assert e.value.code == -3
    assert "no CPU path" in str(e.value) or "device" in str(e.value)


def test_null_arguments_are_errors():
    L = api.lib()
    assert L.pwn_create(None, 0) == -1
    assert L.pwn_sync(None) == -1
    assert L.pwn_shard_config(None, 1, 0) == -1
    assert b"null" in L.pwn_last_error()


def test_host_math_tables_reproduce_the_cpu(oracle):
    """pwn_host_math_probe's tables + the device formula == this CPU's rcpps/rsqrtps."""
    rb, sb, rt, st = api.host_math_probe()
    assert 8 <= rb <= 23 and 8 <= sb <= 23
    rng = np.random.RandomState(3)
    x = np.concatenate([rng.uniform(1e-13, 4.0, 20000), np.exp(rng.uniform(-30, 30, 20000)),
                        [1e-13, 1.0, 2.0, 4.0, 0.5, 38.0 * 38.0 + 1.0]]).astype(np.float32)
    u = x.view(np.uint32)
    e = (u >> 23).astype(np.int64) - 127
    # rcp_emul (trace_kernel.cuh)
    t = rt[(u & 0x7FFFFF) >> (23 - rb)].astype(np.int64)
    rcp = ((t - (e << 23)) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    # rsqrt_emul
    p = e & 1
    idx = ((u & 0x7FFFFF) >> (23 - sb)).astype(np.int64) + (p << sb)
    t = st[idx].astype(np.int64)
    rsq = ((t - (((e - p) >> 1) << 23)) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
    want_rcp = np.array([oracle.lib.orc_kat_rcp(float(v)) for v in x], np.float32)
    want_rsq = np.array([oracle.lib.orc_kat_rsqrt(float(v)) for v in x], np.float32)
    assert (rcp.view(np.uint32) == want_rcp.view(np.uint32)).all()
    assert (rsq.view(np.uint32) == want_rsq.view(np.uint32)).all()


def test_ray_setup_matches_oracle(oracle):
    rng = np.random.RandomState(5)
    for (w, h) in [(320, 240), (320, 200), (1920, 1080), (3840, 2160), (7680, 4320), (333, 77), (17, 1000)]:
        cam = np.eye(4, dtype=np.float32)
        cam[:3, :3] = rng.uniform(-1, 1, (3, 3))
        cam[3, :3] = rng.uniform(0, 30, 3)
        got = api.ray_setup(cam.reshape(16), w, h)
        want = oracle.ray_setup(cam.reshape(16), w, h)
        for g, x in zip(got, want):
            assert (g.view(np.uint32) == x.view(np.uint32)).all()


def libm_samples():
    rng = np.random.RandomState(11)
    x = np.concatenate([rng.uniform(-130, 130, 400000), rng.uniform(-3e4, 3e4, 200000), rng.uniform(-0.9, 0.9, 100000),
                        np.exp(rng.uniform(-12, 14, 100000)) * rng.choice([-1, 1], 100000), -rng.uniform(0, 100, 200000),
                        [0.0, -0.0, 1e-6, -1e-6, 0.78539816, 120.0, -120.0, 119.99999, 2.0**20, -87.0, -88.5, -103.0, -104.5]]).astype(np.float32)
    # plus the inputs on which the two glibc builds round differently
    d = np.array([0x4255b0a9, 0x418a3adb, 0xc27c65d9], np.uint32).view(np.float32)
    return np.concatenate([x, d])


def test_libm_restatement_matches_host_glibc(oracle):
    """csrc/pwn_libm.h (host build of the text the device runs) == this host's sinf/cosf/expf."""
    fma_build, exact = api.host_libm_probe()
    assert exact, "host libm is neither glibc build restated in pwn_libm.h"
    x = libm_samples()
    s, c, e = api.host_libm_eval(x, fma_build)
    hs, hc, he = oracle.libm(x)
    assert (s.view(np.uint32) == hs.view(np.uint32)).all()
    assert (c.view(np.uint32) == hc.view(np.uint32)).all()
    neg = (x <= 0) & (x >= -87)          # the trace passes -0.6*fog; below -88 every result blends to 1
    assert (e.view(np.uint32)[neg] == he.view(np.uint32)[neg]).all()
    tiny = x < -88
    assert (e[tiny] < 1e-37).all() and (he[tiny] < 1e-37).all()
    # the other build differs on the discriminating inputs, so the probe is meaningful
    s2, c2, e2 = api.host_libm_eval(x[-3:], 1 - fma_build)
    assert (s2.view(np.uint32)[0] != hs.view(np.uint32)[-3]) or (c2.view(np.uint32)[1] != hc.view(np.uint32)[-2])


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps each exported entry point to the reference interface it replaces"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = [s for s in api.ABI_SYMBOLS if s not in doc]
    assert not missing, missing

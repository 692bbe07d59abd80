"""oracle/refbind.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libref*.so.

The libraries are the reference's own hot path compiled headless by
oracle/make_ref.py (see ref_harness.c).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

IDS_DTYPE = np.dtype([("cell", "<i4"), ("ldir", "<i2"), ("kind", "<i2"),
                      ("nstep", "<i4"), ("nportal", "<i4")])

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available(variant="libref.so"):
    return os.path.isfile(os.path.join(REFDIR, variant))


class Ref:
    """One loaded variant.  Each variant owns one static level (level.h:83)."""

    def __init__(self, variant="libref_noblur.so"):
        path = os.path.join(REFDIR, variant)
        if not os.path.isfile(path):
            raise FileNotFoundError("%s missing: run python oracle/make_ref.py" % path)
        lib = C.CDLL(path)
        self.lib = lib
        lib.ref_level_load.argtypes = [C.c_char_p]
        lib.ref_level_set.argtypes = [C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_int]
        lib.ref_level_get.argtypes = [_i32p, _i32p]
        lib.ref_level_spawn.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_obj_sphere.argtypes = [C.c_float] * 8
        lib.ref_trace_frame.restype = C.c_double
        lib.ref_trace_frame.argtypes = [_f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_time_frames.restype = C.c_double
        lib.ref_time_frames.argtypes = [_f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_randi.restype = C.c_uint32
        lib.ref_randi.argtypes = [C.POINTER(C.c_uint32)]
        lib.ref_randfu.restype = C.c_float
        lib.ref_randfu.argtypes = [C.POINTER(C.c_uint32)]
        lib.ref_randfs.restype = C.c_float
        lib.ref_randfs.argtypes = [C.POINTER(C.c_uint32)]
        lib.ref_col_ftoint.restype = C.c_uint32
        lib.ref_col_ftoint.argtypes = [_f32p]
        lib.ref_v_normalise.argtypes = [_f32p, _f32p]
        lib.ref_v_dot.restype = C.c_float
        lib.ref_v_dot.argtypes = [_f32p, _f32p]
        lib.ref_mat4_roty.argtypes = [_f32p, C.c_float]
        lib.ref_rcpps.argtypes = [_f32p, _f32p, C.c_size_t]
        lib.ref_rsqrtps.argtypes = [_f32p, _f32p, C.c_size_t]
        lib.ref_libm.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t]
        if hasattr(lib, "ref_player_tick"):
            lib.ref_player_tick.argtypes = [_f32p, _f32p, _i32p, C.c_float]
        if hasattr(lib, "ref_screen_upscale"):
            lib.ref_screen_upscale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.flags = lib.ref_flags()
        self.has_blur = bool(self.flags & 1)
        self.has_ids = bool(self.flags & 2)
        self.wide = bool(self.flags & 4)
        self.dim = lib.ref_dim()
        self.pm_max = lib.ref_pm_max()

    # -- level ---------------------------------------------------------------
    def level_load(self, path):
        if self.lib.ref_level_load(os.fsencode(path)) != 0:
            raise RuntimeError("ref level_load failed: %s" % path)

    def level_set(self, cells, portals, sx, sz):
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        assert cells.shape == (self.dim, self.dim), cells.shape
        pm = np.ascontiguousarray(portals, dtype=np.int32).reshape(-1, 7)
        if self.lib.ref_level_set(self.dim, cells, pm.shape[0], pm if pm.size else np.zeros((1, 7), np.int32), sx, sz) != 0:
            raise RuntimeError("ref_level_set failed")

    def level_get(self):
        cells = np.zeros((self.dim, self.dim), np.int32)
        pm = np.zeros((self.pm_max, 7), np.int32)
        if self.lib.ref_level_get(cells, pm) != 0:
            raise RuntimeError("no level")
        return cells, pm

    def spawn(self):
        sx, sz = C.c_int(), C.c_int()
        self.lib.ref_level_spawn(C.byref(sx), C.byref(sz))
        return sx.value, sz.value

    def objs_clear(self):
        self.lib.ref_objs_clear()

    def obj_sphere(self, r, refl, x, y, z, cb, cg, cr):
        if self.lib.ref_obj_sphere(r, refl, x, y, z, cb, cg, cr) != 0:
            raise RuntimeError("ref_obj_sphere failed")

    # -- frames --------------------------------------------------------------
    def trace_frame(self, cam16, sec, w, h, nthreads=0, want_ids=False):
        cam = np.ascontiguousarray(cam16, dtype=np.float32).reshape(16)
        sbuf = np.zeros((h, w), np.uint32)
        zbuf = np.zeros((h, w), np.float32)
        ids = None
        idp = None
        if want_ids:
            assert self.has_ids
            ids = np.zeros((h, w, 3), IDS_DTYPE)
            idp = ids.ctypes.data
        t = self.lib.ref_trace_frame(cam, float(sec), w, h, nthreads,
                                     sbuf.ctypes.data, zbuf.ctypes.data, idp)
        if t < 0:
            raise RuntimeError("ref_trace_frame failed (%g)" % t)
        return sbuf, zbuf, ids, t

    def time_frames(self, cam16, sec, w, h, nthreads, nframes):
        cam = np.ascontiguousarray(cam16, dtype=np.float32).reshape(16)
        t = self.lib.ref_time_frames(cam, float(sec), w, h, nthreads, nframes)
        if t < 0:
            raise RuntimeError("ref_time_frames failed (%g)" % t)
        return t

    def player_tick(self, cam16, gravity4, keys8, tdiff):
        """main.c:188-379 on (cam, gravity) in place; keys8 = turnleft turnright turnup turndown forward back left right"""
        if self.lib.ref_player_tick(cam16, gravity4, np.ascontiguousarray(keys8, dtype=np.int32), float(tdiff)) != 0:
            raise RuntimeError("ref_player_tick: no level loaded")

    def set_fpenv(self, ftz_daz):
        """MXCSR FTZ|DAZ on (what crtfastmath.o gives the shipped executable) or off, on this thread and the OpenMP workers"""
        return self.lib.ref_set_fpenv(1 if ftz_daz else 0)

    def max_threads(self):
        return self.lib.ref_max_threads()

    def screen_upscale(self, frame, rscale, pitch_words=None):
        """the reference's screen_upscale (screen.h:126-149) of a (h, w) uint32 frame"""
        frame = np.ascontiguousarray(frame, dtype=np.uint32)
        h, w = frame.shape
        pitch_words = w * rscale if pitch_words is None else pitch_words
        out = np.zeros((h * rscale, pitch_words), np.uint32)
        self.lib.ref_screen_upscale(frame.ctypes.data, w, h, rscale, pitch_words * 4, out.ctypes.data)
        return out

    # -- known-answer helpers ---------------------------------------------------
    def rcpps(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.lib.ref_rcpps(x, out, x.size)
        return out

    def rsqrtps(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.lib.ref_rsqrtps(x, out, x.size)
        return out

    def libm(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        s, c, e = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self.lib.ref_libm(x, s, c, e, x.size)
        return s, c, e

    def randfs_seq(self, seed, n):
        s = C.c_uint32(seed)
        out = np.empty(n, np.float32)
        st = np.empty(n, np.uint32)
        for i in range(n):
            out[i] = self.lib.ref_randfs(C.byref(s))
            st[i] = s.value
        return out, st

    def col_ftoint(self, v4):
        return self.lib.ref_col_ftoint(np.ascontiguousarray(v4, dtype=np.float32))

    def v_normalise(self, v4):
        out = np.empty(4, np.float32)
        self.lib.ref_v_normalise(np.ascontiguousarray(v4, dtype=np.float32), out)
        return out

    def mat4_roty(self, m16, ang):
        m = np.ascontiguousarray(m16, dtype=np.float32).reshape(16).copy()
        self.lib.ref_mat4_roty(m, float(ang))
        return m


def spawn_cam(sx, sz):
    """cam as main.c:60-64: identity, pos = spawn + 0.5, y = 0.5."""
    cam = np.eye(4, dtype=np.float32)
    cam[3, 0] = 0.5 + sx
    cam[3, 1] = 0.5
    cam[3, 2] = 0.5 + sz
    return cam.reshape(16)

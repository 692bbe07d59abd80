"""oracle/orcbind.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/libpwnoracle.so
(the CPU restatement in pwn_oracle.c).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpwnoracle.so")
LIB_STRICT = os.path.join(HERE, "libpwnoracle_strict.so")

IDS_DTYPE = np.dtype([("cell", "<i4"), ("ldir", "<i2"), ("kind", "<i2"),
                      ("nstep", "<i4"), ("nportal", "<i4")])

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    src = os.path.join(HERE, "pwn_oracle.c")
    if all(os.path.isfile(p) and os.path.getmtime(p) >= os.path.getmtime(src) for p in (LIB, LIB_STRICT)):
        return
    subprocess.run(["make", "-C", HERE, "libpwnoracle.so", "libpwnoracle_strict.so"], check=True,
                   stdout=subprocess.DEVNULL)


class Oracle:
    """strict=False: randfs as gcc folds it under the reference's shipped -ffast-math (what the shipped
    build and the device compute).  strict=True: the source as written = the reference built without
    -ffast-math."""

    def __init__(self, strict=False):
        build()
        lib = C.CDLL(LIB_STRICT if strict else LIB)
        self.lib = lib
        lib.orc_level_set.argtypes = [C.c_int, C.c_int, _i32p, C.c_int, _i32p]
        lib.orc_spheres_set.argtypes = [C.c_int, C.c_void_p]
        lib.orc_config.argtypes = [C.c_int, C.c_int]
        lib.orc_ray_setup.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        lib.orc_trace_rows.argtypes = [_f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_blur.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_kat_randfs.restype = C.c_float
        lib.orc_kat_randfs.argtypes = [C.POINTER(C.c_uint32)]
        lib.orc_kat_col_ftoint.restype = C.c_uint32
        lib.orc_kat_col_ftoint.argtypes = [_f32p]
        lib.orc_kat_rcp.restype = C.c_float
        lib.orc_kat_rcp.argtypes = [C.c_float]
        lib.orc_kat_rsqrt.restype = C.c_float
        lib.orc_kat_rsqrt.argtypes = [C.c_float]
        lib.orc_kat_libm.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t]
        lib.orc_upscale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.dims = None
        # the strict build of the same restatement, kept in step (level, spheres, config) by this one
        self.twin = None if strict else Oracle(strict=True)

    def level_set(self, cells, portals):
        if self.twin is not None:
            self.twin.level_set(cells, portals)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        h, w = cells.shape
        pm = np.ascontiguousarray(portals, dtype=np.int32).reshape(-1, 7)
        rc = self.lib.orc_level_set(w, h, cells, pm.shape[0], pm if pm.size else np.zeros((1, 7), np.int32))
        if rc != 0:
            raise RuntimeError("orc_level_set failed (%d)" % rc)
        self.dims = (w, h)
        self.spheres_set(None)

    def spheres_set(self, sph):
        """sph: (n,10) float32 rows r refl x y z w b g r a, or None."""
        if self.twin is not None:
            self.twin.spheres_set(sph)
        if sph is None or len(sph) == 0:
            self.lib.orc_spheres_set(0, None)
            return
        sph = np.ascontiguousarray(sph, dtype=np.float32).reshape(-1, 10)
        self.lib.orc_spheres_set(sph.shape[0], sph.ctypes.data)

    def config(self, reflect=2, reflect_blur=True):
        if self.twin is not None:
            self.twin.config(reflect, reflect_blur)
        self.lib.orc_config(reflect, 1 if reflect_blur else 0)

    def ray_setup(self, cam16, dimx, dimy):
        cam = np.ascontiguousarray(cam16, dtype=np.float32).reshape(16)
        rb, rx, ry = (np.zeros(4, np.float32) for _ in range(3))
        self.lib.orc_ray_setup(cam, dimx, dimy, rb, rx, ry)
        return rb, rx, ry

    def trace_frame(self, cam16, sec, w, h, want_ids=False, blur=False, nthreads=0,
                    y0=0, y1=None, rwidth=None, out=None):
        cam = np.ascontiguousarray(cam16, dtype=np.float32).reshape(16)
        rwidth = w if rwidth is None else rwidth
        y1 = h if y1 is None else y1
        if out is None:
            sbuf = np.zeros((h, rwidth), np.uint32)
            zbuf = np.zeros((h, rwidth), np.float32)
        else:
            sbuf, zbuf = out
        ids = None
        idp = None
        if want_ids:
            ids = np.zeros((h, rwidth, 3), IDS_DTYPE)
            ids["cell"] = -1
            ids["ldir"] = -2
            ids["kind"] = -1
            idp = ids.ctypes.data
        rc = self.lib.orc_trace_rows(cam, float(sec), w, h, rwidth, y0, y1,
                                     sbuf.ctypes.data, zbuf.ctypes.data, idp, nthreads)
        if rc != 0:
            raise RuntimeError("orc_trace_rows failed (%d)" % rc)
        if blur:
            ts = sbuf.copy()
            self.lib.orc_blur(w, h, rwidth, y0, y1, sbuf.ctypes.data, zbuf.ctypes.data, ts.ctypes.data, nthreads)
        return sbuf, zbuf, ids

    def blur(self, sbuf, zbuf, w, h, rwidth=None, y0=0, y1=None, nthreads=0):
        rwidth = w if rwidth is None else rwidth
        y1 = h if y1 is None else y1
        ts = np.ascontiguousarray(sbuf).copy()
        out = ts.copy()
        zb = np.ascontiguousarray(zbuf, dtype=np.float32)
        self.lib.orc_blur(w, h, rwidth, y0, y1, out.ctypes.data, zb.ctypes.data, ts.ctypes.data, nthreads)
        return out

    def upscale(self, frame, rscale, pitch_words=None, w=None, h=None):
        """screen_upscale (screen.h:126-149) of the top-left (h, w) of `frame`"""
        frame = np.ascontiguousarray(frame, dtype=np.uint32)
        h = frame.shape[0] if h is None else h
        w = frame.shape[1] if w is None else w
        pitch_words = w * rscale if pitch_words is None else pitch_words
        out = np.zeros((h * rscale, pitch_words), np.uint32)
        if self.lib.orc_upscale(frame.ctypes.data, w, h, frame.shape[1], rscale, pitch_words, out.ctypes.data) != 0:
            raise RuntimeError("orc_upscale failed")
        return out

    def libm(self, x):
        """the host's glibc sinf, cosf, expf (what the reference calls, trace.h:42-46,97)"""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        s, c, e = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self.lib.orc_kat_libm(x, s, c, e, x.size)
        return s, c, e

    def randfs_seq(self, seed, n):
        s = C.c_uint32(seed)
        out = np.empty(n, np.float32)
        st = np.empty(n, np.uint32)
        for i in range(n):
            out[i] = self.lib.orc_kat_randfs(C.byref(s))
            st[i] = s.value
        return out, st

    def col_ftoint(self, v4):
        return self.lib.orc_kat_col_ftoint(np.ascontiguousarray(v4, dtype=np.float32))

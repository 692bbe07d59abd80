"""pwnfps_b200.api -- ctypes binding of the C ABI (include/pwntrace.h) and of the C host
library (pwnfps_b200/host), with the reference's own names for the path:

    level_load(fname)                 level.h:107      -> Level
    Tracer.level_prepare_render(lv)   level.h:64       (uploads the spheres)
    Tracer.trace_screen_centred(lv, x1, y1, x2, y2, cam)   screen.h:31

Everything here forwards to libpwntrace.so; there is no Python or CPU
implementation of the trace.  Importing the module does not need a GPU; creating
a Tracer does, and raises PwnError when no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

PWN_OK = 0
PWN_F_BLUR = 1
PWN_F_IDS = 2
PWN_F_NO_REFLECT_BLUR = 4
PWN_F_LOCAL_BANDS = 8

FXP, FZP, FXN, FZN, FYP, FYN = range(6)

IDS_DTYPE = np.dtype([("cell", "<i4"), ("ldir", "<i2"), ("kind", "<i2"),
                      ("nstep", "<i4"), ("nportal", "<i4")])
PORTAL_DTYPE = np.dtype([("x1", "<i4"), ("z1", "<i4"), ("x2", "<i4"), ("z2", "<i4"),
                         ("rot12", "<i4"), ("c1", "<i4"), ("c2", "<i4")])
SPHERE_DTYPE = np.dtype([("r", "<f4"), ("refl", "<f4"), ("pos", "<f4", (4,)), ("col", "<f4", (4,))])
POSE_DTYPE = np.dtype([("cam", "<f4", (16,)), ("sec_current", "<f4")])


class PwnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pwntrace error %d: %s" % (code, msg))
        self.code = code


def portal_code(p):
    """PWN_PORTAL_CODE: cell code naming portal pair p."""
    return ord("A") + p if p < 26 else 0x100 + p


# ---------------------------------------------------------------------------------
_lib = None
_host = None

# every symbol include/pwntrace.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "pwn_abi_version", "pwn_last_error", "pwn_create", "pwn_destroy",
    "pwn_math_tables", "pwn_math_tables_from_host", "pwn_host_math_probe",
    "pwn_host_libm_probe", "pwn_libm_config", "pwn_host_libm_eval", "pwn_debug_libm",
    "pwn_level_upload", "pwn_parts_upload", "pwn_frame_config", "pwn_frame_clear",
    "pwn_shard_config", "pwn_trace_frame", "pwn_blur_frame", "pwn_trace_cam", "pwn_ray_setup",
    "pwn_sync", "pwn_read_frame", "pwn_trace_screen", "pwn_trace_poses",
    "pwn_frame_device_ptrs", "pwn_frame_redirect", "pwn_frame_ipc_export",
    "pwn_frame_ipc_import", "pwn_frame_ipc_close", "pwn_host_alloc", "pwn_host_free",
    "pwn_launch_count", "pwn_last_opcount", "pwn_last_kernel_ms", "pwn_time_frames", "pwn_fp32_peak", "pwn_stream",
    "pwn_screen_upscale", "pwn_surface_device_ptr", "pwn_last_upscale_ms", "pwn_trace_config",
    "pwn_frame_peers", "pwn_trace_shard_blur", "pwn_frame_peers_status",
    "pwn_host_register", "pwn_host_unregister", "pwn_debug_pack", "pwn_write_frame", "pwn_read_bands",
    "pwn_math_status", "pwn_screen_chunks", "pwn_shard_layout", "pwn_debug_flags", "pwn_shard_blocks", "pwn_join", "pwn_block_shape",
]


def lib():
    """libpwntrace.so (built in-tree on first use)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PWN_LIBPWNTRACE") or _build.build()   # the override is for A/B timing of experimental builds
    L = C.CDLL(path)
    vp, i, f, u = C.c_void_p, C.c_int, C.c_float, C.c_uint
    u64p = C.POINTER(C.c_uint64)
    L.pwn_last_error.restype = C.c_char_p
    L.pwn_create.argtypes = [C.POINTER(vp), i]
    L.pwn_destroy.argtypes = [vp]
    L.pwn_math_tables.argtypes = [vp, vp, i, vp, i]
    L.pwn_math_tables_from_host.argtypes = [vp]
    L.pwn_host_math_probe.argtypes = [C.POINTER(i), C.POINTER(i), vp, C.c_size_t, vp, C.c_size_t]
    L.pwn_host_libm_probe.argtypes = [C.POINTER(i), C.POINTER(i)]
    L.pwn_libm_config.argtypes = [vp, i]
    L.pwn_host_libm_eval.argtypes = [vp, vp, vp, vp, C.c_size_t, i]
    L.pwn_host_libm_eval.restype = None
    L.pwn_debug_libm.argtypes = [vp, vp, vp, vp, vp, i]
    L.pwn_level_upload.argtypes = [vp, i, i, vp, i, vp]
    L.pwn_parts_upload.argtypes = [vp, i, vp]
    L.pwn_frame_config.argtypes = [vp, i, i, i]
    L.pwn_frame_clear.argtypes = [vp]
    L.pwn_shard_config.argtypes = [vp, i, i]
    L.pwn_trace_frame.argtypes = [vp, vp, vp, vp, vp, i, i, f, u]
    L.pwn_blur_frame.argtypes = [vp, i, i]
    L.pwn_trace_cam.argtypes = [vp, vp, i, i, f, u]
    L.pwn_ray_setup.argtypes = [vp, i, i, vp, vp, vp]
    L.pwn_ray_setup.restype = None
    L.pwn_sync.argtypes = [vp]
    L.pwn_read_frame.argtypes = [vp, vp, vp, vp]
    L.pwn_trace_screen.argtypes = [vp, vp, i, i, f, u, vp, vp]
    L.pwn_trace_poses.argtypes = [vp, i, vp, i, i, u, vp, vp]
    L.pwn_frame_device_ptrs.argtypes = [vp, u64p, u64p, u64p]
    L.pwn_frame_redirect.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.pwn_frame_ipc_export.argtypes = [vp, vp, u64p, u64p, u64p]
    L.pwn_frame_ipc_import.argtypes = [vp, vp, u64p]
    L.pwn_frame_ipc_close.argtypes = [vp, C.c_uint64]
    L.pwn_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.pwn_host_free.argtypes = [vp]
    L.pwn_launch_count.argtypes = [vp, u64p]
    L.pwn_last_opcount.argtypes = [vp, u64p]
    L.pwn_last_kernel_ms.argtypes = [vp, C.POINTER(f), C.POINTER(f)]
    L.pwn_time_frames.argtypes = [vp, vp, i, i, f, u, i, C.POINTER(f)]
    L.pwn_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.pwn_stream.argtypes = [vp, u64p]
    L.pwn_screen_upscale.argtypes = [vp, i, i, i, i, vp]
    L.pwn_surface_device_ptr.argtypes = [vp, u64p, u64p]
    L.pwn_last_upscale_ms.argtypes = [vp, C.POINTER(f)]
    L.pwn_trace_config.argtypes = [vp, i]
    L.pwn_frame_peers.argtypes = [vp, i, u64p]
    L.pwn_host_register.argtypes = [vp, C.c_size_t]
    L.pwn_debug_pack.argtypes = [vp, vp, vp, i]
    L.pwn_write_frame.argtypes = [vp, vp, vp]
    L.pwn_read_bands.argtypes = [vp, i, i, vp]
    L.pwn_host_unregister.argtypes = [vp]
    L.pwn_trace_shard_blur.argtypes = [vp, vp, i, i, f, u]
    L.pwn_frame_peers_status.argtypes = [vp, C.POINTER(i)]
    L.pwn_math_status.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.pwn_screen_chunks.argtypes = [vp, i]
    L.pwn_shard_layout.argtypes = [vp, i, i]
    L.pwn_debug_flags.argtypes = [vp, vp, i]
    L.pwn_shard_blocks.argtypes = [vp, i, vp]
    L.pwn_join.argtypes = [vp]
    L.pwn_block_shape.argtypes = [i, i, i, vp, i, vp]
    _lib = L
    return L


class _CLevel(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("cells", C.POINTER(C.c_int32)),
                ("nportals", C.c_int32), ("portals", C.c_void_p), ("sx", C.c_int32), ("sz", C.c_int32)]


def host():
    """libpwnhost.so: the C host side (level loader, camera script)."""
    global _host
    if _host is not None:
        return _host
    H = C.CDLL(_build.build_host())
    H.pwn_level_new.restype = C.POINTER(_CLevel)
    H.pwn_level_new.argtypes = [C.c_int, C.c_int, C.c_int]
    H.pwn_level_free.argtypes = [C.POINTER(_CLevel)]
    H.pwn_level_load.restype = C.POINTER(_CLevel)
    H.pwn_level_load.argtypes = [C.c_char_p]
    H.pwn_level_parse.restype = C.POINTER(_CLevel)
    H.pwn_level_parse.argtypes = [C.c_char_p, C.c_size_t]
    H.pwn_level_link.argtypes = [C.POINTER(_CLevel)]
    H.pwn_level_derive.argtypes = [C.POINTER(_CLevel)]
    H.pwn_level_save_txt.argtypes = [C.POINTER(_CLevel), C.c_char_p]
    H.pwn_level_save_bin.argtypes = [C.POINTER(_CLevel), C.c_char_p]
    H.pwn_level_save_txtx.argtypes = [C.POINTER(_CLevel), C.c_char_p]
    H.pwn_level_load_bin.restype = C.POINTER(_CLevel)
    H.pwn_level_load_bin.argtypes = [C.c_char_p]
    H.pwn_game_init.argtypes = [C.c_void_p]
    H.pwn_game_init.restype = None
    H.pwn_game_tick.argtypes = [C.c_void_p, C.POINTER(_CLevel), C.c_double, C.c_double]
    H.pwn_game_tick.restype = None
    H.pwn_cam_spawn.argtypes = [C.c_int, C.c_int, C.c_void_p]
    H.pwn_cam_spawn.restype = None
    H.pwn_mat4_roty.argtypes = [C.c_void_p, C.c_float]
    H.pwn_mat4_roty.restype = None
    H.pwn_player_tick.argtypes = [C.c_void_p, C.POINTER(_CLevel), C.c_void_p, C.c_float]
    H.pwn_player_tick.restype = None
    H.pwn_flythrough.argtypes = [C.POINTER(_CLevel), C.c_int, C.c_void_p]
    H.pwn_flythrough.restype = C.c_int
    H.pwn_save_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    H.pwn_frame_diff.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    _host = H
    return H


def save_ppm(path, sbuf, w=None, h=None):
    """dump a (rheight, rwidth) uint32 frame (0xAARRGGBB) as a binary PPM"""
    sbuf = np.ascontiguousarray(sbuf, dtype=np.uint32)
    h = sbuf.shape[0] if h is None else h
    w = sbuf.shape[1] if w is None else w
    rc = host().pwn_save_ppm(os.fsencode(path), sbuf.ctypes.data, w, h, sbuf.shape[1])
    if rc != 0:
        raise OSError("pwn_save_ppm(%s) failed (%d)" % (path, rc))


def frame_diff(a, b, gain=64):
    """per colour byte |a - b| * gain, saturated: a frame that shows where two frames differ"""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    assert a.shape == b.shape
    out = np.zeros_like(a)
    if host().pwn_frame_diff(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[1], a.shape[0], a.shape[1], gain) != 0:
        raise ValueError("pwn_frame_diff failed")
    return out


def host_register(arr):
    """page-lock a numpy array in place (pwn_host_register); pair with host_unregister"""
    _check(lib().pwn_host_register(arr.ctypes.data, arr.nbytes))


def host_unregister(arr):
    _check(lib().pwn_host_unregister(arr.ctypes.data))


def _check(rc):
    if rc != PWN_OK:
        raise PwnError(rc, lib().pwn_last_error().decode("utf-8", "replace"))


# ---------------------------------------------------------------------------------
class Level:
    """The reference's `level` as the trace reads it (defs.h:96-111): cells + portals + spawn."""

    def __init__(self, cells, portals, sx=0, sz=0):
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        assert self.cells.ndim == 2
        self.portals = np.ascontiguousarray(portals, dtype=PORTAL_DTYPE).reshape(-1)
        self.sx, self.sz = int(sx), int(sz)
        self.spheres = np.zeros(0, SPHERE_DTYPE)   # lv->objs (defs.h:98)

    @property
    def h(self):
        return self.cells.shape[0]

    @property
    def w(self):
        return self.cells.shape[1]

    def portals_i32(self):
        """(n,7) int32 view: x1 z1 x2 z2 rot12 c1 c2."""
        return self.portals.view(np.int32).reshape(-1, 7)

    @classmethod
    def _from_c(cls, p):
        if not p:
            raise PwnError(-1, "level load failed")
        c = p.contents
        cells = np.ctypeslib.as_array(c.cells, shape=(c.h, c.w)).copy()
        n = c.nportals
        pm = np.zeros(n, PORTAL_DTYPE)
        if n:
            C.memmove(pm.ctypes.data, c.portals, n * PORTAL_DTYPE.itemsize)
        lv = cls(cells, pm, c.sx, c.sz)
        host().pwn_level_free(p)
        return lv

    @classmethod
    def from_grid(cls, cells, nportals, sx=0, sz=0):
        """A level built in memory: portal pairing + rot12/c1/c2 derived as the loader does."""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        h, w = cells.shape
        H = host()
        p = H.pwn_level_new(w, h, nportals)
        C.memmove(p.contents.cells, cells.ctypes.data, cells.nbytes)
        p.contents.sx, p.contents.sz = sx, sz
        if H.pwn_level_link(p) != 0:
            H.pwn_level_free(p)
            raise PwnError(-1, "a portal cell names a pair outside the table")
        return cls._from_c(p)

    def sphere_add(self, r, refl, x, y, z, cb, cg, cr):
        """obj_new + obj_set "sphere" (script.h:1-40)."""
        s = np.zeros(1, SPHERE_DTYPE)
        s["r"], s["refl"] = r, refl
        s["pos"][0] = (x, y, z, 1.0)
        s["col"][0] = (cb, cg, cr, 0.0)
        self.spheres = np.concatenate([self.spheres, s])

    def _to_c(self):
        H = host()
        p = H.pwn_level_new(self.w, self.h, len(self.portals))
        C.memmove(p.contents.cells, self.cells.ctypes.data, self.cells.nbytes)
        if len(self.portals):
            C.memmove(p.contents.portals, self.portals.ctypes.data, self.portals.nbytes)
        p.contents.sx, p.contents.sz = self.sx, self.sz
        return p

    def save_bin(self, path):
        """PWNLV001 binary file (pwn_level.h): any grid size, any number of portal pairs."""
        H = host()
        p = self._to_c()
        rc = H.pwn_level_save_bin(p, os.fsencode(path))
        H.pwn_level_free(p)
        if rc != 0:
            raise PwnError(-1, "cannot write %s" % path)

    def save_txtx(self, path):
        """extended text format "!pwnlv" (pwn_level.h): level.txt rules on any grid, '@' cells for pairs >= 26"""
        H = host()
        p = self._to_c()
        rc = H.pwn_level_save_txtx(p, os.fsencode(path))
        H.pwn_level_free(p)
        if rc != 0:
            raise PwnError(-1, "cannot write %s" % path)

    def save_txt(self, path):
        """level.txt as the reference's loader reads it (64x64 / 26 letters; lower-case marks restored)"""
        H = host()
        p = self._to_c()
        rc = H.pwn_level_save_txt(p, os.fsencode(path))
        H.pwn_level_free(p)
        if rc != 0:
            raise PwnError(-1, "level does not fit the level.txt format")


def level_load(fname):
    """level_load (level.h:107-228) through the C host loader."""
    p = host().pwn_level_load(os.fsencode(fname))
    if not p:
        raise PwnError(-1, "err loading level: %s" % fname)
    return Level._from_c(p)


def level_load_bin(fname):
    p = host().pwn_level_load_bin(os.fsencode(fname))
    if not p:
        raise PwnError(-1, "err loading level: %s" % fname)
    return Level._from_c(p)


def level_parse(text):
    if isinstance(text, str):
        text = text.encode("latin-1")
    return Level._from_c(host().pwn_level_parse(text, len(text)))


def cam_spawn(lv):
    """cam as main.c:60-64."""
    cam = np.zeros(16, np.float32)
    host().pwn_cam_spawn(lv.sx, lv.sz, cam.ctypes.data)
    return cam


def mat4_roty(cam, ang):
    """mat4_roty (util.h:96-110) on a copy."""
    m = np.ascontiguousarray(cam, dtype=np.float32).reshape(16).copy()
    host().pwn_mat4_roty(m.ctypes.data, float(ang))
    return m


def flythrough(lv, nframes):
    """The fixed camera-path script of the headless harness: POSE_DTYPE[nframes]."""
    H = host()
    p = H.pwn_level_new(lv.w, lv.h, len(lv.portals))
    C.memmove(p.contents.cells, lv.cells.ctypes.data, lv.cells.nbytes)
    if len(lv.portals):
        C.memmove(p.contents.portals, lv.portals.ctypes.data, lv.portals.nbytes)
    p.contents.sx, p.contents.sz = lv.sx, lv.sz
    poses = np.zeros(nframes, POSE_DTYPE)
    rc = H.pwn_flythrough(p, nframes, poses.ctypes.data)
    H.pwn_level_free(p)
    if rc != 0:
        raise PwnError(rc, "flythrough failed")
    return poses


class Player:
    """pwn_player (host/pwn_camera.h): `cam` (16 floats, rows x y z w) and `gravity` (4 floats) of main.c:47-48, and
    one game tick of player motion (main.c:188-379) per tick() call."""

    def __init__(self, lv):
        self.state = np.zeros(20, np.float32)
        self._clv = lv._to_c()
        host().pwn_cam_spawn(lv.sx, lv.sz, self.state.ctypes.data)

    @property
    def cam(self):
        return self.state[:16]

    @property
    def gravity(self):
        return self.state[16:]

    def tick(self, keys6, tdiff):
        """keys6 = turnleft turnright moveforward moveback moveleft moveright (pwn_keys)"""
        k = np.ascontiguousarray(keys6, dtype=np.int32)
        assert k.size == 6
        host().pwn_player_tick(self.state.ctypes.data, self._clv, k.ctypes.data, float(tdiff))

    def __del__(self):
        try:
            host().pwn_level_free(self._clv)
        except Exception:
            pass


class Game:
    """The reference's sphere script (game.lua) restated in C (host/pwn_game.c): 14 spheres that
    rotate about a centre which walks through the level.  tick() returns SPHERE_DTYPE[14]."""
    _SIZE = 14 * 8 * 8 + 5 * 8 + 14 * 40

    def __init__(self, lv):
        self.buf = (C.c_char * (self._SIZE + 64))()
        self.lv = lv
        self._clv = lv._to_c()
        host().pwn_game_init(self.buf)

    def spheres(self):
        off = 14 * 8 * 8 + 5 * 8
        return np.frombuffer(self.buf, dtype=SPHERE_DTYPE, count=14, offset=off).copy()

    def tick(self, sec_current, sec_delta):
        host().pwn_game_tick(self.buf, self._clv, float(sec_current), float(sec_delta))
        return self.spheres()

    def __del__(self):
        try:
            host().pwn_level_free(self._clv)
        except Exception:
            pass


def block_shape(dimy, nshards, shard, halo_bands=0, band_start=None):
    """(b0, nb, ht, hb, vlo, vhi) of a shard in the block layout -- pwn_block_shape, host only"""
    out = np.zeros(6, np.int32)
    b = None if band_start is None else np.ascontiguousarray(band_start, dtype=np.int32)
    _check(lib().pwn_block_shape(dimy, nshards, halo_bands, b.ctypes.data if b is not None else None, shard, out.ctypes.data))
    return tuple(int(v) for v in out)


def ray_setup(cam, dimx, dimy):
    """screen.h:43-57 as the library computes it on the host."""
    cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(16)
    rb, rx, ry = (np.zeros(4, np.float32) for _ in range(3))
    lib().pwn_ray_setup(cam.ctypes.data, dimx, dimy, rb.ctypes.data, rx.ctypes.data, ry.ctypes.data)
    return rb, rx, ry


def host_math_probe():
    """(rcp_bits, rsqrt_bits, rcp_tab, rsqrt_tab) of the CPU this runs on. No GPU needed."""
    L = lib()
    rb, sb = C.c_int(), C.c_int()
    rc = L.pwn_host_math_probe(C.byref(rb), C.byref(sb), None, 0, None, 0)
    if rc != 0:
        raise PwnError(rc, "host rcpps/rsqrtps is not a mantissa-bucket table")
    rt = np.zeros(1 << rb.value, np.uint32)
    st = np.zeros(2 << sb.value, np.uint32)
    rc = L.pwn_host_math_probe(C.byref(rb), C.byref(sb), rt.ctypes.data, rt.size, st.ctypes.data, st.size)
    if rc != 0:
        raise PwnError(rc, "host math probe failed")
    return rb.value, sb.value, rt, st


def host_libm_probe():
    """(fma_build, exact): which glibc build of sinf/cosf/expf this host runs. No GPU needed."""
    a, b = C.c_int(), C.c_int()
    lib().pwn_host_libm_probe(C.byref(a), C.byref(b))
    return a.value, bool(b.value)


def host_libm_eval(x, fma_build):
    """sin, cos, exp of x through the host build of csrc/pwn_libm.h (the device runs the same text)."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    s, c, e = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    lib().pwn_host_libm_eval(x.ctypes.data, s.ctypes.data, c.ctypes.data, e.ctypes.data, x.size, int(fma_build))
    return s, c, e


class PinnedBuffer:
    """cudaHostAlloc'd memory viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(lib().pwn_host_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().pwn_host_free(self.ptr)
            self.ptr = None


class Tracer:
    """One GPU: the device side of trace_screen_centred."""

    def __init__(self, device=0, math_tables="host"):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.pwn_create(C.byref(h), device))
        self.h = h
        self.device = device
        self.rwidth = self.rheight = 0
        self.want_ids = False
        self.level = None
        if math_tables == "host":
            _check(self.L.pwn_math_tables_from_host(self.h))
        elif math_tables is not None:
            rb, sb, rt, st = math_tables[:4]
            _check(self.L.pwn_math_tables(self.h, rt.ctypes.data, rb, st.ctypes.data, sb))
            if len(math_tables) > 4:
                self.libm_config(math_tables[4])

    def close(self):
        if self.h:
            self.L.pwn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup -----------------------------------------------------------------
    def level_upload(self, lv):
        pm = lv.portals
        _check(self.L.pwn_level_upload(self.h, lv.w, lv.h, lv.cells.ctypes.data, len(pm),
                                       pm.ctypes.data if len(pm) else None))
        self.level = lv

    def level_prepare_render(self, lv=None):
        """level_prepare_render (level.h:64-81): re-bin the spheres for this frame."""
        lv = lv or self.level
        sp = np.ascontiguousarray(lv.spheres, dtype=SPHERE_DTYPE)
        _check(self.L.pwn_parts_upload(self.h, len(sp), sp.ctypes.data if len(sp) else None))

    def libm_config(self, fma_build):
        _check(self.L.pwn_libm_config(self.h, int(fma_build)))

    def math_status(self):
        """(rcp_bits, rsqrt_bits, libm_fma, libm_exact) -- pwn_math_status"""
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(self.L.pwn_math_status(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def debug_libm(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        s, c, e = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        _check(self.L.pwn_debug_libm(self.h, x.ctypes.data, s.ctypes.data, c.ctypes.data, e.ctypes.data, x.size))
        return s, c, e

    def frame_config(self, rwidth, rheight, want_ids=False):
        _check(self.L.pwn_frame_config(self.h, rwidth, rheight, 1 if want_ids else 0))
        self.rwidth, self.rheight = rwidth, rheight
        self.want_ids = self.want_ids or want_ids

    def frame_clear(self):
        _check(self.L.pwn_frame_clear(self.h))

    def shard_config(self, nshards, shard):
        _check(self.L.pwn_shard_config(self.h, nshards, shard))

    def shard_layout(self, blocks=False, halo_bands=0):
        """PWN_SHARD_BLOCKS (contiguous band blocks, halo exchange for the blur) or the default interleaved bands"""
        _check(self.L.pwn_shard_layout(self.h, 1 if blocks else 0, halo_bands))

    # -- hot path ----------------------------------------------------------------
    def trace_cam(self, cam, dimx, dimy, sec_current=0.0, flags=0):
        cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(16)
        _check(self.L.pwn_trace_cam(self.h, cam.ctypes.data, dimx, dimy, float(sec_current), flags))

    def trace_frame(self, frm, rayb, rdx, rdy, dimx, dimy, sec_current=0.0, flags=0):
        v = [np.ascontiguousarray(a, dtype=np.float32).reshape(4) for a in (frm, rayb, rdx, rdy)]
        _check(self.L.pwn_trace_frame(self.h, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data,
                                      v[3].ctypes.data, dimx, dimy, float(sec_current), flags))

    def blur_frame(self, dimx, dimy):
        _check(self.L.pwn_blur_frame(self.h, dimx, dimy))

    def sync(self):
        _check(self.L.pwn_sync(self.h))

    def join(self):
        """order everything queued so far (also on the library's second stream) before what follows on stream_ptr()"""
        _check(self.L.pwn_join(self.h))

    def debug_pack(self, bgra):
        """the device's col_ftoint (util.h:48-59) on (n,4) float32 BGRA rows -> (n,) uint32"""
        v = np.ascontiguousarray(bgra, dtype=np.float32).reshape(-1, 4)
        out = np.zeros(len(v), np.uint32)
        _check(self.L.pwn_debug_pack(self.h, v.ctypes.data, out.ctypes.data, len(v)))
        return out

    def trace_config(self, reflect=2):
        """REFLECT (defs.h:8): bounce levels, 0..2"""
        _check(self.L.pwn_trace_config(self.h, reflect))

    def screen_upscale(self, dimx, dimy, rscale, pitch_words=None, read=True):
        """screen_upscale (screen.h:126-149) of the frame in sbuf; returns the surface as a
        (dimy*rscale, pitch_words) uint32 array, or None with read=False (surface stays on the device)."""
        pitch_words = dimx * rscale if pitch_words is None else pitch_words
        out = np.zeros((dimy * rscale, pitch_words), np.uint32) if read else None
        _check(self.L.pwn_screen_upscale(self.h, dimx, dimy, rscale, pitch_words * 4,
                                         out.ctypes.data if read else None))
        return out

    def last_upscale_ms(self):
        t = C.c_float()
        _check(self.L.pwn_last_upscale_ms(self.h, C.byref(t)))
        return t.value

    def read_frame(self, want_z=True, want_ids=False):
        sbuf = np.zeros((self.rheight, self.rwidth), np.uint32)
        zbuf = np.zeros((self.rheight, self.rwidth), np.float32) if want_z else None
        ids = np.zeros((self.rheight, self.rwidth, 3), IDS_DTYPE) if want_ids else None
        _check(self.L.pwn_read_frame(self.h, sbuf.ctypes.data,
                                     zbuf.ctypes.data if want_z else None,
                                     ids.ctypes.data if want_ids else None))
        return sbuf, zbuf, ids

    def write_frame(self, sbuf=None, zbuf=None):
        """host -> device: (rheight, rwidth) uint32 sbuf and/or float32 zbuf"""
        if sbuf is not None:
            sbuf = np.ascontiguousarray(sbuf, dtype=np.uint32)
            assert sbuf.shape == (self.rheight, self.rwidth)
        if zbuf is not None:
            zbuf = np.ascontiguousarray(zbuf, dtype=np.float32)
            assert zbuf.shape == (self.rheight, self.rwidth)
        _check(self.L.pwn_write_frame(self.h, sbuf.ctypes.data if sbuf is not None else None,
                                      zbuf.ctypes.data if zbuf is not None else None))

    def trace_screen_centred(self, lv, x1, y1, x2, y2, cam, sec_current=0.0, blur=True, sbuf=None, zbuf=None):
        """trace_screen_centred (screen.h:31): [p1,p2) window, frame returned in host memory.

        As in the reference, x1/y1 only enter through dimx = x2-x1, dimy = y2-y1
        (screen.h:39-40); buffers are indexed from (0,0) with stride rwidth."""
        assert x1 < x2 and y1 < y2   # screen.h:35-36
        if lv is not self.level:
            self.level_upload(lv)
        cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(16)
        if sbuf is None:
            sbuf = np.zeros((self.rheight, self.rwidth), np.uint32)
        _check(self.L.pwn_trace_screen(self.h, cam.ctypes.data, x2 - x1, y2 - y1, float(sec_current),
                                       PWN_F_BLUR if blur else 0, sbuf.ctypes.data,
                                       zbuf.ctypes.data if zbuf is not None else None))
        return sbuf

    def shard_blocks(self, band_start=None):
        """block layout: first band of every shard + the frame's band count (nshards + 1 ints); None = even split"""
        if band_start is None:
            _check(self.L.pwn_shard_blocks(self.h, 0, None))
            return
        b = np.ascontiguousarray(band_start, dtype=np.int32)
        _check(self.L.pwn_shard_blocks(self.h, len(b) - 1, b.ctypes.data))

    def screen_chunks(self, nchunks=0):
        """row chunks per frame of trace_screen_centred (0 = automatic)"""
        _check(self.L.pwn_screen_chunks(self.h, nchunks))

    def trace_poses(self, poses, dimx, dimy, flags=0, out=None, want_sums=False):
        poses = np.ascontiguousarray(poses, dtype=POSE_DTYPE)
        n = len(poses)
        sums = np.zeros(n, np.uint64) if want_sums else None
        optr = None
        if out is not None:
            optr = out.ctypes.data if isinstance(out, np.ndarray) else out
        _check(self.L.pwn_trace_poses(self.h, n, poses.ctypes.data, dimx, dimy, flags, optr,
                                      sums.ctypes.data if want_sums else None))
        return sums

    # -- measurement ---------------------------------------------------------------
    def launch_count(self):
        c = C.c_uint64()
        _check(self.L.pwn_launch_count(self.h, C.byref(c)))
        return c.value

    def last_opcount(self):
        c = C.c_uint64()
        _check(self.L.pwn_last_opcount(self.h, C.byref(c)))
        return c.value

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        _check(self.L.pwn_last_kernel_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_frames(self, cam, dimx, dimy, sec_current, flags, n):
        cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(16)
        t = C.c_float()
        _check(self.L.pwn_time_frames(self.h, cam.ctypes.data, dimx, dimy, float(sec_current), flags, n, C.byref(t)))
        return t.value

    def fp32_peak(self):
        v = C.c_double()
        _check(self.L.pwn_fp32_peak(self.h, C.byref(v)))
        return v.value

    def stream_ptr(self):
        s = C.c_uint64()
        _check(self.L.pwn_stream(self.h, C.byref(s)))
        return s.value

    def device_ptrs(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.L.pwn_frame_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def frame_redirect(self, sbuf_dev, zbuf_dev, tsbuf_dev):
        _check(self.L.pwn_frame_redirect(self.h, sbuf_dev, zbuf_dev, tsbuf_dev))

    def frame_peers(self, bases):
        """bases[s] = device address (in this process) of shard s's frame allocation; None/[] clears"""
        if not bases:
            _check(self.L.pwn_frame_peers(self.h, 0, None))
            return
        arr = (C.c_uint64 * len(bases))(*[int(b or 0) for b in bases])
        _check(self.L.pwn_frame_peers(self.h, len(bases), arr))

    def trace_shard_blur(self, cam, dimx, dimy, sec_current, flags=0):
        cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(16)
        _check(self.L.pwn_trace_shard_blur(self.h, cam.ctypes.data, dimx, dimy, float(sec_current), flags))

    def read_bands(self, dimx, dimy, host_frame):
        """copy this shard's bands into the same rows of a (rheight, rwidth) uint32 host frame (numpy array or address)"""
        ptr = host_frame.ctypes.data if isinstance(host_frame, np.ndarray) else host_frame
        _check(self.L.pwn_read_bands(self.h, dimx, dimy, ptr))

    def debug_flags(self, nwords=128, alt=False):
        out = np.zeros(nwords, np.uint32)
        n = self.L.pwn_debug_flags(self.h, out.ctypes.data, -nwords if alt else nwords)
        if n < 0:
            _check(n)
        return out[:n]

    def peers_timed_out(self):
        v = C.c_int()
        _check(self.L.pwn_frame_peers_status(self.h, C.byref(v)))
        return bool(v.value)

    def ipc_export(self):
        hnd = (C.c_char * 64)()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.L.pwn_frame_ipc_export(self.h, hnd, C.byref(a), C.byref(b), C.byref(c)))
        return bytes(hnd), a.value, b.value, c.value

    def ipc_import(self, handle):
        base = C.c_uint64()
        buf = C.create_string_buffer(handle, 64)
        _check(self.L.pwn_frame_ipc_import(self.h, buf, C.byref(base)))
        return base.value

    def ipc_close(self, base):
        _check(self.L.pwn_frame_ipc_close(self.h, base))


def checksum_ref(sbuf, dimx, dimy):
    """Host restatement of the device frame checksum (pwntrace.h, pwn_trace_poses)."""
    w = np.ascontiguousarray(sbuf[:dimy, :dimx]).astype(np.uint64).reshape(-1)
    i = np.arange(w.size, dtype=np.uint64)
    z = (i << np.uint64(32)) | w
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return np.uint64(z.sum(dtype=np.uint64))

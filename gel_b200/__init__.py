"""gel_b200 -- B200 (sm_100a) implementation of gel's per-frame render path.

The product is two in-tree shared libraries and one executable, all C / CUDA:

  gel_b200/libgelcu.so    hand-written sm_100a kernels behind the C ABI in include/gelcu.h
  gel_b200/libgelhost.so  the host C flow that stays on the CPU (OBJ/BMP load, soup expansion, view basis)
  gel_b200/host/gel       headless `gel` (the reference's main() with lines 505-522 on the GPU)

This Python package is only a ctypes veneer over those libraries for tests/ and bench.py.  It has NO
render fallback: if libgelcu.so is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

GELCU_OK, GELCU_W_CLIPPED = 0, 1
GELCU_E_INVALID, GELCU_E_CUDA, GELCU_E_NOMEM, GELCU_E_NOGPU = -1, -2, -3, -4

# every symbol include/gelcu.h declares (tests check the library exports exactly these)
GELCU_SYMBOLS = [
    "gelcu_device_count", "gelcu_create", "gelcu_set_mesh", "gelcu_set_mesh_indexed", "gelcu_set_texture", "gelcu_render",
    "gelcu_render_rgb8", "gelcu_render_region",
    "gelcu_read_frame", "gelcu_set_option", "gelcu_get_stats", "gelcu_debug_transform", "gelcu_debug_bins",
    "gelcu_tile_grid", "gelcu_host_alloc", "gelcu_host_free", "gelcu_destroy", "gelcu_last_error",
]


class GelcuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gelcu error {code}: {msg}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("kernels_launched", c_uint64), ("views", c_uint64), ("bin_entries", c_uint64),
                ("unique_vertices", c_uint64), ("triangles", c_uint64), ("h2d_bytes", c_uint64),
                ("d2h_bytes", c_uint64), ("ms_transform", c_float), ("ms_bin", c_float), ("ms_raster", c_float), ("ms_dominant", c_float),
                ("ms_total", c_float), ("flags", c_uint32), ("batches", c_uint32), ("pipeline", c_uint32), ("reserved", c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_fp, _u32p, _u64p, _ip = POINTER(c_float), POINTER(c_uint32), POINTER(c_uint64), POINTER(c_int)
_cu = None
_host = None


def _load(name: str) -> ctypes.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is not built -- run `make` (or __graft_entry__.build()); there is no fallback path")
    return ctypes.CDLL(path)


def cu() -> ctypes.CDLL:
    """libgelcu.so with argtypes set."""
    global _cu
    if _cu is None:
        L = _load(os.environ.get("GELCU_LIB", "libgelcu.so"))   # GELCU_LIB: tuning variants built by scripts/
        L.gelcu_device_count.restype = c_int
        L.gelcu_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int]
        L.gelcu_set_mesh.argtypes = [c_void_p, _fp, _fp, _fp, c_int]
        L.gelcu_set_mesh_indexed.argtypes = [c_void_p, _fp, c_int, _fp, c_int, _fp, c_int, _ip, c_int]
        L.gelcu_set_texture.argtypes = [c_void_p, _u32p, c_int, c_int]
        L.gelcu_render_region.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, _fp]
        L.gelcu_render.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, _fp]
        L.gelcu_render_rgb8.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, _fp]
        L.gelcu_read_frame.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
        L.gelcu_set_option.argtypes = [c_void_p, c_char_p, c_int]
        L.gelcu_get_stats.argtypes = [c_void_p, POINTER(Stats)]
        L.gelcu_debug_transform.argtypes = [c_void_p, c_void_p, _fp, _fp]
        L.gelcu_debug_bins.argtypes = [c_void_p, c_void_p, _ip, _ip, c_int, _ip]
        L.gelcu_tile_grid.argtypes = [c_void_p, _ip, _ip, _ip, _ip]
        L.gelcu_host_alloc.argtypes = [POINTER(c_void_p), c_size_t]
        L.gelcu_host_free.argtypes = [c_void_p]
        L.gelcu_destroy.argtypes = [c_void_p]
        L.gelcu_last_error.restype = c_char_p
        _cu = L
    return _cu


class _Mesh(ctypes.Structure):
    _fields_ = [("tv", _fp), ("tn", _fp), ("tt", _fp), ("ntri", c_int), ("nv", c_int), ("nvt", c_int), ("nvn", c_int)]


class _Obj(ctypes.Structure):
    _fields_ = [("v", _fp), ("vt", _fp), ("vn", _fp), ("faces", _ip), ("nv", c_int), ("nvt", c_int), ("nvn", c_int), ("nfaces", c_int),
                ("parse_threads", c_int)]


class _Tex(ctypes.Structure):
    _fields_ = [("pixels", _u32p), ("w", c_int), ("h", c_int)]


def host() -> ctypes.CDLL:
    """libgelhost.so with argtypes set."""
    global _host
    if _host is None:
        L = _load("libgelhost.so")
        L.gel_obj_load.argtypes = [c_char_p, POINTER(_Mesh)]
        L.gel_mesh_free.argtypes = [POINTER(_Mesh)]
        L.gel_obj_parse.argtypes = [c_char_p, POINTER(_Obj)]
        L.gel_obj_free.argtypes = [POINTER(_Obj)]
        L.gel_bmp_load.argtypes = [c_char_p, POINTER(_Tex)]
        L.gel_texture_free.argtypes = [POINTER(_Tex)]
        L.gel_view_basis.argtypes = [c_float, c_float, _fp]
        L.gel_input_step.argtypes = [_fp, _fp, c_int, c_int]
        L.gel_fnv1a64_words.argtypes = [_u32p, c_uint64]
        L.gel_fnv1a64_words.restype = c_uint64
        L.gel_upright.argtypes = [_u32p, c_int, c_int, _u32p]
        _host = L
    return _host


# ---- host flow ------------------------------------------------------------------------------------

def load_obj(path: str):
    """(tv, tn, tt) float32 arrays of shape (ntri, 9) -- gel_obj_load (reference main.c:129-180, 227-286)."""
    m = _Mesh()
    rc = host().gel_obj_load(path.encode(), byref(m))
    if rc != 0:
        raise RuntimeError(f"gel_obj_load({path}) failed with {rc}")
    n = m.ntri
    out = tuple(np.ctypeslib.as_array(p, shape=(max(n, 1), 9))[:n].copy() for p in (m.tv, m.tn, m.tt))
    host().gel_mesh_free(byref(m))
    return out


def load_obj_indexed(path: str):
    """(v (nv, 3), vt (nvt, 3), vn (nvn, 3) float32, faces (nfaces, 9) int32 in the reference's Face layout
    { va,vb,vc, ta,tb,tc, na,nb,nc }, 0-based) -- gel_obj_parse (reference main.c:129-180), no soup expansion."""
    o = _Obj()
    rc = host().gel_obj_parse(path.encode(), byref(o))
    if rc != 0:
        raise RuntimeError(f"gel_obj_parse({path}) failed with {rc}")
    arr = lambda p, n, w, dt: (np.ctypeslib.as_array(p, shape=(n, w)).astype(dt, copy=True) if n else np.zeros((0, w), dt))
    out = (arr(o.v, o.nv, 3, np.float32), arr(o.vt, o.nvt, 3, np.float32), arr(o.vn, o.nvn, 3, np.float32), arr(o.faces, o.nfaces, 9, np.int32))
    host().gel_obj_free(byref(o))
    return out


def load_bmp(path: str) -> np.ndarray:
    """uint32 (h, w) XRGB8888 top-down -- gel_bmp_load (reference main.c:471-484)."""
    t = _Tex()
    rc = host().gel_bmp_load(path.encode(), byref(t))
    if rc != 0:
        raise RuntimeError(f"gel_bmp_load({path}) failed with {rc}")
    out = np.ctypeslib.as_array(t.pixels, shape=(t.h, t.w)).copy()
    host().gel_texture_free(byref(t))
    return out


def view_basis(xt, yt) -> np.ndarray:
    """float32 (12,) = x, y, z, eye -- gel_view_basis (reference main.c:506-512)."""
    b = np.zeros(12, dtype=np.float32)
    host().gel_view_basis(c_float(float(xt)), c_float(float(yt)), b.ctypes.data_as(_fp))
    return b


def view_bases(angles) -> np.ndarray:
    a = np.asarray(angles, dtype=np.float32).reshape(-1, 2)
    return np.stack([view_basis(x, y) for x, y in a]) if len(a) else np.zeros((0, 12), np.float32)


def fnv1a64_words(words: np.ndarray) -> int:
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    return int(host().gel_fnv1a64_words(w.ctypes.data_as(_u32p), w.size))


# ---- device ---------------------------------------------------------------------------------------

def _check(rc: int) -> int:
    if rc < 0:
        raise GelcuError(rc, cu().gelcu_last_error().decode())
    return rc


class PinnedBuffer:
    """numpy view over page-locked host memory from gelcu_host_alloc."""

    def __init__(self, shape, dtype):
        self.ptr = c_void_p()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(cu().gelcu_host_alloc(byref(self.ptr), nbytes))
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            cu().gelcu_host_free(self.ptr)
            self.ptr = None


class Renderer:
    """One gelcu context (one GPU).  Mirrors the reference's frame loop inputs: mesh soups, texture, views."""

    def __init__(self, xres: int, yres: int, device: int = 0):
        self.xres, self.yres, self.device = xres, yres, device
        self._ctx = c_void_p()
        _check(cu().gelcu_create(byref(self._ctx), device, xres, yres))
        self.ntri = 0
        self.last_rc = 0

    def close(self):
        if self._ctx:
            cu().gelcu_destroy(self._ctx)
            self._ctx = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_mesh(self, tv, tn, tt):
        tv, tn, tt = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 9) for a in (tv, tn, tt))
        assert tv.shape == tn.shape == tt.shape
        self.ntri = tv.shape[0]
        _check(cu().gelcu_set_mesh(self._ctx, tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tt.ctypes.data_as(_fp), self.ntri))

    def set_mesh_indexed(self, v, vt, vn, faces):
        """The indexed OBJ arrays (load_obj_indexed); the soups of main.c:242-286 are generated on the device."""
        v, vt, vn = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (v, vt, vn))
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 9)
        self.ntri = faces.shape[0]
        _check(cu().gelcu_set_mesh_indexed(self._ctx, v.ctypes.data_as(_fp), v.shape[0], vt.ctypes.data_as(_fp), vt.shape[0],
                                           vn.ctypes.data_as(_fp), vn.shape[0], faces.ctypes.data_as(_ip), faces.shape[0]))

    def set_texture(self, xrgb):
        t = np.ascontiguousarray(xrgb, dtype=np.uint32)
        _check(cu().gelcu_set_texture(self._ctx, t.ctypes.data_as(_u32p), t.shape[1], t.shape[0]))

    def set_option(self, name: str, value: int):
        _check(cu().gelcu_set_option(self._ctx, name.encode(), int(value)))

    def render(self, bases, *, pixels=True, z=False, hashes=False, pixel_out=None, z_out=None):
        """bases: (n, 12) float32.  Returns dict(pixel=(n, xres*yres) uint32 | None, z=..., hash=(n, 2) uint64 | None,
        device_ms=float, rc=int).  pixel_out / z_out may be preallocated (e.g. PinnedBuffer.array)."""
        b = np.ascontiguousarray(bases, dtype=np.float32).reshape(-1, 12)
        n = b.shape[0]
        frame = self.xres * self.yres
        if pixels and pixel_out is None:
            pixel_out = np.empty((n, frame), dtype=np.uint32)
        if z and z_out is None:
            z_out = np.empty((n, frame), dtype=np.float32)
        h = np.zeros((n, 2), dtype=np.uint64) if hashes else None
        ms = c_float(0.0)
        rc = _check(cu().gelcu_render(self._ctx, b.ctypes.data_as(c_void_p), n,
                                      pixel_out.ctypes.data_as(c_void_p) if pixel_out is not None else None,
                                      z_out.ctypes.data_as(c_void_p) if z_out is not None else None,
                                      h.ctypes.data_as(c_void_p) if h is not None else None, byref(ms)))
        self.last_rc = rc
        return {"pixel": pixel_out, "z": z_out, "hash": h, "device_ms": float(ms.value), "rc": rc}

    def render_rgb8(self, bases, *, hashes=False, rgb_out=None):
        """Frame sink: upright 24-bit frames, (n, yres, xres, 3) uint8 (the body of a binary PPM each).  Returns
        dict(rgb, hash, device_ms, rc)."""
        b = np.ascontiguousarray(bases, dtype=np.float32).reshape(-1, 12)
        n = b.shape[0]
        if rgb_out is None:
            rgb_out = np.empty((n, self.yres, self.xres, 3), dtype=np.uint8)
        h = np.zeros((n, 2), dtype=np.uint64) if hashes else None
        ms = c_float(0.0)
        rc = _check(cu().gelcu_render_rgb8(self._ctx, b.ctypes.data_as(c_void_p), n, rgb_out.ctypes.data_as(c_void_p),
                                           h.ctypes.data_as(c_void_p) if h is not None else None, byref(ms)))
        self.last_rc = rc
        return {"rgb": rgb_out, "hash": h, "device_ms": float(ms.value), "rc": rc}

    def render_region(self, bases, pixel_io, rects, *, z_io=None, rgb8=False, hashes=False):
        """gelcu_render_region: pixel_io (n, xres*yres) uint32 -- or (n, yres, xres, 3) uint8 with rgb8 -- and rects (n, 4)
        int32 { x0, y0, x1, y1 } are updated in place (dirty-rectangle contract, include/gelcu.h).  Returns dict(hash, device_ms, rc)."""
        b = np.ascontiguousarray(bases, dtype=np.float32).reshape(-1, 12)
        n = b.shape[0]
        assert rects.dtype == np.int32 and rects.shape == (n, 4) and rects.flags.c_contiguous and pixel_io.flags.c_contiguous
        h = np.zeros((n, 2), dtype=np.uint64) if hashes else None
        ms = c_float(0.0)
        rc = _check(cu().gelcu_render_region(self._ctx, b.ctypes.data_as(c_void_p), n, pixel_io.ctypes.data_as(c_void_p),
                                             z_io.ctypes.data_as(c_void_p) if z_io is not None else None, rects.ctypes.data_as(c_void_p),
                                             1 if rgb8 else 0, h.ctypes.data_as(c_void_p) if h is not None else None, byref(ms)))
        self.last_rc = rc
        return {"hash": h, "device_ms": float(ms.value), "rc": rc}

    def read_frame(self, slot: int):
        frame = self.xres * self.yres
        px = np.empty(frame, dtype=np.uint32)
        zb = np.empty(frame, dtype=np.float32)
        _check(cu().gelcu_read_frame(self._ctx, slot, px.ctypes.data_as(c_void_p), zb.ctypes.data_as(c_void_p)))
        return px, zb

    def stats(self) -> dict:
        s = Stats()
        _check(cu().gelcu_get_stats(self._ctx, byref(s)))
        return s.as_dict()

    def tile_grid(self):
        v = [c_int() for _ in range(4)]
        _check(cu().gelcu_tile_grid(self._ctx, *[byref(x) for x in v]))
        return tuple(x.value for x in v)

    def debug_transform(self, basis):
        b = np.ascontiguousarray(basis, dtype=np.float32).reshape(12)
        vew = np.empty((self.ntri, 9), dtype=np.float32)
        shade = np.empty((self.ntri, 3), dtype=np.float32)
        _check(cu().gelcu_debug_transform(self._ctx, b.ctypes.data_as(c_void_p), vew.ctypes.data_as(_fp), shade.ctypes.data_as(_fp)))
        return vew, shade

    def debug_bins(self, basis, cap: int = 1 << 24):
        b = np.ascontiguousarray(basis, dtype=np.float32).reshape(12)
        _, _, tx, ty = self.tile_grid()
        counts = np.zeros(tx * ty, dtype=np.int32)
        entries = np.zeros(cap, dtype=np.int32)
        total = c_int()
        _check(cu().gelcu_debug_bins(self._ctx, b.ctypes.data_as(c_void_p), counts.ctypes.data_as(_ip), entries.ctypes.data_as(_ip), cap, byref(total)))
        return counts, entries[:min(cap, total.value)], total.value


def shard_views(nviews: int, world: int, rank: int):
    """Contiguous block of the view list rendered by `rank` of `world` GPUs: no collective
    on the render path, every rank holds a replica of mesh + texture."""
    lo = nviews * rank // world
    hi = nviews * (rank + 1) // world
    return lo, hi

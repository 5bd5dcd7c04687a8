"""TEST INFRASTRUCTURE -- ctypes access to the CPU oracle (oracle/ref_cpu.c -> libgeloracle.so) and to the
unmodified reference binaries under oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or as the reported CPU baseline.  Nothing under gel_b200/ imports it.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import tempfile
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_uint32, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_fp, _u32p, _u64p = POINTER(c_float), POINTER(c_uint32), POINTER(c_uint64)
_lib = None


class Counters(ctypes.Structure):
    _fields_ = [("tested", c_uint64), ("inside", c_uint64), ("zpass", c_uint64), ("lit", c_uint64)]


def build() -> None:
    subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)
    if os.path.exists("/root/reference/main.c"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libgeloracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.ref_view_basis.argtypes = [c_float, c_float, _fp]
        L.ref_transform.argtypes = [_fp, _fp, c_int, _fp, c_int, c_int, _fp, _fp]
        L.ref_render.argtypes = [_fp, _fp, _fp, c_int, _u32p, c_int, c_int, c_int, c_int, _fp, _u32p, _fp, POINTER(Counters)]
        L.ref_render_views.argtypes = [_fp, _fp, _fp, c_int, _u32p, c_int, c_int, c_int, c_int, _fp, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double)]
        L.ref_fnv1a64_words.argtypes = [_u32p, c_uint64]
        L.ref_fnv1a64_words.restype = c_uint64
        L.ref_salted_sum.argtypes = [_u32p, c_uint64]
        L.ref_salted_sum.restype = c_uint64
        L.ref_load_obj.argtypes = [c_char_p, POINTER(_fp), POINTER(_fp), POINTER(_fp)]
        L.ref_load_bmp.argtypes = [c_char_p, POINTER(_u32p), POINTER(c_int), POINTER(c_int)]
        L.ref_free.argtypes = [c_void_p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def view_basis(xt, yt) -> np.ndarray:
    b = np.zeros(12, dtype=np.float32)
    lib().ref_view_basis(c_float(float(xt)), c_float(float(yt)), b.ctypes.data_as(_fp))
    return b


def transform(tv, tn, basis, xres, yres):
    tv, tn, basis = _f32(tv).reshape(-1, 9), _f32(tn).reshape(-1, 9), _f32(basis).reshape(12)
    vew, nrm = np.empty_like(tv), np.empty_like(tv)
    lib().ref_transform(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tv.shape[0], basis.ctypes.data_as(_fp), xres, yres,
                        vew.ctypes.data_as(_fp), nrm.ctypes.data_as(_fp))
    return vew, nrm


def render(tv, tn, tt, tex, xres, yres, basis, counters=False):
    """One frame.  Returns (pixel uint32[xres*yres], zbuff float32[xres*yres], clipped flag[, Counters])."""
    tv, tn, tt = (_f32(a).reshape(-1, 9) for a in (tv, tn, tt))
    tex = np.ascontiguousarray(tex, dtype=np.uint32)
    basis = _f32(basis).reshape(12)
    px = np.empty(xres * yres, dtype=np.uint32)
    zb = np.empty(xres * yres, dtype=np.float32)
    c = Counters()
    rc = lib().ref_render(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tt.ctypes.data_as(_fp), tv.shape[0],
                          tex.ctypes.data_as(_u32p), tex.shape[1], tex.shape[0], xres, yres, basis.ctypes.data_as(_fp),
                          px.ctypes.data_as(_u32p), zb.ctypes.data_as(_fp), byref(c) if counters else None)
    return (px, zb, rc, c) if counters else (px, zb, rc)


def render_views(tv, tn, tt, tex, xres, yres, bases, nthreads=1, *, pixels=True, z=False, hashes=False):
    """Frames-parallel batch on `nthreads` host threads.  Returns dict(pixel, z, hash (n,2), seconds, clipped)."""
    tv, tn, tt = (_f32(a).reshape(-1, 9) for a in (tv, tn, tt))
    tex = np.ascontiguousarray(tex, dtype=np.uint32)
    bases = _f32(bases).reshape(-1, 12)
    n = bases.shape[0]
    px = np.empty((n, xres * yres), dtype=np.uint32) if pixels else None
    zb = np.empty((n, xres * yres), dtype=np.float32) if z else None
    hp = np.zeros(n, dtype=np.uint64) if hashes else None
    hz = np.zeros(n, dtype=np.uint64) if hashes else None
    sec = c_double(0.0)
    ptr = lambda a: a.ctypes.data_as(c_void_p) if a is not None else None
    rc = lib().ref_render_views(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tt.ctypes.data_as(_fp), tv.shape[0],
                                tex.ctypes.data_as(_u32p), tex.shape[1], tex.shape[0], xres, yres,
                                bases.ctypes.data_as(_fp), n, nthreads, ptr(px), ptr(zb), ptr(hp), ptr(hz), byref(sec))
    return {"pixel": px, "z": zb, "hash": np.stack([hp, hz], 1) if hashes else None, "seconds": sec.value, "clipped": rc}


def fnv1a64_words(words) -> int:
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    return int(lib().ref_fnv1a64_words(w.ctypes.data_as(_u32p), w.size))


def salted_sum(words) -> int:
    w = np.ascontiguousarray(words).view(np.uint32).reshape(-1)
    return int(lib().ref_salted_sum(w.ctypes.data_as(_u32p), w.size))


def load_obj(path: str):
    tv, tn, tt = _fp(), _fp(), _fp()
    n = lib().ref_load_obj(path.encode(), byref(tv), byref(tn), byref(tt))
    if n < 0:
        raise RuntimeError(f"ref_load_obj({path}) failed")
    out = tuple(np.ctypeslib.as_array(p, shape=(max(n, 1), 9))[:n].copy() for p in (tv, tn, tt))
    for p in (tv, tn, tt):
        lib().ref_free(p)
    return out


def load_bmp(path: str) -> np.ndarray:
    px, w, h = _u32p(), c_int(), c_int()
    rc = lib().ref_load_bmp(path.encode(), byref(px), byref(w), byref(h))
    if rc != 0:
        raise RuntimeError(f"ref_load_bmp({path}) failed with {rc}")
    out = np.ctypeslib.as_array(px, shape=(h.value, w.value)).copy()
    lib().ref_free(px)
    return out


# ---- the unmodified reference binary ---------------------------------------------------------------

def ref_binary(xres: int, yres: int, shipped: bool = False):
    p = os.path.join(REF_DIR, f"gel_ref_{'shipped_' if shipped else ''}{xres}x{yres}")
    return p if os.path.exists(p) else None


def mouse_angles(frames: int, dx: int, dy: int) -> np.ndarray:
    """(xt, yt) per frame exactly as the reference accumulates them (main.c:394, 408-409)."""
    xt = yt = np.float32(0.0)
    sens = np.float32(0.005)
    out = np.zeros((frames, 2), dtype=np.float32)
    for k in range(frames):
        out[k] = (xt, yt)
        xt = np.float32(xt - np.float32(sens * np.float32(dx)))
        yt = np.float32(yt + np.float32(sens * np.float32(dy)))
    return out


def run_reference(obj_path: str, bmp_path: str, xres: int, yres: int, frames: int = 1, dx: int = 0, dy: int = 0,
                  dump: bool = True, shipped: bool = False, extra_env=None):
    """Runs oracle/_ref/gel_ref_<res> headless.  Returns (list of per-frame dicts, frames uint32 (n, xres*yres) | None)."""
    exe = ref_binary(xres, yres, shipped)
    if exe is None:
        raise FileNotFoundError(f"no reference binary for {xres}x{yres} under {REF_DIR} (make -C oracle ref)")
    env = dict(os.environ, GELSHIM_FRAMES=str(frames), GELSHIM_DX=str(dx), GELSHIM_DY=str(dy))
    if extra_env:
        env.update(extra_env)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "frames.raw")
        if dump:
            env["GELSHIM_DUMP"] = path
        out = subprocess.run([exe, obj_path, bmp_path], env=env, capture_output=True, text=True, check=True).stdout
        lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        px = np.fromfile(path, dtype=np.uint32).reshape(len(lines), xres * yres) if dump else None
    return lines, px

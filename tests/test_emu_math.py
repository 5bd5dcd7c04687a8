"""gel_b200/csrc/gel_math.h compiled for the host (tests/emu/emu.cpp): the device's operation order, the keyed
depth resolve in REVERSE draw order, deferred shading and the exact sign early-out must reproduce the oracle
bit for bit.  (The same header runs on the GPU with _rn intrinsics; the -m gpu tests close that last gap.)"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT, bits, random_soup

_fp, _up = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)


@pytest.fixture(scope="module")
def emu():
    return ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libgelemu.so"))


def emu_render(E, tv, tn, tt, tex, xres, yres, basis, guard=1):
    px, zb = np.empty(xres * yres, np.uint32), np.empty(xres * yres, np.float32)
    tv, tn, tt = (np.ascontiguousarray(a, np.float32) for a in (tv, tn, tt))
    tex = np.ascontiguousarray(tex, np.uint32)
    f = E.emu_render(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tt.ctypes.data_as(_fp), tv.shape[0], tex.ctypes.data_as(_up),
                     tex.shape[1], tex.shape[0], xres, yres, np.ascontiguousarray(basis, np.float32).ctypes.data_as(_fp),
                     px.ctypes.data_as(_up), zb.ctypes.data_as(_fp), guard)
    return px, zb, f


def test_transform_bits(emu, cfg1):
    tv, tn, _, _ = cfg1
    for xt, yt in [(0, 0), (1.1, 0.3), (3.3, -0.2)]:
        basis = oracle.view_basis(xt, yt)
        vew, nrm = oracle.transform(tv, tn, basis, 1920, 1080)
        evew, eshade = np.empty_like(vew), np.empty((tv.shape[0], 3), np.float32)
        emu.emu_transform(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tv.shape[0], basis.ctypes.data_as(_fp), 1920, 1080,
                          evew.ctypes.data_as(_fp), eshade.ctypes.data_as(_fp))
        assert np.array_equal(bits(vew), bits(evew))
        n = nrm.reshape(-1, 3, 3)
        want = (np.float32(0) * n[:, :, 0] + np.float32(0) * n[:, :, 1]) + np.float32(1) * n[:, :, 2]    # main.c:358
        assert np.array_equal(bits(want.astype(np.float32)), bits(eshade))


@pytest.mark.parametrize("res", [(800, 600), (1920, 1080), (101, 67)])
def test_frame_bits_sphere(emu, cfg1, res):
    tv, tn, tt, tex = cfg1
    for xt, yt in [(0, 0), (2.5, -0.3)]:
        basis = oracle.view_basis(xt, yt)
        px, zb, clipped = oracle.render(tv, tn, tt, tex, res[0], res[1], basis)
        for guard in (0, 1):
            epx, ezb, flags = emu_render(emu, tv, tn, tt, tex, res[0], res[1], basis, guard)
            assert flags == clipped == 0
            assert np.array_equal(px, epx) and np.array_equal(bits(zb), bits(ezb))


def test_frame_bits_random_soup_with_ties(emu):
    rng = np.random.default_rng(11)
    tv, tn, tt = random_soup(rng, 600)
    tv, tn = np.vstack([tv, tv[:200]]), np.vstack([tn, tn[:200]])          # exact duplicates later in draw order
    tt = np.vstack([tt, rng.uniform(0, 1, (200, 9)).astype(np.float32)])
    tex = rng.integers(0, 1 << 24, (64, 32), dtype=np.uint32)               # non-square texture
    for xt, yt in [(0, 0), (0.4, 0.1)]:
        basis = oracle.view_basis(xt, yt)
        px, zb, clipped = oracle.render(tv, tn, tt, tex, 640, 480, basis)
        epx, ezb, flags = emu_render(emu, tv, tn, tt, tex, 640, 480, basis)
        assert (flags & 1) == clipped
        assert np.array_equal(px, epx) and np.array_equal(bits(zb), bits(ezb))


def test_salted_sum_definition(emu):
    w = np.random.default_rng(5).integers(0, 2**32, 5000, dtype=np.uint32)
    emu.emu_salted_sum.restype = ctypes.c_uint64
    assert emu.emu_salted_sum(w.ctypes.data_as(_up), ctypes.c_uint64(w.size)) == oracle.salted_sum(w)

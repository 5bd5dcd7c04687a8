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


def emu_render(E, tv, tn, tt, tex, xres, yres, basis, guard=1, trim=2):
    px, zb = np.empty(xres * yres, np.uint32), np.empty(xres * yres, np.float32)
    tv, tn, tt = (np.ascontiguousarray(a, np.float32) for a in (tv, tn, tt))
    tex = np.ascontiguousarray(tex, np.uint32)
    f = E.emu_render(tv.ctypes.data_as(_fp), tn.ctypes.data_as(_fp), tt.ctypes.data_as(_fp), tv.shape[0], tex.ctypes.data_as(_up),
                     tex.shape[1], tex.shape[0], xres, yres, np.ascontiguousarray(basis, np.float32).ctypes.data_as(_fp),
                     px.ctypes.data_as(_up), zb.ctypes.data_as(_fp), guard, trim)
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


def _adversarial_triangles(rng, n, res):
    """Screen-space triangles built to stress bbox_trim: vertices on / next to integer pixel coordinates (the first column and row
    of the bbox then touch the triangle), slivers, sub-pixel and 40-pixel triangles, flat and steep in z (the reference's
    barycentrics use the 3-D Gram matrix, so a steep triangle's coverage is sheared against its 2-D outline)."""
    c = np.stack([rng.uniform(2, res[0] - 3, n), rng.uniform(2, res[1] - 3, n)], 1)
    size = np.exp(rng.uniform(np.log(0.05), np.log(40.0), n))[:, None, None]
    xy = c[:, None, :] + rng.uniform(-1, 1, (n, 3, 2)) * size
    snap = rng.integers(0, 4, (n, 3, 2))                                   # 0 free, 1 exactly integer, 2 integer +- 1 ulp-ish, 3 integer +- 1e-3
    r = np.rint(xy)
    xy = np.where(snap == 1, r, xy)
    xy = np.where(snap == 2, r + rng.choice([-1, 1], (n, 3, 2)) * r * 6e-8, xy)
    xy = np.where(snap == 3, r + rng.uniform(-1e-3, 1e-3, (n, 3, 2)), xy)
    sliver = rng.random(n) < 0.2
    t = rng.uniform(-0.5, 1.5, n)[:, None]
    xy[sliver, 2] = xy[sliver, 0] + (xy[sliver, 1] - xy[sliver, 0]) * t[sliver] + rng.normal(0, 1e-3, (int(sliver.sum()), 2))
    zscale = np.exp(rng.uniform(np.log(1e-4), np.log(30.0), n))[:, None]
    z = 0.5 + rng.uniform(-1, 1, (n, 3)) * zscale
    xy = np.clip(xy, 0.0, np.array([res[0] - 1.001, res[1] - 1.001]))
    return np.concatenate([xy, z[:, :, None]], 2).astype(np.float32).reshape(n, 9)


@pytest.mark.parametrize("res", [(8192, 8192), (640, 480)])
def test_bbox_trim_never_removes_an_inside_pixel(emu, res):
    """Brute force: every pixel that bbox_trim (gel_math.h) removes from a triangle's bbox is evaluated the reference's way
    (tbarycenter + the >= 0 test, main.c:316-332,352) -- none may be inside.  1.5 M adversarial triangles per resolution,
    one and three rounds of trimming; and the trimming must actually remove a large share of the box."""
    emu.emu_trim_soundness.restype = ctypes.c_uint64
    rng = np.random.default_rng(res[0])
    tri = _adversarial_triangles(rng, 1_500_000, res)
    for rounds in (1, 3):
        counts = (ctypes.c_uint64 * 3)(0, 0, 0)
        wrong = emu.emu_trim_soundness(tri.ctypes.data_as(_fp), tri.shape[0], res[0], res[1], rounds, counts)
        assert wrong == 0, f"{wrong} inside pixels were trimmed away ({rounds} rounds)"
        assert counts[2] > 1_000_000 and counts[1] < 0.96 * counts[0]       # real coverage, real trimming (large triangles dominate the count here)


def test_bbox_trim_on_the_cfg3_sphere(emu):
    """The workload it is meant for: one view of the 1 M-triangle sphere at 4K -- nothing inside is trimmed, and most of the tested
    pixels go (the figure quoted in DESIGN.md)."""
    from gel_b200 import synth
    import gel_b200, tempfile
    emu.emu_trim_soundness.restype = ctypes.c_uint64
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "s.obj")
        open(obj, "w").write(synth.sphere_obj_text(200, 200))                  # 80 000 triangles: the cfg-3 geometry at 1/12 of the count
        tv, tn, _ = gel_b200.load_obj(obj)
    vew, _ = oracle.transform(tv, tn, oracle.view_basis(0.49, 0.0), 1100, 620)   # same pixels-per-triangle as 707 x 707 at 3840 x 2160
    for rounds, most in ((1, 0.75), (2, 0.6)):
        counts = (ctypes.c_uint64 * 3)(0, 0, 0)
        assert emu.emu_trim_soundness(vew.ctypes.data_as(_fp), vew.shape[0], 1100, 620, rounds, counts) == 0
        assert counts[1] < most * counts[0]


@pytest.mark.parametrize("res", [(8192, 8192), (640, 480)])
def test_row_trim_never_removes_a_row_that_could_pass(emu, res):
    """Brute force for the band rasteriser's per-column row trimming (gel_math.h: row_trim): every row it removes from a column of a
    triangle's bbox is evaluated the reference's way and the kernels' way -- none is inside, all are rejected by the exact cheap
    tests.  Adversarial triangles (vertices on integer coordinates, slivers, steep z, 0.05 .. 40 px) plus large ones (the tile
    pipeline's customers), and the trimming must remove a large share of the rows."""
    emu.emu_row_trim_soundness.restype = ctypes.c_uint64
    rng = np.random.default_rng(res[1])
    small = _adversarial_triangles(rng, 600_000, res)
    big = _adversarial_triangles(rng, 40_000, res)
    c = big.reshape(-1, 3, 3)[:, :, :2].mean(1, keepdims=True)
    big.reshape(-1, 3, 3)[:, :, :2] = np.clip(c + (big.reshape(-1, 3, 3)[:, :, :2] - c) * rng.uniform(1.0, 8.0, (big.shape[0], 1, 1)), 0.0, np.array([res[0] - 1.001, res[1] - 1.001]))
    for tri, least in ((small, 0.95), (big.astype(np.float32), 0.75)):
        counts = (ctypes.c_uint64 * 3)(0, 0, 0)
        wrong = emu.emu_row_trim_soundness(np.ascontiguousarray(tri).ctypes.data_as(_fp), tri.shape[0], res[0], res[1], counts)
        assert wrong == 0, f"{wrong} rows that pass were trimmed away"
        assert counts[2] > 100_000 and counts[1] < least * counts[0], (counts[0], counts[1], counts[2])
        assert counts[1] >= counts[2]                                           # every inside pixel is on a kept row


def test_row_trim_with_giant_and_degenerate_triangles(emu):
    """Triangles the error analysis must either cover or refuse: vertices 1e2 .. 1e6 pixels outside the frame (the bbox is clipped, the
    Gram terms reach 1e12 .. 1e24 -- beyond the guards the slack terms become infinite and nothing may be trimmed), near-zero area,
    and non-finite vertices (every comparison false)."""
    emu.emu_row_trim_soundness.restype = ctypes.c_uint64
    rng = np.random.default_rng(99)
    res = (640, 480)
    n = 3000
    c = np.stack([rng.uniform(0, res[0], n), rng.uniform(0, res[1], n)], 1)
    size = np.exp(rng.uniform(np.log(1e2), np.log(1e6), n))[:, None, None]
    xy = c[:, None, :] + rng.uniform(-1, 1, (n, 3, 2)) * size
    z = 0.5 + rng.uniform(-1, 1, (n, 3)) * np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))[:, None]
    giant = np.concatenate([xy, z[:, :, None]], 2).astype(np.float32).reshape(n, 9)
    thin = _adversarial_triangles(rng, 20_000, res)
    thin.reshape(-1, 3, 3)[:, 2, :2] = thin.reshape(-1, 3, 3)[:, 0, :2] + (thin.reshape(-1, 3, 3)[:, 1, :2] - thin.reshape(-1, 3, 3)[:, 0, :2]) * rng.uniform(0, 1, (20_000, 1)).astype(np.float32)
    bad = _adversarial_triangles(rng, 2_000, res)
    bad[rng.integers(0, 2_000, 300), rng.integers(0, 9, 300)] = np.array([np.nan, np.inf, -np.inf], np.float32)[rng.integers(0, 3, 300)]
    for tri in (giant, thin, bad):
        counts = (ctypes.c_uint64 * 3)(0, 0, 0)
        wrong = emu.emu_row_trim_soundness(np.ascontiguousarray(tri, np.float32).ctypes.data_as(_fp), tri.shape[0], res[0], res[1], counts)
        assert wrong == 0, f"{wrong} rows that pass were trimmed away"
        assert counts[1] >= counts[2]


def test_row_trim_on_the_cfg5_sphere(emu):
    """The workload it is meant for: views of the 5 000-triangle sphere at 1080p -- no row that passes is removed, and the rows left are
    close to the pixels inside (the figure quoted in DESIGN.md)."""
    from gel_b200 import synth
    import gel_b200, tempfile
    emu.emu_row_trim_soundness.restype = ctypes.c_uint64
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "s.obj")
        open(obj, "w").write(synth.sphere_obj_text(50, 50))
        tv, tn, _ = gel_b200.load_obj(obj)
    for xt in (0.0, 0.0767, 1.3, 3.9):
        vew, _ = oracle.transform(tv, tn, oracle.view_basis(xt, 0.0), 1920, 1080)
        counts = (ctypes.c_uint64 * 3)(0, 0, 0)
        assert emu.emu_row_trim_soundness(vew.ctypes.data_as(_fp), vew.shape[0], 1920, 1080, counts) == 0
        assert counts[2] <= counts[1] < 0.65 * counts[0]

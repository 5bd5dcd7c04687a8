"""The oracle is pinned before it is trusted: restatement == unmodified reference == committed golden vectors."""
import os

import numpy as np
import pytest

import oracle
from conftest import bits


def _angles(case):
    return [(np.uint32(f["xt_bits"]).view(np.float32), np.uint32(f["yt_bits"]).view(np.float32)) for f in case["frames"]]


def test_restatement_matches_golden_vectors(golden, cfg1_paths):
    """Every golden frame (made by the unmodified main.c) is reproduced bit-for-bit by oracle/ref_cpu.c."""
    tv, tn, tt = oracle.load_obj(cfg1_paths[0])
    tex = oracle.load_bmp(cfg1_paths[1])
    assert tv.shape == (5000, 9) and tex.shape == (256, 256)
    for case in golden["cases"]:
        for f, (xt, yt) in zip(case["frames"], _angles(case)):
            px, zb, clipped = oracle.render(tv, tn, tt, tex, case["xres"], case["yres"], oracle.view_basis(xt, yt))
            assert clipped == 0
            assert "%016x" % oracle.fnv1a64_words(px) == f["fnv"]
            assert int((px != 0).sum()) == f["nonzero"]
            assert "%016x" % oracle.salted_sum(px) == f["salted_sum"]
            assert (px >> 24).max() == 0            # output is 0x00RRGGBB


def test_mouse_angle_accumulation_matches_golden(golden):
    for case in golden["cases"]:
        ang = oracle.mouse_angles(len(case["frames"]), case["dx"], case["dy"])
        got = [(int(bits(a[:1])[0]), int(bits(a[1:])[0])) for a in ang]
        assert got == [(f["xt_bits"], f["yt_bits"]) for f in case["frames"]]


@pytest.mark.skipif(oracle.ref_binary(800, 600) is None, reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("res,dx,dy", [((800, 600), -23, 7), ((1920, 1080), 97, -31)])
def test_restatement_matches_reference_binary_live(cfg1_paths, res, dx, dy):
    """Fresh run of the unmodified reference (not the committed vectors) vs the restatement, whole frames."""
    lines, ref_px = oracle.run_reference(*cfg1_paths, res[0], res[1], frames=3, dx=dx, dy=dy)
    tv, tn, tt = oracle.load_obj(cfg1_paths[0])
    tex = oracle.load_bmp(cfg1_paths[1])
    for k, (xt, yt) in enumerate(oracle.mouse_angles(3, dx, dy)):
        px, _, _ = oracle.render(tv, tn, tt, tex, res[0], res[1], oracle.view_basis(xt, yt))
        assert np.array_equal(px, ref_px[k])
        assert "%016x" % oracle.fnv1a64_words(px) == lines[k]["fnv"]


@pytest.mark.skipif(oracle.ref_binary(800, 600, shipped=True) is None, reason="oracle/_ref not built")
def test_shipped_fast_math_build_is_not_the_oracle(cfg1_paths, golden):
    """Negative control: the reference's own -Ofast flags change a handful of edge pixels."""
    lines, _ = oracle.run_reference(*cfg1_paths, 800, 600, frames=1, shipped=True, dump=False)
    assert lines[0]["fnv"] != golden["cases"][0]["frames"][0]["fnv"]


def test_counters_and_threads_agree(cfg1):
    tv, tn, tt, tex = cfg1
    bases = np.stack([oracle.view_basis(x, 0.1 * x) for x in (0.0, 1.0, 2.0, 3.0, 4.0)])
    px0, zb0, _, c = oracle.render(tv, tn, tt, tex, 800, 600, bases[0], counters=True)
    assert (c.tested, c.inside, c.zpass, c.lit) == (648036, 210302, 205698, 105153)   # 0.65 M / 0.21 M / 0.206 M
    assert c.lit == int((zb0 != np.finfo(np.float32).min).sum())
    one = oracle.render_views(tv, tn, tt, tex, 800, 600, bases, nthreads=1, z=True, hashes=True)
    many = oracle.render_views(tv, tn, tt, tex, 800, 600, bases, nthreads=4, z=True, hashes=True)
    assert np.array_equal(one["pixel"], many["pixel"]) and np.array_equal(bits(one["z"]), bits(many["z"]))
    assert np.array_equal(one["hash"], many["hash"])
    assert np.array_equal(one["pixel"][0], px0)
    assert int(one["hash"][0, 0]) == oracle.salted_sum(px0) and int(one["hash"][0, 1]) == oracle.salted_sum(zb0)


def test_draw_order_semantics_of_the_oracle():
    """main.c:356 strict `>`: of two coincident triangles the FIRST submitted wins; zero-area triangles draw nothing."""
    tex = (np.arange(16 * 16, dtype=np.uint32).reshape(16, 16) * 0x010203) & 0xFFFFFF
    tri = np.array([[-0.3, 0.2, 0.0, 0.3, 0.2, 0.0, 0.0, 0.8, 0.0]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (1, 3))
    uv_a = np.array([[0.1, 0.1, 0, 0.2, 0.1, 0, 0.1, 0.2, 0]], np.float32)
    uv_b = np.array([[0.9, 0.9, 0, 0.8, 0.9, 0, 0.9, 0.8, 0]], np.float32)
    basis = oracle.view_basis(0.0, 0.0)
    first, _, _ = oracle.render(tri, nrm, uv_a, tex, 200, 150, basis)
    both, _, _ = oracle.render(np.vstack([tri, tri]), np.vstack([nrm, nrm]), np.vstack([uv_a, uv_b]), tex, 200, 150, basis)
    second_first, _, _ = oracle.render(np.vstack([tri, tri]), np.vstack([nrm, nrm]), np.vstack([uv_b, uv_a]), tex, 200, 150, basis)
    assert (first != 0).sum() > 100
    assert np.array_equal(first, both) and not np.array_equal(first, second_first)
    degenerate = np.array([[-0.3, 0.2, 0.0, 0.3, 0.2, 0.0, 0.0, 0.2, 0.0]], np.float32)
    empty, zb, _ = oracle.render(degenerate, nrm, uv_a, tex, 200, 150, basis)
    assert not empty.any() and (zb == np.finfo(np.float32).min).all()

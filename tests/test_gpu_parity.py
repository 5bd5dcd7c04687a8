"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (include/gelcu.h), against
 (1) the committed golden vectors made by the UNMODIFIED reference, (2) the CPU oracle on the same seeded inputs,
 (3) size-independent properties at BASELINE.json's full sizes.  Bar: bit-exact pixels, z and checksums."""
import os

import numpy as np
import pytest

import gel_b200
import oracle
from conftest import bits, random_soup
from gel_b200 import synth

pytestmark = pytest.mark.gpu
NTHREADS = max(1, min(os.cpu_count() or 1, 64))
FLT_MIN = np.finfo(np.float32).min


PIPELINE = 0      # set per test by the `pipeline` fixture: 0 = the library's choice, 1 = tile pipeline, 2 = direct pipeline


@pytest.fixture(params=[0, 1, 2], ids=["auto", "tile", "direct"], autouse=True)
def pipeline(request):
    """Every test of this module runs three times -- with the pipeline the library picks from the mesh, with the tile
    pipeline forced and with the direct pipeline forced: all three must be bit-exact on every input."""
    global PIPELINE
    PIPELINE = request.param
    yield request.param
    PIPELINE = 0


def make_renderer(xres, yres, tv, tn, tt, tex):
    r = gel_b200.Renderer(xres, yres)
    r.set_mesh(tv, tn, tt)
    r.set_texture(tex)
    if PIPELINE:
        r.set_option("pipeline", PIPELINE)
    return r


def assert_views_match(r, tv, tn, tt, tex, bases, *, expect_rc=0):
    out = r.render(bases, pixels=True, z=True, hashes=True)
    ref = oracle.render_views(tv, tn, tt, tex, r.xres, r.yres, bases, nthreads=NTHREADS, z=True, hashes=True)
    assert out["rc"] == expect_rc
    bad = int((out["pixel"] != ref["pixel"]).sum())
    assert bad == 0, f"{bad} pixels differ"
    assert np.array_equal(bits(out["z"]), bits(ref["z"]))
    assert np.array_equal(out["hash"], ref["hash"])
    plain = r.render(bases, pixels=True, z=True)                        # without checksums: the other variants of the kernels
    assert np.array_equal(plain["pixel"], ref["pixel"]) and np.array_equal(bits(plain["z"]), bits(ref["z"]))
    return out, ref


# ---- golden vectors from the unmodified reference ---------------------------------------------------

def test_golden_vectors(golden, cfg1):
    tv, tn, tt, tex = cfg1
    for case in golden["cases"]:
        ang = [(np.uint32(f["xt_bits"]).view(np.float32), np.uint32(f["yt_bits"]).view(np.float32)) for f in case["frames"]]
        with make_renderer(case["xres"], case["yres"], tv, tn, tt, tex) as r:
            out = r.render(gel_b200.view_bases(ang), hashes=True)
        for k, f in enumerate(case["frames"]):
            assert "%016x" % gel_b200.fnv1a64_words(out["pixel"][k]) == f["fnv"]
            assert int((out["pixel"][k] != 0).sum()) == f["nonzero"]
            assert "%016x" % int(out["hash"][k, 0]) == f["salted_sum"]


# ---- per-stage parity ---------------------------------------------------------------------------------

def test_stage_transform_bits(cfg1):
    tv, tn, tt, tex = cfg1
    with make_renderer(1920, 1080, tv, tn, tt, tex) as r:
        for xt, yt in [(0, 0), (1.1, 0.3), (3.3, -0.2)]:
            basis = gel_b200.view_basis(xt, yt)
            vew, shade = r.debug_transform(basis)
            rvew, rnrm = oracle.transform(tv, tn, basis, 1920, 1080)
            assert np.array_equal(bits(vew), bits(rvew))
            n = rnrm.reshape(-1, 3, 3)
            want = ((np.float32(0) * n[:, :, 0] + np.float32(0) * n[:, :, 1]) + np.float32(1) * n[:, :, 2]).astype(np.float32)
            assert np.array_equal(bits(shade), bits(want))
        uniq = len({bytes(c) for c in np.concatenate([tv.reshape(-1, 3), tn.reshape(-1, 3)], 1).view(np.uint8).reshape(-1, 24)})
        assert r.stats()["unique_vertices"] == uniq <= 2601  # 15 000 corners merge back to the OBJ's vertices (poles/seam collapse further)


def test_stage_bins(cfg1):
    """Per-tile triangle lists = exactly the triangles whose main.c:344-347 bbox overlaps the tile."""
    tv, tn, tt, tex = cfg1
    with make_renderer(800, 600, tv, tn, tt, tex) as r:
        tw, th, ntx, nty = r.tile_grid()
        basis = gel_b200.view_basis(0.7, 0.2)
        counts, entries, total = r.debug_bins(basis)
        vew, _ = oracle.transform(tv, tn, basis, 800, 600)
        v = vew.reshape(-1, 3, 3)
        x0 = np.minimum.reduce(v[:, :, 0], axis=1).astype(np.int32); x1 = np.maximum.reduce(v[:, :, 0], axis=1).astype(np.int32)
        y0 = np.minimum.reduce(v[:, :, 1], axis=1).astype(np.int32); y1 = np.maximum.reduce(v[:, :, 1], axis=1).astype(np.int32)
        want = [[] for _ in range(ntx * nty)]
        for t in range(len(v)):
            for tx in range(x0[t] // tw, x1[t] // tw + 1):
                for ty in range(y0[t] // th, y1[t] // th + 1):
                    want[tx * nty + ty].append(t)
        assert total == sum(len(w) for w in want) == len(entries)
        assert list(counts) == [len(w) for w in want]
        assert list(entries) == [t for w in want for t in w]


# ---- BASELINE.json configs ------------------------------------------------------------------------------

def test_cfg1_default_resolution(cfg1):
    tv, tn, tt, tex = cfg1
    bases = gel_b200.view_bases([(0, 0), (0.2, 0), (0.4, 0), (2.5, -0.3), (4.0, 0.2), (-1.0, 0.45)])
    with make_renderer(800, 600, tv, tn, tt, tex) as r:
        out, ref = assert_views_match(r, tv, tn, tt, tex, bases)
        st = r.stats()
        assert (st["pipeline"], st["kernels_launched"]) in ((1, 4), (2, 8), (2, 9))   # tile: batch init, transform, bin, raster; direct: batch init, transform, region, near, hi-Z, parked, fill, resolve (+ the one-time key-buffer init on a context's first call)
    assert int((out["pixel"][0] != 0).sum()) == 105139


def test_cfg2_360_views_1080p(cfg1):
    tv, tn, tt, tex = cfg1
    bases = gel_b200.view_bases(synth.view_angles(360))
    ref = oracle.render_views(tv, tn, tt, tex, 1920, 1080, bases, nthreads=NTHREADS, pixels=False, hashes=True)
    with make_renderer(1920, 1080, tv, tn, tt, tex) as r:
        out = r.render(bases, pixels=False, hashes=True)
        assert out["rc"] == 0 and np.array_equal(out["hash"], ref["hash"])
        r.set_option("batch_views", 7)                       # ragged batches: 51 x 7 + 3
        out7 = r.render(bases, pixels=False, hashes=True)
        assert np.array_equal(out7["hash"], ref["hash"]) and r.stats()["batches"] == 52
        assert_views_match(r, tv, tn, tt, tex, bases[[0, 45, 90, 179, 180, 271, 359]])


@pytest.fixture(scope="module")
def cfg3_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg3")
    obj = str(d / "sphere707.obj")
    open(obj, "w").write(synth.sphere_obj_text(707, 707))
    bmp = str(d / "tex2048.bmp")
    open(bmp, "wb").write(synth.texture_bmp_bytes(2048))
    tv, tn, tt = gel_b200.load_obj(obj)
    return tv, tn, tt, gel_b200.load_bmp(bmp)


def test_cfg3_1m_triangles_4k(cfg3_inputs):
    tv, tn, tt, tex = cfg3_inputs
    assert tv.shape[0] == 999698 and tex.shape == (2048, 2048)
    bases = gel_b200.view_bases(synth.view_angles(64)[[0, 5, 21, 40]])
    with make_renderer(3840, 2160, tv, tn, tt, tex) as r:
        out, ref = assert_views_match(r, tv, tn, tt, tex, bases)
        assert r.stats()["unique_vertices"] <= 501264
        assert r.stats()["pipeline"] == (PIPELINE or 2)       # a mesh of tiny triangles takes the direct pipeline by default
    assert int((ref["z"][0] != FLT_MIN).sum()) > 1_300_000


def test_cfg4_overdraw_and_z_ties(tmp_path):
    obj = str(tmp_path / "overdraw.obj")
    open(obj, "w").write(synth.overdraw_obj_text(100_000))
    tv, tn, tt = gel_b200.load_obj(obj)
    assert tv.shape[0] == 200_000
    tex = np.random.default_rng(9).integers(0, 1 << 24, (256, 256), dtype=np.uint32)
    bases = gel_b200.view_bases([(0.0, 0.0), (0.05, 0.02), (-0.08, 0.0), (0.1, -0.03)])
    with make_renderer(1920, 1080, tv, tn, tt, tex) as r:
        assert_views_match(r, tv, tn, tt, tex, bases)
        # every pair is coincident: dropping the SECOND copy of each pair must not change a single pixel
        first_only = np.arange(0, 200_000, 2)
        r.set_mesh(tv[first_only], tn[first_only], tt[first_only])
        single = r.render(bases[:2], z=True)
        r.set_mesh(tv, tn, tt)
        double = r.render(bases[:2], z=True)
        assert np.array_equal(single["pixel"], double["pixel"]) and np.array_equal(bits(single["z"]), bits(double["z"]))


def test_cfg5_8192_views_properties(cfg1):
    """Full-size batch: sampled views against the oracle, idempotence, and independence from batch position."""
    tv, tn, tt, tex = cfg1
    n = 8192
    bases = gel_b200.view_bases(synth.view_angles(n))
    with make_renderer(1920, 1080, tv, tn, tt, tex) as r:
        a = r.render(bases, pixels=False, hashes=True)
        b = r.render(bases[::-1].copy(), pixels=False, hashes=True)
        assert a["rc"] == 0
        assert np.array_equal(a["hash"], b["hash"][::-1])                 # a view's frame does not depend on its slot
        sample = np.arange(0, n, 257)
        ref = oracle.render_views(tv, tn, tt, tex, 1920, 1080, bases[sample], nthreads=NTHREADS, pixels=False, hashes=True)
        assert np.array_equal(a["hash"][sample], ref["hash"])
        assert len({int(h) for h in a["hash"][:, 0]}) > n // 2            # the views really differ


# ---- edge cases ------------------------------------------------------------------------------------------

def small_tex(rng, h=16, w=16):
    return rng.integers(0, 1 << 24, (h, w), dtype=np.uint32)


def test_empty_mesh_gives_cleared_frames():
    z = np.zeros((0, 9), np.float32)
    with make_renderer(200, 150, z, z, z, small_tex(np.random.default_rng(0))) as r:
        out = r.render(gel_b200.view_bases([(0, 0), (1, 0)]), z=True, hashes=True)
    assert not out["pixel"].any() and (out["z"] == FLT_MIN).all()
    assert int(out["hash"][0, 0]) == oracle.salted_sum(out["pixel"][0]) and int(out["hash"][1, 1]) == oracle.salted_sum(out["z"][1])


@pytest.mark.parametrize("res", [(200, 150), (101, 67), (33, 31), (1, 1), (640, 480)])
def test_random_soup_odd_resolutions(res):
    rng = np.random.default_rng(res[0] * 1000 + res[1])
    tv, tn, tt = random_soup(rng, 800, size=(0.01, 0.2), xr=(-0.3, 0.3), yr=(-0.2, 1.0))
    tex = small_tex(rng, 64, 32)
    bases = gel_b200.view_bases([(0, 0), (0.3, 0.1), (-0.2, -0.1)])
    with make_renderer(res[0], res[1], tv, tn, tt, tex) as r:
        out = r.render(bases, z=True, hashes=True)
        ref = oracle.render_views(tv, tn, tt, tex, res[0], res[1], bases, z=True, hashes=True)
        assert out["rc"] == (1 if ref["clipped"] else 0)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"]))
        assert np.array_equal(out["hash"], ref["hash"])


@pytest.mark.parametrize("res", [(100, 68), (36, 32), (33, 36), (132, 100), (800, 600), (64, 64), (8, 4)])
def test_tile_reset_through_tma_tensor_stores(res):
    """The tile rasteriser resets untouched tiles with TMA tensor stores (boxes of 32 rows x 4 columns clipped by the copy unit at
    the frame's edges) when no checksums are asked for and the row pitch allows it (yres % 4 == 0): frames whose last tile row /
    column is partial, frames of exactly one tile, a sparse scene (most tiles untouched), several batches into both frame
    buffers -- against the oracle, and against the store loop (option tma_reset = 0) on the same context."""
    rng = np.random.default_rng(res[0] * 7 + res[1])
    tv, tn, tt = random_soup(rng, 60, size=(0.01, 0.1), xr=(-0.3, 0.3), yr=(-0.2, 1.0))
    tex = small_tex(rng, 32, 32)
    bases = gel_b200.view_bases([(0.1 * k, 0.02 * k) for k in range(7)])
    ref = oracle.render_views(tv, tn, tt, tex, res[0], res[1], bases, nthreads=NTHREADS, z=True)
    with make_renderer(res[0], res[1], tv, tn, tt, tex) as r:
        r.set_option("batch_views", 3)                                   # three batches: both frame buffers, a ragged last one
        for tma in (1, 0, 1):
            r.set_option("tma_reset", tma)
            out = r.render(bases, z=True)
            assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])), (res, tma)


def test_ties_degenerates_and_large_triangles():
    rng = np.random.default_rng(21)
    tv, tn, tt = random_soup(rng, 300, size=(0.005, 0.05))
    big_v, big_n, big_t = random_soup(rng, 12, size=(0.3, 0.6), xr=(-0.2, 0.2), yr=(0.2, 0.8))       # many tiles each
    deg = tv[:50].copy(); deg[:, 6:9] = deg[:, 3:6]                                                    # zero area
    tv = np.vstack([tv, big_v, tv[:100], deg, big_v[:4]])
    tn = np.vstack([tn, big_n, tn[:100], tn[:50], big_n[:4]])
    tt = np.vstack([tt, big_t, rng.uniform(0, 1, (100, 9)).astype(np.float32), tt[:50], rng.uniform(0, 1, (4, 9)).astype(np.float32)])
    tex = small_tex(rng, 32, 32)
    with make_renderer(800, 600, tv, tn, tt, tex) as r:
        assert_views_match(r, tv, tn, tt, tex, gel_b200.view_bases([(0, 0), (0.25, 0.05)]))


def test_hundreds_of_large_triangles_in_one_tile():
    """More large (> 256 px in a tile) triangles than the CTA-wide sweep's list holds per round: the overflow goes
    through the per-warp unit path instead; depth order and ties must still come out exactly."""
    rng = np.random.default_rng(77)
    n = 900
    c = np.stack([rng.uniform(-0.02, 0.02, n), rng.uniform(0.48, 0.52, n), rng.uniform(-0.2, 0.2, n)], 1)
    tv = np.empty((n, 3, 3), np.float32)
    for k in range(3):
        ang = rng.uniform(0, 2 * np.pi, n)
        tv[:, k] = c + np.stack([0.12 * np.cos(ang), 0.12 * np.sin(ang), rng.uniform(-0.02, 0.02, n)], 1)
    tv = tv.reshape(n, 9)
    tn = np.tile(np.array([0, 0, 1], np.float32), (n, 3))
    tt = rng.uniform(0, 1, (n, 9)).astype(np.float32)
    tv = np.vstack([tv, tv[:50]]); tn = np.vstack([tn, tn[:50]]); tt = np.vstack([tt, rng.uniform(0, 1, (50, 9)).astype(np.float32)])
    tex = small_tex(rng, 32, 32)
    with make_renderer(640, 480, tv, tn, tt, tex) as r:
        assert_views_match(r, tv, tn, tt, tex, gel_b200.view_bases([(0, 0), (0.1, 0.05)]))


def test_thousands_of_far_triangles_in_one_tile():
    """More parked (far) triangles in a single tile than its scratch holds (4096): the rest is rasterised in the
    near phase.  A near occluder in front makes the hi-Z test reject most of the parked ones."""
    rng = np.random.default_rng(78)
    n = 7000
    c = np.stack([rng.uniform(-0.03, 0.03, n), rng.uniform(0.52, 0.545, n), rng.uniform(-0.45, -0.35, n)], 1)   # far cluster, inside one 32x32 tile
    tv = np.empty((n, 3, 3), np.float32)
    for k in range(3):
        ang = rng.uniform(0, 2 * np.pi, n)
        tv[:, k] = c + np.stack([0.004 * np.cos(ang), 0.004 * np.sin(ang), rng.uniform(-0.001, 0.001, n)], 1)
    tv = tv.reshape(n, 9)
    occluder = np.array([[-0.2, 0.3, 0.4, 0.2, 0.3, 0.4, 0.0, 0.8, 0.4], [-0.2, 0.3, 0.4, 0.0, 0.8, 0.4, -0.3, 0.8, 0.4]], np.float32)
    tv = np.vstack([occluder, tv, tv[:500]])
    m = tv.shape[0]
    tn = np.tile(np.array([0, 0, 1], np.float32), (m, 3))
    tt = rng.uniform(0, 1, (m, 9)).astype(np.float32)
    tex = small_tex(rng, 16, 16)
    with make_renderer(800, 600, tv, tn, tt, tex) as r:
        assert_views_match(r, tv, tn, tt, tex, gel_b200.view_bases([(0, 0), (0.05, 0.0)]))


def test_texture_corners_and_shading_clamp_ends():
    """uv exactly 0 and 1 (texel (0, h-1) .. (w-1, 0)); normals facing away (intensity < 0 -> shading 0)."""
    tv = np.array([[-0.4, 0.1, 0, 0.4, 0.1, 0, 0.0, 0.9, 0], [-0.4, 0.1, -0.2, 0.4, 0.1, -0.2, 0.0, 0.9, -0.2]], np.float32)
    tn = np.array([[0, 0, 1] * 3, [0, 0, -1] * 3], np.float32)
    tt = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 1, 1, 0, 1, 0, 0]], np.float32)
    tex = small_tex(np.random.default_rng(4), 8, 8)
    with make_renderer(320, 240, tv, tn, tt, tex) as r:
        out, ref = assert_views_match(r, tv, tn, tt, tex, gel_b200.view_bases([(0, 0), (np.pi, 0)]))
    assert out["pixel"][0].any()


def test_offscreen_triangles_are_clipped_and_flagged():
    rng = np.random.default_rng(33)
    tv, tn, tt = random_soup(rng, 200, size=(0.05, 0.3), xr=(-1.6, 1.6), yr=(-0.9, 1.9))
    tex = small_tex(rng)
    with make_renderer(320, 240, tv, tn, tt, tex) as r:
        out, ref = assert_views_match(r, tv, tn, tt, tex, gel_b200.view_bases([(0, 0)]), expect_rc=gel_b200.GELCU_W_CLIPPED)
        assert ref["clipped"] == 1 and r.stats()["flags"] & 1


def test_texel_out_of_range_is_clamped_and_flagged():
    tv = np.array([[-0.4, 0.1, 0, 0.4, 0.1, 0, 0.0, 0.9, 0]], np.float32)
    tn = np.array([[0, 0, 1] * 3], np.float32)
    tt = np.array([[0, 0, 0, 1.5, 0, 0, 0, 1.5, 0]], np.float32)
    with make_renderer(320, 240, tv, tn, tt, small_tex(np.random.default_rng(1))) as r:
        bases = gel_b200.view_bases([(0, 0)])
        out = r.render(bases, z=True)
        assert out["rc"] == gel_b200.GELCU_W_CLIPPED and r.stats()["flags"] & 2
        ref = oracle.render_views(tv, tn, tt, small_tex(np.random.default_rng(1)), 320, 240, bases, z=True)   # the oracle defines the case the same way
        assert ref["clipped"] & 2 and np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"]))


def test_batches_read_frame_and_mesh_replacement(cfg1):
    tv, tn, tt, tex = cfg1
    bases = gel_b200.view_bases([(0.1 * k, 0.02 * k) for k in range(10)])
    with make_renderer(800, 600, tv, tn, tt, tex) as r:
        r.set_option("batch_views", 3)
        out, ref = assert_views_match(r, tv, tn, tt, tex, bases)
        assert r.stats()["batches"] == 4
        r.render(bases, pixels=False)                      # frames stay on the device
        px, zb = r.read_frame(0)                           # slot 0 of the last batch = view 9
        assert np.array_equal(px, ref["pixel"][9]) and np.array_equal(bits(zb), bits(ref["z"][9]))
        rng = np.random.default_rng(2)
        tv2, tn2, tt2 = random_soup(rng, 100)
        r.set_mesh(tv2, tn2, tt2)
        assert_views_match(r, tv2, tn2, tt2, tex, bases[:2])


def test_state_carried_between_calls_is_clean():
    """One context, many calls: frame slots and the direct pipeline's key buffer (left "no winner" by the resolve pass,
    never cleared per batch) are reused by views whose screen regions differ and by calls with fewer views."""
    rng = np.random.default_rng(404)
    tv, tn, tt = random_soup(rng, 1500, size=(0.004, 0.03), xr=(-0.1, 0.5), yr=(0.2, 0.9), zspread=0.4)   # lopsided: the region moves with the view
    tex = small_tex(rng, 64, 64)
    views = [(0, 0), (0.5, 0.1), (-0.6, -0.1), (1.2, 0.05), (2.4, 0.0), (3.1, 0.1), (-1.5, 0.0), (0.2, 0.3)]
    with make_renderer(480, 360, tv, tn, tt, tex) as r:
        r.set_option("batch_views", 3)                                       # slots are reused inside one call, too
        for order in (views, views[::-1], views[3:5], views[1:2], views):
            bases = gel_b200.view_bases(order)
            out = r.render(bases, z=True, hashes=True)
            ref = oracle.render_views(tv, tn, tt, tex, 480, 360, bases, nthreads=NTHREADS, z=True, hashes=True)
            assert out["rc"] == (1 if ref["clipped"] else 0)
            assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"]))
            assert np.array_equal(out["hash"], ref["hash"])


def upright_rgb(canvas, xres, yres):
    """Host restatement of the presentation step: schurn's -90 degree un-rotation (main.c:424-432; the same mapping as
    gel_upright in gel_host.c) followed by dropping the X byte."""
    up = canvas.reshape(xres, yres).T[::-1]                                 # up[wy, wx] = canvas[(yres-1-wy) + wx*yres]
    return np.stack([(up >> 16) & 0xFF, (up >> 8) & 0xFF, up & 0xFF], axis=-1).astype(np.uint8)


@pytest.mark.parametrize("res", [(800, 600), (203, 97), (64, 32), (66, 35), (1, 1), (4, 3)])
def test_frame_sink_rgb8(cfg1, res):
    """gelcu_render_rgb8: upright 24-bit frames (word path when xres % 4 == 0, byte path otherwise) equal the host
    transformation of the oracle's sideways frames; several batches so both frame buffers are used."""
    tv, tn, tt, tex = cfg1
    bases = gel_b200.view_bases([(0, 0), (0.4, 0.1), (2.0, 0.0), (-1.0, -0.05), (3.0, 0.2)])
    with make_renderer(res[0], res[1], tv, tn, tt, tex) as r:
        r.set_option("batch_views", 2)
        out = r.render_rgb8(bases, hashes=True)
        ref = oracle.render_views(tv, tn, tt, tex, res[0], res[1], bases, nthreads=NTHREADS, hashes=True)
        assert out["rgb"].shape == (5, res[1], res[0], 3)
        for k in range(5):
            assert np.array_equal(out["rgb"][k], upright_rgb(ref["pixel"][k], res[0], res[1])), f"view {k}"
        assert np.array_equal(out["hash"], ref["hash"])
        again = r.render(bases, pixels=True)                                 # the plain path still works on the same context
        assert np.array_equal(again["pixel"], ref["pixel"])


def test_wide_and_compact_shading_records_agree(cfg1):
    """The direct pipeline's resolve pass reads one static record per triangle: 32 bytes when the mesh has fewer than 2^21
    distinct vertices, 64 bytes otherwise.  Force each form on the same mesh."""
    tv, tn, tt, tex = cfg1
    bases = gel_b200.view_bases([(0, 0), (1.3, 0.1)])
    ref = oracle.render_views(tv, tn, tt, tex, 400, 300, bases, nthreads=NTHREADS, z=True, hashes=True)
    for compact in (0, 1):
        with gel_b200.Renderer(400, 300) as r:
            r.set_option("compact_records", compact)
            r.set_mesh(tv, tn, tt); r.set_texture(tex); r.set_option("pipeline", 2)
            out = r.render(bases, z=True, hashes=True)
            assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])) and np.array_equal(out["hash"], ref["hash"])


# ---- indexed mesh entry: soups generated on the device ---------------------------------

def test_indexed_mesh_equals_the_soup_mesh(cfg1_paths, cfg1):
    """gelcu_set_mesh_indexed on the OBJ's arrays == gelcu_set_mesh on the soups the host expands from them
    (vmaxlen / (int) scale / tvgen / ttgen / tngen, main.c:233-286): same frames, same merged vertex count."""
    tv, tn, tt, tex = cfg1
    v, vt, vn, faces = gel_b200.load_obj_indexed(cfg1_paths[0])
    assert faces.shape == (5000, 9) and v.shape == (2601, 3)
    bases = gel_b200.view_bases([(0, 0), (0.9, 0.2), (3.5, -0.1)])
    ref = oracle.render_views(tv, tn, tt, tex, 640, 480, bases, nthreads=NTHREADS, z=True, hashes=True)
    with gel_b200.Renderer(640, 480) as r:
        r.set_mesh_indexed(v, vt, vn, faces); r.set_texture(tex)
        if PIPELINE:
            r.set_option("pipeline", PIPELINE)
        out = r.render(bases, z=True, hashes=True)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])) and np.array_equal(out["hash"], ref["hash"])
        vew, shade = r.debug_transform(bases[1])
        rvew, _ = oracle.transform(tv, tn, bases[1], 640, 480)
        assert np.array_equal(bits(vew), bits(rvew))
        assert r.stats()["unique_vertices"] == len({(int(a), int(b)) for a, b in zip(faces[:, 0:3].ravel(), faces[:, 6:9].ravel())})
        r.set_mesh(tv, tn, tt)                                            # and back to the soup entry on the same context
        again = r.render(bases, hashes=True)
        assert np.array_equal(again["hash"], ref["hash"])


def test_indexed_mesh_hard_edges_scale_and_errors():
    """A position paired with several normals (hard edges) becomes several merged vertices; max|v| = 2.9 scales by
    1.0f / (int) 2 as tvgen does (main.c:244,253); bad indices and max|v| < 1 are refused."""
    rng = np.random.default_rng(12)
    nv, nn, nt, nf = 40, 25, 30, 300
    v = rng.uniform(-1, 1, (nv, 3)).astype(np.float32); v[:, 1] = np.abs(v[:, 1]) * 0.8 + 0.1; v[:, 2] *= 0.3
    v[7] = (2.0, 2.0, 0.7)                                                # |v| = 2.91 -> scale 2
    vn = rng.normal(size=(nn, 3)).astype(np.float32); vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    vt = np.zeros((nt, 3), np.float32); vt[:, :2] = rng.uniform(0, 1, (nt, 2)); vt[:, 2] = 9.0      # uv.z is never read
    faces = np.concatenate([rng.integers(0, nv, (nf, 3)), rng.integers(0, nt, (nf, 3)), rng.integers(0, nn, (nf, 3))], 1).astype(np.int32)
    inv = np.float32(1.0) / np.float32(int(np.sqrt((v.astype(np.float32) ** 2).sum(1)).max()))
    tv = (v[faces[:, 0:3]] * inv).astype(np.float32).reshape(nf, 9)
    tt = vt[faces[:, 3:6]].reshape(nf, 9); tn = vn[faces[:, 6:9]].reshape(nf, 9)
    tex = small_tex(rng, 32, 32)
    bases = gel_b200.view_bases([(0, 0), (0.3, 0.1)])
    ref = oracle.render_views(tv, tn, tt, tex, 320, 240, bases, z=True, hashes=True)
    with gel_b200.Renderer(320, 240) as r:
        r.set_mesh_indexed(v, vt, vn, faces); r.set_texture(tex)
        if PIPELINE:
            r.set_option("pipeline", PIPELINE)
        out = r.render(bases, z=True, hashes=True)
        assert out["rc"] == (1 if ref["clipped"] else 0)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])) and np.array_equal(out["hash"], ref["hash"])
        assert r.stats()["unique_vertices"] == len({(int(a), int(b)) for a, b in zip(faces[:, 0:3].ravel(), faces[:, 6:9].ravel())}) > nv
        bad = faces.copy(); bad[5, 7] = nn                                # normal index one past the end
        with pytest.raises(gel_b200.GelcuError) as e:
            r.set_mesh_indexed(v, vt, vn, bad)
        assert e.value.code == gel_b200.GELCU_E_INVALID
        with pytest.raises(gel_b200.GelcuError):                          # a failed load leaves NO mesh behind
            r.render(bases)
        with pytest.raises(gel_b200.GelcuError):
            r.set_mesh_indexed(v * np.float32(0.2), vt, vn, faces)        # max|v| < 1: (int) maxlen == 0
        r.set_mesh_indexed(v, vt, vn, faces[:0])                          # no faces: cleared frames
        assert not r.render(bases)["pixel"].any()


def test_indexed_mesh_cfg3(cfg3_inputs, tmp_path):
    """Full size: the 1 M-triangle OBJ through the indexed entry, device checksums against the soup entry."""
    tv, tn, tt, tex = cfg3_inputs
    obj = str(tmp_path / "sphere707.obj")
    open(obj, "w").write(synth.sphere_obj_text(707, 707))
    v, vt, vn, faces = gel_b200.load_obj_indexed(obj)
    bases = gel_b200.view_bases(synth.view_angles(64)[[3, 30]])
    with make_renderer(3840, 2160, tv, tn, tt, tex) as r:
        want = r.render(bases, pixels=False, hashes=True)
        nu = r.stats()["unique_vertices"]
        r.set_mesh_indexed(v, vt, vn, faces)
        got = r.render(bases, pixels=False, hashes=True)
        assert np.array_equal(got["hash"], want["hash"])
        assert nu <= r.stats()["unique_vertices"] <= 501264              # the soup entry also merges equal-valued corners (seam, poles)


# ---- region output (dirty-rectangle contract) --------------------------------------------------------------

@pytest.mark.parametrize("rgb8", [False, True])
def test_render_region_keeps_frames_complete(rgb8):
    """gelcu_render_region into ONE reused set of frame slots: views whose regions move, shrink and grow; a slot that starts
    as garbage with an "anything" rectangle; an empty mesh at the end.  After every call each frame must equal the oracle's."""
    rng = np.random.default_rng(505)
    tv, tn, tt = random_soup(rng, 900, size=(0.004, 0.05), xr=(-0.1, 0.45), yr=(0.25, 0.8), zspread=0.4)      # lopsided: the region moves with the view
    tex = small_tex(rng, 64, 64)
    W, H, n = 456, 344, 5
    calls = [[(0, 0), (0.5, 0.1), (-0.6, -0.1), (1.2, 0.05), (2.4, 0.0)], [(3.1, 0.1), (-1.5, 0.0), (0.2, 0.3), (0.0, 0.0), (0.7, -0.2)],
             [(0.1, 0.0), (0.1, 0.0), (2.0, 0.1), (-2.0, 0.1), (1.0, 0.0)]]
    with make_renderer(W, H, tv, tn, tt, tex) as r:
        r.set_option("batch_views", 2)                                   # several batches per call, both frame buffers
        frames = np.full((n, H, W, 3), 0xAB, np.uint8) if rgb8 else np.full((n, W * H), 0xDEADBEEF, np.uint32)
        zs = None if rgb8 else np.full((n, W * H), 7.0, np.float32)
        rects = np.tile(np.array([0, 0, W - 1, H - 1], np.int32), (n, 1))       # "anything": the library resets the whole frame
        for angles in calls:
            bases = gel_b200.view_bases(angles)
            ref = oracle.render_views(tv, tn, tt, tex, W, H, bases, nthreads=NTHREADS, z=True, hashes=True)
            out = r.render_region(bases, frames, rects, z_io=zs, rgb8=rgb8, hashes=True)
            assert out["rc"] == (1 if ref["clipped"] else 0) and np.array_equal(out["hash"], ref["hash"])
            for k in range(n):
                lit = np.argwhere(ref["z"][k].reshape(W, H) != FLT_MIN)
                x0, y0, x1, y1 = rects[k]
                assert x0 % 8 == 0 and y0 % 8 == 0 and x0 <= lit[:, 0].min() and lit[:, 0].max() <= x1 and y0 <= lit[:, 1].min() and lit[:, 1].max() <= y1
                if rgb8:
                    assert np.array_equal(frames[k], upright_rgb(ref["pixel"][k], W, H)), f"view {k}"
                else:
                    assert np.array_equal(frames[k], ref["pixel"][k]) and np.array_equal(bits(zs[k]), bits(ref["z"][k])), f"view {k}"
            assert r.stats()["d2h_bytes"] < 0.8 * n * W * H * (3 if rgb8 else 8)      # less than the full frames crossed PCIe
        empty = np.zeros((0, 9), np.float32)
        r.set_mesh(empty, empty, empty)
        out = r.render_region(gel_b200.view_bases(calls[0]), frames, rects, z_io=zs, rgb8=rgb8)
        assert not frames.any() and (rects[:, 2] < rects[:, 0]).all() and (zs is None or (zs == FLT_MIN).all())


def test_render_region_cfg2_sweep(cfg1):
    """The cfg-2 sweep through one reused canvas (the reference's own pattern: one streaming texture, main.c:504,523):
    every frame complete, a fraction of the bytes copied."""
    tv, tn, tt, tex = cfg1
    ang = synth.view_angles(360)[::9]                                     # 40 views
    W, H = 1920, 1080
    ref = oracle.render_views(tv, tn, tt, tex, W, H, gel_b200.view_bases(ang), nthreads=NTHREADS)
    with make_renderer(W, H, tv, tn, tt, tex) as r:
        canvas = np.zeros((1, W * H), np.uint32)
        rect = np.array([[0, 0, -1, -1]], np.int32)                       # a zeroed canvas holds nothing
        moved = 0
        for k in range(len(ang)):
            r.render_region(gel_b200.view_bases(ang[k:k + 1]), canvas, rect)
            assert np.array_equal(canvas[0], ref["pixel"][k]), f"view {k}"
            moved += r.stats()["d2h_bytes"]
        assert moved < 0.5 * len(ang) * 4 * W * H


def test_contexts_sharded_like_ranks_match_one_context(cfg1):
    """N contexts, each rendering its shard of the view list (contiguous blocks, gel_b200.shard_views) from its own
    host thread -- on N devices when the box has them, else all on device 0 -- give the per-view checksums of ONE context
    rendering the whole list: a frame depends only on mesh, texture and (xt, yt) (main.c:509-522)."""
    import threading
    tv, tn, tt, tex = cfg1
    n, world = 96, 4
    bases = gel_b200.view_bases(synth.view_angles(n))
    ndev = max(1, gel_b200.cu().gelcu_device_count())
    with make_renderer(1024, 768, tv, tn, tt, tex) as r:
        whole = r.render(bases, pixels=False, hashes=True)["hash"]
    parts, errors = [None] * world, []

    def shard(rank):
        try:
            lo, hi = gel_b200.shard_views(n, world, rank)
            with gel_b200.Renderer(1024, 768, device=rank % ndev) as q:
                q.set_mesh(tv, tn, tt); q.set_texture(tex)
                if PIPELINE:
                    q.set_option("pipeline", PIPELINE)
                parts[rank] = q.render(bases[lo:hi], pixels=False, hashes=True)["hash"]
        except Exception as e:                                            # noqa: BLE001 -- surfaced below
            errors.append(e)

    ts = [threading.Thread(target=shard, args=(k,)) for k in range(world)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors
    assert np.array_equal(np.concatenate(parts), whole)
    ref = oracle.render_views(tv, tn, tt, tex, 1024, 768, bases[::16], nthreads=NTHREADS, pixels=False, hashes=True)
    assert np.array_equal(whole[::16], ref["hash"])


def test_small_calls_replay_a_graph_and_stay_exact(cfg1):
    """The interactive pattern: one view per call, many calls (the library replays a captured CUDA graph for calls of up to 4
    views).  Shapes alternate (1, 3, 1, 2 views, with / without checksums, rgb8, a pageable destination), options and the mesh
    change in between -- each change re-captures; several contexts do this at once from their own host threads."""
    import threading
    tv, tn, tt, tex = cfg1
    angles = [(0.1 * k, 0.02 * (k % 5)) for k in range(24)]
    bases = gel_b200.view_bases(angles)
    W, H = 320, 240
    ref = oracle.render_views(tv, tn, tt, tex, W, H, bases, nthreads=NTHREADS, z=True, hashes=True)
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            with make_renderer(W, H, tv, tn, tt, tex) as r:
                k = 0
                for step in range(40):
                    n = int(rng.choice([1, 1, 1, 3, 2]))
                    sel = (np.arange(n) + k) % len(bases); k += n
                    mode = int(rng.integers(0, 4))
                    if mode == 0:
                        out = r.render(bases[sel], z=True, hashes=True)
                        ok = np.array_equal(out["pixel"], ref["pixel"][sel]) and np.array_equal(bits(out["z"]), bits(ref["z"][sel])) and np.array_equal(out["hash"], ref["hash"][sel])
                    elif mode == 1:
                        out = r.render(bases[sel])
                        ok = np.array_equal(out["pixel"], ref["pixel"][sel])
                    elif mode == 2:
                        out = r.render_rgb8(bases[sel])
                        ok = all(np.array_equal(out["rgb"][j], upright_rgb(ref["pixel"][sel[j]], W, H)) for j in range(n))
                    else:
                        r.render(bases[sel], pixels=False)
                        px, zb = r.read_frame(n - 1)
                        ok = np.array_equal(px, ref["pixel"][sel[-1]]) and np.array_equal(bits(zb), bits(ref["z"][sel[-1]]))
                    assert ok, (seed, step, n, mode)
                    if step == 15:
                        r.set_option("graph_small_calls", 0)
                    if step == 25:
                        r.set_option("graph_small_calls", 1); r.set_mesh(tv, tn, tt)
        except Exception as e:                                            # noqa: BLE001 -- surfaced below
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(s,)) for s in range(3)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors


def test_non_finite_vertices_match_the_oracle():
    """NaN / inf / huge coordinates in a few triangles: a NaN depth never passes `z > zbuff` (main.c:356) and den = NaN never
    draws; the rest of the scene is untouched.  Compared where x86's and CUDA's float->int conversions agree (no NaN in x / y)."""
    rng = np.random.default_rng(99)
    tv, tn, tt = random_soup(rng, 400, size=(0.01, 0.12), xr=(-0.3, 0.3), yr=(0.1, 0.9))
    tv = tv.reshape(-1, 3, 3).copy()
    tv[3, 0, 2] = np.inf; tv[9, 1, 2] = -np.inf; tv[20, 2, 2] = np.nan; tv[31, :, 2] = np.nan
    tv[40, 0, 2] = 3.0e38; tv[40, 1, 2] = -3.0e38                          # finite, but the three-term depth sum can overflow
    tv = tv.reshape(-1, 9)
    tex = small_tex(rng, 16, 16)
    bases = gel_b200.view_bases([(0, 0), (0.2, 0.05)])
    with make_renderer(320, 240, tv, tn, tt, tex) as r:
        out = r.render(bases, z=True, hashes=True)
        ref = oracle.render_views(tv, tn, tt, tex, 320, 240, bases, z=True, hashes=True)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"]))
        assert not np.isnan(out["z"]).any()


@pytest.mark.parametrize("opts", [{"fill_mode": 1, "fill_sleep_ns": 300}, {"fill_mode": 2, "fill_ctas_per_sm": 2}, {"fill_mode": 3}, {"fill_mode": 4, "fill_ctas_per_sm": 4},
                                  {"fill_mode": 5, "fill_ctas_per_sm": 2}, {"fill_after": 1, "store_hint": 3}, {"fill_after": 2, "red_hint": 0}])
def test_direct_pipeline_fill_variants(opts):
    """The background reset of the direct pipeline in its measured variants: persistent grids with L2 cache-policy
    hints and pacing, TMA bulk stores (cp.async.bulk, shared -> global) before / after the near pass, later start points, store
    hints.  Every one must leave exactly the reference's frames -- at a height that allows the 16-byte bulk granules and at one
    that does not (falls back to the store loop), with and without checksums."""
    if PIPELINE == 1:
        pytest.skip("direct pipeline only")
    rng = np.random.default_rng(606)
    tv, tn, tt = random_soup(rng, 2500, size=(0.004, 0.05), xr=(-0.15, 0.45), yr=(0.2, 0.8), zspread=0.4)
    tex = small_tex(rng, 32, 32)
    bases = gel_b200.view_bases([(0, 0), (0.6, 0.1), (-1.0, 0.0), (2.0, -0.1), (3.0, 0.2)])
    for (W, H) in ((640, 480), (333, 250)):
        ref = oracle.render_views(tv, tn, tt, tex, W, H, bases, nthreads=NTHREADS, z=True, hashes=True)
        with gel_b200.Renderer(W, H) as r:
            r.set_mesh(tv, tn, tt); r.set_texture(tex); r.set_option("pipeline", 2); r.set_option("batch_views", 2)
            for k, v in opts.items():
                r.set_option(k, v)
            for hashes in (False, True, False):
                out = r.render(bases, z=True, hashes=hashes)
                assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])), (W, H, opts, hashes)
                assert not hashes or np.array_equal(out["hash"], ref["hash"])


@pytest.mark.parametrize("opts", [{}, {"near_carveout": 100}, {"band_carveout": 85, "raster_ctas_per_sm": 7}, {"raster_mode": 0}])
def test_session3_knobs_leave_the_frames_alone(opts):
    """Round 2, session 3: the shared-memory carve-out hints and the CTA-per-tile rasteriser (a call of four batches, both frame
    buffers, copies overlapping renders) -- the reference's frames, through the whole-frame and the region entry points."""
    rng = np.random.default_rng(7707)
    tv, tn, tt = random_soup(rng, 1200, size=(0.01, 0.2), xr=(-0.4, 0.4), yr=(0.1, 0.9), zspread=0.4)
    tex = small_tex(rng, 32, 32)
    W, H = 320, 200
    bases = gel_b200.view_bases([(0.21 * k, 0.01 * k) for k in range(30)])
    ref = oracle.render_views(tv, tn, tt, tex, W, H, bases, nthreads=NTHREADS, z=True)
    with make_renderer(W, H, tv, tn, tt, tex) as r:
        for k, v in opts.items():
            r.set_option(k, v)
        out = r.render(bases, z=True)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"])), opts
        assert r.stats()["batches"] == 4
        canvas = np.full((len(bases), W * H), 0xDEADBEEF, np.uint32)
        rects = np.zeros((len(bases), 4), np.int32); rects[:] = (0, 0, W - 1, H - 1)        # garbage everywhere: everything may be stale
        r.render_region(bases, canvas, rects)
        assert np.array_equal(canvas, ref["pixel"]), opts


def test_call_order_and_argument_errors():
    r = gel_b200.Renderer(64, 64)
    with pytest.raises(gel_b200.GelcuError) as e:
        r.render(gel_b200.view_bases([(0, 0)]))
    assert e.value.code == gel_b200.GELCU_E_INVALID
    with pytest.raises(gel_b200.GelcuError):
        r.set_option("no_such_option", 1)
    r.close()
    with pytest.raises(gel_b200.GelcuError):
        gel_b200.Renderer(0, 10)
    with pytest.raises(gel_b200.GelcuError):
        gel_b200.Renderer(64, 64, device=99)


def test_headless_gel_matches_reference_output_format(cfg1_paths, golden):
    """The host C program end to end: same per-frame FNV / non-zero lines as the unmodified reference printed."""
    import json, subprocess
    from conftest import ROOT
    case = golden["cases"][1]                              # 800x600, mouse (-37, 11), 4 frames
    exe = os.path.join(ROOT, "gel_b200", "host", "gel")
    out = subprocess.run([exe, *cfg1_paths, "--frames", "4", "--mouse", "-37,11"], capture_output=True, text=True, check=True).stdout
    lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    for f, l in zip(case["frames"], lines):
        assert l["fnv"] == f["fnv"] and l["nonzero"] == f["nonzero"] and l["checksum"] == f["salted_sum"]
    assert lines[-1]["summary"] and lines[-1]["views"] == 4


def test_headless_gel_ppm_through_the_device_sink(cfg1_paths, tmp_path):
    """`gel --ppm` writes the same files whether the host un-rotates and packs the XRGB frame (gel_write_ppm) or the
    device does (--sink rgb8 -> gelcu_render_rgb8)."""
    import json, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "gel_b200", "host", "gel")
    common = [exe, *cfg1_paths, "--res", "320x200", "--frames", "3", "--mouse", "-50,7", "--batch", "2"]
    a = subprocess.run([*common, "--ppm", str(tmp_path / "host")], capture_output=True, text=True, check=True).stdout
    b = subprocess.run([*common, "--ppm", str(tmp_path / "dev"), "--sink", "rgb8"], capture_output=True, text=True, check=True).stdout
    la, lb = ([json.loads(l) for l in o.splitlines() if l.startswith("{")] for o in (a, b))
    for k in range(3):
        assert la[k]["checksum"] == lb[k]["checksum"] and la[k]["nonzero"] == lb[k]["nonzero"]
        host, dev = (open(tmp_path / f"{p}{k:04d}.ppm", "rb").read() for p in ("host", "dev"))
        assert host == dev and host.startswith(b"P6\n320 200\n255\n") and len(host) == 15 + 320 * 200 * 3


def test_headless_gel_indexed_soups_and_region_agree(cfg1_paths, tmp_path):
    """`gel` loads through gelcu_set_mesh_indexed by default (soups generated on the device); --soups takes the reference's
    host-side tvgen / ttgen / tngen + gelcu_set_mesh; --region returns frames through gelcu_render_region into reused slots.
    All three must print the same per-frame lines and dump the same frames."""
    import json, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "gel_b200", "host", "gel")
    common = [exe, *cfg1_paths, "--res", "400x304", "--sweep", "12", "--batch", "5"]
    outs = {}
    for name, extra in (("indexed", []), ("soups", ["--soups"]), ("region", ["--region"]), ("region_rgb8", ["--region", "--sink", "rgb8"]), ("rgb8", ["--sink", "rgb8"])):
        dump = tmp_path / f"{name}.raw"
        o = subprocess.run([*common, *extra, "--dump", str(dump)], capture_output=True, text=True, check=True).stdout
        outs[name] = ([json.loads(l) for l in o.splitlines() if l.startswith("{") and "summary" not in l], dump.read_bytes())
    assert outs["indexed"] == outs["soups"] == outs["region"] and len(outs["indexed"][0]) == 12
    assert outs["region_rgb8"] == outs["rgb8"]


def test_differential_fuzz_sample(monkeypatch):
    """A slice of scripts/gpu_fuzz.py (adversarial random scenes: slivers, sub-pixel and screen-filling triangles,
    duplicates, coplanar stacks, off-screen and out-of-texture inputs; both pipelines, two calls per context)."""
    import importlib.util, sys
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("gpu_fuzz", os.path.join(ROOT, "scripts", "gpu_fuzz.py"))
    fuzz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fuzz)
    monkeypatch.setattr(sys, "argv", ["gpu_fuzz.py", "24", "1000"])
    assert fuzz.main() == 0


def test_maximum_resolution_8192():
    """The largest frame the ABI accepts (8192 x 8192, 13-bit coordinates in the packed records): two views, small and
    screen-sized triangles, through the oracle's own full-size render."""
    rng = np.random.default_rng(8192)
    tv, tn, tt = random_soup(rng, 400, size=(0.002, 0.05), xr=(-0.4, 0.4), yr=(0.1, 0.9))
    bv, bn, bt = random_soup(rng, 6, size=(0.3, 0.6), xr=(-0.2, 0.2), yr=(0.3, 0.7))
    tv, tn, tt = np.vstack([tv, bv]), np.vstack([tn, bn]), np.vstack([tt, bt])
    tex = small_tex(rng, 64, 64)
    bases = gel_b200.view_bases([(0.1, 0.05), (3.0, -0.1)])
    with make_renderer(8192, 8192, tv, tn, tt, tex) as r:
        out = r.render(bases, z=True, hashes=True)
        ref = oracle.render_views(tv, tn, tt, tex, 8192, 8192, bases, nthreads=2, z=True, hashes=True)
        assert out["rc"] == (1 if ref["clipped"] else 0)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(bits(out["z"]), bits(ref["z"]))
        assert np.array_equal(out["hash"], ref["hash"])
        assert int((ref["pixel"] != 0).sum()) > 1_000_000

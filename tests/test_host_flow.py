"""The product's host C flow (gel_b200/host/gel_host.c) against the oracle's restatement of main.c:84-180, 227-286,
471-484, 506-512 -- and the ABI surface of libgelcu.so (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import gel_b200
import oracle
from conftest import ROOT, bits


def test_obj_and_bmp_loaders_match_oracle(cfg1_paths):
    tv, tn, tt = gel_b200.load_obj(cfg1_paths[0])
    otv, otn, ott = oracle.load_obj(cfg1_paths[0])
    assert tv.shape == (5000, 9)
    for a, b in ((tv, otv), (tn, otn), (tt, ott)):
        assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(gel_b200.load_bmp(cfg1_paths[1]), oracle.load_bmp(cfg1_paths[1]))


def test_obj_grammar_quirks(tmp_path):
    """Dispatch order vn / vt / v / f, 1-based indices, int-truncated scale (main.c:244): max|v| = 2.9 -> scale 2."""
    p = tmp_path / "q.obj"
    p.write_text("# comment\n\nvn 0 0 1\nvn 0 1 0\nvt 0.25 0.75 0\nvt 1 0 0\nvt 0 1 0\n"
                 "v 2.9 0 0\nv 0 1.5 0\nv 0 0 -1\nv 0.5 0.5 0.5\ng group\ns off\n"
                 "f 1/1/1 2/2/2 3/3/1\nf 4/3/2 3/2/1 2/1/1 1/1/1\n")
    tv, tn, tt = gel_b200.load_obj(str(p))
    otv, otn, ott = oracle.load_obj(str(p))
    assert tv.shape == (2, 9)
    assert np.array_equal(bits(tv), bits(otv)) and np.array_equal(bits(tn), bits(otn)) and np.array_equal(bits(tt), bits(ott))
    assert tv[0, 0] == np.float32(2.9) * np.float32(0.5)
    assert list(tt[1, :3]) == [0.0, 1.0, 0.0] and list(tn[1, :3]) == [0.0, 1.0, 0.0]


def test_obj_errors(tmp_path):
    with pytest.raises(RuntimeError):
        gel_b200.load_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 1 0 0\nvt 0 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 1/1/1\n")       # index 2 out of range
    with pytest.raises(RuntimeError):
        gel_b200.load_obj(str(bad))
    small = tmp_path / "small.obj"
    small.write_text("v 0.5 0 0\nvt 0 0 0\nvn 0 0 1\nf 1/1/1 1/1/1 1/1/1\n")   # (int)maxlen == 0 (Q5)
    with pytest.raises(RuntimeError):
        gel_b200.load_obj(str(small))


def test_parallel_obj_parse_is_identical(tmp_path, monkeypatch):
    """Files above 4 MB are parsed in chunks by several threads: same soups for every thread
    count, equal to the reference-style loader; a bad face in a late chunk is still an error."""
    from gel_b200 import synth
    text = synth.sphere_obj_text(150, 150)                                  # 45 000 triangles, ~4.9 MB
    assert len(text) > (4 << 20)
    path = tmp_path / "big.obj"
    path.write_text(text)
    ref = oracle.load_obj(str(path))
    for n in ("1", "3", "8", "32"):
        monkeypatch.setenv("GEL_PARSE_THREADS", n)
        got = gel_b200.load_obj(str(path))
        assert all(np.array_equal(bits(a), bits(b)) for a, b in zip(got[:2], ref[:2])), n
        assert np.array_equal(bits(got[2])[:, [0, 1, 3, 4, 6, 7]], bits(ref[2])[:, [0, 1, 3, 4, 6, 7]]), n
    bad = tmp_path / "bad.obj"
    bad.write_text(text + "f 1/1/1 2/2\n")
    monkeypatch.setenv("GEL_PARSE_THREADS", "4")
    with pytest.raises(RuntimeError):
        gel_b200.load_obj(str(bad))


def test_bmp_padding_and_topdown(tmp_path):
    import struct
    w, h = 3, 2                                  # row = 9 bytes + 3 pad
    rows = [bytes([10 * y + x for x in range(9)]) + b"\0\0\0" for y in range(h)]
    for sign in (1, -1):
        hdr = struct.pack("<2sIHHI", b"BM", 54 + 24, 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, sign * h, 1, 24, 0, 24, 0, 0, 0, 0)
        p = tmp_path / f"t{sign}.bmp"
        p.write_bytes(hdr + b"".join(rows))
        t = gel_b200.load_bmp(str(p))
        assert np.array_equal(t, oracle.load_bmp(str(p)))
        file_row0 = [(2 << 16) | (1 << 8) | 0, (5 << 16) | (4 << 8) | 3, (8 << 16) | (7 << 8) | 6]
        assert list(t[h - 1 if sign == 1 else 0]) == file_row0
    with pytest.raises(RuntimeError):
        gel_b200.load_bmp(str(tmp_path / "nope.bmp"))


def _bmp(w, h, bpp, body, *, palette=b"", comp=0, dib=40, off=None, ncol=0):
    import struct
    off = 14 + dib + len(palette) if off is None else off
    info = struct.pack("<IiiHHIIiiII", dib, w, h, 1, bpp, comp, len(body), 0, 0, ncol, 0) + b"\0" * (dib - 40)
    return struct.pack("<2sIHHI", b"BM", off + len(body), 0, 0, off) + info + palette + body


def test_bmp_32bit_and_8bit_palettised_equal_the_24bit_image(tmp_path):
    """The reference converts whatever IMG_Load returns to RGB888 (main.c:473-480): the same picture stored as 24-bit,
    32-bit (alpha dropped; BI_RGB and BI_BITFIELDS headers) and 8-bit palettised BMP must decode to the same texels."""
    rng = np.random.default_rng(3)
    w, h = 5, 4
    idx = rng.integers(0, 7, (h, w), dtype=np.uint8)                      # image as palette indices, top-down
    pal = rng.integers(0, 256, (7, 3), dtype=np.uint8)                    # B, G, R
    bgr = pal[idx]                                                        # (h, w, 3)
    want = (bgr[..., 2].astype(np.uint32) << 16) | (bgr[..., 1].astype(np.uint32) << 8) | bgr[..., 0]
    rows24 = b"".join(bgr[y].tobytes() + b"\0" * (-(3 * w) % 4) for y in range(h - 1, -1, -1))
    rows32 = b"".join(np.concatenate([bgr[y], np.full((w, 1), 0xAB, np.uint8)], 1).tobytes() for y in range(h - 1, -1, -1))
    rows8 = b"".join(idx[y].tobytes() + b"\0" * (-w % 4) for y in range(h - 1, -1, -1))
    palette = b"".join(bytes(c) + b"\0" for c in pal)
    files = {"a24": _bmp(w, h, 24, rows24), "a32": _bmp(w, h, 32, rows32), "a32bf": _bmp(w, h, 32, rows32, comp=3, dib=56),
             "a8": _bmp(w, h, 8, rows8, palette=palette, ncol=7)}
    for name, data in files.items():
        p = tmp_path / f"{name}.bmp"
        p.write_bytes(data)
        assert np.array_equal(gel_b200.load_bmp(str(p)), want), name
    assert np.array_equal(oracle.load_bmp(str(tmp_path / "a24.bmp")), want)


def test_bmp_headers_are_validated_before_use(tmp_path):
    """Sizes and offsets come from the file: none may drive an allocation or a seek before being checked."""
    good = _bmp(4, 4, 24, b"\x11" * 48)
    cases = {
        "huge_w": _bmp(1 << 30, 4, 24, b"\x11" * 48),                     # w*h would overflow int arithmetic
        "huge_h": _bmp(4, -(1 << 31), 24, b"\x11" * 48),                  # INT32_MIN has no absolute value
        "offset_past_end": _bmp(4, 4, 24, b"\x11" * 48, off=1 << 20),
        "truncated": good[:-20],
        "rle": _bmp(4, 4, 8, b"\x11" * 16, palette=b"\0" * 1024, comp=1),
        "sixteen_bit": _bmp(4, 4, 16, b"\x11" * 32),
        "not_bmp": b"PNG" + good[3:],
    }
    for name, data in cases.items():
        p = tmp_path / f"{name}.bmp"
        p.write_bytes(data)
        with pytest.raises(RuntimeError):
            gel_b200.load_bmp(str(p))
    p = tmp_path / "good.bmp"
    p.write_bytes(good)
    assert gel_b200.load_bmp(str(p)).shape == (4, 4)


def test_view_basis_matches_oracle_bitwise():
    rng = np.random.default_rng(7)
    for xt, yt in [(0, 0), (0.2, 0), (np.pi, 0.1), (-1.3, 0.7)] + list(rng.uniform(-7, 7, (200, 2))):
        assert np.array_equal(bits(gel_b200.view_basis(xt, yt)), bits(oracle.view_basis(xt, yt)))
    b = gel_b200.view_basis(0, 0)
    assert list(b) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1]


def test_input_step_matches_reference_accumulation():
    xt, yt = ctypes.c_float(0), ctypes.c_float(0)
    ang = oracle.mouse_angles(50, -37, 11)
    for k in range(50):
        assert (np.float32(xt.value), np.float32(yt.value)) == (ang[k, 0], ang[k, 1])
        gel_b200.host().gel_input_step(ctypes.byref(xt), ctypes.byref(yt), -37, 11)


def test_fnv_and_upright(cfg1):
    rng = np.random.default_rng(3)
    w = rng.integers(0, 2**32, 1000, dtype=np.uint32)
    assert gel_b200.fnv1a64_words(w) == oracle.fnv1a64_words(w)
    xres, yres = 5, 3
    canvas = np.arange(xres * yres, dtype=np.uint32)
    up = np.zeros(xres * yres, np.uint32)
    gel_b200.host().gel_upright(canvas.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), xres, yres, up.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    for wy in range(yres):
        for wx in range(xres):
            assert up[wy * xres + wx] == canvas[(yres - 1 - wy) + wx * yres]


def test_abi_exports_every_declared_symbol():
    """libgelcu.so loads without a GPU and exports exactly the entry points include/gelcu.h declares."""
    hdr = open(os.path.join(ROOT, "include", "gelcu.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|const char\*)\s+(gelcu_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared == sorted(gel_b200.GELCU_SYMBOLS)
    L = gel_b200.cu()
    for s in declared:
        assert hasattr(L, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "gel_b200", "libgelcu.so")], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (gelcu_[a-z0-9_]+)", nm)))
    assert exported == declared


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a device, context creation reports GELCU_E_NOGPU (skipped where a GPU exists)."""
    L = gel_b200.cu()
    if L.gelcu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gel_b200.GelcuError) as e:
        gel_b200.Renderer(800, 600)
    assert e.value.code == gel_b200.GELCU_E_NOGPU and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """Nothing under gel_b200/ or include/ may import, link or open oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "gel_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "geloracle" not in text and "ref_cpu" not in text, f
    ldd = subprocess.run(["ldd", os.path.join(ROOT, "gel_b200", "libgelcu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_headless_gel_usage_and_errors(tmp_path):
    exe = os.path.join(ROOT, "gel_b200", "host", "gel")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("args: path/to/obj path/to/bmp")      # main.c:488-492
    r = subprocess.run([exe, str(tmp_path / "none.obj"), str(tmp_path / "none.bmp")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("could not open")                      # main.c:463-467


def test_indexed_parse_expands_to_the_same_soups(cfg1_paths):
    """gel_obj_parse returns the reference's Obj (v / vt / vn lines, Face = { va,vb,vc, ta,tb,tc, na,nb,nc } 0-based, main.c:22-28,
    165-170); expanding it the way tvgen / ttgen / tngen do (main.c:242-286) gives gel_obj_load's soups bit for bit."""
    v, vt, vn, faces = gel_b200.load_obj_indexed(cfg1_paths[0])
    tv, tn, tt = gel_b200.load_obj(cfg1_paths[0])
    assert faces.min() == 0 and faces[:, 0:3].max() == len(v) - 1
    inv = np.float32(1.0) / np.float32(int(np.sqrt((v * v).sum(1, dtype=np.float32)).max()))
    assert np.array_equal(bits((v[faces[:, 0:3]] * inv).reshape(-1, 9)), bits(tv))
    assert np.array_equal(bits(vt[faces[:, 3:6]].reshape(-1, 9)), bits(tt))
    assert np.array_equal(bits(vn[faces[:, 6:9]].reshape(-1, 9)), bits(tn))


def test_bench_reference_arm_and_mouse_script(cfg1_paths):
    """`bench.py --impl reference` drives the unmodified binary through the workload's own view list: the scripted integer
    mouse steps land on the nearest multiples of 0.005 rad, and both arms print the same `config`."""
    import bench
    steps = np.array([[int(x) for x in l.split()] for l in bench.mouse_script("cfg1", 5, 4).splitlines()])
    xt = -0.005 * np.cumsum(steps[:, 0])
    assert np.allclose(xt, 2 * np.pi * (np.arange(4) + 5) / 64, atol=0.0026) and not steps[:, 1].any()
    if oracle.ref_binary(800, 600) is None:
        pytest.skip("oracle/_ref not built")
    cb = bench.run_cpu_reference("cfg1", cfg1_paths[0], cfg1_paths[1], 5000, steps=1, warmup=1, frames_per_step=2, budget_s=5.0)
    assert cb["kind"] == "reference" and cb["value"] > 0 and set(cb["single_thread_ms_per_frame"]) == {"strict_O2", "shipped_Ofast"}
    assert bench.workload_config("cfg3", 999698)["resolution"] == "3840x2160"

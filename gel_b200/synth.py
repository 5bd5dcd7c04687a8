"""Deterministic synthetic inputs for the gel render path.

The reference ships no assets (its .gitignore drops obj/), so every mesh and texture used by the tests
and by bench.py is generated here, as the same text OBJ / 24-bit BMP files the reference's loaders read
(/root/reference/main.c:129-180 for OBJ, :471-484 for the image).

Input rules every mesh obeys: max|v| in [1,2) so `(int)maxlen == 1` (main.c:244),
model on y >= 0 so every projected vertex stays inside the viewport, unit non-zero `vn`, `vt` in [0,1]^2,
faces are `f v/t/n v/t/n v/t/n` triangles.
"""
from __future__ import annotations

import math
import random
import struct

import numpy as np


def _fmt_block(prefix: str, arr: np.ndarray) -> str:
    """Lines `prefix a b c` with C-style %.6f formatting (Python's % is correctly rounded like glibc's)."""
    return "".join(f"{prefix} %.6f %.6f %.6f\n" % tuple(r) for r in arr.tolist())


def sphere_obj_text(nlat: int, nlon: int) -> str:
    """UV sphere, radius 0.6 centred on (0, 0.6, 0): (nlat+1)(nlon+1) vertices, 2*nlat*nlon triangles.

    cfg 1/2/5: nlat = nlon = 50  -> 2 601 vertices, 5 000 triangles.
    cfg 3:     nlat = nlon = 707 -> 501 264 vertices, 999 698 triangles.
    """
    i = np.arange(nlat + 1, dtype=np.float64)[:, None]
    j = np.arange(nlon + 1, dtype=np.float64)[None, :]
    theta = math.pi * i / nlat
    phi = 2.0 * math.pi * j / nlon
    nx = np.sin(theta) * np.cos(phi)
    ny = np.cos(theta) * np.ones_like(phi)
    nz = np.sin(theta) * np.sin(phi)
    n = np.stack([nx, ny, nz], axis=-1).reshape(-1, 3)
    v = np.stack([0.6 * nx, 0.6 + 0.6 * ny, 0.6 * nz], axis=-1).reshape(-1, 3)
    t = np.stack([(j / nlon) * np.ones_like(i), (1.0 - i / nlat) * np.ones_like(j), np.zeros((nlat + 1, nlon + 1))],
                 axis=-1).reshape(-1, 3)
    out = [_fmt_block("v", v), _fmt_block("vt", t), _fmt_block("vn", n)]
    ii, jj = np.meshgrid(np.arange(nlat), np.arange(nlon), indexing="ij")
    a = (ii * (nlon + 1) + jj + 1).reshape(-1)
    b = ((ii + 1) * (nlon + 1) + jj + 1).reshape(-1)
    c = ((ii + 1) * (nlon + 1) + jj + 2).reshape(-1)
    d = (ii * (nlon + 1) + jj + 2).reshape(-1)
    faces = np.empty((a.size * 2, 3), dtype=np.int64)
    faces[0::2] = np.stack([a, b, c], axis=1)
    faces[1::2] = np.stack([a, c, d], axis=1)
    out.append("".join("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (p, p, p, q, q, q, r, r, r) for p, q, r in faces.tolist()))
    return "".join(out)


def overdraw_obj_text(npairs: int = 100_000, seed: int = 4242, yres: int = 1080) -> str:
    """cfg 4: `npairs` random small triangles (edge ~4-12 px at `yres`) in the slab x in [-.5,.5],
    y in [.1,1], z in [-.05,.05]; every triangle is emitted TWICE with different `vt` (exact z ties,
    the lower face index must win, main.c:356) and random winding (no back-face cull, main.c:352).
    One extra far vertex (0,1.2,0) keeps max|v| in [1,2) (main.c:244); it is referenced by no face.
    """
    rng = np.random.default_rng(seed)
    px = 1.5 / yres  # one pixel in model units at z ~ 0 (viewport scale yres/1.5, main.c:290)
    centre = np.stack([rng.uniform(-0.5, 0.5, npairs), rng.uniform(0.1, 1.0, npairs), rng.uniform(-0.05, 0.05, npairs)], 1)
    corners = []
    for _ in range(3):
        ang = rng.uniform(0, 2 * math.pi, npairs)
        rad = rng.uniform(2.0, 6.0, npairs) * px
        dz = rng.uniform(-0.004, 0.004, npairs)
        corners.append(centre + np.stack([rad * np.cos(ang), rad * np.sin(ang), dz], 1))
    verts = np.stack(corners, 1).reshape(-1, 3)  # 3 per pair
    verts = np.vstack([verts, [[0.0, 1.2, 0.0]]])
    uv = rng.uniform(0.0, 1.0, (npairs * 6, 2))
    uv3 = np.concatenate([uv, np.zeros((uv.shape[0], 1))], 1)
    nrm = rng.normal(size=(npairs * 3, 3)) * 0.3 + np.array([0.0, 0.0, 1.0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    flip = rng.integers(0, 2, npairs)
    out = [_fmt_block("v", verts), _fmt_block("vt", uv3), _fmt_block("vn", nrm)]
    lines = []
    for k in range(npairs):
        v0 = 3 * k + 1
        order = (0, 1, 2) if flip[k] == 0 else (0, 2, 1)
        for dup in range(2):
            t0 = 6 * k + 3 * dup + 1
            lines.append("f " + " ".join(f"{v0 + o}/{t0 + o}/{v0 + o}" for o in order) + "\n")
    out.append("".join(lines))
    return "".join(out)


def texture_bmp_bytes(n: int, seed: int = 1234) -> bytes:
    """n x n uncompressed 24-bit bottom-up BMP.  File bytes per pixel (B,G,R) =
    (x-ramp ^ 0x55*checker16, y-ramp, Random(seed) byte), the RNG drawn in file row order."""
    rng = random.Random(seed)
    rowbytes = (3 * n + 3) & ~3
    pad = b"\0" * (rowbytes - 3 * n)
    xs = np.arange(n)
    rows = []
    rnd = bytes(rng.randrange(256) for _ in range(n * n)) if n <= 512 else None
    if rnd is None:  # large textures: same recipe, one getrandbits call per row (still deterministic)
        rnd = b"".join(rng.getrandbits(8 * n).to_bytes(n, "little") for _ in range(n))
    rnd = np.frombuffer(rnd, dtype=np.uint8).reshape(n, n)
    for y in range(n):
        chk = ((xs // 16) + (y // 16)) & 1
        b = ((xs * 255 // (n - 1)) ^ (0x55 * chk)).astype(np.uint8)
        g = np.full(n, y * 255 // (n - 1), dtype=np.uint8)
        r = rnd[y]
        rows.append(np.stack([b, g, r], 1).tobytes() + pad)
    body = b"".join(rows)
    hdr = struct.pack("<2sIHHI", b"BM", 54 + len(body), 0, 0, 54)
    hdr += struct.pack("<IiiHHIIiiII", 40, n, n, 1, 24, 0, len(body), 2835, 2835, 0, 0)
    return hdr + body


def write_inputs(obj_path: str, bmp_path: str, *, kind: str = "sphere", nlat: int = 50, nlon: int = 50,
                 tex: int = 256, **kw) -> None:
    text = sphere_obj_text(nlat, nlon) if kind == "sphere" else overdraw_obj_text(**kw)
    with open(obj_path, "w") as f:
        f.write(text)
    with open(bmp_path, "wb") as f:
        f.write(texture_bmp_bytes(tex))


def view_angles(n: int, *, dtype=np.float32) -> np.ndarray:
    """`n` views xt_k = 2*pi*k/n (rounded to float32), yt = 0 -- the sweeps of cfg 2/3/5."""
    xt = (2.0 * math.pi * np.arange(n, dtype=np.float64) / n).astype(dtype)
    return np.stack([xt, np.zeros(n, dtype=dtype)], 1)


if __name__ == "__main__":
    import sys
    nlat, nlon, tex = (int(a) for a in sys.argv[1:4])
    write_inputs(sys.argv[4], sys.argv[5], nlat=nlat, nlon=nlon, tex=tex)

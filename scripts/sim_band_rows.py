#!/usr/bin/env python
"""Host estimate for the band rasteriser (gel_band.cuh): row iterations of the column-unit loop per view with the plain bbox row range
and with each unit's rows trimmed to the rows that can pass the cheap tests (per-unit row trimming).  Units are grouped 32 at a time in
entry order; an iteration count is the maximum over the group (warp-uniform trip count).  Oracle transform, float64 setup: an estimate."""
import sys, os, tempfile, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from gel_b200 import synth

SLACK = int(sys.argv[2]) if len(sys.argv) > 2 else 0
xres, yres, k, nv = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 100, 8192
d = tempfile.mkdtemp()
synth.write_inputs(d + "/m.obj", d + "/t.bmp", nlat=50, nlon=50, tex=256)
tv, tt, tn = oracle.load_obj(d + "/m.obj")
basis = oracle.view_basis(2 * math.pi * k / nv, 0.0)
vew, _ = oracle.transform(tv, tn, basis, xres, yres)
T = vew.reshape(-1, 3, 3).astype(np.float64)
zmax = T[:, :, 2].max(1)
zs = vew.reshape(-1, 3)[:, 2]
zthr = zs.min() + 0.4 * (zs.max() - zs.min())
a, b, c = T[:, 0], T[:, 1], T[:, 2]
v0, v1 = b - a, c - a
d00, d01, d11 = (v0 * v0).sum(1), (v0 * v1).sum(1), (v1 * v1).sum(1)
den = d00 * d11 - d01 * d01
x0 = np.trunc(T[:, :, 0].min(1)).astype(int); x1 = np.trunc(T[:, :, 0].max(1)).astype(int)
y0 = np.trunc(T[:, :, 1].min(1)).astype(int); y1 = np.trunc(T[:, :, 1].max(1)).astype(int)
near = (zmax >= zthr) & (np.abs(den) > 0)
bands = {}
for t in np.nonzero(near)[0]:
    for tx in range(x0[t] // 32, x1[t] // 32 + 1):
        for ty in range(y0[t] // 32, y1[t] // 32 + 1):
            for bnd in range(4):
                px0 = tx * 32 + bnd * 8
                if x1[t] < px0 or x0[t] > px0 + 7: continue
                bands.setdefault((tx, ty, bnd), []).append(t)
it_before = it_after = units = groups = inside_px = 0
for (tx, ty, bnd), tris in bands.items():
    px0, py0 = tx * 32 + bnd * 8, ty * 32
    ulist = []
    for t in tris:
        cx0, cx1 = max(x0[t], px0), min(x1[t], px0 + 7)
        cy0, cy1 = max(y0[t], py0), min(y1[t], py0 + 31)
        if cy0 > cy1: continue
        ys = np.arange(cy0, cy1 + 1)
        for x in range(cx0, cx1 + 1):
            v2x, v2y, v2z = x - a[t, 0], ys - a[t, 1], -a[t, 2]
            d20 = v2x * v0[t, 0] + v2y * v0[t, 1] + v2z * v0[t, 2]
            d21 = v2x * v1[t, 0] + v2y * v1[t, 1] + v2z * v1[t, 2]
            sg = 1.0 if den[t] > 0 else -1.0
            nvv = (d11[t] * d20 - d01[t] * d21) * sg; nww = (d00[t] * d21 - d01[t] * d20) * sg
            ok = (nvv >= 0) & (nww >= 0) & (nvv + nww <= abs(den[t]))
            inside_px += int(ok.sum())
            if ok.any():
                idx = np.nonzero(ok)[0]; trimmed = idx[-1] - idx[0] + 1 + SLACK
                trimmed = min(trimmed, cy1 - cy0 + 1)
            else: trimmed = 0
            ulist.append((cy1 - cy0 + 1, trimmed))
    units += len(ulist)
    # batches of 32 triangles are approximated by 32 consecutive units
    for g in range(0, len(ulist), 32):
        grp = ulist[g:g + 32]; groups += 1
        it_before += max(u[0] for u in grp); it_after += max(u[1] for u in grp)
print(f"view {k}: near tris {int(near.sum())}, bands {len(bands)}, units {units}, groups {groups}, inside px {inside_px}")
print(f"row iterations: bbox {it_before}  trimmed {it_after}  ({100.0 * it_after / it_before:.1f} %), per group {it_before / groups:.1f} -> {it_after / groups:.1f}")

#!/usr/bin/env python
"""Differential fuzz on the GPU box: random in-domain but adversarial scenes, both pipelines, bit-exact against the CPU
oracle (pixels, z bit patterns, checksums, clipped flag).   python scripts/gpu_fuzz.py [seeds] [first_seed]
Scene ingredients (mixed per seed): slivers, sub-pixel triangles, screen-filling triangles, exact duplicates (z ties),
coplanar stacks, vertices on integer pixel coordinates after projection is NOT attempted (projection is fp32), zero-area
triangles, triangles crossing the screen edges (clip flag), uv outside [0,1] (texel clamp flag), tiny / odd resolutions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gel_b200, oracle

NT = max(1, min(os.cpu_count() or 1, 32))


def scene(rng):
    parts_v, parts_n, parts_t = [], [], []

    def add(n, size, xr, yr, zr, uvr=(0.0, 1.0)):
        c = np.stack([rng.uniform(*xr, n), rng.uniform(*yr, n), rng.uniform(*zr, n)], 1)
        tv = np.empty((n, 3, 3))
        for k in range(3):
            ang = rng.uniform(0, 2 * np.pi, n); rad = rng.uniform(*size, n)
            tv[:, k] = c + np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-1, 1, n) * rad * rng.uniform(0, 1)], 1)
        tn = rng.normal(size=(n, 3, 3)) + np.array([0, 0, rng.uniform(0, 2)])
        tn /= np.maximum(np.linalg.norm(tn, axis=2, keepdims=True), 1e-9)
        tt = np.zeros((n, 3, 3)); tt[:, :, :2] = rng.uniform(*uvr, (n, 3, 2))
        parts_v.append(tv.reshape(n, 9)); parts_n.append(tn.reshape(n, 9)); parts_t.append(tt.reshape(n, 9))

    kinds = rng.integers(0, 2, 9)
    wide = rng.integers(0, 2)
    add(int(rng.integers(50, 2500)), (0.002, 0.03), (-0.5, 0.5) if wide else (-0.2, 0.2), (0.0, 1.0) if wide else (0.3, 0.7), (-0.3, 0.3) if wide else (-0.15, 0.15))   # small
    if kinds[0] and wide: add(int(rng.integers(1, 30)), (0.3, 1.5), (-0.3, 0.3), (0.2, 0.8), (-0.2, 0.2))          # huge, some off screen
    if kinds[1]: add(int(rng.integers(10, 400)), (1e-5, 5e-4), (-0.2, 0.2), (0.3, 0.7), (-0.15, 0.15))      # sub-pixel
    if kinds[2]: add(int(rng.integers(10, 300)), (0.01, 0.1), (-0.2, 0.2), (0.3, 0.7), (0.0, 0.0))        # coplanar (z = 0 before view)
    if kinds[3] and wide: add(int(rng.integers(5, 100)), (0.01, 0.2), (-0.9, 0.9), (-0.4, 1.4), (-0.3, 0.3))       # crossing the screen edges
    if kinds[4] and wide: add(int(rng.integers(5, 100)), (0.01, 0.1), (-0.2, 0.2), (0.3, 0.7), (-0.15, 0.15), uvr=(-0.3, 1.3))   # texel out of range
    if kinds[8] and wide: add(int(rng.integers(1, 12)), (5.0, 400.0), (-0.3, 0.3), (0.2, 0.8), (-0.05, 0.05))        # giant: edges of 1e3..1e5 px, den beyond the guard range
    tv, tn, tt = (np.vstack(p).astype(np.float32) for p in (parts_v, parts_n, parts_t))
    n = tv.shape[0]
    if kinds[5]:                                                                                           # exact duplicates, different uv
        k = int(rng.integers(1, max(2, n // 3))); sel = rng.integers(0, n, k)
        tv = np.vstack([tv, tv[sel]]); tn = np.vstack([tn, tn[sel]])
        t2 = np.zeros((k, 9), np.float32); t2.reshape(k, 3, 3)[:, :, :2] = rng.uniform(0, 1, (k, 3, 2)); tt = np.vstack([tt, t2])
    if kinds[6]:                                                                                           # slivers and zero-area
        k = int(rng.integers(1, 60)); sel = rng.integers(0, tv.shape[0], k); s = tv[sel].copy()
        s[:, 6:9] = s[:, 3:6] + (s[:, 3:6] - s[:, 0:3]) * rng.choice([0.0, 1e-4, 1.0, -0.5], (k, 1)).astype(np.float32)
        tv = np.vstack([tv, s]); tn = np.vstack([tn, tn[sel]]); tt = np.vstack([tt, tt[sel]])
    if kinds[7]:                                                                                           # shuffle the draw order
        p = rng.permutation(tv.shape[0]); tv, tn, tt = tv[p], tn[p], tt[p]
    return tv, tn, tt


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time(); bad = 0
    for seed in range(first, first + seeds):
        rng = np.random.default_rng(seed)
        tv, tn, tt = scene(rng)
        res = [(320, 240), (200, 150), (97, 203), (64, 64), (33, 17), (640, 360), (8, 8), (1, 1)][int(rng.integers(0, 8))]
        tex = rng.integers(0, 1 << 24, (int(rng.integers(1, 70)), int(rng.integers(1, 70))), dtype=np.uint32)
        bases = gel_b200.view_bases([(float(rng.uniform(-3.2, 3.2)), float(rng.uniform(-0.3, 0.3))) for _ in range(3)])
        ref = oracle.render_views(tv, tn, tt, tex, res[0], res[1], bases, nthreads=NT, z=True, hashes=True)
        for pl in (1, 2):
            with gel_b200.Renderer(*res) as r:
                r.set_mesh(tv, tn, tt); r.set_texture(tex); r.set_option("pipeline", pl); r.set_option("batch_views", int(rng.integers(1, 4)))
                for rep in range(2):                                                                        # twice: state carried between calls
                    try:
                        out = r.render(bases, z=True, hashes=True)
                    except gel_b200.GelcuError as e:
                        print(f"seed {seed} pipeline {pl}: {e}"); bad += 1; break
                    ok = (np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(out["z"].view(np.uint32), ref["z"].view(np.uint32))
                          and np.array_equal(out["hash"], ref["hash"]))
                    # flag 1 (bbox off screen) must agree; flag 2 (texel clamped) is raised by the product for final winners only, by the
                    # oracle for every fragment that passed the depth test when it was drawn -- a subset
                    fl = int(r.stats()["flags"])
                    ok = ok and (fl & 1) == (int(ref["clipped"]) & 1) and (fl & 2) <= (int(ref["clipped"]) & 2) and (out["rc"] != 0) == (fl != 0)
                    if ok and rep == 1:                                                                      # and without checksums (the other variants of the kernels)
                        plain = r.render(bases, z=True)
                        ok = np.array_equal(plain["pixel"], ref["pixel"]) and np.array_equal(plain["z"].view(np.uint32), ref["z"].view(np.uint32))
                    if not ok:
                        bad += 1
                        print(f"seed {seed} pipeline {pl} rep {rep} res {res} tris {tv.shape[0]}: MISMATCH pixels {(out['pixel'] != ref['pixel']).sum()} "
                              f"z {(out['z'].view(np.uint32) != ref['z'].view(np.uint32)).sum()} rc {out['rc']} clipped {ref['clipped']}")
                        break
                # region output into frame slots that hold garbage / another view's frame, with "anything" or stale rectangles
                nv = len(bases); W, H = res
                frames = rng.integers(0, 1 << 32, (nv, W * H), dtype=np.uint32); zs = rng.uniform(-1, 1, (nv, W * H)).astype(np.float32)
                rects = np.tile(np.array([0, 0, W - 1, H - 1], np.int32), (nv, 1))
                for order in (np.arange(nv), np.arange(nv)[::-1].copy()):
                    out = r.render_region(bases[order], frames, rects, z_io=zs, hashes=True)
                    if not (np.array_equal(frames, ref["pixel"][order]) and np.array_equal(zs.view(np.uint32), ref["z"][order].view(np.uint32))
                            and np.array_equal(out["hash"], ref["hash"][order])):
                        bad += 1
                        print(f"seed {seed} pipeline {pl} res {res}: REGION MISMATCH pixels {(frames != ref['pixel'][order]).sum()}"); break
                # the same scene through the indexed entry: corners become (position, normal) index pairs of a random small vertex pool
                if pl == 1 and seed % 3 == 0:
                    pv, inv = np.unique(tv.reshape(-1, 3), axis=0, return_inverse=True)
                    pn, inn = np.unique(tn.reshape(-1, 3), axis=0, return_inverse=True)
                    pt, itt = np.unique(tt.reshape(-1, 3), axis=0, return_inverse=True)
                    scale = np.float32(int(np.sqrt((pv.astype(np.float32) ** 2).sum(1, dtype=np.float32)).max())) if len(pv) else np.float32(0)
                    if scale == 1:                                           # (int) max|v| == 1: the scaled positions are the positions
                        faces = np.concatenate([inv.reshape(-1, 3), itt.reshape(-1, 3), inn.reshape(-1, 3)], 1).astype(np.int32)
                        r.set_mesh_indexed(pv, pt, pn, faces)
                        out = r.render(bases, z=True, hashes=True)
                        if not (np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(out["hash"], ref["hash"])):
                            bad += 1; print(f"seed {seed} res {res}: INDEXED MISMATCH pixels {(out['pixel'] != ref['pixel']).sum()}")
    print(f"fuzz: {seeds} seeds x 2 pipelines x 2 calls, {bad} failures, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

#!/bin/bash
# compute-sanitizer on small invocations of both pipelines: memcheck (out-of-bounds / misaligned), racecheck (shared-memory
# hazards in the tile rasteriser and the per-warp scratch), initcheck (uninitialised global reads)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys, gzip, tempfile, numpy as np
sys.path.insert(0, os.getcwd())
import gel_b200, oracle
g = "tests/golden"
td = tempfile.mkdtemp()
open(td + "/s.obj", "wb").write(gzip.open(g + "/sphere50.obj.gz").read()); open(td + "/t.bmp", "wb").write(gzip.open(g + "/tex256.bmp.gz").read())
tv, tn, tt = gel_b200.load_obj(td + "/s.obj"); tex = gel_b200.load_bmp(td + "/t.bmp")
rng = np.random.default_rng(5)
sys.path.insert(0, "tests"); from conftest import random_soup
sv, sn, st = random_soup(rng, 3000, size=(0.003, 0.08))
for pl in (1, 2):
    for (W, H, mesh) in ((320, 240, (tv, tn, tt)), (203, 131, (sv, sn, st))):
        r = gel_b200.Renderer(W, H); r.set_mesh(*mesh); r.set_texture(tex); r.set_option("pipeline", pl); r.set_option("batch_views", 2)
        bases = gel_b200.view_bases([(0, 0), (0.4, 0.1), (2.0, -0.2)])
        out = r.render(bases, z=True, hashes=True)
        ref = oracle.render_views(*mesh, tex, W, H, bases, nthreads=3, z=True, hashes=True)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(out["hash"], ref["hash"]), (pl, W, H)
        out = r.render(bases[::-1], z=True, hashes=True)                     # second call on the same context (key buffer restored by the resolve pass)
        assert np.array_equal(out["pixel"], ref["pixel"][::-1]), (pl, W, H, "second call")
        up = ref["pixel"][0].reshape(W, H).T[::-1]                           # frame sink: upright 24-bit frames
        rgb = r.render_rgb8(bases)["rgb"]
        assert np.array_equal(rgb[0], np.stack([(up >> 16) & 255, (up >> 8) & 255, up & 255], -1).astype(np.uint8)), (pl, W, H, "sink")
        # round 2 entry points: region output into reused slots (sideways + rgb8), indexed mesh (device soup generation),
        # the persistent overlapped fill with cache hints
        frames = np.zeros((3, W * H), np.uint32); rects = np.tile(np.array([0, 0, -1, -1], np.int32), (3, 1))
        r.render_region(bases, frames, rects, hashes=True)
        assert np.array_equal(frames, ref["pixel"]), (pl, W, H, "region")
        r.render_region(bases[::-1], frames, rects)
        assert np.array_equal(frames, ref["pixel"][::-1]), (pl, W, H, "region, second call")
        r.set_option("fill_mode", 1); r.set_option("red_hint", 1)
        out = r.render(bases, z=True, hashes=True)
        assert np.array_equal(out["pixel"], ref["pixel"]) and np.array_equal(out["hash"], ref["hash"]), (pl, W, H, "fill_mode 1")
        r.close()
v, vt, vn, faces = gel_b200.load_obj_indexed(td + "/s.obj")
r = gel_b200.Renderer(320, 240); r.set_mesh_indexed(v, vt, vn, faces); r.set_texture(tex)
bases = gel_b200.view_bases([(0, 0), (0.4, 0.1)])
ref = oracle.render_views(tv, tn, tt, tex, 320, 240, bases, nthreads=2, hashes=True)
assert np.array_equal(r.render(bases, hashes=True)["hash"], ref["hash"]), "indexed"
r.close()
print("sanitizer workload ok")
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done

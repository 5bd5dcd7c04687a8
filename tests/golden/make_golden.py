"""Regenerates tests/golden/*: the cfg-1 inputs and known answers produced by the UNMODIFIED reference.

Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_golden.py
Every hash below comes from oracle/_ref/gel_ref_<res> -- /root/reference/main.c compiled as is (800x600) or
with only its resolution literal substituted in a pipe (1920x1080) -- never from the restatement or the GPU.
"""
import gzip, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from gel_b200 import synth

subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all", "ref"], check=True, capture_output=True)
obj_text = synth.sphere_obj_text(50, 50)
bmp = synth.texture_bmp_bytes(256)
with gzip.GzipFile(os.path.join(HERE, "sphere50.obj.gz"), "wb", mtime=0) as f:
    f.write(obj_text.encode())
with gzip.GzipFile(os.path.join(HERE, "tex256.bmp.gz"), "wb", mtime=0) as f:
    f.write(bmp)
tmp_obj, tmp_bmp = "/tmp/_golden.obj", "/tmp/_golden.bmp"
open(tmp_obj, "w").write(obj_text); open(tmp_bmp, "wb").write(bmp)

cases = []
for (w, h, frames, dx, dy) in [(800, 600, 4, -40, 0), (800, 600, 4, -37, 11), (800, 600, 3, 211, -19),
                               (1920, 1080, 4, -37, 11), (1920, 1080, 2, 500, 40)]:
    lines, px = oracle.run_reference(tmp_obj, tmp_bmp, w, h, frames, dx, dy)
    ang = oracle.mouse_angles(frames, dx, dy)
    cases.append({"xres": w, "yres": h, "dx": dx, "dy": dy,
                  "frames": [{"xt_bits": int(np.float32(a[0]).view(np.uint32)), "yt_bits": int(np.float32(a[1]).view(np.uint32)),
                              "fnv": l["fnv"], "nonzero": l["nonzero"], "salted_sum": "%016x" % oracle.salted_sum(px[k])}
                             for k, (a, l) in enumerate(zip(ang, lines))]})
json.dump({"source": "oracle/_ref/gel_ref_<res> = unmodified /root/reference/main.c + headless SDL shim, -std=c99 -O2 -ffp-contract=off",
           "inputs": {"obj": "sphere50.obj.gz (synth.sphere_obj_text(50,50))", "bmp": "tex256.bmp.gz (synth.texture_bmp_bytes(256))"},
           "cases": cases}, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
print("wrote", len(cases), "cases")

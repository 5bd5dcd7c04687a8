#!/usr/bin/env python
"""Where a single-view call spends its time (cfg1, 800x600): destination pageable / pinned / none, region output, stage timing off."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, gel_b200
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
wd = os.path.join(tempfile.gettempdir(), "gel_b200_bench"); os.makedirs(wd, exist_ok=True)
inp = bench.build_inputs(name, wd)
_, _, _, xres, yres, *_ = bench.WORKLOADS[name]
r = gel_b200.Renderer(xres, yres); r.set_mesh(inp["tv"], inp["tn"], inp["tt"]); r.set_texture(inp["tex"])
bases = bench.step_bases(name, 400, 0)
pageable = np.empty((1, xres * yres), np.uint32)
pinned = gel_b200.PinnedBuffer((1, xres * yres), np.uint32); pinned.array[...] = 0
rect = np.array([[0, 0, -1, -1]], np.int32)
def run(label, fn, n=300):
    for k in range(20): fn(bases[k:k + 1])
    t = []
    for k in range(20, 20 + n):
        t0 = time.perf_counter(); fn(bases[k:k + 1]); t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e3
    print(f"{label:46s} median {np.median(t):.3f} ms  p10 {np.percentile(t, 10):.3f}  p99 {np.percentile(t, 99):.3f}   device_ms {r.stats()['ms_total']:.3f}")
run("gelcu_render -> pageable", lambda b: r.render(b, pixel_out=pageable))
run("gelcu_render -> pinned", lambda b: r.render(b, pixel_out=pinned.array))
run("gelcu_render, frames stay on the device", lambda b: r.render(b, pixels=False))
run("gelcu_render_region -> pinned (reused canvas)", lambda b: r.render_region(b, pinned.array, rect))
pageable[...] = 0; rect[:] = (0, 0, -1, -1)
run("gelcu_render_region -> pageable (reused canvas)", lambda b: r.render_region(b, pageable, rect))
r.set_option("stage_timing", 0)
run("gelcu_render -> pinned, stage_timing 0", lambda b: r.render(b, pixel_out=pinned.array))
for o in sys.argv[2:]:
    r.set_option(o.split("=")[0], int(o.split("=")[1]))
    run(f"gelcu_render -> pinned, {o}", lambda b: r.render(b, pixel_out=pinned.array))
    run(f"gelcu_render -> pageable, {o}", lambda b: r.render(b, pixel_out=pageable))
    run(f"gelcu_render_region -> pinned, {o}", lambda b: r.render_region(b, pinned.array, rect))

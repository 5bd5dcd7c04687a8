#!/usr/bin/env python
"""Load-time costs on the GPU box for a workload (default cfg3): OBJ parse, BMP load, gelcu_set_mesh, gelcu_set_texture."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, gel_b200
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
d = tempfile.mkdtemp()
t = time.time(); inp = bench.build_inputs(name, d); t_build = time.time() - t
t = time.time(); tv, tn, tt = gel_b200.load_obj(inp["obj"]); t_obj = time.time() - t
_, _, _, xres, yres, *_ = bench.WORKLOADS[name]
with gel_b200.Renderer(xres, yres) as r:
    t = time.time(); r.set_mesh(tv, tn, tt); t_mesh = time.time() - t
    t = time.time(); r.set_mesh(tv, tn, tt); t_mesh2 = time.time() - t
    t = time.time(); r.set_texture(inp["tex"]); t_tex = time.time() - t
    t = time.time(); v, vt, vn, faces = gel_b200.load_obj_indexed(inp["obj"]); t_parse = time.time() - t
    t = time.time(); r.set_mesh_indexed(v, vt, vn, faces); t_idx = time.time() - t
    t = time.time(); r.set_mesh_indexed(v, vt, vn, faces); t_idx2 = time.time() - t
print(f"{name}: gel_obj_parse {t_parse:.3f}s  set_mesh_indexed {t_idx:.3f}s (again {t_idx2:.3f}s)  [indexed path: no host soup expansion, no corner hashing]")
print(f"{name}: generate+load {t_build:.2f}s  gel_obj_load {t_obj:.3f}s  set_mesh {t_mesh:.3f}s (again {t_mesh2:.3f}s)  set_texture {t_tex:.3f}s  triangles {tv.shape[0]}")

#!/bin/bash
# round 2, call Q: the background fill as TMA bulk stores (fill_mode 3: before the near pass, 4: after it, 5: 3 + evict-first hint)
mkdir -p gpurun_out
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import gel_b200, oracle
from conftest import random_soup
rng = np.random.default_rng(3)
tv, tn, tt = random_soup(rng, 4000, size=(0.003, 0.04), xr=(-0.2, 0.4), yr=(0.2, 0.8))
tex = rng.integers(0, 1 << 24, (32, 32), dtype=np.uint32)
for (W, H) in ((640, 480), (1000, 604), (320, 8200 // 4 * 4)):
    bases = gel_b200.view_bases([(0, 0), (0.6, 0.1), (-1.0, 0.0), (2.0, -0.1), (3.0, 0.2)])
    ref = oracle.render_views(tv, tn, tt, tex, W, H, bases, nthreads=8, z=True)
    for mode in (3, 4, 5):
        for ctas in (1, 2):
            with gel_b200.Renderer(W, H) as r:
                r.set_mesh(tv, tn, tt); r.set_texture(tex); r.set_option("pipeline", 2); r.set_option("fill_mode", mode); r.set_option("fill_ctas_per_sm", ctas); r.set_option("batch_views", 2)
                for rep in range(2):
                    out = r.render(bases if rep == 0 else bases[::-1], z=True)
                    want = ref if rep == 0 else {"pixel": ref["pixel"][::-1], "z": ref["z"][::-1]}
                    assert np.array_equal(out["pixel"], want["pixel"]) and np.array_equal(out["z"].view(np.uint32), want["z"].view(np.uint32)), (W, H, mode, ctas, rep)
print("bulk fill parity ok")
PY
for v in "fill_mode=0" "fill_mode=3" "fill_mode=5" "fill_mode=4" "fill_mode=3 fill_ctas_per_sm=2" "fill_mode=4 fill_ctas_per_sm=2" "fill_mode=5 fill_ctas_per_sm=4" "fill_mode=4 fill_ctas_per_sm=4"; do
  opts=""; for o in $v; do opts="$opts --opt $o"; done
  timeout 600 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/bf.json 2> gpurun_out/bf.err; tail -1 gpurun_out/bf.err
  python - <<PY
import json
for l in open("gpurun_out/bf.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$v fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done

#!/bin/bash
# session 3, call K: end-to-end (gelcu_render_region into pinned frames) against the library's batch size: the first batch of a call is rendered
# before any copy can start, every other batch renders under the previous batch's copies
mkdir -p gpurun_out
for w in cfg3 cfg5 cfg2; do
for b in 0 8 4 2; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e region,region_rgb8 --batch $b > gpurun_out/k_${w}_$b.json 2> gpurun_out/k_${w}_$b.err; tail -1 gpurun_out/k_${w}_$b.err
  python - gpurun_out/k_${w}_$b.json $w $b <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print(sys.argv[2], "batch", sys.argv[3], "device fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d.get("e2e_variants", {}).items()}, "batches/step", d["run"]["library_batches_per_step"])
PY
done
done

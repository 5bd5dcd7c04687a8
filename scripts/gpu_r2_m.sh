#!/bin/bash
# round 2, call M: evict-first hints on the write-once frame stores (fill / resolve), cfg3 and cfg4
mkdir -p gpurun_out
for opt in "store_hint=0" "store_hint=1" "store_hint=2" "store_hint=3" "store_hint=3 --opt red_hint=0"; do
 for w in cfg3; do
  timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" --opt $opt > gpurun_out/sh.json 2> gpurun_out/sh.err; tail -1 gpurun_out/sh.err
  python - <<PY
import json
for l in open("gpurun_out/sh.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$w $opt fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
 done
done
timeout 600 python -m pytest tests -m gpu -q -x -k "direct and (cfg3 or cfg4 or state_carried or soup)" 2>&1 | tail -1

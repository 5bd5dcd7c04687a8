#!/bin/bash
# round 2, call R: when the trailing fill starts (as the near pass drains / after it / after the parked pass), plain and bulk
mkdir -p gpurun_out
for v in "fill_after=0" "fill_after=1" "fill_after=2" "fill_mode=4 fill_ctas_per_sm=8 fill_after=1" "fill_mode=4 fill_ctas_per_sm=8 fill_after=0" "fill_after=1 store_hint=1" "fill_after=2 store_hint=1"; do
  opts=""; for o in $v; do opts="$opts --opt $o"; done
  timeout 600 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/bf.json 2> gpurun_out/bf.err; tail -1 gpurun_out/bf.err
  python - <<PY
import json
for l in open("gpurun_out/bf.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$v fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done

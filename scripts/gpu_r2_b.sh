#!/bin/bash
# round 2, call B: paced fill under the near pass (cfg3)
mkdir -p gpurun_out
for v in "fill_mode=0" "fill_mode=1 fill_sleep_ns=500" "fill_mode=1 fill_sleep_ns=1000" "fill_mode=1 fill_sleep_ns=1500" "fill_mode=1 fill_sleep_ns=2000" "fill_mode=1 fill_sleep_ns=3000" "fill_mode=1 fill_sleep_ns=1000 red_hint=1" "fill_mode=2 fill_sleep_ns=1000" "fill_mode=1 fill_sleep_ns=2000 fill_ctas_per_sm=2" "fill_mode=1 fill_sleep_ns=4000 fill_ctas_per_sm=2"; do
  opts=""; for o in $v; do opts="$opts --opt $o"; done
  tag=$(echo $v | tr ' =' '__')
  timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/fill_$tag.json 2> gpurun_out/fill_$tag.err; echo "== $v rc=$?"; tail -1 gpurun_out/fill_$tag.err
  python - <<PY
import json
for l in open("gpurun_out/fill_$tag.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, d["clocks"])
PY
done

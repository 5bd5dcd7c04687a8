#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "direct and not 8192" 2>&1 | tail -1
for w in cfg4 cfg3 "cfg3 --views 8" "cfg1 --opt pipeline=2"; do
  timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/v.json 2> gpurun_out/v.err; tail -1 gpurun_out/v.err
  python - <<PY
import json
for l in open("gpurun_out/v.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done

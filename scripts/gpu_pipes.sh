#!/bin/bash
mkdir -p gpurun_out
for w in cfg4 cfg1 cfg3; do for pl in 1 2; do
  echo "== $w pipeline=$pl"
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-extra --opt pipeline=$pl > gpurun_out/pipe_${w}_$pl.json 2> gpurun_out/pipe.err; tail -2 gpurun_out/pipe.err
  python - <<PY
import json
for l in open("gpurun_out/pipe_${w}_$pl.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print(" ", round(d["value"]), "fps", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "views/step", d["config"]["views_per_step_per_gpu"], d["config"].get("pipeline"))
PY
done; done

#!/bin/bash
# tile-pipeline build variants (CTA size / tile width): parity on the tile-forced subset, then cfg5 / cfg1 device-timed
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  GELCU_LIB=$lib timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "tile and not 8192" > gpurun_out/pytest_$lib.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_$lib.log
  for w in cfg5 cfg1; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/tv_${lib}_$w.json 2> gpurun_out/tv_${lib}_$w.err; tail -1 gpurun_out/tv_${lib}_$w.err
    python - <<PY
import json
for l in open("gpurun_out/tv_${lib}_$w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  $w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
  done
done

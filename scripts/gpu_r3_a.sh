#!/bin/bash
# round-2 (second session) call A: tile pipeline with per-view triangle records from K2 + rotated key slots.
# parity on the tile / auto parametrisations with the new library, then cfg5 / cfg2 / cfg1 device-timed for: base (HEAD before the change), records without the rotation, records + rotation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "tile or auto" > gpurun_out/pytest_tile.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tile.log
for lib in libgelcu_base.so libgelcu_nosw.so libgelcu.so; do
  echo "== $lib"
  for w in cfg5 cfg2 cfg1; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/tv_${lib}_$w.json 2> gpurun_out/tv_${lib}_$w.err; tail -1 gpurun_out/tv_${lib}_$w.err
    python - <<PY
import json
for l in open("gpurun_out/tv_${lib}_$w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  $w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
  done
done

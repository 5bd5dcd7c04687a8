#!/bin/bash
# tile-pipeline iteration: parity (tile-forced parametrisation + auto) and quick device-timed numbers on cfg5 / cfg2 / cfg1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "tile or auto" > gpurun_out/pytest_tile.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tile.log
for w in cfg5 cfg2 cfg1; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/tile_$w.json 2> gpurun_out/tile_$w.err; tail -1 gpurun_out/tile_$w.err
  python - <<PY
import json
for l in open("gpurun_out/tile_$w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done

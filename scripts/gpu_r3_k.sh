#!/bin/bash
# call K: warp-per-band rasteriser (raster_mode 1, default) vs CTA-per-tile (raster_mode 0): parity on the tile / auto parametrisations, then cfg5 / cfg2 / cfg1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "tile or auto" > gpurun_out/pytest_tile.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tile.log
for opt in "raster_mode=0" "raster_mode=1" "raster_mode=0" "raster_mode=1"; do
  echo "== $opt"
  for w in cfg5 cfg2 cfg1; do
    timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" --opt $opt > gpurun_out/tv_${opt}_$w.json 2> gpurun_out/tv_${opt}_$w.err; tail -1 gpurun_out/tv_${opt}_$w.err
    python - <<PY
import json
for l in open("gpurun_out/tv_${opt}_$w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  $w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
  done
done

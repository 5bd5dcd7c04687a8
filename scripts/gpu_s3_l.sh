#!/bin/bash
# session 3, call L: small first batch for calls that return frames: parity, then end to end on every workload (default bench legs)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_l.log
for w in cfg3 cfg5 cfg2 cfg1 cfg4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e region,full,rgb8,region_rgb8 > gpurun_out/l_$w.json 2> gpurun_out/l_$w.err; tail -1 gpurun_out/l_$w.err
  python - gpurun_out/l_$w.json $w <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print(sys.argv[2], "device fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d.get("e2e_variants", {}).items()})
PY
done

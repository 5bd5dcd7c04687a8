#!/bin/bash
# round 2, call I: exact bbox trimming in the direct pipeline (0 / 1 / 2 rounds): parity on the direct-forced subset, cfg3 + cfg4 device-timed
mkdir -p gpurun_out
for lib in libgelcu_trim0.so libgelcu_trim1.so libgelcu_trim2.so; do
  echo "== $lib"
  GELCU_LIB=$lib timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "direct and not 8192" > gpurun_out/pytest_$lib.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_$lib.log
  for w in cfg3 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/tr_${lib}_$w.json 2> gpurun_out/tr_${lib}_$w.err; tail -1 gpurun_out/tr_${lib}_$w.err
    python - <<PY
import json
for l in open("gpurun_out/tr_${lib}_$w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  $w fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
  done
done
GELCU_LIB=libgelcu_trim2.so timeout 900 python scripts/gpu_fuzz.py 200 9000 | tail -1

#!/bin/bash
# call G: whole GPU suite on the current library, the multi-threaded graph test five more times (the capture / device-sync race),
# then K2 with staged (coalesced) record writes vs the scattered ones (m8)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "small_calls_replay" 2>&1 | tail -1; done
for lib in libgelcu_m8.so libgelcu.so libgelcu_m8.so libgelcu.so; do
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

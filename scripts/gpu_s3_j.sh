#!/bin/bash
# session 3, call J: parity with the per-view constants of K1 computed once in batch_init_kernel; cfg3 / cfg5 / cfg1 against the session's start
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_j.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu_head.so libgelcu.so; do
  echo "== $lib"
  for w in cfg3 cfg5 cfg1 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/j_${lib}_$w.json 2> gpurun_out/j_${lib}_$w.err; tail -1 gpurun_out/j_${lib}_$w.err
    show gpurun_out/j_${lib}_$w.json
  done
done
done
python scripts/latency_probe.py 2>&1 | tail -5

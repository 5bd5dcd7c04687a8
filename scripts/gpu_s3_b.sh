#!/bin/bash
# session 3, call B: parity of (1) TMA reset with 32x16 boxes issued lane-per-tile, (2) 32-bit element indices in the resolve pass,
# (3) survivor carry-over in the direct raster kernels; then HEAD vs new, same call
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu_head.so libgelcu_nocarry.so libgelcu.so; do
  echo "== $lib cfg3"
  GELCU_LIB=$lib timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/b_${lib}_$rep.json 2> gpurun_out/b_${lib}_$rep.err; tail -1 gpurun_out/b_${lib}_$rep.err
  show gpurun_out/b_${lib}_$rep.json
done
done
for lib in libgelcu_head.so libgelcu.so libgelcu_head.so libgelcu.so; do
  echo "== $lib tile"
  for w in cfg5 cfg2 cfg1 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/b_${lib}_$w.json 2> gpurun_out/b_${lib}_$w.err; tail -1 gpurun_out/b_${lib}_$w.err
    show gpurun_out/b_${lib}_$w.json
  done
done

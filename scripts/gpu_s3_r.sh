#!/bin/bash
# session 3, call R: race-free depth merge of the band rasteriser (MATCH.ANY, atomic only on a real conflict) vs the racing stores; near pass with
# vertex prefetch one batch ahead (CCTL.E.PF1); parity; racecheck
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu_merge0.so libgelcu.so; do
  echo "== $lib tile"
  for w in cfg5 cfg2 cfg1; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/r_${lib}_$w.json 2> gpurun_out/r_${lib}_$w.err; tail -1 gpurun_out/r_${lib}_$w.err
    show gpurun_out/r_${lib}_$w.json
  done
done
for lib in libgelcu.so libgelcu_pf.so; do
  echo "== $lib direct"
  for w in cfg3 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/r_${lib}_$w.json 2> gpurun_out/r_${lib}_$w.err; tail -1 gpurun_out/r_${lib}_$w.err
    show gpurun_out/r_${lib}_$w.json
  done
done
done
GELCU_LIB=libgelcu_pf.so timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "direct or golden or cfg3 or cfg4" > gpurun_out/pytest_r_pf.log 2>&1; echo "pytest pf rc=$?"; tail -2 gpurun_out/pytest_r_pf.log
bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r02_compute_sanitizer_s3.txt | tail -12

#!/bin/bash
# session 3, call D: parity of the band rasteriser's per-column row trimming (+ K2 slack terms) and of the direct pipeline's final form;
# HEAD~1 (libgelcu_head) vs no-trim vs new on the tile workloads, head vs r0 (per-view base pointers) vs new on cfg3; capture of the band kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_d.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu_head.so libgelcu_notrim.so libgelcu.so; do
  echo "== $lib tile"
  for w in cfg5 cfg2 cfg1; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/d_${lib}_$w.json 2> gpurun_out/d_${lib}_$w.err; tail -1 gpurun_out/d_${lib}_$w.err
    show gpurun_out/d_${lib}_$w.json
  done
done
for lib in libgelcu_head.so libgelcu_r0.so libgelcu.so; do
  echo "== $lib cfg3"
  GELCU_LIB=$lib timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/d_${lib}_$rep.json 2> gpurun_out/d_${lib}_$rep.err; tail -1 gpurun_out/d_${lib}_$rep.err
  show gpurun_out/d_${lib}_$rep.json
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_band_kernel -s 3 -c 1 -o gpurun_out/s3_band_cfg5_trim -f \
     python bench.py --workload cfg5 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_band_trim.log 2>&1; echo "cap rc=$?"

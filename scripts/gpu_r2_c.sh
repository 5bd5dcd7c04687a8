#!/bin/bash
# round 2, call C: resolve occupancy variants on cfg3; full ncu capture (with source) of the tile rasteriser on cfg5
mkdir -p gpurun_out
for lib in libgelcu.so libgelcu_rm6.so libgelcu_rm5.so; do
  GELCU_LIB=$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/var_$lib.json 2> gpurun_out/var_$lib.err; echo "== $lib rc=$?"; tail -1 gpurun_out/var_$lib.err
  python - <<PY
import json
for l in open("gpurun_out/var_$lib.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/r02_raster_cfg5 -f \
   python bench.py --workload cfg5 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_raster_cfg5.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# session 3, call A: state check of HEAD (quick parity subset, cfg3 / cfg5 bench), D1 occupancy variants (2- and 4-warp CTAs at 48 registers),
# ncu captures with source of the band rasteriser (cfg5) and of the direct pipeline's near pass / resolve (cfg3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "golden or cfg1 or cfg4" > gpurun_out/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_subset.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu.so libgelcu_t64b20.so libgelcu_t128b10.so; do
  echo "== $lib cfg3"
  GELCU_LIB=$lib timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/a_${lib}_$rep.json 2> gpurun_out/a_${lib}_$rep.err; tail -1 gpurun_out/a_${lib}_$rep.err
  show gpurun_out/a_${lib}_$rep.json
done
done
for w in cfg5 cfg2 cfg1; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/a_base_$w.json 2> gpurun_out/a_base_$w.err; tail -1 gpurun_out/a_base_$w.err
  show gpurun_out/a_base_$w.json
done
cap() { # cap <workload> <views> <regex> <outname> <skip> <count>
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $5 -c $6 -o gpurun_out/$4 -f \
     python bench.py --workload $1 --steps 1 --warmup 3 --views $2 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_$4.log 2>&1; echo "$4 rc=$?"; }
cap cfg5 64 raster_band_kernel s3_band_cfg5 3 1
cap cfg3 64 direct_raster_kernel s3_d1 6 2
cap cfg3 64 direct_resolve_kernel s3_d5 3 1
cap cfg3 64 transform_kernel s3_xf 3 1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# session 3, call T: 256-bit record loads (ld.global.nc.v8.b32 = LDG.E.ENL2.256) in the resolve pass, the band rasteriser's record fetch and its shade pass
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_t.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for lib in libgelcu_ld128.so libgelcu.so; do
  echo "== $lib"
  for w in cfg3 cfg5 cfg2 cfg1 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/t_${lib}_$w.json 2> gpurun_out/t_${lib}_$w.err; tail -1 gpurun_out/t_${lib}_$w.err
    show gpurun_out/t_${lib}_$w.json
  done
done
done
for lib in libgelcu_ld128.so libgelcu.so; do
GELCU_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sectors.sum,l1tex__t_requests.sum --clock-control none --csv --log-file gpurun_out/t_launch_$lib.csv \
     python bench.py --workload cfg3 --steps 1 --warmup 1 --no-extra --no-cpu --e2e "" > gpurun_out/t_launch.log 2>&1
grep "direct_resolve_kernel" gpurun_out/t_launch_$lib.csv | tail -3 | awk -F'","' '{print $5, $(NF-2), $NF}'
done

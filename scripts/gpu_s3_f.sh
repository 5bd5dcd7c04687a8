#!/bin/bash
# session 3, call F: parity of the committed state + PHASE-specific D1 scratch; near pass with 4 224 B of shared memory per warp under
# different carve-out hints (what is left of the SM's 256 KB is L1); tile workloads with the PTX queue tickets
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_f.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
  echo "== head cfg3"
  GELCU_LIB=libgelcu_head.so timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/f_head_$rep.json 2> gpurun_out/f_head_$rep.err; tail -1 gpurun_out/f_head_$rep.err
  show gpurun_out/f_head_$rep.json
  for co in -1 71 85 100 50; do
    echo "== new cfg3 near_carveout=$co"
    timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" --opt near_carveout=$co > gpurun_out/f_co${co}_$rep.json 2> gpurun_out/f_co${co}_$rep.err; tail -1 gpurun_out/f_co${co}_$rep.err
    show gpurun_out/f_co${co}_$rep.json
  done
  for w in cfg5 cfg2 cfg1 cfg4; do
    timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/f_new_$w.json 2> gpurun_out/f_new_$w.err; tail -1 gpurun_out/f_new_$w.err
    show gpurun_out/f_new_$w.json
  done
done
for co in -1 71; do
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,launch__shared_mem_config_size --clock-control none --csv --log-file gpurun_out/f_launch_$co.csv \
     python bench.py --workload cfg3 --steps 1 --warmup 1 --no-extra --no-cpu --e2e "" --opt near_carveout=$co > gpurun_out/f_launch_$co.log 2>&1
grep "direct_raster_kernel<0" gpurun_out/f_launch_$co.csv | tail -3 | cut -d, -f5,13-
done

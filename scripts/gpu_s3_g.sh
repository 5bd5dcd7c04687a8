#!/bin/bash
# session 3, call G: shared-memory carve-out hints.  The resolve pass is L1-bound and has no shared memory of its own: it inherits the carve-out of the
# raster kernel before it (196 KB -> 60 KB of L1); 0 asks for all 256 KB as L1.  The near pass ran FASTER at the 228 KB carve-out than at its own 164 KB.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
  for combo in "" "resolve_carveout=0" "resolve_carveout=0 near_carveout=100" "resolve_carveout=0 near_carveout=100 parked_carveout=100" "resolve_carveout=0 parked_carveout=71" "resolve_carveout=25" "near_carveout=100 resolve_carveout=100"; do
    echo "== cfg3 [$combo]"
    opts=""; for o in $combo; do opts="$opts --opt $o"; done
    timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/g_$rep.json 2> gpurun_out/g_$rep.err; tail -1 gpurun_out/g_$rep.err
    show gpurun_out/g_$rep.json
  done
done
for combo in "" "resolve_carveout=0" "resolve_carveout=0 near_carveout=100"; do
  opts=""; for o in $combo; do opts="$opts --opt $o"; done
  timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,launch__shared_mem_config_size --clock-control none --csv --log-file gpurun_out/g_launch.csv \
     python bench.py --workload cfg3 --steps 1 --warmup 1 --no-extra --no-cpu --e2e "" $opts > gpurun_out/g_launch.log 2>&1
  echo "== launches [$combo]"
  python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/g_launch.csv")) if len(r) > 10 and r[0].isdigit()]
ids = sorted({int(r[0]) for r in rows})[-7:]
for i in ids:
    rr = [r for r in rows if int(r[0]) == i]
    print("  ", rr[0][4].split("(")[0][-34:], {r[-3].split("__")[-1][:22]: r[-1] for r in rr})
PY
done

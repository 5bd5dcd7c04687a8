#!/bin/bash
# session 3, call C: which of the direct-pipeline changes costs what -- per-kernel durations (ncu launch list of one step) and bench per variant
# head = last commit; S = slab split only; R1 / R2 = S + resolve pass on 32-bit element indices (2: restrict bases); C = S + survivor carry-over
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for v in head S R1 R2 C; do
  lib=libgelcu_$v.so
  GELCU_LIB=$lib timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/c_${v}_$rep.json 2> gpurun_out/c_${v}_$rep.err; tail -1 gpurun_out/c_${v}_$rep.err
  echo "== $v"; show gpurun_out/c_${v}_$rep.json
done
done
for v in head S R1 R2 C; do
  GELCU_LIB=libgelcu_$v.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launch_$v.csv \
     python bench.py --workload cfg3 --steps 1 --warmup 1 --no-extra --no-cpu --e2e "" > gpurun_out/c_launch_$v.log 2>&1
  python - gpurun_out/c_launch_$v.csv $v <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows[-8:]:
    name = r[4].split("(")[0][-40:]; acc[name] = acc.get(name, 0) + float(r[-1].replace(",", ""))
print(sys.argv[2], {k: round(v / 1e6 if v > 1e5 else v / 1e3, 4) for k, v in acc.items()})
PY
done

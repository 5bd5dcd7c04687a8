#!/bin/bash
# round 2, call L: near/far split and RED cache hint with bbox trimming in place (cfg3)
mkdir -p gpurun_out
for lib in libgelcu.so libgelcu_z030.so libgelcu_z035.so libgelcu_z045.so libgelcu_z050.so; do
 for opt in "red_hint=0" "red_hint=1"; do
  GELCU_LIB=$lib timeout 600 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" --opt $opt > gpurun_out/zs.json 2> gpurun_out/zs.err; tail -1 gpurun_out/zs.err
  python - <<PY
import json
for l in open("gpurun_out/zs.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$lib $opt fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
 done
done

#!/bin/bash
# round 2, call W: split background fill (part 1 beside the transform kernel from a host-side guess of the region)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/gpu_fuzz.py 200 13000 | tail -1
for v in "split_fill=0" "split_fill=1" "split_fill=1 --opt store_hint=1"; do
 for w in cfg3 cfg4; do
  timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-extra --no-cpu --e2e "" --opt $v > gpurun_out/w.json 2> gpurun_out/w.err; tail -1 gpurun_out/w.err
  python - <<PY
import json
for l in open("gpurun_out/w.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("$w $v fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
 done
done

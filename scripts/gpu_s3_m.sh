#!/bin/bash
# session 3, call M: small first batch (taper) on / off, same box, alternating; cfg4 with 64 views per step
mkdir -p gpurun_out
for rep in 1 2 3; do
for t in 0 1; do
for w in cfg3 cfg1; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e region,region_rgb8,full --opt taper=$t > gpurun_out/m_$w.json 2> gpurun_out/m_$w.err; tail -1 gpurun_out/m_$w.err
  python - gpurun_out/m_$w.json $w $t <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print(sys.argv[2], "taper", sys.argv[3], "device fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d.get("e2e_variants", {}).items()})
PY
done
done
done
for v in 8 64; do
  timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extra --no-cpu --e2e region --views $v > gpurun_out/m_cfg4_$v.json 2> gpurun_out/m_cfg4_$v.err; tail -1 gpurun_out/m_cfg4_$v.err
  python - gpurun_out/m_cfg4_$v.json $v <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("cfg4 views", sys.argv[2], "device fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
done

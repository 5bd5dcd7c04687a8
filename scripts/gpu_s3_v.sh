#!/bin/bash
# session 3, call V: retune of the direct pipeline's compile-time knobs on the final kernels: resolve warp footprint (columns x rows of the 32 pixels),
# near / far depth split, triangles per warp
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
for v in "" _wc2 _wc8 _zs35 _zs45 _tpw512 _tpw2048; do
  lib=libgelcu$v.so
  echo "== $lib"
  for w in cfg3 cfg4; do
    GELCU_LIB=$lib timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" > gpurun_out/v_$w.json 2> gpurun_out/v_$w.err; tail -1 gpurun_out/v_$w.err
    show gpurun_out/v_$w.json
  done
done
done

#!/bin/bash
# session 3, call X: the last views of a batch in shorter runs (finer-grained end of the near / parked passes' grids): parity on the direct pipeline, then cfg3 / cfg4
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "direct or golden or cfg3 or cfg4" > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_x.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); print("  ", d["config"]["workload"][:4], "fps", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
}
for rep in 1 2; do
  for combo in "tail_views=0" "tail_views=4 tail_div=4" "tail_views=8 tail_div=4" "tail_views=16 tail_div=4" "tail_views=8 tail_div=2" "tail_views=4 tail_div=8" "tail_views=8 tail_div=8"; do
    echo "== [$combo]"
    opts=""; for o in $combo; do opts="$opts --opt $o"; done
    timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/x.json 2> gpurun_out/x.err; tail -1 gpurun_out/x.err
    show gpurun_out/x.json
  done
done
for combo in "tail_views=0" "tail_views=8 tail_div=4"; do
  opts=""; for o in $combo; do opts="$opts --opt $o"; done
  timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/x.json 2> gpurun_out/x.err; tail -1 gpurun_out/x.err
  echo "== [$combo]"; show gpurun_out/x.json
done

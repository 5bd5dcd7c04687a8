#!/bin/bash
# round 2, call A: new ABI entries + parametrised suite + new bench line + the fill-overlap experiment (cfg3)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for v in "fill_mode=0" "fill_mode=1" "fill_mode=2" "fill_mode=1 red_hint=1" "fill_mode=0 red_hint=1" "fill_mode=1 fill_ctas_per_sm=2" "fill_mode=1 fill_ctas_per_sm=2 red_hint=1" "fill_mode=1 fill_ctas_per_sm=4 red_hint=1"; do
  opts=""; for o in $v; do opts="$opts --opt $o"; done
  tag=$(echo $v | tr ' =' '__')
  timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu --e2e "" $opts > gpurun_out/fill_$tag.json 2> gpurun_out/fill_$tag.err; echo "== $v rc=$?"; tail -1 gpurun_out/fill_$tag.err
  python - <<PY
import json
for l in open("gpurun_out/fill_$tag.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print("  fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, d["clocks"])
PY
done
( time timeout 1500 python bench.py --steps 6 --warmup 3 ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_full.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err; cut -c1-600 gpurun_out/bench_reference.json

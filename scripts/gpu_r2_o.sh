#!/bin/bash
# round 2, call O: CUDA-graph path for small calls + single-launch batch init: whole GPU suite, latency probe, bench sanity
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/latency_probe.py cfg1 graph_small_calls=0 2>&1 | tail -10
timeout 900 python scripts/gpu_fuzz.py 150 11000 | tail -1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -2 gpurun_out/bench_o.err; python scripts/show_bench.py gpurun_out/bench_o.json | cut -c1-420
python - <<'PY'
import json
for l in open("gpurun_out/bench_o.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l); print(d["other_workloads"]["cfg1"]["single_view_latency"])
PY

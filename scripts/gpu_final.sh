#!/bin/bash
# end-of-round check: what the driver runs (smoke, gpu tests, both bench arms), on one box
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err; python scripts/show_bench.py gpurun_out/bench_default.json

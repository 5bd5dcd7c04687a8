#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; "$@" > gpurun_out/exp.json 2> gpurun_out/exp.err; tail -2 gpurun_out/exp.err; python scripts/show_bench.py gpurun_out/exp.json; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-extra"
run $B
run $B --opt debug_skip_clear=1
run $B --opt raster_ctas_per_sm=4
run $B --opt raster_ctas_per_sm=6
run $B --batch 8
run $B --batch 16
GELCU_LIB=libgelcu_u2.so run $B
GELCU_LIB=libgelcu_u2.so timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -2

#!/bin/bash
# iteration script: parity suite, short bench, ncu of raster+bin on cfg3 (8 views)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then bash scripts/gpu_prof2.sh cfg3 8; fi

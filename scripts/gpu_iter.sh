#!/bin/bash
# iteration script: parity suite (every test runs under the auto / tile / direct pipelines), short bench, optional ncu of the kernels on cfg3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then bash scripts/gpu_prof3.sh cfg3 8 direct_raster_kernel d1 8 2; fi

#!/bin/bash
# round 2, call H: whole GPU suite, extended differential fuzz (region + indexed entries), compute-sanitizer, load-time figures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python scripts/gpu_fuzz.py 400 7000 > gpurun_out/fuzz_r02.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/fuzz_r02.log
bash scripts/gpu_sanitize.sh 2>&1 | tail -12
timeout 600 python scripts/time_load.py cfg3 2>&1 | tail -2

#!/bin/bash
# session 3, call W: what the driver runs at round end (smoke, pytest -m gpu, both bench arms) on the committed code, then 4000 more fuzz seeds
mkdir -p gpurun_out
bash scripts/gpu_final.sh 2>&1 | tail -25
timeout 1500 python scripts/gpu_fuzz.py 4000 120000 > gpurun_out/r02_fuzz_s3b.txt 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/r02_fuzz_s3b.txt

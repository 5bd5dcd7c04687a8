#!/bin/bash
# session 3, call P: differential fuzz and compute-sanitizer on the final code (row trimming, TMA resets of 32x16 boxes, band rasteriser, 32-bit indices)
mkdir -p gpurun_out
timeout 1500 python scripts/gpu_fuzz.py 2500 90000 > gpurun_out/r02_fuzz_s3.txt 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r02_fuzz_s3.txt
bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/r02_compute_sanitizer_s3.txt

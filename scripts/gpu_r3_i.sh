#!/bin/bash
# call I: compute-sanitizer over both pipelines and every entry point on the current library (the tile pipeline's K2 now aliases
# its staging slabs with the grouping arrays in shared memory), then the differential fuzz
bash scripts/gpu_sanitize.sh
timeout 1500 python scripts/gpu_fuzz.py 600 91000 2>&1 | tail -3

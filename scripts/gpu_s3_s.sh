#!/bin/bash
# session 3, call S: racecheck of the race-free merge build (GEL_BAND_MERGE=1) and of the shipped build; parity of the shipped build
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s.log
bash scripts/gpu_sanitize.sh > gpurun_out/san_shipped.txt 2>&1; cat gpurun_out/san_shipped.txt | tail -12
GELCU_LIB=libgelcu_merge1.so bash scripts/gpu_sanitize.sh > gpurun_out/san_merge1.txt 2>&1; cat gpurun_out/san_merge1.txt | tail -12
GELCU_LIB=libgelcu_merge1.so timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "tile or auto" > gpurun_out/pytest_s_merge1.log 2>&1; echo "pytest merge1 rc=$?"; tail -2 gpurun_out/pytest_s_merge1.log

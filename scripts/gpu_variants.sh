#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for lib in libgelcu.so "$@"; do
  echo "== $lib"
  GELCU_LIB=$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err; tail -2 gpurun_out/bench_$lib.err
  python scripts/show_bench.py gpurun_out/bench_$lib.json
done
GELCU_LIB=libgelcu_t128_g32.so timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3

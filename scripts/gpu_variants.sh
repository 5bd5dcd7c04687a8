#!/bin/bash
# compare tuning variants: bench per library;  WORKLOAD=cfg5 bash scripts/gpu_variants.sh lib...
mkdir -p gpurun_out
W=${WORKLOAD:-cfg3}
for lib in "$@"; do
  echo "== $lib"
  GELCU_LIB=$lib timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err; tail -2 gpurun_out/bench_$lib.err
  python scripts/show_bench.py gpurun_out/bench_$lib.json
done

#!/bin/bash
# compare tuning variants: bench (cfg3 + cfg2/5) per library, then the parity suite on the last one
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  GELCU_LIB=$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err; tail -2 gpurun_out/bench_$lib.err
  python scripts/show_bench.py gpurun_out/bench_$lib.json
done

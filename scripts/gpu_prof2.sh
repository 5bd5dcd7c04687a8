#!/bin/bash
mkdir -p gpurun_out
W=${1:-cfg3}; V=${2:-8}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_$W -f \
   python bench.py --workload $W --steps 1 --warmup 3 --views $V --no-extra --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 3 -c 1 -o gpurun_out/bin_$W -f \
   python bench.py --workload $W --steps 1 --warmup 3 --views $V --no-extra --no-cpu > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"

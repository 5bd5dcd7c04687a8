#!/bin/bash
# full ncu capture of named kernels:  gpu_prof3.sh <workload> <views> <regex> <outname> [skip]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s ${5:-3} -c ${6:-1} -o gpurun_out/$4 -f \
   python bench.py --workload $1 --steps 1 --warmup 3 --views $2 --no-extra --no-cpu > gpurun_out/ncu_$4.log 2>&1; echo "$4 rc=$?"

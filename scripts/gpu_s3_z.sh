#!/bin/bash
# session 3, call Z: full capture of the near pass on cfg4 (the one named shape without a traffic record)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:direct_raster_kernel -s 6 -c 2 -o gpurun_out/r02_d1_cfg4 -f \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_r02_d1_cfg4.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_d1_cfg4.ncu-rep

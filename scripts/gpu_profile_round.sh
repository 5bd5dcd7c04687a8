#!/bin/bash
# Evidence for profiles/: bench line, ncu launch list of the same command, full captures of the kernels (cfg3).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv \
   python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_cfg3 -f \
   python bench.py --steps 1 --warmup 3 --views 8 --no-extra --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 3 -c 1 -o gpurun_out/bin_cfg3 -f \
   python bench.py --steps 1 --warmup 3 --views 8 --no-extra --no-cpu > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transform_kernel -s 3 -c 1 -o gpurun_out/transform_cfg3 -f \
   python bench.py --steps 1 --warmup 3 --views 8 --no-extra --no-cpu > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_cfg5 -f \
   python bench.py --workload cfg5 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu > gpurun_out/ncu_full4.log 2>&1; echo "rc=$?"

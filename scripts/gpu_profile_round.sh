#!/bin/bash
# Evidence for profiles/: bench line, ncu launch list of the same command, full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv \
   python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
bash scripts/gpu_prof3.sh cfg3 64 direct_raster_kernel d1 6 2
bash scripts/gpu_prof3.sh cfg3 64 direct_resolve_kernel d5 3 1
bash scripts/gpu_prof3.sh cfg3 64 direct_fill_kernel d5a 3 1
bash scripts/gpu_prof3.sh cfg3 64 direct_hiz_kernel d2 3 1
bash scripts/gpu_prof3.sh cfg3 8 transform_kernel transform_cfg3 3 1
bash scripts/gpu_prof3.sh cfg5 64 raster_kernel raster_cfg5 3 1

#!/bin/bash
# round 2 evidence for profiles/: the bench line, the ncu launch list of the same command, full ncu captures (with source) of every kernel of
# the cfg-3 step (direct pipeline) and of the tile pipeline's kernels on cfg5 / cfg2 / cfg1
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_bench_cfg3.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -3 gpurun_out/r02_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg3_bench_cmd.csv \
   python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --e2e region > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
cap() { # cap <workload> <views> <regex> <outname> <skip> <count>
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $5 -c $6 -o gpurun_out/$4 -f \
     python bench.py --workload $1 --steps 1 --warmup 3 --views $2 --no-extra --no-cpu --e2e "" > gpurun_out/ncu_$4.log 2>&1; echo "$4 rc=$?"; }
cap cfg3 64 direct_raster_kernel r02_d1 6 2
cap cfg3 64 direct_resolve_kernel r02_d5 3 1
cap cfg3 64 direct_fill_kernel r02_d5a 3 1
cap cfg3 64 direct_hiz_kernel r02_d2 3 1
cap cfg3 64 direct_clear_kernel r02_d0 3 1
cap cfg3 64 transform_kernel r02_transform_cfg3 3 1
cap cfg5 64 raster_band_kernel r02_raster_cfg5 3 1
cap cfg5 64 bin_kernel r02_bin_cfg5 3 1
cap cfg2 64 raster_band_kernel r02_raster_cfg2 3 1
cap cfg1 64 raster_band_kernel r02_raster_cfg1 3 1
ls -la gpurun_out/r02_*.ncu-rep

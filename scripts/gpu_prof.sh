#!/bin/bash
# parity suite + ncu launch list + full capture of the rasteriser on cfg3 and cfg2
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== ncu launch list (cfg3, 1 step of 64 views after 3 warm-ups)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv \
   python bench.py --steps 1 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full raster cfg3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_cfg3 -f \
   python bench.py --steps 1 --warmup 3 --views 8 --no-extra --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full bin kernels cfg3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_ -s 9 -c 3 -o gpurun_out/bin_cfg3 -f \
   python bench.py --steps 1 --warmup 3 --views 8 --no-extra --no-cpu > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
echo "== ncu full raster cfg2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_cfg2 -f \
   python bench.py --workload cfg2 --steps 1 --warmup 3 --views 64 --no-extra --no-cpu > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
ls -la gpurun_out

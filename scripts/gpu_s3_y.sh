#!/bin/bash
# session 3, call Y: last check of the committed code: GPU suite, then the bench line and the reference line as the driver runs them (traffic stamp current)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_y.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -3 gpurun_out/r02_bench_reference.err
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_bench_cfg3.err

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err; python scripts/show_bench.py gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --workload cfg1 --gpus $N --steps 3 --warmup 3 > gpurun_out/ref_n$N.json 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/ref_n$N.json
python - <<'PY'
import gzip, os, subprocess, json
g='tests/golden'
open('/tmp/s.obj','wb').write(gzip.open(g+'/sphere50.obj.gz').read()); open('/tmp/t.bmp','wb').write(gzip.open(g+'/tex256.bmp.gz').read())
PY
timeout 300 gel_b200/host/gel /tmp/s.obj /tmp/t.bmp --res 1920x1080 --sweep 512 --gpus $N --no-readback | tail -1
timeout 300 gel_b200/host/gel /tmp/s.obj /tmp/t.bmp --res 1920x1080 --sweep 512 --gpus 1 --no-readback | tail -1

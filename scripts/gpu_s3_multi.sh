#!/bin/bash
# session 3, multi-GPU call:  gpu_s3_multi.sh N   -- the bench under torchrun with N ranks (rank parity of the 8192-view cfg-5 list), and the reference arm the same way
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l)
        print("cfg3 fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k,v in d.get("e2e_variants",{}).items()}, "parity", d.get("parity_ranks"), d.get("parity_detail"))
        for k,v in d.get("other_workloads",{}).items(): print(" ", k, round(v["value"]), "fps frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]), v["scaling"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "reference N=$N rc=$?"; tail -2 gpurun_out/bench_ref_n$N.json | cut -c1-300

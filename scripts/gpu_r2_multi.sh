#!/bin/bash
# round 2, multi-GPU call:  gpu_r2_multi.sh N   -- host-ingest ceiling (pure D2H sweep), then the bench under torchrun with N ranks
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> gpurun_out/topo_n$N.txt; (numactl -H >> gpurun_out/topo_n$N.txt 2>&1 || true)
timeout 600 python scripts/d2h_sweep.py --mb 512 --reps 6 > gpurun_out/d2h_sweep_n$N.json 2> gpurun_out/d2h_sweep_n$N.err; echo "sweep rc=$?"; grep aggregate gpurun_out/d2h_sweep_n$N.json | head -20
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith('{"metric"'):
        d=json.loads(l)
        print("cfg3 fps", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k,v in d.get("e2e_variants",{}).items()}, "parity", d.get("parity_ranks"), d.get("parity_detail"))
        for k,v in d.get("other_workloads",{}).items(): print(" ", k, round(v["value"]), "fps frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]), v["scaling"])
PY

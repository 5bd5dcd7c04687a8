#!/bin/bash
mkdir -p gpurun_out
W=${1:-cfg3}; V=${2:-8}; shift; shift
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv \
   python bench.py --workload $W --steps 1 --warmup 3 --views $V --no-extra --no-cpu "$@" > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$W.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
data={}
for r in rows[1:]:
    data.setdefault(r[iid],{"k":r[ik][:60]})[r[im]]=r[iv]
ids=sorted(data,key=int)[-8:]
for i in ids:
    d=data[i]; print(i, d["k"], "us=%.1f"%(float(d.get("gpu__time_duration.sum","0").replace(",",""))/1e3), "Minst=%.1f"%(float(d.get("smsp__inst_executed.sum","0").replace(",",""))/1e6), "rdMB=", d.get("dram__bytes_read.sum"), "wrMB=", d.get("dram__bytes_write.sum"))
PY

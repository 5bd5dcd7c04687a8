#!/usr/bin/env python
"""Turns the captures of scripts/gpu_r2_evidence.sh (gpurun_out/r02_*) into the committed evidence under profiles/:
summaries (ncu_summary + ncu_lines) of every .ncu-rep, the bench lines, the launch list, and traffic.json stamped with the hash of
the kernel sources the captures were taken on.  Run here (no GPU):  python scripts/make_profiles.py"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
HEAD = "# ncu --set full --clock-control none --import-source on  (scripts/gpu_r2_evidence.sh), summarised by scripts/ncu_summary.py + scripts/ncu_lines.py\n"
CAPS = {"r02_d1": "r02_ncu_d1.txt", "r02_d5": "r02_ncu_d5.txt", "r02_d5a": "r02_ncu_d5a.txt", "r02_d2": "r02_ncu_d2.txt", "r02_d0": "r02_ncu_d0.txt", "r02_transform_cfg3": "r02_ncu_transform_cfg3.txt",
        "r02_raster_cfg5": "r02_ncu_raster_cfg5.txt", "r02_bin_cfg5": "r02_ncu_bin_cfg5.txt", "r02_raster_cfg2": "r02_ncu_raster_cfg2.txt", "r02_raster_cfg1": "r02_ncu_raster_cfg1.txt", "r02_d1_cfg4": "r02_ncu_d1_cfg4.txt"}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    def val(r, k):
        v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return [(r[hdr.index("Kernel Name")], val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")) for r in rows[2:]]


traffic = {}
for cap, txt in CAPS.items():
    rep = os.path.join(G, cap + ".ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep); continue
    a = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "4"], capture_output=True, text=True).stdout
    open(os.path.join(P, txt), "w").write(HEAD + a + "\n## hottest source lines (warp-instructions executed, share, avg active threads, share of stall samples)\n" + b)
    traffic[cap] = raw(rep)
    print(txt, [(k[:40], round((r + w) / 64 / 1e6, 2)) for k, r, w in traffic[cap]])

for src, dst in (("r02_bench_cfg3.json", "r02_bench_cfg3.json"), ("r02_bench_reference.json", "r02_bench_reference.json"), ("r02_launches_cfg3_bench_cmd.csv", "r02_launches_cfg3_bench_cmd.csv")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

def per_frame(cap, pick=0): r = traffic[cap][pick]; return int((r[1] + r[2]) / 64), f"{r[1] / 1e6:.1f} MB read + {r[2] / 1e6:.1f} MB written over 64 frames"
recs = []
for kernel, wl, cap, txt in (("direct_raster_kernel<0>", "cfg3", "r02_d1", "r02_ncu_d1.txt"), ("direct_raster_kernel<0>", "cfg4", "r02_d1_cfg4", "r02_ncu_d1_cfg4.txt"), ("raster_band_kernel", "cfg5", "r02_raster_cfg5", "r02_ncu_raster_cfg5.txt"),
                             ("raster_band_kernel", "cfg2", "r02_raster_cfg2", "r02_ncu_raster_cfg2.txt"), ("raster_band_kernel", "cfg1", "r02_raster_cfg1", "r02_ncu_raster_cfg1.txt")):
    if cap in traffic:
        b, how = per_frame(cap); recs.append({"kernel": kernel, "workload": wl, "bytes_per_frame": b, "from": f"{txt}: {how}"})
whole = sum((r + w) for cap in ("r02_transform_cfg3", "r02_d1", "r02_d5a", "r02_d2", "r02_d5") if cap in traffic for _, r, w in traffic[cap]) / 64
tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernels from `ncu --set full` (profiles/r02_ncu_*.txt), per frame; bench.py multiplies by the frames per launch and reports it only while kernel_sources_sha matches the kernel sources it runs (bench.kernel_sources_sha)",
      "kernel_sources_sha": bench.kernel_sources_sha(), "captured": "round 2, scripts/gpu_r2_evidence.sh, 64 views per launch", "records": recs,
      "whole_step_cfg3": {"bytes_per_frame": int(whole), "from": "sum over r02_ncu_{transform_cfg3,d1,d5a,d2,d5}.txt: transform + near pass + fill + hi-Z + parked pass + resolve; against B_alg = 167.8 MB per frame"}}
json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("traffic.json stamped", tj["kernel_sources_sha"], "whole step", round(whole / 1e6, 1), "MB per frame")

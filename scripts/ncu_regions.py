#!/usr/bin/env python
"""Region view of an .ncu-rep (needs -lineinfo + --import-source on): warp-instructions, lane use and stall samples per phase
of one kernel.  ncu_regions.py file.ncu-rep main_file.cuh line:name [line:name ...]
Every SASS instruction is attributed by ADDRESS ORDER: instructions whose line info points into `main_file` open the region
containing that line; instructions inlined from other files (gel_math.h, CUDA headers) -- and helper functions of main_file
defined above the first region -- belong to the region of the nearest preceding anchored instruction."""
import csv, subprocess, sys, collections
rep, mainf = sys.argv[1], sys.argv[2]
marks = sorted((int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[3:])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, line, sass = "", None, 0, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    if r[0] != "": line = int(r[0]); continue
    if not r[2].startswith("0x"): continue
    a = int(r[2], 16)
    ex = int(r[hdr.index("Instructions Executed")] or 0); tex = int(r[hdr.index("Thread Instructions Executed")] or 0); smp = int(r[hdr.index("# Samples")] or 0)
    bar = int(r[hdr.index("stall_barrier")] or 0)
    # an address can be listed under several files (inlining): keep the main file's attribution when present
    if a not in sass or fname == mainf: sass[a] = (fname, line, ex, tex, smp, bar, r[3].strip())
def region_of(ln):
    name = None
    for l, n in marks:
        if ln >= l: name = n
    return name
acc = collections.OrderedDict((n, [0, 0, 0, 0]) for _, n in marks); acc["(before)"] = [0, 0, 0, 0]
cur = "(before)"
for a in sorted(sass):
    f, ln, ex, tex, smp, bar, txt = sass[a]
    if f == mainf:
        r = region_of(ln)
        if r: cur = r
    acc[cur][0] += ex; acc[cur][1] += tex; acc[cur][2] += smp; acc[cur][3] += bar
tot = sum(v[0] for v in acc.values()); ts = sum(v[2] for v in acc.values())
print(f"total {tot/1e6:.1f} M warp-instr, {ts} samples")
for n, v in acc.items():
    if v[0] or v[2]: print(f"{n:28s} instr {v[0]/1e6:8.2f}M {100*v[0]/tot:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}  samples {100*v[2]/max(ts,1):5.1f}%  (barrier {100*v[3]/max(ts,1):4.1f}%)")

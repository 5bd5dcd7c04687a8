#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on):
   ncu_lines.py file.ncu-rep [min_exec_millions] -> file:line, warp-instr executed (M, %), avg active threads, stall samples, source"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = "", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "" and len(r) == len(hdr):
        try:
            ex = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        tex = int(r[hdr.index("Thread Instructions Executed")] or 0)
        smp = int(r[hdr.index("# Samples")] or 0)
        lines.append((fname, int(r[0]), ex, tex, smp, r[1].strip()))
tot = sum(l[2] for l in lines); tsmp = sum(l[4] for l in lines)
print(f"total warp-instr {tot/1e6:.1f} M, samples {tsmp}")
for f, ln, ex, tex, smp, src in lines:
    if ex / 1e6 >= thr or smp > tsmp * 0.01:
        print(f"{f[:16]:16s}:{ln:4d} {ex/1e6:8.2f}M {100*ex/tot:5.1f}% thr={tex/max(ex,1):5.1f} smp={100*smp/max(tsmp,1):5.1f}%  {src[:110]}")

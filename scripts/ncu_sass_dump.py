#!/usr/bin/env python
"""ncu_sass_dump.py file.ncu-rep [kernel-index] -> every SASS instruction of the capture in address order with its executed count, active
lanes, stall samples and the source line it belongs to (needs -lineinfo + --import-source on).  For reading a kernel phase by phase."""
import csv, subprocess, sys
rep = sys.argv[1]
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
    if a not in sass or not fname.startswith(("sm_", "device_", "math_", "gel_math")): sass[a] = (fname, line, ex, tex, smp, r[3].strip())
base = min(sass)
for a in sorted(sass):
    f, ln, ex, tex, smp, txt = sass[a]
    print(f"{(a - base) // 16:5d} {f[:16]:16s}:{ln:4d} ex={ex/1e6:8.3f}M thr={tex/max(ex,1):5.1f} smp={smp:5d}  {txt}")

#!/usr/bin/env python
"""Compact view of `ncu --page source --print-source sass --csv`: index, executed warp-instructions (M), avg active
threads, stall samples, SASS text.  usage: ncu_sass.py file.csv [min_exec_millions]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[1]
ia, isrc, isamp, iex, iavg = (hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
tot = sum(int(r[iex]) for r in rows[2:]); tsamp = sum(int(r[isamp]) for r in rows[2:])
print(f"total warp-instr {tot/1e6:.1f} M, samples {tsamp}")
for k, r in enumerate(rows[2:]):
    ex = int(r[iex])
    if ex / 1e6 >= thr:
        print(f"{k:5d} {ex/1e6:9.2f}M thr={float(r[iavg]):5.1f} smp={int(r[isamp]):6d}  {r[isrc].strip()[:90]}")

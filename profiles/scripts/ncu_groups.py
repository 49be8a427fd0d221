#!/usr/bin/env python
"""Instruction / stall-sample shares of an ncu report grouped by source-line ranges.
usage: ncu_groups.py report.ncu-rep file.cu name:first-last [name:first-last ...]"""
import csv, subprocess, io, collections, sys
rep, main = sys.argv[1], sys.argv[2]
groups = [(g.split(':')[0], int(g.split(':')[1].split('-')[0]), int(g.split(':')[1].split('-')[1])) for g in sys.argv[3:]]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hl = None
per = collections.defaultdict(lambda: [0, 0, 0])
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hl = r; continue
    if hl is None or len(r) < len(hl) or r[0] == '': continue
    ln = num(r[0])
    v = per[(cur, ln)]
    v[0] += num(r[hl.index("Instructions Executed")]); v[1] += num(r[hl.index("# Samples")]); v[2] += num(r[hl.index("Thread Instructions Executed")])
N = sum(v[0] for v in per.values()); S = sum(v[1] for v in per.values())
print(f"warp instructions {N}, samples {S}")
def show(name, sel):
    n = sum(v[0] for k, v in per.items() if sel(k)); s = sum(v[1] for k, v in per.items() if sel(k)); tn = sum(v[2] for k, v in per.items() if sel(k))
    print(f"{name:28} ins={100*n/N:5.1f}% smp={100*s/S:5.1f}% lanes={tn/max(n,1):4.1f}")
for name, a, b in groups:
    show(name, lambda k: k[0] == main and a <= k[1] <= b)
for ff in sorted(set(f for f, l in per)):
    if ff != main: show(ff, lambda k: k[0] == ff)

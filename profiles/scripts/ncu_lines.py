#!/usr/bin/env python
"""Per-source-line summary of an ncu report: share of stall samples and of warp instructions per CUDA line.
usage: ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; hdr_list = r; continue
    if hdr is None or len(r) < len(hdr_list) or r[0] == '': continue  # SASS rows repeat what the line row sums
    key = (cur, r[0])
    a = agg.setdefault(key, dict(src=r[1], s=0, n=0, tn=0, lsb=0, ssb=0, wait=0, br=0, notsel=0, math=0))
    def g(name):
        i = hdr_list.index(name); v = r[i]
        try: return int(v)
        except: return 0
    a["s"] += g("# Samples"); a["n"] += g("Instructions Executed"); a["tn"] += g("Thread Instructions Executed")
    a["lsb"] += g("stall_long_sb"); a["ssb"] += g("stall_short_sb"); a["wait"] += g("stall_wait"); a["br"] += g("stall_branch_resolving")
    a["math"] += g("stall_math")
S = sum(a["s"] for a in agg.values()); N = sum(a["n"] for a in agg.values())
print(f"samples {S} warp-instructions {N}")
for (f, ln), a in agg.items():
    if a["s"] > S * thr / 100 or a["n"] > N * thr / 100:
        lanes = a["tn"] / a["n"] if a["n"] else 0
        print(f"{f}:{ln:>4} smp={100*a['s']/S:5.1f}% ins={100*a['n']/N:5.1f}% lanes={lanes:4.1f} lsb={100*a['lsb']/S:4.1f} ssb={100*a['ssb']/S:4.1f} wait={100*a['wait']/S:4.1f} br={100*a['br']/S:4.1f} | {a['src'].strip()[:100]}")

#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries of ncu_r2.sh: DRAM bytes of one analyze_reads_kernel launch per workload
and resident form, keyed by the hash of the kernel sources they were captured from (bench.py reports a figure only for the
sources it belongs to).  usage: make_traffic.py <dir with prof_r2_<workload>_<form>.txt>"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

d = sys.argv[1]
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE analyze_reads_kernel launch (1 Mi fragments) from the ncu "
                   "--set full capture named in 'source'; bench.py reports it as roofline.traffic when kernel_source_hash matches"}
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for w in ("c2", "c3", "c4"):
    for form in ("text", "packed"):
        p = os.path.join(d, "prof_r2_%s_%s.txt" % (w, form))
        if not os.path.exists(p):
            continue
        tot, name, dur = 0.0, None, None
        for ln in open(p):
            f = ln.split()
            if ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum"):
                tot += float(f[2].replace(",", "")) * unit.get(f[1], 1)
            if ln.startswith("Kernel Name"):
                name = ln.split(None, 2)[-1].strip()[:120]
            if ln.startswith("gpu__time_duration.sum"):
                dur = " ".join(f[1:3])
        if tot:
            out["%s_%s" % (w, form)] = {"dram_bytes_per_launch": int(tot), "kernel_source_hash": bench.kernel_source_hash(),
                                        "kernel": name, "duration_under_ncu": dur,
                                        "source": "profiles/prof_r2_%s_%s.txt (ncu --set full, profiles/scripts/ncu_r2.sh)" % (w, form)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

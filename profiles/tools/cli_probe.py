#!/usr/bin/env python
"""Measurement helper (GPU box): the CLI on synthetic FASTQ files in /dev/shm with SHK_TIMING stamps, for a few
settings of the host pipeline's knobs.  usage: cli_probe.py [c2|c4] [reads] [variant ...]   variant = NAME:ENV=V,ENV=V"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from shark_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
variants = sys.argv[3:] or ["default:"]
wl = bench.WORKLOADS[name]
d = "/dev/shm/cli_probe"
os.makedirs(d, exist_ok=True)
names, bases, rec_off = synth.make_reference(wl["genes"], seed=1)
fa, f1, f2 = d + "/ref.fa", d + "/s_1.fq", d + "/s_2.fq"
synth.write_fasta(fa, names, bases, rec_off)
want_q = wl["q"] > 0
t0 = time.time()
done = 0
while done < n:
    m = min(4 << 20, n - done)
    seq, qual, _ = synth.make_reads(bases, wl["genes"], m, wl["L"], wl["paired"], seed=2, varied_qual=want_q, want_qual=want_q,
                                    first_read=done // synth.BLOCK * synth.BLOCK)
    synth.write_fastq_fast(f1, f2, seq, qual, m, wl["L"], wl["paired"], first_name=done, append=done > 0)
    done += m
print("wrote %d reads in %.1fs" % (n, time.time() - t0), flush=True)
flags = ["-k", str(wl["k"]), "-c", str(wl["c"]), "-b", str(wl["b"])] + (["-q", str(wl["q"])] if wl["q"] else []) + (["-s"] if wl["single"] else [])
for v in variants:
    label, _, envs = v.partition(":")
    env = dict(os.environ, SHK_TIMING="1")
    for kv in envs.split(","):
        if kv:
            k_, _, v_ = kv.partition("=")
            env[k_] = v_
    for rep in range(2):
        time.sleep(float(os.environ.get("PROBE_SLEEP", "1.5")))  # the previous run's device process may still be tearing down
        cmd = [bench.CLI_BIN, "-r", fa, "-1", f1, "-o", d + "/o1.fq"] + flags + (["-2", f2, "-p", d + "/o2.fq"] if wl["paired"] else [])
        t0 = time.perf_counter()
        with open(d + "/o.ssv", "wb") as fo:
            p = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, env=env)
        secs = time.perf_counter() - t0
        print("== %s run %d: %.3fs wall, %.2f M fragments/s, rc %d" % (label, rep, secs, n / secs / 1e6, p.returncode), flush=True)
        if rep == 1:
            for ln in p.stderr.decode().splitlines():
                if "timing" in ln:
                    print("   " + ln)

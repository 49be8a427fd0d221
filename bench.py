#!/usr/bin/env python
"""bench.py - reads/s of Shark's k-mer Bloom-filter hot path on B200 (Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

A step = one pass of the hot path over this rank's whole batch of synthetic reads (workload C2 by
default: SYN(1000 genes), 10 M single-end 100 bp reads, k=17, 1 GiB Bloom filter).
  value   reads/s with the reads already resident in HBM when the timed region starts
  e2e     reads/s through the public API (Shark.analyze_chunks -> shk_reads_submit/collect) from
          pinned HOST buffers of read TEXT, H2D and D2H inside the timed region.  --upload split (default):
          shk_reads_submit sends part of every chunk as text and packs the rest to 3 bits per base on the
          host cores while that copy runs (SHK_F_HOST_PACK; the packing is inside the timed region and the
          time is max(device stopwatch, host clock)); --upload plain: text only.  The mode not chosen is
          measured too (e2e_other).
  roofline  analyze_reads_kernel: 32 B x (k-mer windows probed) / its CUDA-event time, against the
          measured HBM copy rate in MEASURED_PEAKS.json (and the measured random-sector ceiling)
  cpu_baseline  the unmodified reference (oracle/_ref/shark -t <cores>) on a bounded prefix of the
          same reads, on this box's host cores
N > 1: one process per GPU (torchrun); reads are sharded (weak scaling: every rank gets its own
batch of the same size), the index is built on rank 0 and replicated with an NCCL broadcast
(--index sharded|both: every rank indexes one gene shard, filters OR-merged by the library's P2P kernel).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: genes, reads, L, paired, k, b, q, single, c
    "c2": dict(genes=1000, reads=10_000_000, L=100, paired=False, k=17, b=1, q=0, single=False, c=0.6,
               desc="C2: SYN(1000 genes x 3 kbp), 10M single-end 100 bp reads, k=17, c=0.6, 1 GiB Bloom filter"),
    "c3": dict(genes=5000, reads=8_000_000, L=150, paired=True, k=21, b=1, q=20, single=True, c=0.6,
               desc="C3 (8M-pair slice of 50M): SYN(5000 genes), paired 150 bp, k=21, -q 20, -s, 1 GiB Bloom filter"),
    "c4": dict(genes=20000, reads=8_000_000, L=150, paired=True, k=31, b=4, q=0, single=False, c=0.6,
               desc="C4 (8M-pair slice of 100M): SYN(20000 genes ~60 Mbp), paired 150 bp, k=31, 4 GiB Bloom filter"),
}
CHUNK_READS = 1 << 20


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def probe_kmers(bases, k, n, seed=7):
    """n canonical k-mers for the stand-alone probe kernel (BF::get_index): half are windows of the
    reference (hits), half are random (misses, up to the filter's false-positive rate)."""
    rng = np.random.default_rng(seed)
    code = np.full(256, 4, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = code[ch | 0x20] = i
    c = code[np.asarray(bases[: min(len(bases), 4_000_000)])].astype(np.uint64)
    m = len(c) - k + 1
    fwd = np.zeros(m, np.uint64)
    rc = np.zeros(m, np.uint64)
    bad = np.zeros(m, bool)
    for j in range(k):
        cj = c[j:j + m]
        bad |= cj > 3
        fwd = (fwd << np.uint64(2)) | (cj & np.uint64(3))
        rc |= (np.uint64(3) - (cj & np.uint64(3))) << np.uint64(2 * j)
    canon = np.minimum(fwd, rc)[~bad]
    hits = canon[rng.integers(0, len(canon), n // 2)]
    miss = rng.integers(0, 1 << 62, n - n // 2, dtype=np.uint64) & np.uint64((1 << (2 * k)) - 1)
    out = np.concatenate([hits, miss])
    rng.shuffle(out)
    return out


def make_workload(wl, rank, n_reads):
    """-> names, bases, rec_off, pinned chunk list [(seq, qual, off32, n)], keepalive buffers."""
    from shark_b200 import capi, synth
    names, bases, rec_off = synth.make_reference(wl["genes"], seed=1)
    W = 2 * wl["L"] + 1 if wl["paired"] else wl["L"]
    want_q = wl["q"] > 0
    blocks_per_rank = (wl["reads"] + synth.BLOCK - 1) // synth.BLOCK
    first = rank * blocks_per_rank * synth.BLOCK
    pin_seq = capi.PinnedBuffer(n_reads * W)
    pin_qual = capi.PinnedBuffer(n_reads * W) if want_q else None
    t0 = time.time()
    synth.make_reads(bases, wl["genes"], n_reads, wl["L"], wl["paired"], seed=2, varied_qual=want_q, want_qual=want_q,
                     out_seq=pin_seq.u8, out_qual=pin_qual.u8 if want_q else None, first_read=first)
    log("[bench] rank %d generated %d reads in %.1fs" % (rank, n_reads, time.time() - t0))
    pin_off = capi.PinnedBuffer((CHUNK_READS + 1) * 4)
    off32 = pin_off.view(np.uint32, CHUNK_READS + 1)
    off32[:] = np.arange(CHUNK_READS + 1, dtype=np.uint32) * np.uint32(W)
    chunks = []
    for a in range(0, n_reads, CHUNK_READS):
        n = min(CHUNK_READS, n_reads - a)
        chunks.append((pin_seq.u8[a * W:(a + n) * W], pin_qual.u8[a * W:(a + n) * W] if want_q else None, off32[:n + 1], n))
    return names, bases, rec_off, chunks, (pin_seq, pin_qual, pin_off), W


def reference_arm_run(wl, sample_reads, workdir, threads):
    """Runs oracle/_ref/shark -t threads on the first `sample_reads` reads; returns
    (sample-stage seconds, total seconds, ssv path).  The sample stage is the wall-clock
    difference to a run over one read (index build is common to both)."""
    from shark_b200 import synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "shark")
    if not os.path.exists(ref_bin):
        return None
    fa = os.path.join(workdir, "ref.fa")
    f1, f2 = os.path.join(workdir, "s_1.fq"), os.path.join(workdir, "s_2.fq")
    t1, t2 = os.path.join(workdir, "t_1.fq"), os.path.join(workdir, "t_2.fq")
    if not os.path.exists(fa):
        names, bases, rec_off = synth.make_reference(wl["genes"], seed=1)
        synth.write_fasta(fa, names, bases, rec_off)
        want_q = wl["q"] > 0
        seq, qual, _ = synth.make_reads(bases, wl["genes"], sample_reads, wl["L"], wl["paired"], seed=2,
                                        varied_qual=want_q, want_qual=want_q)
        synth.write_fastq(f1, f2, seq, qual, sample_reads, wl["L"], wl["paired"])
        synth.write_fastq(t1, t2, seq, qual, 1, wl["L"], wl["paired"])
    flags = ["-k", str(wl["k"]), "-c", str(wl["c"]), "-b", str(wl["b"]), "-t", str(threads)]
    if wl["q"]:
        flags += ["-q", str(wl["q"])]
    if wl["single"]:
        flags += ["-s"]

    def run(a, b, out):
        cmd = [ref_bin, "-r", fa, "-1", a, "-o", os.path.join(workdir, "o1.fq")] + flags
        if wl["paired"]:
            cmd += ["-2", b, "-p", os.path.join(workdir, "o2.fq")]
        t0 = time.perf_counter()
        with open(out, "wb") as fo:
            subprocess.run(cmd, stdout=fo, stderr=subprocess.DEVNULL, check=True)
        return time.perf_counter() - t0

    t_tiny = run(t1, t2, os.path.join(workdir, "tiny.ssv"))
    t_full = run(f1, f2, os.path.join(workdir, "full.ssv"))
    return max(t_full - t_tiny, 1e-6), t_full, os.path.join(workdir, "full.ssv")


def capi_pack_info():
    from shark_b200 import capi
    return capi.host_pack_info()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (testing)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="reads given to the CPU reference (0 = 6M for the cpu_baseline leg, about 10 s on 16 cores; "
                         "2M per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extend", default="auto", choices=["auto", "on", "off"],
                    help="anchor-and-extend: automatic (on for DRAM-sized front tables), or forced (A/B runs)")
    ap.add_argument("--upload", default="split",
                    help="how e2e moves the reads: 'plain' (text over PCIe), 'split' (part of every chunk packed to 3 bits "
                         "per base by the host cores while the rest is in flight, auto-balanced) or a fixed packed share "
                         "in (0, 1]; the other mode is measured too and reported as e2e_other")
    ap.add_argument("--e2e-slots", type=int, default=3,
                    help="chunks in flight in the e2e leg (slots reused round-robin by Shark.analyze_chunks)")
    ap.add_argument("--index", default="broadcast", choices=["broadcast", "sharded", "both"],
                    help="N > 1: build on rank 0 + NCCL broadcast (default), or every rank indexes one gene shard and the "
                         "filters are OR-merged by the library's P2P kernel; 'both' times both and checks they are identical")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    n_reads = args.reads or wl["reads"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if not args.cpu_sample:
        # the reference hands out batches of 50 000 reads to its threads (main.cpp:215): both samples keep every
        # thread busy; the cpu_baseline leg is one run of ~10 s, the reference arm repeats a shorter one K+W times
        args.cpu_sample = 2_000_000 if args.impl == "reference" else 6_000_000
    config = {"workload": wl["desc"], "reads_per_gpu": n_reads, "read_len": wl["L"], "paired": wl["paired"], "k": wl["k"],
              "bf_gib": wl["b"], "min_quality": wl["q"], "single": wl["single"], "chunk_reads": CHUNK_READS,
              "sharding": "reads sharded by rank, index replicated (NCCL broadcast)" if world > 1 else "single GPU",
              "upload": ("e2e: plain text over PCIe" if args.upload == "plain" else
                         "e2e: split upload through shk_reads_submit (SHK_F_HOST_PACK): text in pinned host memory, part of "
                         "every chunk packed to 3 bits/base by the host cores inside the timed region, share = %s"
                         % ("auto-balanced" if args.upload == "split" else args.upload)),
              "l2": "inputs larger than L2: every step streams %d MB of reads and probes a %.1f GB filter at random"
                    % (n_reads * (2 * wl["L"] + 1 if wl["paired"] else wl["L"]) // 1_000_000, wl["b"] * 8 / 7 * 1.0737)}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = min(args.cpu_sample, n_reads)
        with tempfile.TemporaryDirectory() as tmp:
            times = []
            for i in range(args.warmup + args.steps):
                r = reference_arm_run(wl, sample, tmp, cores)
                if r is None:
                    print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/shark is not built"}))
                    return 0
                if i >= args.warmup:
                    times.append(r[0])
                log("[bench/reference] run %d: sample stage %.2fs (whole run %.2fs)" % (i, r[0], r[1]))
            t = float(np.mean(times))
        v = sample / t
        sample_desc = "first %d reads of the workload (oracle/_ref/shark -t %d; sample stage = wall(full) - wall(1 read))" % (sample, cores)
        print(json.dumps({
            "impl": "reference", "metric": "reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample_desc},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ---------------------------------------------------------------- our arm
    # libraries (NCCL's version banner, ...) may write to fd 1: keep the real stdout for the ONE
    # JSON line and send everything else to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from shark_b200.engine import Shark
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the split upload packs with host threads: share the box's cores between the ranks of this node
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("SHK_PACK_THREADS", str(max(2, min(32, cores // max(local_world, 1)))))
    names, bases, rec_off, chunks, keep_alive, W = make_workload(wl, rank, n_reads)
    n_chunks = len(chunks)
    sh = Shark(k=wl["k"], c=wl["c"], bf_bits=wl["b"] << 33, min_quality=wl["q"], single=wl["single"], device=local_rank,
               n_slots=max(n_chunks, 2), max_reads_per_chunk=CHUNK_READS, max_bytes_per_chunk=CHUNK_READS * W,
               extend={"auto": None, "on": True, "off": False}[args.extend])
    # index: build on rank 0, replicate over NVLink
    bcast_ms = 0.0
    index_extra = {"mode": "single GPU" if world == 1 else args.index}
    if world == 1 or args.index in ("broadcast", "both"):
        if rank == 0:
            info = sh.build_index(bases, rec_off)
            log("[bench] index: %d genes, %d set bits, %d ids, %.2f ms on device (%.1f ms wall)" %
                (info.n_genes, info.n_set_bits, info.tot_ids, info.build_ms, info.build_wall_ms))
        if world > 1:
            from shark_b200 import dist_index
            bcast_ms = dist_index.broadcast_index(sh, src=0)
    if world > 1 and args.index in ("sharded", "both"):
        from shark_b200 import dist_index
        before = sh.export_index() if args.index == "both" else None
        b_info = sh.info
        barrier()
        s_info, s_secs = dist_index.build_index_sharded(sh, bases, rec_off)
        index_extra.update({"sharded_wall_ms": s_secs * 1e3, "sharded_device_ms": s_info.build_ms, "n_shards": s_info.n_shards,
                            "sharded_steps_wall_ms": getattr(dist_index.build_index_sharded, "last_steps_ms", None)})
        if before is not None:
            after = sh.export_index()
            same = all(np.array_equal(a, b) for a, b in zip(before, after)) and \
                (b_info.n_genes, b_info.n_set_bits, b_info.tot_ids) == (s_info.n_genes, s_info.n_set_bits, s_info.tot_ids)
            index_extra["sharded_equals_broadcast"] = bool(same)
            if not same:
                raise SystemExit("sharded index differs from the broadcast index on rank %d" % rank)
        log("[bench] rank %d sharded index build: %.1f ms wall, %.2f ms on device" % (rank, s_secs * 1e3, s_info.build_ms))
    info = sh.info
    launches0 = sh.kernel_launches()

    # random-sector ceiling (kernel B0) for the secondary roofline denominator
    rs_ms = sh.random_sector_bench(1 << 28)
    rs_gbs = (1 << 28) * 32 / rs_ms / 1e6

    # stand-alone probe kernel (K5 = BF::get_index: filter word -> sector rank -> entry) on 2^25 canonical k-mers,
    # half hits / half misses: the "BF probe GB/s vs the random-sector roofline" half of the metric
    pk = probe_kmers(bases, wl["k"], 1 << 25)
    pb_ms, pb_hits = sh.probe_bench(pk, reps=4)
    pb_gbs = len(pk) * 32 / pb_ms / 1e6
    probe = {"kernel": "probe_bench_kernel (hash -> filter word -> sector + rank -> entry)", "probes": len(pk),
             "ms": pb_ms, "gprobes_per_s": len(pk) / pb_ms / 1e6, "achieved_gbs_32B_per_probe": pb_gbs,
             "random_sector_ceiling_gbs": rs_gbs, "frac_of_random_sector_ceiling": pb_gbs / rs_gbs if rs_gbs else None,
             "hit_fraction": pb_hits / len(pk)}
    del pk

    # ---- device-resident: upload once, then time K passes of the kernels + result read-back
    for i, (s, q, o, n) in enumerate(chunks):
        sh.upload(i, s, q, o, n)

    def resident_step():
        for i in range(n_chunks):
            sh.analyze_resident(i)
        out = [sh.collect(i, copy=False) for i in range(n_chunks)]
        return out

    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    sh.timer_start()  # device stopwatch (CUDA events over all slot streams); the host clock is the cross-check
    n_probes = n_hits = n_assoc = n_slow = n_ext = n_loads = 0
    for _ in range(args.steps):
        for r in resident_step():
            n_probes += r["n_probes"]
            n_hits += r["n_hits"]
            n_assoc += r["n_assoc"]
            n_slow += r["n_slow_reads"]
            n_ext += r["n_extended"]
            n_loads += r["n_table_loads"]
    t_res = sh.timer_stop() * 1e-3
    barrier()
    t_res_wall = time.perf_counter() - t0
    launches_res = sh.kernel_launches() - launches0

    # ---- roofline pass: the same launches one at a time (no overlap between the slots' streams),
    #      so that the CUDA-event time of analyze_reads_kernel is the time of that kernel alone
    probe_ms = 0.0
    for _ in range(args.steps):
        for i in range(n_chunks):
            sh.analyze_resident(i)
            probe_ms += sh.collect(i, copy=False)["probe_kernel_ms"]

    # ---- end to end through the public API: pinned host chunks -> H2D -> kernels -> D2H
    sh2 = sh  # same context; slots 0/1 are reused round-robin by analyze_chunks
    sh2.n_slots = max(2, min(args.e2e_slots, n_chunks))
    d2h = [0]

    def on_result(r):
        d2h[0] += r["n_assoc"] * 8 + r["n_reads"] + 48

    def upload_mode(name):
        return False if name == "plain" else (True if name == "split" else float(name))

    leg_stats = {}

    def e2e_leg(mode):
        sh2.set_upload_mode(upload_mode(mode))
        d2h[0] = 0
        for _ in range(args.warmup):
            sh2.analyze_chunks(chunks, copy=False, on_result=lambda r: None)
        barrier()
        h0, dd0 = sh2.h2d_bytes(), sh2.d2h_bytes()
        t0 = time.perf_counter()
        sh2.timer_start()
        for _ in range(args.steps):
            sh2.analyze_chunks(chunks, copy=False, on_result=on_result)
        t_dev = sh2.timer_stop() * 1e-3
        # the split upload packs on the host BEFORE a chunk's first device operation: the device stopwatch would
        # miss the packing of the first chunk of a step, the host clock around the same region does not
        barrier()
        t_wall = time.perf_counter() - t0
        leg_stats[mode] = sh2.upload_stats()
        return t_dev, t_wall, (sh2.h2d_bytes() - h0) // max(args.steps, 1), (sh2.d2h_bytes() - dd0) // max(args.steps, 1)

    other_mode = "split" if args.upload == "plain" else "plain"
    o_dev, o_wall, o_h2d, _ = e2e_leg(other_mode)
    t_e2e, t_e2e_wall, h2d_step, d2h_step = e2e_leg(args.upload)
    if args.upload != "plain":
        t_e2e = max(t_e2e, t_e2e_wall)
    if other_mode != "plain":
        o_dev = max(o_dev, o_wall)
    clocks = sampler.stop()

    # max over ranks (device times; the wall-clock figures ride along as a cross-check)
    times = torch.tensor([t_res, t_e2e, probe_ms, t_res_wall, t_e2e_wall, o_dev], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(n_probes), float(n_hits), float(n_assoc)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t_res_max, t_e2e_max, probe_ms_max, t_res_wall_max, t_e2e_wall_max, o_dev_max = times.tolist()
    total_reads = n_reads * world * args.steps
    value = total_reads / t_res_max
    e2e_value = total_reads / t_e2e_max

    peak, peak_src = measured_peaks()
    launches_per_step = n_chunks  # analyze_reads_kernel launches per step on this rank
    probes_per_launch = n_probes / max(args.steps * n_chunks, 1)
    probe_ms_per_launch = probe_ms / max(args.steps * n_chunks, 1)
    achieved = 32.0 * probes_per_launch / (probe_ms_per_launch * 1e-3) / 1e9 if probe_ms_per_launch > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "analyze_reads_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 32.0 * probes_per_launch, "probes_per_launch": probes_per_launch,
                "kernel_ms_per_launch": probe_ms_per_launch, "launches_per_step": launches_per_step,
                "hit_fraction": n_hits / max(n_probes, 1),
                "extend": bool(info.extend), "extended_fraction": n_ext / max(n_probes, 1),
                "table_loads_per_probe": (n_loads / max(n_probes, 1)) if info.extend else 1.0,
                "random_sector_ceiling_gbs": rs_gbs, "frac_of_random_sector_ceiling": achieved / rs_gbs if rs_gbs else None}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample, n_reads)
        with tempfile.TemporaryDirectory() as tmp:
            r = reference_arm_run(wl, sample, tmp, cores)
            if r is not None:
                cpu_baseline = {"value": sample / r[0], "unit": "reads/s", "cores": cores, "kind": "reference",
                                "sample": "first %d reads of the workload, oracle/_ref/shark -t %d, sample stage %.2fs "
                                          "(whole run %.2fs)" % (sample, cores, r[0], r[1])}
                # parity on the same prefix: identical read->gene pairs after sorting (north star)
                pref, left = [], sample
                for c0 in chunks:
                    if left <= 0:
                        break
                    m = min(left, c0[3])
                    pref.append((c0[0][: m * W], None if c0[1] is None else c0[1][: m * W], c0[2][: m + 1], m))
                    left -= m
                ours, base = [b""], 0
                for res, c0 in zip(sh.analyze_chunks(pref), pref):
                    ours += [b"r%09d %s" % (base + int(a), names[int(g)]) for a, g in zip(res["read_idx"], res["gene_idx"])]
                    base += c0[3]
                ref_lines = sorted(open(r[2], "rb").read().split(b"\n"))
                cpu_baseline["parity_on_sample"] = bool(sorted(ours) == ref_lines)
                cpu_baseline["ssv_lines"] = len(ref_lines) - 1
            else:
                cpu_baseline = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference",
                                "sample": "oracle/_ref/shark is not built"}

    if rank == 0:
        out = {
            "metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_res_max / args.steps * 1e3,
            "timing": "CUDA events over all slot streams (shk_device_timer_*), max over ranks",
            "wall_ms_per_step": t_res_wall_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "upload": args.upload, "slots": sh2.n_slots,
                    "packed_share": leg_stats[args.upload][0], "pack_gbases_per_s": leg_stats[args.upload][1],
                    "ms_per_step": t_e2e_max / args.steps * 1e3, "wall_ms_per_step": t_e2e_wall_max / args.steps * 1e3},
            "e2e_other": {"upload": other_mode, "value": total_reads / o_dev_max, "unit": "reads/s",
                          "h2d_bytes_per_step": int(o_h2d), "ms_per_step": o_dev_max / args.steps * 1e3,
                          "pack": "%s, %d threads" % capi_pack_info(),
                          "packed_share": leg_stats[other_mode][0], "pack_gbases_per_s": leg_stats[other_mode][1]},
            "gpu_launches": int(launches_res), "roofline": roofline, "probe": probe, "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "index": {"n_genes": info.n_genes, "n_set_bits": info.n_set_bits, "tot_ids": info.tot_ids,
                      "build_ms": info.build_ms, "build_wall_ms": info.build_wall_ms, "broadcast_ms": bcast_ms,
                      "device_bytes": info.device_bytes, **index_extra},
            "associations_per_step": sums.tolist()[2] / args.steps / world, "slow_reads_per_step": n_slow / args.steps,
        }
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    sh.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

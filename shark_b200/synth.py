"""Deterministic synthetic inputs of the benchmark configurations.

Reference  SYN(G, seed): G genes `gene%05d`, 3000 bp i.i.d. uniform ACGT; every 10th gene copies
           bases 500..2000 of the previous gene (multi-gene lists, ties); every 50th gene has one
           'N' at offset 1500.
Reads      block b (1 Mi reads) is a pure function of (seed, b): 90 % drawn from a uniformly
           chosen gene (uniform start, random strand, 1 % substitutions), 10 % i.i.d. background,
           every base 'N' with p = 0.001.  Paired: mate 2 is the reverse complement of a window
           200..400 bp downstream in the same gene (clamped).  A prefix of the reads of any run is
           therefore the same reads, which is what the CPU reference arm is fed.
Layout     the joined SoA layout of the C ABI: text = mate1 [+ 'N' + mate2], offsets.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b
BLOCK = 1 << 20


def make_reference(n_genes, seed=1, gene_len=3000):
    """-> (names list[bytes], bases uint8[n_genes*gene_len], rec_off uint64[n_genes+1])"""
    rng = np.random.default_rng([seed, 0xA11CE])
    g = ACGT[rng.integers(0, 4, (n_genes, gene_len), dtype=np.uint8)]
    for i in range(10, n_genes, 10):
        g[i, 500:2000] = g[i - 1, 500:2000]
    for i in range(0, n_genes, 50):
        g[i, 1500] = ord("N")
    names = [b"gene%05d" % i for i in range(n_genes)]
    off = (np.arange(n_genes + 1, dtype=np.uint64) * np.uint64(gene_len))
    return names, g.reshape(-1), off


def _mate(rng, flat, gene_len, gidx, start, L, bg_mask, strand):
    n = len(gidx)
    idx = (gidx.astype(np.int64) * gene_len + start)[:, None] + np.arange(L, dtype=np.int64)[None, :]
    m = flat[idx]
    # N in the reference stays N in the read
    if bg_mask.any():
        m[bg_mask] = ACGT[rng.integers(0, 4, (int(bg_mask.sum()), L), dtype=np.uint8)]
    rc = strand & ~bg_mask
    if rc.any():
        m[rc] = _COMP[m[rc][:, ::-1]]
    total = n * L
    k_sub = rng.binomial(total, 0.01)
    pos = rng.integers(0, total, k_sub)
    m.reshape(-1)[pos] = ACGT[rng.integers(0, 4, k_sub, dtype=np.uint8)]
    k_n = rng.binomial(total, 0.001)
    m.reshape(-1)[rng.integers(0, total, k_n)] = ord("N")
    return m


def _quals(rng, n, L, varied):
    if not varied:
        return np.full((n, L), ord("I"), np.uint8)
    q = rng.integers(35, 41, (n, L), dtype=np.uint8)
    t = L - L // 5
    q[:, t:] = rng.integers(2, 41, (n, L - t), dtype=np.uint8)
    return q + np.uint8(33)


def make_read_block(ref_flat, n_genes, gene_len, block_idx, n, L, paired, seed=2, varied_qual=False, want_qual=False):
    """One block of reads -> (text uint8[n, W], qual uint8[n, W] | None) with W = L or 2L+1."""
    rng = np.random.default_rng([seed, block_idx])
    gidx = rng.integers(0, n_genes, n)
    bg = rng.random(n) < 0.1
    strand = rng.random(n) < 0.5
    start = rng.integers(0, gene_len - L + 1, n)
    m1 = _mate(rng, ref_flat, gene_len, gidx, start, L, bg, strand)
    if not paired:
        text = m1
        qual = _quals(rng, n, L, varied_qual) if want_qual else None
        return text, qual
    gap = rng.integers(200, 401, n)
    start2 = np.minimum(start + gap, gene_len - L)
    m2 = _mate(rng, ref_flat, gene_len, gidx, start2, L, bg, ~strand)
    text = np.empty((n, 2 * L + 1), np.uint8)
    text[:, :L] = m1
    text[:, L] = ord("N")  # FastqSplitter.hpp:63,83 joiner
    text[:, L + 1:] = m2
    qual = None
    if want_qual:
        qual = np.empty((n, 2 * L + 1), np.uint8)
        qual[:, :L] = _quals(rng, n, L, varied_qual)
        qual[:, L] = 0x1B  # FastqSplitter.hpp:84
        qual[:, L + 1:] = _quals(rng, n, L, varied_qual)
    return text, qual


def make_reads(ref_flat, n_genes, n_reads, L, paired, seed=2, gene_len=3000, varied_qual=False, want_qual=False,
               out_seq=None, out_qual=None, first_read=0):
    """Reads [first_read, first_read + n_reads) in the joined SoA layout.
    -> (seq uint8[n_reads*W], qual | None, off uint64[n_reads+1]).  first_read must be a multiple
    of BLOCK.  out_seq/out_qual let the caller fill pinned buffers in place."""
    assert first_read % BLOCK == 0
    W = 2 * L + 1 if paired else L
    seq = out_seq if out_seq is not None else np.empty(n_reads * W, np.uint8)
    qual = (out_qual if out_qual is not None else np.empty(n_reads * W, np.uint8)) if want_qual else None
    done = 0
    b = first_read // BLOCK
    while done < n_reads:
        # always generate the whole block so that a prefix is independent of n_reads
        t, q = make_read_block(ref_flat, n_genes, gene_len, b, BLOCK, L, paired, seed, varied_qual, want_qual)
        m = min(BLOCK, n_reads - done)
        seq[done * W:(done + m) * W] = t[:m].reshape(-1)
        if want_qual:
            qual[done * W:(done + m) * W] = q[:m].reshape(-1)
        done += m
        b += 1
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(W)
    return seq[: n_reads * W], (qual[: n_reads * W] if want_qual else None), off


def write_fasta(path, names, bases, rec_off, width=80):
    with open(path, "wb") as f:
        for i, nm in enumerate(names):
            s = bases[int(rec_off[i]):int(rec_off[i + 1])].tobytes()
            f.write(b">" + nm + b"\n")
            f.write(b"\n".join(s[j:j + width] for j in range(0, len(s), width)) + b"\n")


def write_fastq(path1, path2, seq, qual, n_reads, L, paired, first_name=0):
    """Dumps reads of the joined layout as FASTQ files (names r%09d, both mates share names)."""
    W = 2 * L + 1 if paired else L
    t = np.asarray(seq[: n_reads * W]).reshape(n_reads, W)
    q = np.asarray(qual[: n_reads * W]).reshape(n_reads, W) if qual is not None else None
    const_q = b"I" * L
    f1 = open(path1, "wb")
    f2 = open(path2, "wb") if paired else None
    step = 1 << 16
    for a in range(0, n_reads, step):
        b = min(a + step, n_reads)
        o1, o2 = [], []
        for i in range(a, b):
            nm = b"@r%09d\n" % (first_name + i)
            row = t[i].tobytes()
            o1.append(nm + row[:L] + b"\n+\n" + (q[i, :L].tobytes() if q is not None else const_q) + b"\n")
            if paired:
                o2.append(nm + row[L + 1:] + b"\n+\n" + (q[i, L + 1:].tobytes() if q is not None else const_q) + b"\n")
        f1.write(b"".join(o1))
        if paired:
            f2.write(b"".join(o2))
    f1.close()
    if f2:
        f2.close()


def fastq_records(mate, qual, first_name):
    """Fixed-length reads (uint8[n, L]) -> uint8[n, R] of FASTQ records `@r%09d\nSEQ\n+\nQUAL\n`, built with
    array operations (the per-read loop of write_fastq manages ~0.5 M reads/s; this writes tens of millions)."""
    n, L = mate.shape
    R = 12 + L + 3 + L + 1
    rec = np.empty((n, R), np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    idx = np.arange(first_name, first_name + n, dtype=np.int64)
    for d in range(9):
        rec[:, 10 - d] = (idx % 10 + ord("0")).astype(np.uint8)
        idx //= 10
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + L] = mate
    rec[:, 12 + L:15 + L] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 15 + L:15 + 2 * L] = qual if qual is not None else ord("I")
    rec[:, 15 + 2 * L] = ord("\n")
    return rec


def write_fastq_fast(path1, path2, seq, qual, n_reads, L, paired, first_name=0, append=False, piece=1 << 20):
    """Same files as write_fastq, vectorised; `append` adds to existing files."""
    W = 2 * L + 1 if paired else L
    t = np.asarray(seq[: n_reads * W]).reshape(n_reads, W)
    q = np.asarray(qual[: n_reads * W]).reshape(n_reads, W) if qual is not None else None
    mode = "ab" if append else "wb"
    with open(path1, mode) as f1:
        f2 = open(path2, mode) if paired else None
        for a in range(0, n_reads, piece):
            b = min(a + piece, n_reads)
            fastq_records(t[a:b, :L], None if q is None else q[a:b, :L], first_name + a).tofile(f1)
            if paired:
                fastq_records(t[a:b, L + 1:], None if q is None else q[a:b, L + 1:], first_name + a).tofile(f2)
        if f2:
            f2.close()


def fastq_bytes(seq, qual, n_reads, L, paired, first_name=0):
    """-> (bytes of mate-1 FASTQ, bytes of mate-2 FASTQ | None) for n_reads reads, in memory."""
    W = 2 * L + 1 if paired else L
    t = np.asarray(seq[: n_reads * W]).reshape(n_reads, W)
    q = np.asarray(qual[: n_reads * W]).reshape(n_reads, W) if qual is not None else None
    b1 = fastq_records(t[:, :L], None if q is None else q[:, :L], first_name).tobytes()
    b2 = fastq_records(t[:, L + 1:], None if q is None else q[:, L + 1:], first_name).tobytes() if paired else None
    return b1, b2

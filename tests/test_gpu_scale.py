"""GPU tests at the sizes BASELINE.json names, through properties that do not need the oracle to
finish the whole input:

  * an oracle-checked prefix of the same synthetic stream (bit-exact);
  * chunking invariance: the association list does not depend on how the reads are cut into
    chunks (1 Mi-read chunks vs ragged small ones), compared through an order-independent
    checksum of (read, gene) pairs and the keep-flag count;
  * permutation invariance: shuffling the reads permutes the result and nothing else;
  * strand symmetry: a single-end read and its reverse complement get the same genes (canonical
    k-mers; the coverage of ReadAnalyzer.hpp:58,81 is the size of the union of the hit windows);
  * keep flags == reads that own at least one association; genes of a read strictly ascending.

C2 runs at its full 10 M reads; C3 and C4 use their reference, flags and read shape on a 1 Mi-pair
slice (the full 50 M / 100 M pairs are the same kernel launches repeated).
"""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

MIX = np.uint64(0x9E3779B97F4A7C15)


def _mix(read_idx, gene_idx):
    """order-independent checksum of a set of (read, gene) pairs"""
    with np.errstate(over="ignore"):
        x = (read_idx.astype(np.uint64) << np.uint64(20)) ^ gene_idx.astype(np.uint64)
        x = (x ^ (x >> np.uint64(31))) * MIX
        x = (x ^ (x >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
        return int(np.bitwise_xor.reduce(x ^ (x >> np.uint64(32)))) if len(x) else 0, int(x.sum(dtype=np.uint64)) if len(x) else 0


def _run(sh, text, qual, W, chunk):
    """all reads of a fixed-width block through analyze_chunks in chunks of `chunk` reads"""
    n = text.shape[0]
    chunks = []
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        off = (np.arange(m + 1, dtype=np.uint32) * np.uint32(W))
        chunks.append((text[a:a + m].reshape(-1), None if qual is None else qual[a:a + m].reshape(-1), off, m))
    ar, ag, kept, slow = [], [], 0, 0
    base = 0
    for res, c in zip(sh.analyze_chunks(chunks), chunks):
        ar.append(res["read_idx"].astype(np.int64) + base)
        ag.append(res["gene_idx"].astype(np.int64))
        kept += int(res["keep"].sum())
        slow += res["n_slow_reads"]
        base += c[3]
    return np.concatenate(ar), np.concatenate(ag), kept, slow


def _check_lists(ar, ag, kept):
    assert kept == len(np.unique(ar))
    same = ar[1:] == ar[:-1]
    assert np.all(ar[1:] >= ar[:-1])                 # reads in input order
    assert np.all(ag[1:][same] > ag[:-1][same])      # genes of one read strictly ascending


CONFIGS = {
    # name: genes, reads, L, paired, k, b, q, single
    "c2_full": dict(genes=1000, reads=10_000_000, L=100, paired=False, k=17, b=1, q=0, single=False),
    "c3_slice": dict(genes=5000, reads=1 << 20, L=150, paired=True, k=21, b=1, q=20, single=True),
    "c4_slice": dict(genes=20000, reads=1 << 20, L=150, paired=True, k=31, b=4, q=0, single=False),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_scale_properties(name):
    from shark_b200 import synth
    from shark_b200.engine import Shark
    cfg = CONFIGS[name]
    names, bases, rec_off = synth.make_reference(cfg["genes"], seed=1)
    want_q = cfg["q"] > 0
    seq, qual, _ = synth.make_reads(bases, cfg["genes"], cfg["reads"], cfg["L"], cfg["paired"], seed=2,
                                    varied_qual=want_q, want_qual=want_q)
    W = 2 * cfg["L"] + 1 if cfg["paired"] else cfg["L"]
    text = seq.reshape(-1, W)
    q2 = qual.reshape(-1, W) if want_q else None
    n = text.shape[0]
    with Shark(k=cfg["k"], c=0.6, bf_bits=cfg["b"] << 33, min_quality=cfg["q"], single=cfg["single"],
               max_reads_per_chunk=1 << 20, max_bytes_per_chunk=(1 << 20) * W) as sh:
        info = sh.build_index(bases, rec_off)
        assert info.n_genes == cfg["genes"]
        ar, ag, kept, slow = _run(sh, text, q2, W, 1 << 20)
        assert len(ar) > 0.5 * n * (0.2 if cfg["single"] else 1)   # not vacuous
        _check_lists(ar, ag, kept)
        ref_sum = _mix(ar, ag)

        # oracle on a prefix of the same stream (bit-exact)
        m = 20000
        off = np.arange(m + 1, dtype=np.uint64) * np.uint64(W)
        ora = po.Index(bases, rec_off, cfg["k"], cfg["b"] << 33)
        cnt0, ar0, ag0 = ora.analyze(text[:m].reshape(-1), off, 0.6, qual=None if q2 is None else q2[:m].reshape(-1),
                                     min_quality=cfg["q"], single=cfg["single"])
        pre = ar < m
        assert np.array_equal(ar[pre], ar0) and np.array_equal(ag[pre], ag0)

        # chunking invariance (ragged chunk size, several chunks in flight)
        sub = min(n, 3_000_000)
        ar2, ag2, kept2, _ = _run(sh, text[:sub], None if q2 is None else q2[:sub], W, 333_333)
        pre = ar < sub
        assert np.array_equal(ar2, ar[pre]) and np.array_equal(ag2, ag[pre])

        # permutation invariance
        sub = min(n, 2_000_000)
        perm = np.random.default_rng(5).permutation(sub)
        ar3, ag3, kept3, _ = _run(sh, text[:sub][perm], None if q2 is None else q2[:sub][perm], W, 1 << 20)
        pre = ar < sub
        assert _mix(perm[ar3], ag3) == _mix(ar[pre], ag[pre])
        assert kept3 == len(np.unique(ar[pre]))

        # strand symmetry (single-end, no quality masking)
        if not cfg["paired"] and not want_q:
            comp = np.arange(256, dtype=np.uint8)
            for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
                comp[a] = b
            rcs = comp[text[:sub][:, ::-1]]
            ar4, ag4, kept4, _ = _run(sh, np.ascontiguousarray(rcs), None, W, 1 << 20)
            assert np.array_equal(ar4, ar[pre]) and np.array_equal(ag4, ag[pre])
        assert ref_sum == _mix(ar, ag)

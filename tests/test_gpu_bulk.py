"""GPU parity of the bulk classification kernel (shk_bulk.cu: diagonals, match runs, shared lookups).

The kernel decides per window whether its gene ids are looked up or copied from the window before; every such
decision must be invisible in the results.  The reads here are the shapes that stress those decisions: both
strands, substitutions / insertions / deletions (the diagonal shifts), chimeras (the diagonal jumps to another
gene or strand), reads that overhang the ends of the reference or span two records, copies of a region in several
genes (the anchor names another gene than the read's), lower case, invalid bytes, runs of one base, lengths from 0
to 1024 at every alignment of the packed stream.  Each case is classified by the bulk kernel, by
analyze_reads_kernel (SHK_BULK=0) and by the oracle; all three must agree bit for bit."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import ACGT, quirky_reference, rnd_genes, to_soa

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def rc(s):
    return s.translate(COMP)[::-1]


def mutate(rng, s, sub=0.01, ins=0.0, dele=0.0, pn=0.0):
    out = bytearray()
    for ch in s:
        u = rng.random()
        if u < dele:
            continue
        if u < dele + ins:
            out.append(int(ACGT[rng.integers(0, 4)]))
        u = rng.random()
        if u < sub:
            out.append(int(ACGT[rng.integers(0, 4)]))
        elif u < sub + pn:
            out.append(ord("N"))
        else:
            out.append(ch)
    return bytes(out)


def stress_reads(rng, genes, n, kmax):
    """n reads of the shapes listed in the module docstring."""
    flat = b"".join(genes)                       # reads cut from the concatenation span record boundaries
    usable = [g for g in genes if len(g) >= 200]
    texts = []
    for i in range(n):
        kind = i % 12
        g = usable[int(rng.integers(0, len(usable)))]
        L = int(rng.integers(1, 320))
        st = int(rng.integers(0, max(1, len(g) - L)))
        piece = g[st:st + L]
        if kind == 0:                              # clean, either strand
            t = piece
        elif kind == 1:                            # substitutions and N
            t = mutate(rng, piece, sub=0.02, pn=0.004)
        elif kind == 2:                            # insertions and deletions: the diagonal shifts
            t = mutate(rng, piece, sub=0.005, ins=0.01, dele=0.01)
        elif kind == 3:                            # chimera of two genes, second part on the other strand half the time
            g2 = usable[int(rng.integers(0, len(usable)))]
            st2 = int(rng.integers(0, max(1, len(g2) - 150)))
            p2 = g2[st2:st2 + int(rng.integers(20, 150))]
            t = piece + (rc(p2) if rng.random() < 0.5 else p2)
        elif kind == 4:                            # overhang: random bases in front of / behind a prefix or suffix of a gene
            junk = ACGT[rng.integers(0, 4, int(rng.integers(1, 120)))].tobytes()
            t = junk + g[:L] if rng.random() < 0.5 else g[-L:] + junk
        elif kind == 5:                            # across a record boundary of the concatenated reference
            st = int(rng.integers(0, max(1, len(flat) - L)))
            t = flat[st:st + L]
        elif kind == 6:                            # paired layout: mate + 'N' + reverse-complemented mate downstream
            m2 = g[min(st + 60, max(0, len(g) - L)):][:L]
            t = mutate(rng, piece) + b"N" + mutate(rng, rc(m2))
        elif kind == 7:                            # background
            t = ACGT[rng.integers(0, 4, L)].tobytes()
        elif kind == 8:                            # lower case and invalid bytes in a gene read
            t = bytearray(mutate(rng, piece).lower())
            for j in rng.integers(0, max(1, len(t)), 2):
                if len(t):
                    t[int(j)] = int(rng.integers(0, 256))
            t = bytes(t)
        elif kind == 9:                            # long reads (up to kMaxFastLen and beyond)
            Lb = int(rng.integers(600, 1100))
            stb = int(rng.integers(0, max(1, len(flat) - Lb)))
            t = mutate(rng, flat[stb:stb + Lb], sub=0.01)
        elif kind == 10:                           # the same read twice in a row, second copy reverse-complemented
            t = piece + rc(piece)
        else:                                      # runs of one base around a gene piece
            t = b"A" * int(rng.integers(0, 80)) + piece + b"T" * int(rng.integers(0, 80))
        if rng.random() < 0.5:
            t = rc(t)
        texts.append(t)
    return texts


def classify(monkeypatch, bulk, bases, rec_off, seq, off, **kw):
    from shark_b200.engine import Shark
    monkeypatch.setenv("SHK_BULK", "1" if bulk else "0")
    with Shark(max_reads_per_chunk=1500, extend=True, compact=True, **kw) as sh:
        sh.build_index(bases, rec_off)
        keep, ar, ag, stats = sh.analyze(seq, off, None, packed=True)
    return keep, ar, ag, stats


@pytest.mark.parametrize("k,c,single,bf_bits", [(31, 0.6, False, 1 << 30), (21, 0.5, True, 1 << 28), (17, 0.6, False, 1 << 26),
                                                 (11, 0.3, False, 1 << 22), (5, 0.8, False, 1 << 20), (1, 0.5, False, 1 << 16),
                                                 (25, 0.0, False, 1000003), (13, 0.6, True, 3 << 33)])
def test_bulk_stress_parity(monkeypatch, k, c, single, bf_bits):
    rng = np.random.default_rng(7000 + k)
    genes = quirky_reference(rng)
    genes += rnd_genes(rng, 20, 1500, 3000)
    genes[45] = genes[44][:700] + genes[45][700:]           # copies: the anchor of a shared window names the first gene
    genes[47] = genes[46][300:1200] + genes[47][900:]
    bases, rec_off = po.concat_records(genes)
    texts = stress_reads(rng, [g.upper() for g in genes], 6000, k)
    texts += [b"", b"A", b"ACGT" * 8]
    seq, off = to_soa(texts)
    ref = po.Index(bases, rec_off, k, bf_bits)
    cnt0, ar0, ag0 = ref.analyze(seq, off, c, single=single)
    out = {}
    for bulk in (True, False):
        keep, ar, ag, stats = classify(monkeypatch, bulk, bases, rec_off, seq, off, k=k, c=c, bf_bits=bf_bits, single=single)
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8)), bulk
        assert np.array_equal(ar, ar0), bulk
        assert np.array_equal(ag, ag0), bulk
        out[bulk] = stats
    # both kernels count the same valid windows and the same windows with ids
    assert out[True]["n_probes"] == out[False]["n_probes"]
    assert out[True]["n_hits"] == out[False]["n_hits"]
    if k >= 17 and bf_bits >= 1 << 26:   # (a filter of 10^6 bits is dense: neighbouring windows rarely share a list)
        assert out[True]["n_extended"] > 0.3 * out[True]["n_hits"]


def test_bulk_single_diagonal_reads(monkeypatch):
    """One gene, reads that are exact pieces of it on both strands starting at every offset of the first 130 bases:
    every alignment of read word against reference word, every strand, no lookups but the anchors."""
    rng = np.random.default_rng(12)
    gene = ACGT[rng.integers(0, 4, 2000)].tobytes()
    bases, rec_off = po.concat_records([gene, ACGT[rng.integers(0, 4, 500)].tobytes()])
    texts = []
    for st in range(130):
        for L in (31, 32, 33, 63, 64, 65, 100, 151):
            texts.append(gene[st:st + L])
            texts.append(rc(gene[st:st + L]))
            texts.append(gene[2000 - st - L:2000 - st])
    seq, off = to_soa(texts)
    for k in (31, 16, 7):
        ref = po.Index(bases, rec_off, k, 1 << 26)
        cnt0, ar0, ag0 = ref.analyze(seq, off, 0.6)
        keep, ar, ag, stats = classify(monkeypatch, True, bases, rec_off, seq, off, k=k, c=0.6, bf_bits=1 << 26)
        assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
        assert stats["n_extended"] > 0.8 * stats["n_hits"]


def test_auto_mode_small_table(monkeypatch):
    """Without a forcing flag an index whose front table fits L2 still carries the extension structures: packed reads go
    to the bulk kernel, text reads to the thread-per-read kernel over the slots-only copy of the table
    (info.plain_front); both equal the oracle.  A forcing flag keeps the single table of its path."""
    from shark_b200.engine import Shark
    monkeypatch.delenv("SHK_BULK", raising=False)
    rng = np.random.default_rng(99)
    genes = rnd_genes(rng, 30, 800, 2500)
    genes[11] = genes[10][:600] + genes[11][600:]
    bases, rec_off = po.concat_records(genes)
    texts = stress_reads(rng, [g.upper() for g in genes], 4000, 21)
    seq, off = to_soa(texts)
    ref = po.Index(bases, rec_off, 21, 1 << 28)
    cnt0, ar0, ag0 = ref.analyze(seq, off, 0.5)
    for extend, want_ext, want_plain in ((None, 1, 1), (True, 1, 0), (False, 0, 0)):
        with Shark(max_reads_per_chunk=1500, extend=extend, compact=True, k=21, c=0.5, bf_bits=1 << 28) as sh:
            info = sh.build_index(bases, rec_off)
            assert (info.extend, info.plain_front) == (want_ext, want_plain), extend
            for packed in (False, True):
                keep, ar, ag, stats = sh.analyze(seq, off, None, packed=packed)
                assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0), (extend, packed)
                assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8)), (extend, packed)
                # who extends: the bulk kernel whenever the structures exist, the thread-per-read kernel only when forced
                extends = (packed and want_ext) or (want_ext and not want_plain)
                assert (stats["n_extended"] > 0) == bool(extends), (extend, packed, stats["n_extended"])

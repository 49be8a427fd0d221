"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same
inputs.  Bit-exact (everything here is integer/byte/index work; the one floating-point
operation, `max >= c*len`, is an IEEE double multiply on both sides)."""
import os

import numpy as np
import pytest

from helpers import ACGT, GOLDEN, edge_cases, example_cases, flags_to_kwargs, gz_read, md5, stage_edge, stage_example
from oracle import pyoracle as po
from oracle import shark_text as st

pytestmark = pytest.mark.gpu


def _shark(**kw):
    from shark_b200.engine import Shark
    return Shark(**kw)


@pytest.fixture(params=[False, True], ids=["lookup", "extend"])
def extend(request):
    """Every classification test runs twice: table lookups only, and with anchor-and-extend forced
    on (it is automatic only for DRAM-sized tables).  Results must be identical."""
    return request.param


def rnd_genes(rng, n, lo=200, hi=900):
    return [ACGT[rng.integers(0, 4, int(rng.integers(lo, hi)))].tobytes() for _ in range(n)]


def quirky_reference(rng):
    g = [bytearray(s) for s in rnd_genes(rng, 40)]
    g[3][100:400] = g[2][50:350]
    g[5] = bytearray(g[4])
    g[6][200:203] = b"NNN"
    g[7] = bytearray(b"N" * 64)          # no valid window, len >= k  (nidx quirk)
    g[8] = bytearray(b"ACGTACG")         # shorter than k
    g[9] = bytearray(bytes(g[9]).lower())
    g[10][0:60] = b"A" * 60
    g[11][300:360] = b"A" * 60
    g[12] = bytearray(b"")               # empty record
    for i in range(20, 34):              # a family sharing a 120-bp segment -> long lists, >8 genes per read
        g[i][50:170] = g[19][50:170]
    g[35] = bytearray(b"N" * 10 + bytes(g[35][:100]) + b"-" + bytes(g[35][100:]))
    return [bytes(x) for x in g]


def sample_reads(rng, genes, n, L, paired=False, err=0.02, pn=0.004, bg=0.15):
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    texts = []
    usable = [s.upper() for s in genes if len(s) >= L + 60]
    for _ in range(n):
        def mate(src, st_):
            m = bytearray(src[st_:st_ + L])
            for j in range(len(m)):
                u = rng.random()
                if u < err:
                    m[j] = ACGT[rng.integers(0, 4)]
                elif u < err + pn:
                    m[j] = ord("N")
            return bytes(m)
        if rng.random() < bg:
            a, b = ACGT[rng.integers(0, 4, L)].tobytes(), ACGT[rng.integers(0, 4, L)].tobytes()
        else:
            s = usable[int(rng.integers(0, len(usable)))]
            st_ = int(rng.integers(0, len(s) - L - 50))
            a = mate(s, st_)
            b = mate(s, st_ + 50).translate(comp)[::-1]
            if rng.random() < 0.5:
                a, b = b, a
        texts.append(a + b"N" + b if paired else a)
    return texts


def to_soa(texts):
    seq, off = po.concat_records(texts)
    return seq, off


def quals_for(rng, texts, paired_L=None):
    qs = []
    for t in texts:
        q = np.where(rng.random(len(t)) < 0.05, rng.integers(35, 53, len(t)), rng.integers(60, 74, len(t))).astype(np.uint8)
        if paired_L is not None:
            q[paired_L] = 0x1B
        qs.append(q.tobytes())
    return np.frombuffer(b"".join(qs), dtype=np.uint8).copy()


# ------------------------------------------------------------------------------------------
# index build: identical set bits, identical CSR
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,bf_bits", [(17, 1 << 33), (5, 1 << 20), (31, 1 << 24), (11, 1000003), (21, 3 << 33),
                                        (1, 1 << 16), (13, 4099)])
def test_index_build_parity(k, bf_bits):
    rng = np.random.default_rng(k * 1000 + bf_bits % 997)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    ref = po.Index(bases, rec_off, k, bf_bits)
    with _shark(k=k, bf_bits=bf_bits, max_reads_per_chunk=1 << 12) as sh:
        info = sh.build_index(bases, rec_off)
        assert info.n_records == len(genes)
        assert info.n_genes == ref.n_genes
        assert info.n_set_bits == ref.n_set
        assert info.tot_ids == ref.tot_ids
        pos, off, ids = sh.export_index()
        assert np.array_equal(pos, ref.pos)
        assert np.array_equal(off, ref.off)
        assert np.array_equal(ids, ref.ids)


def test_index_empty_and_degenerate():
    for genes in ([], [b""], [b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"], [b"ACG"]):
        bases, rec_off = po.concat_records(genes)
        ref = po.Index(bases, rec_off, 17, 1 << 22)
        with _shark(k=17, bf_bits=1 << 22, max_reads_per_chunk=1 << 10) as sh:
            info = sh.build_index(bases, rec_off)
            assert (info.n_genes, info.n_set_bits, info.tot_ids) == (ref.n_genes, ref.n_set, ref.tot_ids)
            seq, off = to_soa([b"ACGTACGTACGTACGTACGTACGTACGT", b""])
            keep, ar, ag, _ = sh.analyze(seq, off)
            assert keep.tolist() == [0, 0] and len(ar) == 0


# ------------------------------------------------------------------------------------------
# probe = BF::get_index
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,bf_bits", [(17, 1 << 33), (9, 1 << 18), (21, 3 << 33), (15, 1000003)])
def test_probe_parity(k, bf_bits):
    rng = np.random.default_rng(5 + k)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    ref = po.Index(bases, rec_off, k, bf_bits)
    kmers = [rng.integers(0, 1 << (2 * k), 5000, dtype=np.uint64)]
    for s in genes[:12]:
        e = po.enumerate_kmers(s, k)
        if e is not None:
            kmers.append(e[0])
    kmers = np.concatenate(kmers)
    with _shark(k=k, bf_bits=bf_bits, max_reads_per_chunk=1 << 10) as sh:
        sh.build_index(bases, rec_off)
        rank, begin, ln = sh.get_index(kmers)
    r0, b0, l0 = ref.probe(kmers)
    assert np.array_equal(rank, r0)
    assert np.array_equal(begin, b0)
    assert np.array_equal(ln, l0)
    assert (rank >= 0).sum() > 1000


# ------------------------------------------------------------------------------------------
# read classification
# ------------------------------------------------------------------------------------------
CASES = [
    dict(k=17, c=0.6, bf_bits=1 << 33, q=0, single=False, paired=False, L=100),
    dict(k=17, c=0.6, bf_bits=1 << 33, q=0, single=False, paired=True, L=100),
    dict(k=21, c=0.6, bf_bits=1 << 33, q=20, single=True, paired=True, L=150),
    dict(k=31, c=0.9, bf_bits=1 << 26, q=0, single=True, paired=True, L=150),
    dict(k=5, c=0.3, bf_bits=1 << 20, q=0, single=False, paired=True, L=100),   # many genes per read
    dict(k=7, c=0.5, bf_bits=4099, q=0, single=False, paired=False, L=76),      # heavy false positives
    dict(k=11, c=0.7, bf_bits=1000003, q=25, single=False, paired=False, L=64),
    dict(k=13, c=0.0, bf_bits=1 << 22, q=0, single=False, paired=False, L=33),
    dict(k=13, c=1.0, bf_bits=1 << 22, q=0, single=False, paired=True, L=50),
    dict(k=1, c=1.0, bf_bits=1 << 16, q=0, single=False, paired=False, L=40),
    dict(k=21, c=0.6, bf_bits=3 << 33, q=95, single=False, paired=False, L=100),  # `char mq` wraps
    dict(k=19, c=0.6, bf_bits=1 << 30, q=200, single=False, paired=True, L=90),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_c%g_b%d_q%d_%s%s_L%d" % (
    c["k"], c["c"], c["bf_bits"], c["q"], "s" if c["single"] else "m", "pe" if c["paired"] else "se", c["L"]))
def test_analyze_parity(case, extend):
    rng = np.random.default_rng(case["k"] * 7 + case["L"])
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    texts = sample_reads(rng, genes, 3000, case["L"], paired=case["paired"])
    texts += [b"", b"A", b"N" * 50, b"ACGT" * 10 + b"N" + b"ACGT" * 10]
    seq, off = to_soa(texts)
    qual = None
    if case["q"]:
        qual = quals_for(rng, texts)
        if case["paired"]:
            for i in range(3000):
                qual[int(off[i]) + case["L"]] = 0x1B
    ref = po.Index(bases, rec_off, case["k"], case["bf_bits"])
    cnt0, ar0, ag0 = ref.analyze(seq, off, case["c"], qual=qual, min_quality=case["q"], single=case["single"])
    with _shark(k=case["k"], c=case["c"], bf_bits=case["bf_bits"], min_quality=case["q"], single=case["single"],
                max_reads_per_chunk=1024, extend=extend) as sh:   # several chunks -> exercises both slots
        info = sh.build_index(bases, rec_off)
        assert info.extend == int(extend)
        keep, ar, ag, stats = sh.analyze(seq, off, qual)
    assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
    assert np.array_equal(ar, ar0)
    assert np.array_equal(ag, ag0)
    assert stats["chunks"] >= 3
    if case["k"] >= 11:
        assert len(ar0) > 500  # the case is not vacuous
        if extend:  # ... and neither is the extension: a good share of the hits came without a table access
            assert stats["n_extended"] > 0.1 * stats["n_hits"], (stats["n_extended"], stats["n_hits"])
    if not extend:
        assert stats["n_extended"] == 0


def test_exact_path_long_reads_and_wide_lists(extend):
    """Reads longer than 1024 bytes, reads touching > 8 genes and lists longer than 8 all take
    the exact path and must agree with the oracle."""
    rng = np.random.default_rng(99)
    core = ACGT[rng.integers(0, 4, 400)].tobytes()
    genes = [core + ACGT[rng.integers(0, 4, 300)].tobytes() for _ in range(300)]  # 300 genes share 400 bp
    genes += rnd_genes(rng, 20, 2500, 3000)
    bases, rec_off = po.concat_records(genes)
    texts = []
    for i in range(200):
        texts.append(core[i:i + 120])                                   # list length 300 per k-mer
        g = genes[300 + i % 20]
        texts.append(g[:1500 + i])                                      # > 1024 bytes
        texts.append(genes[i][350:500])                                 # crosses shared/unique boundary
    seq, off = to_soa(texts)
    for k, c, single in ((17, 0.6, False), (17, 0.6, True), (25, 0.2, False)):
        ref = po.Index(bases, rec_off, k, 1 << 28)
        cnt0, ar0, ag0 = ref.analyze(seq, off, c, single=single)
        with _shark(k=k, c=c, bf_bits=1 << 28, single=single, max_reads_per_chunk=256,
                    max_bytes_per_chunk=1 << 20, extend=extend) as sh:
            sh.build_index(bases, rec_off)
            keep, ar, ag, stats = sh.analyze(seq, off)
        assert stats["n_slow_reads"] >= 300
        assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
        if not single:
            assert (cnt0 >= 300).sum() >= 100  # 300-way ties are emitted in ascending gene order


# ------------------------------------------------------------------------------------------
# golden fixtures: the reference's own example + reference-binary goldens, end to end
# ------------------------------------------------------------------------------------------
def run_pipeline(ref_path, s1, s2, k=17, c=0.6, b=1, q=0, single=False, batch=50000, extend=None):
    """Same host stages as oracle.shark_text.run_shark, but classification on the GPU."""
    legend, seqs = st.load_reference(ref_path)
    bases, rec_off = po.concat_records(seqs)
    pairs, bid = st.load_sample(s1, s2, batch)
    seq, qual, off = st.join_reads(pairs, q > 0)
    with _shark(k=k, c=c, bf_bits=b << 33, min_quality=q, single=single, max_reads_per_chunk=4096, extend=extend) as sh:
        sh.build_index(bases, rec_off)
        keep, ar, ag, _ = sh.analyze(seq, off, qual)
    ssv, o1, o2 = [], [], []
    prev, prev_batch = b"", 0
    for r, g in zip(ar.tolist(), ag.tolist()):
        a, bb = pairs[r]
        if bid[r] != prev_batch:
            prev, prev_batch = b"", bid[r]
        ssv.append(a[0] + b" " + legend[g] + b"\n")
        if prev != a[0]:
            o1.append(b"@" + a[0] + b"\n" + a[1] + b"\n+\n" + a[2] + b"\n")
            if bb is not None:
                o2.append(b"@" + bb[0] + b"\n" + bb[1] + b"\n+\n" + bb[2] + b"\n")
        prev = a[0]
    return b"".join(ssv), b"".join(o1), (b"".join(o2) if s2 else None)


@pytest.mark.parametrize("case", sorted(example_cases()["cases"]))
def test_example_goldens(tmp_path, case, extend):
    info = example_cases()["cases"][case]
    f = stage_example(tmp_path)
    ssv, o1, o2 = run_pipeline(f["ENSG00000277117.fa"], f["sample_1.fq"], f["sample_2.fq"] if info["paired"] else None,
                               extend=extend, **flags_to_kwargs(info["flags"]))
    assert ssv.count(b"\n") == info["ssv_lines"]
    assert md5(ssv) == info["ssv_md5"]
    assert md5(o1) == info["o1_md5"]
    if info["paired"]:
        assert md5(o2) == info["o2_md5"]
    if case == "default":
        d = os.path.join(GOLDEN, "example")
        assert ssv == gz_read(os.path.join(d, "ENSG00000277117.truth.ssv.gz"))
        assert o1 == gz_read(os.path.join(d, "sharked.sample_1.truth.fq.gz"))
        assert o2 == gz_read(os.path.join(d, "sharked.sample_2.truth.fq.gz"))


@pytest.mark.parametrize("scenario,case", [(s, c) for s, cs in sorted(edge_cases().items()) for c in sorted(cs)])
def test_edge_goldens(tmp_path, scenario, case, extend):
    info = edge_cases()[scenario][case]
    f = stage_edge(tmp_path, scenario)
    ssv, o1, o2 = run_pipeline(f["ref.fa"], f["r1.fq"], f.get("r2.fq") if info["paired"] else None,
                               extend=extend, **flags_to_kwargs(info["flags"]))
    d = os.path.join(GOLDEN, "edge", scenario)
    assert ssv == gz_read(os.path.join(d, case + ".ssv.gz"))
    assert o1 == gz_read(os.path.join(d, case + ".o1.fq.gz"))
    if info["paired"]:
        assert o2 == gz_read(os.path.join(d, case + ".o2.fq.gz"))

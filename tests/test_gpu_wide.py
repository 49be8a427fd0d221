"""the reference's limits lifted behind an opt-in flag.  SHK_F_WIDE_IDS keeps gene ids in 32 bits
(the reference stores them in 16: small_vector.hpp:46, bloomfilter.h:45) and bounds the id total by 32-bit offsets
instead of an `int` (bloomfilter.h:130).  The checker is the oracle's restatement with exactly those two types
widened (oracle.pyoracle.Index(wide=True)); for inputs inside the reference's limits the wide build must agree with
the default one, and the default must keep failing loudly beyond them."""
import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import ACGT, CASES, quirky_reference, rnd_genes, sample_reads, to_soa

pytestmark = pytest.mark.gpu


def _tiny_genes(rng, n, lo=36, hi=60):
    return [ACGT[rng.integers(0, 4, int(rng.integers(lo, hi)))].tobytes() for _ in range(n)]


def test_wide_ids_70000_genes_against_the_widened_oracle():
    from shark_b200 import capi
    from shark_b200.engine import Shark
    rng = np.random.default_rng(70)
    genes = _tiny_genes(rng, 70000)
    # families across the 16-bit boundary: gene i + 65536 shares a segment with gene i (lists {i, i + 65536}, which
    # 16-bit ids would collapse), a 300-way family of high ids (long lists -> the bitmap sort of wide ids, the exact
    # path of the classification), records without a window (the nidx quirk shifts every later index)
    for i in range(0, 3000, 7):
        genes[65536 + i] = genes[i][:30] + genes[65536 + i][30:]
    core = ACGT[rng.integers(0, 4, 40)].tobytes()
    for i in range(66000, 66300):
        genes[i] = core + genes[i][:12]
    genes[100] = b"N" * 40
    genes[69000] = b"ACGT"
    bases, rec_off = po.concat_records(genes)
    k, bf_bits, c = 15, 1 << 31, 0.5
    ref = po.Index(bases, rec_off, k, bf_bits, wide=True)
    assert ref.n_genes == 69999 and int(ref.ids32.max()) > 65535
    texts = [genes[i] for i in rng.integers(0, 70000, 3000)]
    texts += [genes[65536 + i] for i in range(0, 3000, 7)] + [genes[i] for i in range(0, 3000, 7)]
    texts += [genes[i][:30] for i in range(0, 3000, 7)]                  # ties {i, i + 65536}
    texts += [core[i:i + 30] for i in range(10)] + [core] * 30           # 300-way ties of ids above 65535
    texts += [core + ACGT[rng.integers(0, 4, 10)].tobytes() for _ in range(50)] + [b"", b"N" * 30, b"ACGTACGTAC"]
    texts += [genes[69999], genes[69998], genes[65535], genes[65536], genes[65537]]
    seq, off = to_soa(texts)
    cnt0, ar0, ag0 = ref.analyze(seq, off, c)
    assert int(ag0.max()) > 65535 and (cnt0 == 300).sum() >= 40 and (cnt0 == 2).sum() > 300
    with Shark(k=k, c=c, bf_bits=bf_bits, max_reads_per_chunk=1000, wide_ids=True) as sh:
        info = sh.build_index(bases, rec_off)
        assert (info.n_genes, info.n_set_bits, info.tot_ids, info.id_bits) == (ref.n_genes, ref.n_set, ref.tot_ids, 32)
        assert info.front_entries == 0 and info.extend == 0
        pos, coff, ids = sh.export_index()
        assert ids.dtype == np.uint32
        assert np.array_equal(pos, ref.pos) and np.array_equal(coff, ref.off) and np.array_equal(ids, ref.ids32)
        with pytest.raises(capi.SharkError):   # 16-bit export of a wide index
            sh.export_index(wide=False)
        for packed in (False, True):
            keep, ar, ag, stats = sh.analyze(seq, off, packed=packed)
            assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
            assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
        rank, begin, ln = sh.get_index(np.concatenate([po.enumerate_kmers(genes[65536 + 7], k)[0],
                                                       rng.integers(0, 1 << (2 * k), 500, dtype=np.uint64)]))
        r0, b0, l0 = ref.probe(np.concatenate([po.enumerate_kmers(genes[65536 + 7], k)[0],
                                               np.zeros(0, np.uint64)]))
        assert np.array_equal(rank[: len(r0)], r0) and np.array_equal(begin[: len(r0)], b0) and np.array_equal(ln[: len(r0)], l0)
    # the default keeps the reference's limit, loudly
    with Shark(k=k, c=c, bf_bits=bf_bits, max_reads_per_chunk=1000) as sh:
        with pytest.raises(capi.SharkError) as ei:
            sh.build_index(bases, rec_off)
        assert ei.value.code == -5 and "SHK_F_WIDE_IDS" in str(ei.value)


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[4], CASES[5], CASES[10]], ids=lambda c: "k%d_q%d" % (c["k"], c["q"]))
def test_wide_ids_equal_the_default_inside_the_reference_limits(case):
    """Same inputs, both id widths: identical index (ids widened) and identical associations - through the oracle too."""
    from shark_b200.engine import Shark
    rng = np.random.default_rng(case["k"] * 3 + 1)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    texts = sample_reads(rng, genes, 2000, case["L"], paired=case["paired"]) + [b"", b"A", b"N" * 40]
    seq, off = to_soa(texts)
    qual = None
    if case["q"]:
        qual = rng.integers(33, 75, len(seq)).astype(np.uint8)
    ref = po.Index(bases, rec_off, case["k"], case["bf_bits"])
    cnt0, ar0, ag0 = ref.analyze(seq, off, case["c"], qual=qual, min_quality=case["q"], single=case["single"])
    out = {}
    for wide in (False, True):
        with Shark(k=case["k"], c=case["c"], bf_bits=case["bf_bits"], min_quality=case["q"], single=case["single"],
                   max_reads_per_chunk=700, wide_ids=wide, compact=wide) as sh:
            sh.build_index(bases, rec_off)
            out[wide] = (sh.export_index(wide=True), sh.analyze(seq, off, qual)[:3])
    for a, b in zip(out[False][0], out[True][0]):
        assert np.array_equal(a, b)
    for w in (False, True):
        keep, ar, ag = out[w][1]
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8)) and np.array_equal(ar, ar0) and np.array_equal(ag, ag0)


def test_wide_ids_replicate_save_load(tmp_path):
    """The wide index travels through the view / adopt / finalize protocol and through a file; a context of the
    other id width refuses it."""
    from shark_b200 import capi
    from shark_b200.engine import Shark
    rng = np.random.default_rng(9)
    genes = _tiny_genes(rng, 66000)
    bases, rec_off = po.concat_records(genes)
    texts = [genes[i] for i in (0, 5, 65535, 65536, 65999)]
    seq, off = to_soa(texts)
    path = str(tmp_path / "wide.idx")
    with Shark(k=13, bf_bits=1 << 28, max_reads_per_chunk=64, wide_ids=True) as a:
        a.build_index(bases, rec_off)
        want = a.analyze(seq, off)[:3]
        assert want[2].tolist() == [0, 5, 65535, 65536, 65999]
        a.save_index(path)
        with Shark(k=13, bf_bits=1 << 28, max_reads_per_chunk=64, wide_ids=True) as b:
            rc = b.lib.shk_index_replicate(a.ctx, b.ctx)
            assert rc == 0
            import ctypes as C
            i = capi.IndexInfo()
            b.lib.shk_index_info_get(b.ctx, C.byref(i))
            b.info = i
            got = b.analyze(seq, off)[:3]
            assert all(np.array_equal(x, y) for x, y in zip(want, got))
    with Shark(k=13, bf_bits=1 << 28, max_reads_per_chunk=64, wide_ids=True) as c:
        c.load_index(path)
        got = c.analyze(seq, off)[:3]
        assert all(np.array_equal(x, y) for x, y in zip(want, got))
    with Shark(k=13, bf_bits=1 << 28, max_reads_per_chunk=64) as d:
        with pytest.raises(capi.SharkError):
            d.load_index(path)

"""GPU parity of the packed read path and of the compact result form.

shk_reads_submit_packed takes a chunk already reduced to what the kernels use of a text byte (2-bit code +
validity bit per base, the -q masking rule of FastqSplitter.hpp:104-109 folded in) and the PACKED variant of
the classification kernels reads that form directly.  SHK_F_COMPACT_RESULTS leaves the results in the form in
which they cross the link (one 16-bit word per read + a list for ties).  Every classification case of
test_gpu_parity.py is repeated through both and must equal the oracle bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import ACGT, CASES, quals_for, quirky_reference, rnd_genes, sample_reads, to_soa

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[False, True], ids=["lookup", "extend"])
def extend(request):
    return request.param


def _case_inputs(case, seed_mul=13):
    rng = np.random.default_rng(case["k"] * seed_mul + case["L"])
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    texts = sample_reads(rng, genes, 3000, case["L"], paired=case["paired"])
    # odd lengths (reads start at every alignment of the packed stream), empty reads, invalid bytes of all kinds
    texts += [b"", b"A", b"N" * 50, b"ACGT" * 10 + b"N" + b"ACGT" * 10, b"acgt" * 12 + b"n.-*" + b"ACGT" * 10,
              bytes([0x81, 0xA1, 0xC3, 0xE7]) * 20, bytes(rng.integers(0, 256, 300, dtype=np.uint8))]
    texts += [ACGT[rng.integers(0, 4, int(n))].tobytes() for n in rng.integers(1, 70, 64)]
    seq, off = to_soa(texts)
    qual = None
    if case["q"]:
        qual = quals_for(rng, texts)
        tail = int(off[3000])
        qual[tail:] = rng.integers(0, 256, len(qual) - tail, dtype=np.uint8)
        if case["paired"]:
            for i in range(3000):
                qual[int(off[i]) + case["L"]] = 0x1B
    return bases, rec_off, seq, off, qual


@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_c%g_b%d_q%d_%s%s_L%d" % (
    c["k"], c["c"], c["bf_bits"], c["q"], "s" if c["single"] else "m", "pe" if c["paired"] else "se", c["L"]))
def test_packed_analyze_parity(case, extend):
    from shark_b200.engine import Shark
    bases, rec_off, seq, off, qual = _case_inputs(case)
    ref = po.Index(bases, rec_off, case["k"], case["bf_bits"])
    cnt0, ar0, ag0 = ref.analyze(seq, off, case["c"], qual=qual, min_quality=case["q"], single=case["single"])
    with Shark(k=case["k"], c=case["c"], bf_bits=case["bf_bits"], min_quality=case["q"], single=case["single"],
               max_reads_per_chunk=1000, extend=extend, compact=True) as sh:
        sh.build_index(bases, rec_off)
        keep, ar, ag, stats = sh.analyze(seq, off, qual, packed=True)
        packed_bytes = sh.h2d_bytes()
    assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
    assert np.array_equal(ar, ar0)
    assert np.array_equal(ag, ag0)
    assert stats["chunks"] >= 3
    assert packed_bytes < 0.45 * len(seq) + 4 * (len(off) + 64)   # 0.375 bytes per base + offsets
    if case["k"] >= 11:
        assert len(ar0) > 500
        if extend:
            assert stats["n_extended"] > 0.1 * stats["n_hits"]


def test_packed_exact_path(extend):
    """Reads longer than 1024 bytes, 300-way ties and lists of 300 ids through the packed middle and exact paths."""
    from shark_b200.engine import Shark
    rng = np.random.default_rng(99)
    core = ACGT[rng.integers(0, 4, 400)].tobytes()
    genes = [core + ACGT[rng.integers(0, 4, 300)].tobytes() for _ in range(300)]
    genes += rnd_genes(rng, 20, 2500, 3000)
    bases, rec_off = po.concat_records(genes)
    texts = []
    for i in range(200):
        texts.append(core[i:i + 120])
        texts.append(genes[300 + i % 20][:1500 + i])
        texts.append(genes[i][350:500])
    seq, off = to_soa(texts)
    for k, c, single in ((17, 0.6, False), (25, 0.2, True)):
        ref = po.Index(bases, rec_off, k, 1 << 28)
        cnt0, ar0, ag0 = ref.analyze(seq, off, c, single=single)
        with Shark(k=k, c=c, bf_bits=1 << 28, single=single, max_reads_per_chunk=256, max_bytes_per_chunk=1 << 20,
                   extend=extend, compact=True) as sh:
            sh.build_index(bases, rec_off)
            keep, ar, ag, stats = sh.analyze(seq, off, packed=True)
        assert stats["n_slow_reads"] >= 300
        assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))


def test_compact_equals_expanded():
    """The same chunk collected in both result forms: shk_result_expand (inside shk_reads_collect) and the
    vectorised expansion of the compact arrays give the same lists; marker collisions (gene indices 0xFFFE and
    0xFFFF as single winners) travel through the multi list."""
    from shark_b200 import capi
    from shark_b200.engine import Shark
    rng = np.random.default_rng(5)
    # 65536 tiny records: the last two gene indices are 0xFFFE and 0xFFFF; reads of them must be reported
    n_rec = 65536
    genes = [ACGT[rng.integers(0, 4, 40)].tobytes() for _ in range(n_rec)]
    bases, rec_off = po.concat_records(genes)
    texts = [genes[i] for i in (0, 1, 65533, 65534, 65535, 65534, 65535, 7)] + [b"N" * 40]
    seq, off = to_soa(texts)
    ref = po.Index(bases, rec_off, 15, 1 << 30)
    cnt0, ar0, ag0 = ref.analyze(seq, off, 0.6)
    assert {65534, 65535} <= set(ag0.tolist())
    out = []
    for compact in (False, True):
        with Shark(k=15, c=0.6, bf_bits=1 << 30, max_reads_per_chunk=1 << 10, compact=compact) as sh:
            sh.build_index(bases, rec_off)
            keep, ar, ag, _ = sh.analyze(seq, off)
            out.append((keep, ar, ag))
            if compact:
                o32 = (off - off[0]).astype(np.uint32)
                pin = capi.PinnedBuffer(len(seq) + 64)
                pin.u8[:len(seq)] = seq
                sh.submit(0, pin.u8, None, o32, len(texts))
                r = sh.collect(0)
                assert r["read_idx"] is None and r["keep"] is None
                g16 = r["gene16"]
                assert (g16 == capi.GENE_MULTI).sum() >= 4 and g16[-1] == capi.GENE_NONE
                assert set(r["multi"][:, 1].tolist()) >= {65534, 65535}
    for (k0, a0, g0) in out:
        assert np.array_equal(k0, (cnt0 > 0).astype(np.uint8))
        assert np.array_equal(a0, ar0) and np.array_equal(g0, ag0)


def test_slot_protocol_is_guarded():
    """A slot must be collected before it is submitted to again: SHK_E_STATE, not a silent overwrite."""
    from shark_b200 import capi
    from shark_b200.engine import Shark
    rng = np.random.default_rng(3)
    genes = rnd_genes(rng, 4)
    bases, rec_off = po.concat_records(genes)
    seq, off = to_soa([genes[0][:100], genes[1][:100]])
    o32 = off.astype(np.uint32)
    pin = capi.PinnedBuffer(len(seq) + 64)
    pin.u8[:len(seq)] = seq
    with Shark(k=17, bf_bits=1 << 24, max_reads_per_chunk=64) as sh:
        sh.build_index(bases, rec_off)
        sh.submit(0, pin.u8, None, o32, 2)
        for call in (lambda: sh.submit(0, pin.u8, None, o32, 2), lambda: sh.upload(0, pin.u8, None, o32, 2),
                     lambda: sh.analyze_resident(0)):
            with pytest.raises(capi.SharkError) as ei:
                call()
            assert ei.value.code == -3
        r = sh.collect(0)
        assert r["n_reads"] == 2 and r["n_assoc"] == 2
        sh.submit(0, pin.u8, None, o32, 2)   # fine after the collect
        sh.collect(0)

"""GPU parity of the split upload (SHK_F_HOST_PACK): part of every chunk crosses PCIe as text, the rest as
3 bits per base packed by the host cores and expanded on the device.  Results must equal the oracle's for
every split point - the masking rule (FastqSplitter.hpp:104-109) is folded into the packed validity bits."""
import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import CASES, quals_for, quirky_reference, sample_reads, to_soa

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("share", [True, 1.0, 0.37, 0.05], ids=["auto", "all", "0.37", "0.05"])
@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[6], CASES[10], CASES[11]], ids=lambda c: "k%d_q%d_%s" % (
    c["k"], c["q"], "pe" if c["paired"] else "se"))
def test_split_upload_parity(case, share):
    from shark_b200.engine import Shark
    rng = np.random.default_rng(case["k"] * 11 + case["L"])
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    texts = sample_reads(rng, genes, 3000, case["L"], paired=case["paired"])
    # bytes that only differ in how they are invalid, lowercase bases, bytes >= 128 (0x81 - 64 = 'A' under masking)
    texts += [b"", b"A", b"N" * 50, b"acgt" * 12 + b"n.-*" + b"ACGT" * 10, bytes([0x81, 0xA1, 0xC3, 0xE7]) * 20,
              bytes(rng.integers(0, 256, 300, dtype=np.uint8))]
    seq, off = to_soa(texts)
    qual = None
    if case["q"]:
        qual = quals_for(rng, texts)
        tail = int(off[3000])
        qual[tail:] = rng.integers(0, 256, len(qual) - tail, dtype=np.uint8)   # every quality byte value
        if case["paired"]:
            for i in range(3000):
                qual[int(off[i]) + case["L"]] = 0x1B
    ref = po.Index(bases, rec_off, case["k"], case["bf_bits"])
    cnt0, ar0, ag0 = ref.analyze(seq, off, case["c"], qual=qual, min_quality=case["q"], single=case["single"])
    with Shark(k=case["k"], c=case["c"], bf_bits=case["bf_bits"], min_quality=case["q"], single=case["single"],
               max_reads_per_chunk=1024, host_pack=share) as sh:
        sh.build_index(bases, rec_off)
        keep, ar, ag, stats = sh.analyze(seq, off, qual)
        text_bytes = len(seq) * (2 if case["q"] else 1) + 4 * (len(off) - 1 + stats["chunks"])
        assert sh.h2d_bytes() < text_bytes            # something was packed ...
        if share is not True and share == 1.0:
            assert sh.h2d_bytes() < 0.45 * text_bytes  # ... everything, here
    assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
    assert np.array_equal(ar, ar0)
    assert np.array_equal(ag, ag0)
    assert stats["chunks"] >= 3 and len(ar0) > 500


def test_split_upload_equals_plain_on_big_chunks():
    """Chunks large enough for the pack pool (several blocks per thread), auto-balanced share, -q 20 with
    varied qualities: identical associations and keep flags as the plain upload."""
    from shark_b200 import synth
    from shark_b200.engine import Shark
    names, bases, rec_off = synth.make_reference(100, seed=1)
    text, qual = synth.make_read_block(bases, 100, 3000, 0, 1 << 18, 150, True, seed=2, varied_qual=True, want_qual=True)
    seq, qual = text.reshape(-1), qual.reshape(-1)
    off = np.arange(text.shape[0] + 1, dtype=np.uint64) * np.uint64(text.shape[1])
    out = []
    for hp in (False, True):
        with Shark(k=21, c=0.6, bf_bits=1 << 30, min_quality=20, single=True, max_reads_per_chunk=1 << 16,
                   max_bytes_per_chunk=(1 << 16) * 320, host_pack=hp) as sh:
            sh.build_index(bases, rec_off)
            out.append(sh.analyze(seq, off, qual)[:3])
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    assert len(out[0][1]) > 100000

"""GPU tests of the two index paths beside the one-call build (SURVEY.md 8e second mode, 8f.3):

* sharded build: every context indexes one shard of the reference records, the filters are
  OR-merged by the P2P kernel, every context finishes the same index.  On a one-GPU box the
  contexts share device 0 (the peer pointers are then plain device pointers); with more GPUs the
  same test spreads them over the devices, so peer access over NVLink is exercised too.
* index serialisation: save -> load into a fresh context -> identical index and identical results.

The yardstick is the oracle (bit-exact), as in test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import quirky_reference, sample_reads, to_soa

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return max(torch.cuda.device_count(), 1)


def _sharks(n, **kw):
    from shark_b200.engine import Shark
    nd = _n_devices()
    return [Shark(device=i % nd, max_reads_per_chunk=1 << 12, **kw) for i in range(n)]


def _close(sharks):
    for s in sharks:
        s.close()


def _check_index(sh, ref, n_records):
    assert sh.info.n_records == n_records
    assert sh.info.n_genes == ref.n_genes
    assert sh.info.n_set_bits == ref.n_set
    assert sh.info.tot_ids == ref.tot_ids
    pos, off, ids = sh.export_index()
    assert np.array_equal(pos, ref.pos)
    assert np.array_equal(off, ref.off)
    assert np.array_equal(ids, ref.ids)


@pytest.mark.parametrize("n_shards", [2, 3, 5])
@pytest.mark.parametrize("k,bf_bits", [(17, 1 << 30), (5, 1 << 20), (31, 1000003), (21, 3 << 28)])
def test_sharded_build_equals_oracle(n_shards, k, bf_bits):
    """One call (shk_index_build_sharded): n contexts, one host thread each."""
    from shark_b200.engine import Shark
    rng = np.random.default_rng(k * 100 + n_shards)
    genes = quirky_reference(rng)   # includes the nidx quirk: has_window flags must be merged too
    bases, rec_off = po.concat_records(genes)
    ref = po.Index(bases, rec_off, k, bf_bits)
    sharks = _sharks(n_shards, k=k, bf_bits=bf_bits)
    try:
        info = Shark.build_index_sharded(sharks, bases, rec_off)
        assert info.n_shards == n_shards
        for sh in sharks:
            _check_index(sh, ref, len(genes))
    finally:
        _close(sharks)


def test_sharded_protocol_step_by_step_and_classification(tmp_path):
    """The protocol call by call (what one-process-per-GPU drivers do, dist_index.build_index_sharded),
    with the anchor-and-extend structures forced on (they are derived from the gathered window array),
    followed by read classification on every context against the oracle."""
    from shark_b200 import capi
    k, bf_bits, c = 17, 1 << 30, 0.6
    rng = np.random.default_rng(77)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    ref = po.Index(bases, rec_off, k, bf_bits)
    texts = sample_reads(rng, genes, 3000, 100, paired=True)
    seq, off = to_soa(texts)
    cnt0, ar0, ag0 = ref.analyze(seq, off, c)
    n = 3
    sharks = _sharks(n, k=k, bf_bits=bf_bits, c=c, extend=True)
    try:
        mems = [sh.shard_begin(bases, rec_off, i, n) for i, sh in enumerate(sharks)]
        assert all(m.shard == i and m.pid == os.getpid() for i, m in enumerate(mems))
        seen = []
        for sh in sharks:
            arr = (capi.ShardMem * n)()
            for j in range(n):
                arr[j] = sh.shard_open(mems[j])
            seen.append(arr)
        # out-of-order calls fail loudly
        with pytest.raises(capi.SharkError) as ei:
            sharks[0].shard_merge(2, seen[0])
        assert ei.value.code == -3
        with pytest.raises(capi.SharkError):
            sharks[0].shard_rank()
        for i, sh in enumerate(sharks):
            sh.shard_merge(1, seen[i])
        for i, sh in enumerate(sharks):
            sh.shard_merge(2, seen[i])
            sh.shard_rank()
        for i, sh in enumerate(sharks):
            info = sh.shard_finish(seen[i])
            assert info.extend == 1 and info.n_shards == n
        for i, sh in enumerate(sharks):
            for j in range(n):
                sh.shard_close(seen[i][j])
            sh.shard_end()
        for sh in sharks:
            _check_index(sh, ref, len(genes))
            keep, ar, ag, stats = sh.analyze(seq, off)
            assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
            assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
            assert stats["n_extended"] > 0
    finally:
        _close(sharks)


def test_sharded_build_degenerate_inputs():
    """More shards than records, an empty reference, one shard."""
    from shark_b200.engine import Shark
    k, bf_bits = 11, 1 << 22
    rng = np.random.default_rng(5)
    for genes, n in (([b"ACGTTGCAAGGCTTAACCGGATATCG", b"", b"NNNNNNNNNNNNNNNN"], 5), ([], 2), ([b""], 3),
                     (quirky_reference(rng), 1)):
        bases, rec_off = po.concat_records(genes)
        ref = po.Index(bases, rec_off, k, bf_bits)
        sharks = _sharks(n, k=k, bf_bits=bf_bits)
        try:
            Shark.build_index_sharded(sharks, bases, rec_off)
            for sh in sharks:
                _check_index(sh, ref, len(genes))
        finally:
            _close(sharks)


def test_shard_limits():
    from shark_b200 import capi
    sharks = _sharks(1, k=11, bf_bits=1 << 20)
    try:
        bases, rec_off = po.concat_records([b"ACGTACGTACGTAAC"])
        with pytest.raises(capi.SharkError) as ei:
            sharks[0].shard_begin(bases, rec_off, 0, 17)
        assert ei.value.code == -1
        with pytest.raises(capi.SharkError):
            sharks[0].shard_begin(bases, rec_off, 2, 2)
    finally:
        _close(sharks)


@pytest.mark.parametrize("extend", [False, True], ids=["lookup", "extend"])
def test_index_save_load_roundtrip(tmp_path, extend):
    from shark_b200 import capi
    from shark_b200.engine import Shark
    k, bf_bits, c = 21, 1 << 28, 0.5
    rng = np.random.default_rng(9)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    ref = po.Index(bases, rec_off, k, bf_bits)
    texts = sample_reads(rng, genes, 2000, 120)
    seq, off = to_soa(texts)
    cnt0, ar0, ag0 = ref.analyze(seq, off, c)
    path = str(tmp_path / "idx.shk")
    with Shark(k=k, bf_bits=bf_bits, c=c, extend=extend, max_reads_per_chunk=1 << 12) as a:
        a.build_index(bases, rec_off)
        a.save_index(path)
        views = a.index_views()
        assert os.path.getsize(path) > sum(views.bytes)
    with Shark(k=k, bf_bits=bf_bits, c=c, max_reads_per_chunk=1 << 12) as b:
        info = b.load_index(path)
        assert info.extend == int(extend)
        _check_index(b, ref, len(genes))
        keep, ar, ag, _ = b.analyze(seq, off)
        assert np.array_equal(ar, ar0) and np.array_equal(ag, ag0)
        assert np.array_equal(keep, (cnt0 > 0).astype(np.uint8))
    # mismatches fail loudly and leave no index behind
    with Shark(k=k - 2, bf_bits=bf_bits, max_reads_per_chunk=1 << 12) as w:
        with pytest.raises(capi.SharkError) as ei:
            w.load_index(path)
        assert ei.value.code == -1 and "another k" in str(ei.value)
    with Shark(k=k, bf_bits=bf_bits << 1, max_reads_per_chunk=1 << 12) as w:
        with pytest.raises(capi.SharkError):
            w.load_index(path)
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x40
    bad = str(tmp_path / "bad.shk")
    open(bad, "wb").write(raw)
    trunc = str(tmp_path / "trunc.shk")
    open(trunc, "wb").write(raw[: len(raw) // 3])
    with Shark(k=k, bf_bits=bf_bits, max_reads_per_chunk=1 << 12) as w:
        with pytest.raises(capi.SharkError) as ei:
            w.load_index(bad)
        assert "checksum" in str(ei.value)
        with pytest.raises(capi.SharkError):
            w.load_index(trunc)
        with pytest.raises(capi.SharkError):
            w.load_index(str(tmp_path / "missing.shk"))
        with pytest.raises(capi.SharkError):   # no index in this context
            w.analyze(seq, off)
    with Shark(k=k, bf_bits=bf_bits, max_reads_per_chunk=1 << 12) as w:
        with pytest.raises(capi.SharkError):   # nothing to save yet
            w.save_index(str(tmp_path / "none.shk"))


def test_build_ms_is_device_time():
    """build_ms is the sum of CUDA-event segments (no allocation gaps); build_wall_ms is the host clock."""
    from shark_b200.engine import Shark
    rng = np.random.default_rng(3)
    genes = quirky_reference(rng)
    bases, rec_off = po.concat_records(genes)
    with Shark(k=17, bf_bits=1 << 30, max_reads_per_chunk=1 << 12) as sh:
        sh.build_index(bases, rec_off)
        info = sh.build_index(bases, rec_off)
        assert 0 < info.build_ms <= info.build_wall_ms * 1.05
        assert info.n_shards == 1

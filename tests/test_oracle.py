"""Pins the oracle (oracle/shark_oracle.c + oracle/shark_text.py) against
 (1) the known answers of SURVEY.md App. B.3 (from the reference's kmer_utils.hpp),
 (2) the reference's own example truth files,
 (3) golden outputs of the unmodified reference binary (tests/golden/make_golden.py).
CPU only."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, edge_cases, example_cases, flags_to_kwargs, gz_read, md5, stage_edge, stage_example
from oracle import pyoracle as po
from oracle import shark_text as st

KAT = [  # canonical k-mer, XXH64, % 2^33, % 3*2^33   
    (0x0000000000000000, 0x34C96ACDCADB1BBB, 7698324411, 16288259003),
    (0x0000000000000001, 0x9F29CB17A2A49995, 7023663509, 15613598101),
    (0x0000000000000002, 0xEAC73E4044E82DB0, 1156066736, 18335935920),
    (0x0000000000000003, 0x87B8166DA7EC4841, 7112247361, 15702181953),
    (0x0123456789ABCDEF, 0xEA3C52081E9843EC, 513295340, 513295340),
    (0x3FFFFFFFFFFFFFFF, 0xCC4E8923C52E58A0, 7603116192, 7603116192),
    (0x00000003FFFFFFFF, 0x31EB3411F5FEF9D8, 8422095320, 25601964504),
]


@pytest.mark.parametrize("kmer,h,m1,m3", KAT)
def test_xxh64_known_answers(kmer, h, m1, m3):
    assert po.xxh64_u64(kmer) == h
    assert h % (1 << 33) == m1
    assert h % (3 << 33) == m3


def test_kmer_codec_known_answers():
    assert po.build_kmer("ACGT" * 8 + "AC", 0, 17) == (0x6C6C6C6C, 17)
    assert po.revcompl(0x6C6C6C6C, 17) == 0x31B1B1B1B
    assert po.build_kmer("ACGNNACGTTGCATGCAAGGTTCCAAGG", 0, 17) == (0x6F9390AF, 22)
    assert po.build_kmer("ACGTNACGT", 0, 5) == (-1, 9)
    L = po.lib()
    assert L.shko_lsappend(0x3FFFFFFFF, 2, 17) == 0x3FFFFFFFE
    assert L.shko_rsprepend(0x3FFFFFFFF, 1, 17) == 0x1FFFFFFFF


def test_enumerate_matches_bruteforce():
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"ACGTacgtNn-", dtype=np.uint8)
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    for k in (1, 3, 5, 17, 31):
        for _ in range(40):
            n = int(rng.integers(0, 120))
            s = alphabet[rng.choice(len(alphabet), n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .01, .01])].tobytes()
            exp = []
            for e in range(k - 1, n):
                w = s[e - k + 1:e + 1]
                if all(c in code for c in w):
                    f = 0
                    for c in w:
                        f = (f << 2) | code[c]
                    r = 0
                    for c in reversed(w):
                        r = (r << 2) | (3 - code[c])
                    exp.append((min(f, r), e))
            got = po.enumerate_kmers(s, k)
            if got is None:
                assert n >= k and not exp
            else:
                assert list(zip(got[0].tolist(), got[1].tolist())) == exp


def test_example_truth_files(tmp_path):
    f = stage_example(tmp_path)
    ssv, o1, o2 = st.run_shark(f["ENSG00000277117.fa"], f["sample_1.fq"], f["sample_2.fq"])
    d = os.path.join(GOLDEN, "example")
    assert ssv == gz_read(os.path.join(d, "ENSG00000277117.truth.ssv.gz"))
    assert o1 == gz_read(os.path.join(d, "sharked.sample_1.truth.fq.gz"))
    assert o2 == gz_read(os.path.join(d, "sharked.sample_2.truth.fq.gz"))


@pytest.mark.parametrize("case", sorted(example_cases()["cases"]))
def test_example_flag_variants(tmp_path, case):
    info = example_cases()["cases"][case]
    f = stage_example(tmp_path)
    ssv, o1, o2 = st.run_shark(f["ENSG00000277117.fa"], f["sample_1.fq"], f["sample_2.fq"] if info["paired"] else None,
                               **flags_to_kwargs(info["flags"]))
    assert ssv.count(b"\n") == info["ssv_lines"]
    assert md5(ssv) == info["ssv_md5"]
    assert md5(o1) == info["o1_md5"]
    if info["paired"]:
        assert md5(o2) == info["o2_md5"]


def _edge_params():
    return [(s, c) for s, cs in sorted(edge_cases().items()) for c in sorted(cs)]


@pytest.mark.parametrize("scenario,case", _edge_params())
def test_edge_goldens(tmp_path, scenario, case):
    info = edge_cases()[scenario][case]
    f = stage_edge(tmp_path, scenario)
    ssv, o1, o2 = st.run_shark(f["ref.fa"], f["r1.fq"], f.get("r2.fq") if info["paired"] else None,
                               **flags_to_kwargs(info["flags"]))
    d = os.path.join(GOLDEN, "edge", scenario)
    assert ssv == gz_read(os.path.join(d, case + ".ssv.gz"))
    assert o1 == gz_read(os.path.join(d, case + ".o1.fq.gz"))
    if info["paired"]:
        assert o2 == gz_read(os.path.join(d, case + ".o2.fq.gz"))


def test_index_structure_invariants():
    rng = np.random.default_rng(3)
    seqs = [bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300)]) for _ in range(20)]
    seqs[5] = seqs[4]
    bases, off = po.concat_records(seqs)
    ix = po.Index(bases, off, 11, 1 << 33)
    assert ix.n_genes == 20
    assert np.all(np.diff(ix.pos.astype(np.int64)) > 0)
    assert ix.off[0] == 0 and ix.off[-1] == ix.tot_ids
    assert np.all(np.diff(ix.off.astype(np.int64)) >= 1)  # no set bit has an empty list (App. A.4)
    for r in range(0, ix.n_set, 97):
        lst = ix.ids[ix.off[r]:ix.off[r + 1]]
        assert np.all(np.diff(lst.astype(np.int64)) > 0)
    # every k-mer of gene 4 lists both 4 and 5
    canon, _ = po.enumerate_kmers(seqs[4], 11)
    rank, begin, ln = ix.probe(canon)
    assert np.all(rank >= 0)
    for b, l in zip(begin, ln):
        lst = ix.ids[b:b + l].tolist()
        assert 4 in lst and 5 in lst

"""Host ingest (shark_b200/csrc/host/ingest.hpp): the block scanner must yield exactly the kseq_read()
outcome stream of the record-at-a-time reader (fastx.hpp, itself pinned against the reference binary
by the CLI goldens) on well-formed and on hostile inputs, for any block size."""
import gzip
import os
import random
import subprocess

import pytest

from helpers import ROOT

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shark_b200", "csrc", "host")


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "host_tools")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-pthread", os.path.join(HOST, "host_tools.cpp"), "-o", out, "-lz"])
    return out


def _check(tool, path, block):
    r = subprocess.run([tool, "scan-check", path, str(block)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (path, block, r.stdout, r.stderr)
    return int(r.stdout.split()[1])


def _rec(rng, L=None, qual_delta=0, crlf=False, comment=False, plus_text=False, wrap=0):
    L = rng.randint(1, 120) if L is None else L
    name = "r%d" % rng.randint(0, 10 ** 6)
    if comment:
        name += rng.choice([" c=1", "\tx y", " "])
    seq = "".join(rng.choice("ACGTNacgtn") for _ in range(L))
    qual = "".join(chr(rng.randint(33, 74)) for _ in range(max(0, L + qual_delta)))
    nl = "\r\n" if crlf else "\n"
    if wrap:
        seq = nl.join(seq[i:i + wrap] for i in range(0, L, wrap))
        qual = nl.join(qual[i:i + wrap] for i in range(0, len(qual), wrap))
    return "@" + name + nl + seq + nl + "+" + (name if plus_text else "") + nl + qual + nl


def test_strict_fastq_all_block_sizes(tool, tmp_path):
    rng = random.Random(1)
    p = str(tmp_path / "a.fq")
    open(p, "w").write("".join(_rec(rng, comment=rng.random() < 0.3, plus_text=rng.random() < 0.2) for _ in range(3000)))
    n = None
    for block in (1 << 23, 4096, 257, 64, 17, 1):
        m = _check(tool, p, block)
        assert n is None or m == n
        n = m
    assert n == 3000 + 3  # records + the sticky end-of-file outcomes the checker consumes


def test_quality_chars_that_look_like_headers(tool, tmp_path):
    rng = random.Random(2)
    recs = []
    for i in range(2000):
        L = rng.randint(5, 60)
        q = "".join(rng.choice("@+>I5") for _ in range(L))   # '@' and '+' at line starts of the quality
        recs.append("@n%d\n%s\n+\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(L)), q))
    p = str(tmp_path / "b.fq")
    open(p, "w").write("".join(recs))
    for block in (1 << 20, 100, 7):
        assert _check(tool, p, block) == 2003


@pytest.mark.parametrize("seed", range(12))
def test_hostile_inputs_match_the_record_reader(tool, tmp_path, seed):
    rng = random.Random(100 + seed)
    parts = []
    for _ in range(400):
        u = rng.random()
        if u < 0.45:
            parts.append(_rec(rng))
        elif u < 0.55:
            parts.append(_rec(rng, crlf=True))
        elif u < 0.63:
            parts.append(_rec(rng, qual_delta=rng.choice([-3, -1, 1, 4])))          # kseq returns -2 / resynchronises
        elif u < 0.70:
            parts.append(_rec(rng, L=rng.randint(30, 200), wrap=rng.choice([10, 60])))  # wrapped FASTQ
        elif u < 0.78:
            parts.append(">fa%d desc\n%s\n" % (rng.randint(0, 99), "\n".join("ACGTAC" * rng.randint(0, 9) for _ in range(rng.randint(0, 3)))))
        elif u < 0.84:
            parts.append(rng.choice(["\n", "\n\n", "garbage line\n", "+\n", "@\n", "@x\n\n+\n\n", " \t\n"]))
        elif u < 0.90:
            parts.append("@nul%d\nAC\x00GT\n+\nII\x00II\n" % rng.randint(0, 9))
        elif u < 0.95:
            parts.append("@big\n" + "ACGT" * rng.randint(500, 3000) + "\n+\n" + "IIII" * rng.randint(500, 3000) + "\n")
        else:
            parts.append(_rec(rng)[: rng.randint(1, 30)])                               # truncated record in the middle
    data = "".join(parts)
    if seed % 3 == 0:
        data = data.rstrip("\n")                                                         # no newline at end of file
    if seed % 4 == 1:
        data = data[: len(data) - rng.randint(1, 50)]                                    # file cut inside the last record
    p = str(tmp_path / "h.fq")
    open(p, "wb").write(data.encode("latin-1"))
    for block in (1 << 20, 1024, 61, 3):
        _check(tool, p, block)
    pz = p + ".gz"
    with gzip.open(pz, "wb") as f:
        f.write(data.encode("latin-1"))
    for block in (1 << 20, 500):
        _check(tool, pz, block)


def test_empty_and_missing_files(tool, tmp_path):
    p = str(tmp_path / "e.fq")
    open(p, "w").close()
    assert _check(tool, p, 100) == 3
    r = subprocess.run([tool, "scan-check", str(tmp_path / "nope.fq")], capture_output=True, text=True)
    assert r.returncode == 0 and "OPEN_FAILED" in r.stdout


def test_host_pack_matches_numpy_restatement():
    """shk_host_pack (pure host code of the split upload): validity after the masking rule
    (FastqSplitter.hpp:104-109: `seq[i] -= 64` where `(char)qual[i] < (char)(q+33)`; to_int kmer_utils.hpp:29-41)
    and 2-bit codes, against a numpy restatement - AVX2 and scalar paths, single and parallel, ragged sizes,
    every byte value, `char` wrap-around of q."""
    import subprocess
    import sys
    import numpy as np
    from shark_b200 import capi

    def ref_pack(seq, qual, q):
        ch = seq.astype(np.int64)
        if qual is not None and (q & 0xFF) != 0:
            mq = np.int64(np.int8(np.uint8(((q & 0xFF) + 33) & 0xFF)))
            ch = np.where(qual.astype(np.int8).astype(np.int64) < mq, (ch - 64) & 0xFF, ch)
        ok = np.isin(ch | 0x20, [0x61, 0x63, 0x67, 0x74])
        code = np.where(ok, (ch >> 1) & 3, 0).astype(np.uint64)
        g = (len(seq) + 31) // 32
        pad = g * 32 - len(seq)
        ok = np.concatenate([ok, np.zeros(pad, bool)]).reshape(g, 32).astype(np.uint64)
        code = np.concatenate([code, np.zeros(pad, np.uint64)]).reshape(g, 32)
        sh = np.arange(32, dtype=np.uint64)
        return (code << (2 * sh)).sum(1).astype(np.uint64), (ok << sh).sum(1).astype(np.uint32)

    isa, threads = capi.host_pack_info()
    assert isa in ("avx512", "avx2", "scalar") and threads >= 1
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 1000, (1 << 19) + 77):
        seq = rng.integers(0, 256, n, dtype=np.uint8)
        m = rng.random(n) < 0.7
        seq[m] = np.frombuffer(b"ACGTacgtNn", np.uint8)[rng.integers(0, 10, int(m.sum()))]
        qual = rng.integers(0, 256, n, dtype=np.uint8)
        for q in (0, 20, 94, 95, 200, 223):
            c0, v0 = ref_pack(seq, qual, q)
            for par in (False, True):
                c, v = capi.host_pack(seq, qual, q, par)
                assert np.array_equal(c, c0) and np.array_equal(v, v0), (n, q, par)
    # the other instruction sets, each in a fresh process (the choice is made once): byte-identical output
    code = ("import numpy as np; from shark_b200 import capi; print(capi.host_pack_info()[0]);"
            "rng = np.random.default_rng(3); s = rng.integers(0, 256, 100003, dtype=np.uint8);"
            "m = rng.random(len(s)) < 0.7; s[m] = np.frombuffer(b'ACGTacgtNn', np.uint8)[rng.integers(0, 10, int(m.sum()))];"
            "q = rng.integers(0, 256, len(s), dtype=np.uint8);"
            "import hashlib; h = hashlib.md5();"
            "[h.update(a.tobytes()) for mq in (0, 20, 95) for a in capi.host_pack(s, q, mq)]; print(h.hexdigest())")
    outs = {}
    for name, extra in (("default", {}), ("avx512", {"SHK_PACK_AVX512": "1"}), ("avx2", {"SHK_PACK_AVX2": "1"}),
                        ("scalar", {"SHK_PACK_SCALAR": "1"})):
        env = dict(os.environ, PYTHONPATH=ROOT, **extra)
        isa_used, digest = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                          check=True).stdout.decode().split()
        outs[name] = digest
        if name == "scalar":
            assert isa_used == "scalar"
        if name == "avx2":
            assert isa_used in ("avx2", "scalar")
    assert outs["default"] == outs["avx512"] == outs["avx2"] == outs["scalar"], outs

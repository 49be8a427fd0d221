"""Host ingest (shark_b200/csrc/host/ingest.hpp, fastpipe.hpp, pipeline.hpp): the block scanner and the parallel
scanner of mapped files must yield exactly the kseq_read() outcome stream of the record-at-a-time reader
(tests/host_tools/fastx.hpp, itself pinned against the reference binary by the CLI goldens) on well-formed and on
hostile inputs, for any block / segment size; the batcher and the writer must produce the packed chunks and
the output bytes that the text-level oracle gives for the same results."""
import gzip
import os
import random
import subprocess

import pytest

from helpers import ROOT

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shark_b200", "csrc", "host")
TOOLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_tools")


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "host_tools")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-pthread", "-I", HOST, os.path.join(TOOLS, "host_tools.cpp"),
                           "-o", out, "-lz"])
    return out


def _check(tool, path, block):
    r = subprocess.run([tool, "scan-check", path, str(block)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (path, block, r.stdout, r.stderr)
    n = int(r.stdout.split()[1])
    if not path.endswith(".gz"):
        # the same file through the parallel scanner of the mapped file: segment size = the block size (tiny
        # segments put a guessed record start on every line), outcome by outcome and with bulk takes in between
        for seg, bulk in ((max(block, 16), 0), (max(block, 16), 37), (1 << 22, 5)):
            r = subprocess.run([tool, "pscan-check", path, str(seg), str(bulk)], capture_output=True, text=True, timeout=120,
                               env=dict(os.environ, SHK_HOST_THREADS="4"))
            assert r.returncode == 0 and r.stdout.startswith("OK"), (path, seg, bulk, r.stdout, r.stderr)
            assert int(r.stdout.split()[1]) >= n - 2   # (the checkers stop after three end-of-file outcomes; bulk takes skip none)
    return n


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


# ---- blocked gzip (BGZF): parallel inflate must yield gzread's bytes and gzread's errors --------------------
def _bgzf_block(data, extra_subfields=b"", level=6):
    import struct
    import zlib
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    payload = co.compress(data) + co.flush()
    extra = b"BC" + struct.pack("<H", 2) + struct.pack("<H", 0)  # BSIZE patched below
    extra = extra_subfields + extra
    hdr_len = 12 + len(extra)
    bsize = hdr_len + len(payload) + 8 - 1
    assert bsize < 65536
    extra = extra_subfields + b"BC" + struct.pack("<H", 2) + struct.pack("<H", bsize)
    hdr = b"\x1f\x8b\x08\x04" + b"\x00\x00\x00\x00" + b"\x00\xff" + struct.pack("<H", len(extra)) + extra
    return hdr + payload + struct.pack("<I", zlib.crc32(data) & 0xFFFFFFFF) + struct.pack("<I", len(data))


def _bgzf(data, rng, eof_marker=True, max_block=65280, subfields=False):
    out, i = [], 0
    while i < len(data):
        n = rng.randint(1, max_block) if rng.random() < 0.3 else max_block
        extra = b"XY" + b"\x03\x00abc" if subfields and rng.random() < 0.5 else b""
        out.append(_bgzf_block(data[i:i + n], extra, level=rng.choice([1, 6])))
        i += n
        if rng.random() < 0.02:
            out.append(_bgzf_block(b""))  # empty blocks may appear anywhere
    if eof_marker:
        out.append(_bgzf_block(b""))
    return out


def _fastq_bytes(rng, n_rec, hostile=False):
    parts = []
    for _ in range(n_rec):
        u = rng.random()
        if hostile and u < 0.1:
            parts.append(_rec(rng, qual_delta=rng.choice([-2, 1])))
        elif hostile and u < 0.2:
            parts.append(_rec(rng, L=rng.randint(30, 200), wrap=rng.choice([10, 60])))
        elif hostile and u < 0.25:
            parts.append(rng.choice(["\n", "garbage line\n", "@x\n\n+\n\n"]))
        else:
            parts.append(_rec(rng, comment=rng.random() < 0.2))
    return "".join(parts).encode("latin-1")


@pytest.mark.parametrize("threads", ["1", "5"])
def test_bgzf_parallel_inflate_matches_gzread(tool, tmp_path, threads, monkeypatch):
    monkeypatch.setenv("SHK_INGEST_THREADS", threads)
    rng = random.Random(7)
    data = _fastq_bytes(rng, 12000)
    assert len(data) > 1500000
    cases = {
        "plain": b"".join(_bgzf(data, rng)),
        "no_eof_marker": b"".join(_bgzf(data, rng, eof_marker=False)),
        "small_blocks": b"".join(_bgzf(data[:300000], rng, max_block=700)),
        "extra_subfields": b"".join(_bgzf(data[:400000], rng, subfields=True)),
        "hostile_text": b"".join(_bgzf(_fastq_bytes(rng, 3000, hostile=True), rng)),
        "empty_payload": b"".join(_bgzf(b"", rng)),
    }
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".fq.gz"))
        open(p, "wb").write(blob)
        assert gzip.decompress(blob) is not None  # a valid multi-member gzip file for any reader
        for block in (1 << 23, 1 << 20, 70001, 500):
            n = _check(tool, p, block)
        if name == "plain":
            assert n == 12000 + 3


def _dump(tool, path, block, hash_first):
    r = subprocess.run([tool, "scan-dump", path, str(block), str(hash_first)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("DUMP"), (path, block, r.stdout, r.stderr)
    f = dict(kv.split("=") for kv in r.stdout.split()[1:])
    return int(f["count"]), int(f["last"]), f["hash"]


def _fnv_records(recs):
    h = 0xCBF29CE484222325
    for rec in recs:
        for field in rec:
            for c in field:
                h = ((h ^ c) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.mark.parametrize("seed", range(6))
def test_bgzf_foreign_members_and_trailing_bytes_behave_like_gzread(tool, tmp_path, seed, monkeypatch):
    """Ordinary gzip members before / between / after the blocks, concatenated files, trailing garbage: valid input
    for gzread, so the outcome stream must equal the record reader's exactly (the parallel path hands over to zlib
    at the first member that is not a BGZF block)."""
    monkeypatch.setenv("SHK_INGEST_THREADS", "4")
    rng = random.Random(50 + seed)
    data = _fastq_bytes(rng, 6000)
    blocks = _bgzf(data, rng, max_block=rng.choice([65280, 9000]))
    whole = b"".join(blocks)
    k = rng.randrange(1, len(blocks) - 1)
    variants = {
        "trailing_garbage": whole + b"this is not gzip",
        "trailing_1f": whole + b"\x1f",
        "gzip_member_between": b"".join(blocks[:k]) + gzip.compress(b"@mid\nACGT\n+\nIIII\n") + b"".join(blocks[k:]),
        "gzip_member_first": gzip.compress(data[:5000]) + whole,
        "gzip_member_last": whole + gzip.compress(b"@last\nACGT\n+\nIIII\n"),
        "concatenated_files": whole + whole,
    }
    for name, blob in variants.items():
        p = str(tmp_path / ("%s_%d.fq.gz" % (name, seed)))
        open(p, "wb").write(blob)
        for block in (1 << 20, 3000):
            _check(tool, p, block)


@pytest.mark.parametrize("seed", range(6))
def test_bgzf_damaged_files_fail_cleanly(tool, tmp_path, seed, monkeypatch):
    """Truncated or corrupted blocks.  What gzread delivers before it reports such an error depends on its caller's
    buffer sizes (it drops the bytes of the failing call), so there is no byte-exact yardstick - the reference itself
    does not detect I/O errors.  Required here: every record of the undamaged prefix arrives
    intact and in order, nothing is invented, the stream ends in an error or end-of-file outcome, no crash or hang."""
    monkeypatch.setenv("SHK_INGEST_THREADS", "4")
    rng = random.Random(90 + seed)
    n_rec = 6000
    recs = []
    for i in range(n_rec):
        L = rng.randint(20, 150)
        recs.append((b"r%d" % i, "".join(rng.choice("ACGT") for _ in range(L)).encode(), bytes(rng.randint(33, 74) for _ in range(L))))
    data = b"".join(b"@" + n + b"\n" + s + b"\n+\n" + q + b"\n" for n, s, q in recs)
    ends = []
    pos = 0
    for n, s, q in recs:
        pos += len(n) + len(s) + len(q) + 6
        ends.append(pos)
    blocks = _bgzf(data, rng, max_block=rng.choice([65280, 9000]))
    whole = b"".join(blocks)
    k = rng.randrange(1, len(blocks) - 1)
    at = sum(len(b) for b in blocks[:k])
    good_bytes = sum(int.from_bytes(b[-4:], "little") for b in blocks[:k])
    bl = len(blocks[k])
    variants = {
        "trailing_magic_then_junk": (whole + b"\x1f\x8b\x08\x04 junk that is no member", len(data)),
        "cut_in_header": (whole[: at + rng.randint(1, 17)], good_bytes),
        "cut_in_payload": (whole[: at + 18 + rng.randint(1, max(2, bl - 30))], good_bytes),
        "cut_in_trailer": (whole[: at + bl - rng.randint(1, 7)], good_bytes),
        "bad_crc": (whole[: at + bl - 8] + b"\x00\x01\x02\x03" + whole[at + bl - 4:], good_bytes),
        "bad_isize": (whole[: at + bl - 4] + b"\x05\x00\x00\x00" + whole[at + bl:], good_bytes),
        "bad_payload": (whole[: at + 40] + bytes([whole[at + 40] ^ 0x55]) + whole[at + 41:], good_bytes),
        "bad_bsize": (whole[: at + 16] + b"\x10\x00" + whole[at + 18:], good_bytes),
    }
    import bisect
    for name, (blob, intact) in variants.items():
        p = str(tmp_path / ("%s_%d.fq.gz" % (name, seed)))
        open(p, "wb").write(blob)
        whole_records = bisect.bisect_right(ends, intact)   # records that lie entirely in the undamaged prefix
        sure = max(whole_records - 1, 0)   # the record that straddles the damage may arrive cut short
        for block in (1 << 20, 3000):
            count, last, h = _dump(tool, p, block, sure)
            assert last in (-1, -2, -3), (name, last)
            assert sure <= count <= n_rec, (name, block, count, whole_records)
            assert h == _fnv_records(recs[:sure]), (name, block, count)


def test_host_pack_against_the_oracle_masking_and_base_table():
    """The split upload must carry exactly what the oracle's (= the reference's) rules make of every byte: masking
    by shko_mask (FastqSplitter.hpp:104-109), validity and base by shko_base_code (to_int, kmer_utils.hpp:29-41)."""
    import numpy as np
    from oracle import pyoracle as po
    from shark_b200 import capi
    L = po.lib()
    code_of = np.array([L.shko_base_code(b) for b in range(256)], np.int64)   # -1 invalid, else A C G T = 0..3
    raw_of_code = np.array([0, 1, 3, 2])                                       # the packed code is (byte >> 1) & 3: A 0 C 1 T 2 G 3
    rng = np.random.default_rng(21)
    n = 20000
    seq = rng.integers(0, 256, n, dtype=np.uint8)
    m = rng.random(n) < 0.6
    seq[m] = np.frombuffer(b"ACGTacgtN", np.uint8)[rng.integers(0, 9, int(m.sum()))]
    qual = rng.integers(0, 256, n, dtype=np.uint8)
    sh = np.arange(32, dtype=np.uint64)
    for q in (0, 1, 20, 41, 94, 95, 127, 200, 222, 223, 255):
        masked = po.mask(seq, qual, q) if q else seq
        code = code_of[masked]
        ok = code >= 0
        want_codes = np.where(ok, raw_of_code[np.maximum(code, 0)], 0).astype(np.uint64)
        g = (n + 31) // 32
        pad = g * 32 - n
        want_v = (np.concatenate([ok, np.zeros(pad, bool)]).reshape(g, 32).astype(np.uint64) << sh).sum(1).astype(np.uint32)
        want_c = (np.concatenate([want_codes, np.zeros(pad, np.uint64)]).reshape(g, 32) << (2 * sh)).sum(1).astype(np.uint64)
        c, v = capi.host_pack(seq, qual, q, parallel=bool(q & 1))
        assert np.array_equal(v, want_v), q
        assert np.array_equal(c, want_c), q


# ---------------------------------------------------------------------------------------------------------
# batcher + writer (pipeline.hpp) against the text-level oracle: packed chunks and output bytes
# ---------------------------------------------------------------------------------------------------------
def _expected_pipe(path1, path2, q):
    """What host_tools pipe-check must print: the reference's ReadOutput (ReadOutput.hpp:37-50, previd reset per
    50 000-read batch) for the faked results `read i -> gene gA; every 7th (i % 7 == 3) dropped; i % 5 == 1 ->
    gA and gB`, plus the joined text of all reads for the packed chunks."""
    from oracle import shark_text as st
    pairs, bid = st.load_sample(path1, path2)
    seq, qual, off = st.join_reads(pairs, q > 0)
    ssv, o1, o2 = [], [], []
    prev, prev_batch = b"", 0
    for i, (a, b) in enumerate(pairs):
        if i % 7 == 3:
            continue
        if bid[i] != prev_batch:
            prev, prev_batch = b"", bid[i]
        name = st._cstr(a[0])
        for g in ((b"gA", b"gB") if i % 5 == 1 else (b"gA",)):
            ssv.append(name + b" " + g + b"\n")
        if prev != name:
            o1.append(b"@" + name + b"\n" + st._cstr(a[1]) + b"\n+\n" + st._cstr(a[2]) + b"\n")
            if b is not None:
                o2.append(b"@" + st._cstr(b[0]) + b"\n" + st._cstr(b[1]) + b"\n+\n" + st._cstr(b[2]) + b"\n")
        prev = name
    return b"".join(ssv), b"".join(o1), b"".join(o2), seq, qual, off, bid


def _fnv(h, data):
    for byte in data:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _run_pipe(tool, tmp_path, f1, f2, q, chunk, max_bytes, env_extra):
    o1, o2 = str(tmp_path / "o1.fq"), str(tmp_path / "o2.fq")
    for p in (o1, o2):
        if os.path.exists(p):
            os.unlink(p)
    cmd = [tool, "pipe-check", f1] + ([f2] if f2 else []) + ["--qual", str(q), "--chunk", str(chunk), "--bytes", str(max_bytes)]
    ssv_path = str(tmp_path / "o.ssv")
    with open(ssv_path, "wb") as fo:
        r = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, timeout=300, env=dict(os.environ, OUT1=o1, OUT2=o2, **env_extra))
    assert r.returncode == 0, r.stderr.decode()[-500:]
    chunks = []
    for ln in r.stderr.decode().splitlines():
        if ln.startswith("CHUNK"):
            kv = dict(x.split("=") for x in ln.split()[1:])
            chunks.append((int(kv["n"]), int(kv["bytes"]), int(kv["bulk"]), [int(x) for x in kv["batches"].split(",") if x],
                           int(kv["hash"], 16)))
    return open(ssv_path, "rb").read(), open(o1, "rb").read(), (open(o2, "rb").read() if f2 else b""), chunks


def _check_pipe(tool, tmp_path, f1, f2, q, chunk, max_bytes=10 ** 9, env_extra=None, want_bulk=None):
    import numpy as np
    from shark_b200 import capi
    ssv0, a0, b0, seq, qual, off, bid = _expected_pipe(f1, f2, q)
    ssv, a, b, chunks = _run_pipe(tool, tmp_path, f1, f2, q, chunk, max_bytes, env_extra or {})
    assert ssv == ssv0
    assert a == a0
    if f2:
        assert b == b0
    # the chunks cover the reads in order, respect both bounds, mark the batch starts, and carry the packed text
    r = 0
    for n, nbytes, bulk, batches, h in chunks:
        assert 0 < n <= chunk and nbytes <= max_bytes
        lo, hi = int(off[r]), int(off[r + n])
        assert nbytes == hi - lo
        assert batches == [i for i in range(n) if r + i == 0 or bid[r + i] != bid[r + i - 1]]
        codes, valid = capi.host_pack(seq[lo:hi], qual[lo:hi] if q else None, q)
        o32 = (off[r:r + n + 1] - off[r]).astype(np.uint32)
        assert h == _fnv(_fnv(_fnv(0xCBF29CE484222325, o32.tobytes()), codes.tobytes()), valid.tobytes())
        r += n
    assert r == len(off) - 1
    if want_bulk is not None:
        assert (sum(c[0] for c in chunks if c[2]) > 0) == want_bulk
    return chunks


def _write_fastq(path, rng, n, name_run=3, Lmin=18, Lmax=40, seed_names=0):
    with open(path, "w") as f:
        for i in range(n):
            L = rng.randint(Lmin, Lmax)
            f.write("@r%d%s\n%s\n+\n%s\n" % (seed_names + i // name_run, " x" if i % 11 == 0 else "",
                                               "".join(rng.choice("ACGTN") for _ in range(L)),
                                               "".join(chr(rng.randint(33, 74)) for _ in range(L))))


@pytest.mark.parametrize("paired,q", [(False, 0), (True, 0), (True, 25), (False, 30)])
def test_pipeline_bulk_chunks_batches_and_dedup(tool, tmp_path, paired, q):
    """120 000 short reads whose names repeat in runs of three (ReadOutput's consecutive-name dedup, also across
    range, chunk and batch boundaries), through chunk sizes that do and do not divide the 50 000-read batch, with a
    byte bound that cuts chunks short, in parallel and with one thread."""
    rng = random.Random(7 + q + paired)
    f1, f2 = str(tmp_path / "a_1.fq"), (str(tmp_path / "a_2.fq") if paired else None)
    _write_fastq(f1, rng, 120000)
    if paired:
        _write_fastq(f2, rng, 120000, name_run=2, seed_names=10 ** 6)
    for chunk, max_bytes, env in ((1000000, 10 ** 9, {}), (30000, 10 ** 9, {"SHK_HOST_THREADS": "4"}), (50000, 10 ** 9, {"SHK_HOST_THREADS": "1"}),
                                  (17777, 700000, {"SHK_HOST_THREADS": "3", "SHK_OUT": "map", "PREFAULT": "1"}), (40000, 10 ** 9, {"SHK_NO_BULK": "1", "SHK_OUT": "write"})):
        chunks = _check_pipe(tool, tmp_path, f1, f2, q, chunk, max_bytes, env, want_bulk="SHK_NO_BULK" not in env)
        assert len(chunks) >= 120000 // chunk


@pytest.mark.parametrize("seed", range(6))
def test_pipeline_hostile_inputs(tool, tmp_path, seed):
    """Failed reads, NUL bytes, FASTA records, wrapped lines, truncated tails, files of different lengths: the
    exact path takes over wherever the bulk path may not run, output and packed text stay those of the oracle."""
    rng = random.Random(300 + seed)

    def hostile(n):
        parts = []
        for _ in range(n):
            u = rng.random()
            if u < 0.80:
                parts.append(_rec(rng, comment=rng.random() < 0.2))
            elif u < 0.84:
                parts.append(_rec(rng, crlf=True))
            elif u < 0.88:
                parts.append(_rec(rng, qual_delta=rng.choice([-2, 3])))
            elif u < 0.91:
                parts.append(_rec(rng, L=rng.randint(30, 90), wrap=20))
            elif u < 0.94:
                parts.append(">fa%d d\nACGTTGCA\nACG\n" % rng.randint(0, 9))
            elif u < 0.97:
                parts.append("@nul%d\nAC\x00GT\n+\nII\x00II\n" % rng.randint(0, 9))
            else:
                parts.append(rng.choice(["\n", "junk\n", "@\n"]))
        return "".join(parts)

    f1, f2 = str(tmp_path / "h_1.fq"), str(tmp_path / "h_2.fq")
    d1, d2 = hostile(3000), hostile(2500 + 200 * seed)
    if seed % 2:
        d1 = d1[: len(d1) - rng.randint(1, 40)]
    open(f1, "wb").write(d1.encode("latin-1"))
    open(f2, "wb").write(d2.encode("latin-1"))
    for q in (0, 20):
        _check_pipe(tool, tmp_path, f1, None, q, 700, env_extra={"SHK_SCAN_SEGMENT": "4096", "SHK_HOST_THREADS": "4", "SHK_OUT": "map", "PREFAULT": "1"})
        _check_pipe(tool, tmp_path, f1, f2, q, 1000, env_extra={"SHK_SCAN_SEGMENT": "1000", "SHK_HOST_THREADS": "3"})


def test_pipeline_read_longer_than_chunk_fails_cleanly(tool, tmp_path):
    """a read that cannot fit a chunk must end the run with a message, not overflow the 32-bit offsets."""
    f1 = str(tmp_path / "long.fq")
    open(f1, "w").write("@a\nACGT\n+\nIIII\n@b\n" + "ACGT" * 500 + "\n+\n" + "IIII" * 500 + "\n@c\nAC\n+\nII\n")
    r = subprocess.run([tool, "pipe-check", f1, "--chunk", "100", "--bytes", "1000"], capture_output=True, timeout=60)
    assert r.returncode == 3 and b"exceeds the chunk capacity" in r.stderr

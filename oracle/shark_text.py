"""Text-level restatement of the reference's host stages around the hot path: kseq record
parsing, FastaSplitter/FastqSplitter batching and ReadOutput formatting.

TEST INFRASTRUCTURE ONLY (see shark_oracle.c header).  Small inputs only - pure Python loops.
Citations are relative to /root/reference.
"""
import gzip

import numpy as np

from . import pyoracle

_SPACE = b" \t\n\v\f\r"


def _read_all(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] == b"\x1f\x8b":  # gzread is transparent for plain files (kseq over zlib, main.cpp:32)
        data = gzip.decompress(data)
    return data


class _Stream:
    """kstream_t over an in-memory buffer: ks_getc (kseq.h:67-79), ks_getuntil2 (kseq.h:93-144)."""

    def __init__(self, data):
        self.d, self.i, self.n = data, 0, len(data)

    def getc(self):
        if self.i >= self.n:
            return -1
        c = self.d[self.i]
        self.i += 1
        return c

    def getuntil(self, line, buf, append=False):
        """-> (return value, delimiter found or 0)"""
        if not append:
            del buf[:]
        if self.i >= self.n:
            return -1, 0  # !gotany && eof
        j = self.i
        if line:
            j = self.d.find(b"\n", self.i)
            j = self.n if j < 0 else j
        else:
            while j < self.n and self.d[j] not in _SPACE:
                j += 1
        buf += self.d[self.i:j]
        dret = self.d[j] if j < self.n else 0
        self.i = min(j + 1, self.n)
        if line and len(buf) > 1 and buf[-1] == 0x0D:  # kseq.h:141
            buf.pop()
        return len(buf), dret


class FastxReader:
    """kseq_t over an in-memory buffer.  read() = kseq_read (kseq.h:177-218): returns
    (name, seq, qual) bytes (qual b'' for FASTA records) or None for ANY negative return.
    Like the reference's, the reader stays usable after a failure: after a length mismatch
    (-2, kseq.h:216) the next read() resynchronises at the next '>'/'@' byte."""

    HDR = (0x3E, 0x40)  # '>' '@'

    def __init__(self, data):
        self.ks = _Stream(data)
        self.last_char = 0

    def read(self):
        ks, HDR = self.ks, self.HDR
        if self.last_char == 0:  # kseq.h:181-185: jump to the next '>' or '@' anywhere
            c = ks.getc()
            while c >= 0 and c not in HDR:
                c = ks.getc()
            if c < 0:
                return None
            self.last_char = c
        name, comment, seq, qual = bytearray(), bytearray(), bytearray(), bytearray()
        r, c = ks.getuntil(False, name)  # kseq.h:188
        if r < 0:
            return None
        if c != 0x0A:
            ks.getuntil(True, comment)  # kseq.h:189 (dropped by Shark)
        c = ks.getc()
        while c >= 0 and c not in HDR and c != 0x2B:  # kseq.h:194-198
            if c != 0x0A:
                seq.append(c)
                ks.getuntil(True, seq, append=True)
            c = ks.getc()
        if c in HDR:
            self.last_char = c
        if c != 0x2B:  # FASTA record (kseq.h:205)
            return bytes(name), bytes(seq), b""
        c = ks.getc()
        while c >= 0 and c != 0x0A:  # kseq.h:210
            c = ks.getc()
        if c == -1:
            return None  # -2, at EOF
        while True:  # kseq.h:212
            r, _ = ks.getuntil(True, qual, append=True)
            if not (r >= 0 and len(qual) < len(seq)):
                break
        self.last_char = 0
        if len(seq) != len(qual):
            return None  # -2, stream continues
        return bytes(name), bytes(seq), bytes(qual)


def parse_fastx(data):
    """All records up to the first negative kseq_read (the `while (kseq_read(seq) >= 0)` loop
    of main.cpp:159)."""
    rd, recs = FastxReader(data), []
    while True:
        r = rd.read()
        if r is None:
            return recs
        recs.append(r)


def batches(data1, data2=None, maxnum=50000):
    """FastqSplitter::operator() called until it returns an empty batch (main.cpp:66-77,
    FastqSplitter.hpp:47-93).  A failed kseq_read ends the CURRENT batch only; the next call
    keeps reading (so parsing resumes after a truncated record unless the batch was empty).
    In paired mode file 2 is only read when file 1 succeeded.  Yields lists of
    (rec1, rec2|None)."""
    r1 = FastxReader(data1)
    r2 = FastxReader(data2) if data2 is not None else None
    while True:
        batch = []
        while len(batch) < maxnum:
            a = r1.read()
            if a is None:
                break
            b = None
            if r2 is not None:
                b = r2.read()
                if b is None:
                    break
            batch.append((a, b))
        if not batch:
            return
        yield batch


def _cstr(b):
    """Shark handles record fields as C strings (FastqSplitter.hpp:56): cut at the first NUL."""
    k = b.find(b"\x00")
    return b if k < 0 else b[:k]


def load_reference(path):
    """FastaSplitter (FastaSplitter.hpp:42-54): every record, in file order -> legend + seqs."""
    recs = parse_fastx(_read_all(path))
    legend = [_cstr(r[0]) for r in recs]
    seqs = [_cstr(r[1]) for r in recs]
    return legend, seqs


def load_sample(path1, path2=None, maxnum=50000):
    """-> (pairs, batch id per pair)"""
    pairs, bid = [], []
    for i, batch in enumerate(batches(_read_all(path1), _read_all(path2) if path2 else None, maxnum)):
        pairs += batch
        bid += [i] * len(batch)
    return pairs, bid


def join_reads(pairs, with_qual):
    """-> (seq uint8, qual uint8|None, off uint64) in the joined layout of FastqSplitter.hpp:63,83-84."""
    texts, quals = [], []
    for a, b in pairs:
        s, q = _cstr(a[1]), _cstr(a[2])
        if b is not None:
            s = s + b"N" + _cstr(b[1])
            q = q + b"\x1b" + _cstr(b[2])
        texts.append(s)
        # mask_seq loops over the QUAL length (FastqSplitter.hpp:104-108); pad/cut to the text
        quals.append((q + b"\x7f" * len(s))[: len(s)] if with_qual else b"")  # 0x7f is never < mq
    seq, off = pyoracle.concat_records(texts)
    qual = np.frombuffer(b"".join(quals), dtype=np.uint8).copy() if with_qual else None
    return seq, qual, off


def run_shark(ref, s1, s2=None, k=17, c=0.6, b=1, q=0, single=False, batch=50000):
    """Whole-program restatement with -t 1 ordering -> (ssv, out1, out2|None) bytes.
    main.cpp:83-240; output format ReadOutput.hpp:37-50 (FASTQ dedup by consecutive name,
    previd reset per 50 000-read batch: main.cpp:215)."""
    legend, seqs = load_reference(ref)
    bases, rec_off = pyoracle.concat_records(seqs)
    ix = pyoracle.Index(bases, rec_off, k, b << 33)
    pairs, bid = load_sample(s1, s2, batch)
    seq, qual, off = join_reads(pairs, q > 0)
    cnt, ar, ag = ix.analyze(seq, off, c, qual=qual, min_quality=q, single=single)
    ssv, o1, o2 = [], [], []
    prev = b""  # ReadOutput.hpp:39 `previd = ""`
    prev_batch = 0
    for r, g in zip(ar.tolist(), ag.tolist()):
        a, bb = pairs[r]
        if bid[r] != prev_batch:
            prev, prev_batch = b"", bid[r]
        name = _cstr(a[0])
        ssv.append(name + b" " + legend[g] + b"\n")
        if prev != name:
            o1.append(b"@" + name + b"\n" + _cstr(a[1]) + b"\n+\n" + _cstr(a[2]) + b"\n")
            if bb is not None:
                o2.append(b"@" + _cstr(bb[0]) + b"\n" + _cstr(bb[1]) + b"\n+\n" + _cstr(bb[2]) + b"\n")
        prev = name
    return b"".join(ssv), b"".join(o1), (b"".join(o2) if s2 is not None else None)

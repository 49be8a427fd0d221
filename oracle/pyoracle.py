"""ctypes view of oracle/libshark_oracle.so (the CPU restatement, shark_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg.  Nothing under shark_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libshark_oracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "shark")
HYBRID_BIN = os.path.join(_HERE, "_ref", "shark_hybrid")


def build(force=False):
    """Compile the restatement (gcc, <1 s).  Also builds oracle/_ref/shark when the reference
    tree is present (this container only; the GPU box uses the prebuilt binary)."""
    src = os.path.join(_HERE, "shark_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libshark_oracle.so"])
    if os.path.isdir("/root/reference") and not os.path.exists(REF_BIN):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    # the reference's main.cpp on top of our functor header (needs libshark_b200.so; make checks dates)
    if os.path.isdir("/root/reference") and os.path.exists(os.path.join(_HERE, "..", "shark_b200", "libshark_b200.so")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "hybrid"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u16p, u32p, u64p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int64))
        L.shko_base_code.restype = C.c_int
        L.shko_base_code.argtypes = [C.c_uint8]
        L.shko_xxh64_u64.restype = C.c_uint64
        L.shko_xxh64_u64.argtypes = [C.c_uint64]
        L.shko_revcompl.restype = C.c_uint64
        L.shko_revcompl.argtypes = [C.c_uint64, C.c_int]
        L.shko_lsappend.restype = C.c_uint64
        L.shko_lsappend.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.shko_rsprepend.restype = C.c_uint64
        L.shko_rsprepend.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.shko_build_kmer.restype = C.c_int64
        L.shko_build_kmer.argtypes = [C.c_char_p, C.c_int64, i64p, C.c_int]
        L.shko_enumerate.restype = C.c_int64
        L.shko_enumerate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.shko_index_build.restype = C.c_void_p
        L.shko_index_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint64]
        L.shko_index_build_wide.restype = C.c_void_p
        L.shko_index_build_wide.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint64]
        L.shko_index_free.argtypes = [C.c_void_p]
        for f, t in (("n_set", C.c_uint64), ("tot_ids", C.c_uint64), ("n_genes", C.c_uint32)):
            getattr(L, "shko_index_" + f).restype = t
            getattr(L, "shko_index_" + f).argtypes = [C.c_void_p]
        L.shko_index_pos.restype = u64p
        L.shko_index_off.restype = u32p
        L.shko_index_ids.restype = u16p
        L.shko_index_ids32.restype = u32p
        for f in ("pos", "off", "ids", "ids32"):
            getattr(L, "shko_index_" + f).argtypes = [C.c_void_p]
        L.shko_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.shko_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.shko_analyze.restype = C.c_uint64
        L.shko_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.shko_read_table.restype = C.c_int
        L.shko_read_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def xxh64_u64(v):
    return lib().shko_xxh64_u64(C.c_uint64(v & 0xFFFFFFFFFFFFFFFF))


def revcompl(kmer, k):
    return lib().shko_revcompl(kmer, k)


def build_kmer(s, p, k):
    """-> (kmer or -1, new p)   (kmer_utils.hpp:57-71)"""
    if isinstance(s, str):
        s = s.encode()
    pp = C.c_int64(p)
    r = lib().shko_build_kmer(s, len(s), C.byref(pp), k)
    return r, pp.value


def enumerate_kmers(s, k):
    """-> (canonical uint64[], end positions int64[]) or None when the reference `continue`s."""
    if isinstance(s, str):
        s = s.encode()
    b = np.frombuffer(s, dtype=np.uint8)
    canon = np.zeros(max(len(b), 1), dtype=np.uint64)
    endp = np.zeros(max(len(b), 1), dtype=np.int64)
    n = lib().shko_enumerate(_p(b) if len(b) else None, len(b), k, _p(canon), _p(endp))
    if n < 0:
        return None
    return canon[:n].copy(), endp[:n].copy()


def concat_records(seqs):
    """list of bytes -> (uint8 bases, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return bases, off


class Index:
    """Sparse restatement of class BF after switch_mode(2) (bloomfilter.h:36-203)."""

    def __init__(self, bases, rec_off, k, bf_bits, wide=False):
        """wide=True: the widened restatement (32-bit gene ids, SHK_F_WIDE_IDS / SURVEY.md 8f.4)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        self._keep = (bases, rec_off)
        self.k, self.bf_bits = k, bf_bits
        L = lib()
        build = L.shko_index_build_wide if wide else L.shko_index_build
        self._h = build(_p(bases), _p(rec_off), len(rec_off) - 1, k, C.c_uint64(bf_bits))
        self.n_set = L.shko_index_n_set(self._h)
        self.tot_ids = L.shko_index_tot_ids(self._h)
        self.n_genes = L.shko_index_n_genes(self._h)
        self.pos = np.ctypeslib.as_array(L.shko_index_pos(self._h), shape=(self.n_set + 1,))[: self.n_set]
        self.off = np.ctypeslib.as_array(L.shko_index_off(self._h), shape=(self.n_set + 1,))
        self.ids = np.ctypeslib.as_array(L.shko_index_ids(self._h), shape=(self.tot_ids + 1,))[: self.tot_ids]
        self.ids32 = np.ctypeslib.as_array(L.shko_index_ids32(self._h), shape=(self.tot_ids + 1,))[: self.tot_ids]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().shko_index_free(self._h)
            self._h = None

    def probe(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        n = len(kmers)
        rank = np.zeros(n, np.int64)
        begin = np.zeros(n, np.uint32)
        ln = np.zeros(n, np.uint32)
        lib().shko_probe(self._h, _p(kmers), n, _p(rank), _p(begin), _p(ln))
        return rank, begin, ln

    def analyze(self, seq, off, c, qual=None, min_quality=0, single=False):
        """-> (count per read uint32[n], assoc_read uint64[m], assoc_gene uint32[m])"""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        q = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
        cnt = np.zeros(max(n, 1), np.uint32)
        m = lib().shko_analyze(self._h, _p(seq), None if q is None else _p(q), _p(off), n, c, min_quality, int(single),
                               _p(cnt), None, None, 0)
        ar = np.zeros(max(m, 1), np.uint64)
        ag = np.zeros(max(m, 1), np.uint32)
        m2 = lib().shko_analyze(self._h, _p(seq), None if q is None else _p(q), _p(off), n, c, min_quality,
                                int(single), _p(cnt), _p(ar), _p(ag), m)
        assert m2 == m
        return cnt[:n], ar[:m], ag[:m]

    def read_table(self, text, cap=4096):
        if isinstance(text, str):
            text = text.encode()
        b = np.frombuffer(text, dtype=np.uint8)
        g = np.zeros(cap, np.int32)
        cv = np.zeros(cap, np.uint32)
        h = np.zeros(cap, np.uint32)
        n = lib().shko_read_table(self._h, _p(b), len(b), _p(g), _p(cv), _p(h), cap)
        n = min(n, cap)
        return g[:n], cv[:n], h[:n]


def mask(seq, qual, min_quality):
    seq = np.array(seq, dtype=np.uint8, copy=True)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    lib().shko_mask(_p(seq), _p(qual), len(seq), min_quality)
    return seq

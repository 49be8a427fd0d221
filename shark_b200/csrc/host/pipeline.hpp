// Host pipeline stages of shark-b200 around the device calls:
//
//   RecordStream x {1,2}  ->  Batcher  ->  [shk_reads_submit / shk_reads_collect]  ->  Writer
//   (ingest.hpp, a thread     (a thread)     (the caller's thread)                     (a thread)
//    per input file)
//
// Batcher = FastqSplitter::operator() called until it returns an empty batch (main.cpp:66-77,
// FastqSplitter.hpp:47-93): whole 50 000-read batches are packed into one chunk in the SoA layout
// of shk_reads_submit; what ReadOutput prints later (names, qualities) is kept as pointers into
// the scanner blocks, which the chunk keeps alive.  Writer = ReadOutput::operator()
// (ReadOutput.hpp:37-50) per batch, with its own buffers and write(2).
// `Alloc` provides the staging memory (pinned, from the library, in the CLI).
#pragma once
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ingest.hpp"

namespace shkhost {

constexpr unsigned kBatch = 50000;  // FastqSplitter batch (main.cpp:215): ReadOutput's dedup resets per batch

struct AssocPair {  // layout of shk_assoc
    uint32_t read_idx, gene_idx;
};

template <class Alloc>
struct Staging {
    uint8_t *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n, size_t keep_bytes)
    {
        if (n <= cap) return;
        size_t want = cap ? cap : (1u << 20);
        while (want < n) want *= 2;
        uint8_t *q = (uint8_t *)Alloc::alloc(want);
        if (p) {
            if (keep_bytes) memcpy(q, p, keep_bytes);
            Alloc::free(p);
        }
        p = q;
        cap = want;
    }
    ~Staging()
    {
        if (p) Alloc::free(p);
    }
};

// What ReadOutput needs for one read besides the sequence text (which is in Chunk::seq).
struct ReadMeta {
    const char *name1 = nullptr, *qual1 = nullptr, *name2 = nullptr, *qual2 = nullptr;
    uint32_t nlen1 = 0, qlen1 = 0, nlen2 = 0, qlen2 = 0;
    uint32_t len1 = 0;  // length of mate 1 in the joined text (mate 2 = total - len1 - 1)
};

template <class Alloc>
struct Chunk {
    Staging<Alloc> seq, qual, off;      // text (mate1 [+ 'N' + mate2]), qualities (+ 0x1B), uint32 offsets
    std::vector<ReadMeta> meta;
    std::vector<uint32_t> batch_start;  // read indices where a 50 000-read batch begins
    std::vector<std::shared_ptr<Block>> keep;
    std::vector<AssocPair> assoc;       // results, copied out of the slot before the slot is reused
    uint32_t n = 0;
    uint64_t bytes = 0;
    bool last = false;
    uint64_t index = 0;
    void clear()
    {
        meta.clear();
        batch_start.clear();
        keep.clear();
        assoc.clear();
        n = 0;
        bytes = 0;
        last = false;
    }
};

template <class Alloc>
class Batcher {
public:
    Batcher(const char *path1, const char *path2, bool with_qual_gpu) : s1_(path1), with_qual_gpu_(with_qual_gpu)
    {
        if (path2) s2_.reset(new RecordStream(path2));
    }
    bool files_ok() const { return s1_.ok() && (!s2_ || s2_->ok()); }
    void start()
    {
        s1_.start();
        if (s2_) s2_->start();
    }

    // Fills one chunk with whole batches; returns false when the input is exhausted (the chunk
    // may still hold reads).  A failed read ends the CURRENT batch only (FastqSplitter.hpp:53,61).
    bool fill(Chunk<Alloc> &ch, unsigned max_reads, uint64_t max_bytes)
    {
        ch.clear();
        max_reads_ = max_reads;
        max_bytes_ = max_bytes;
        uint64_t last_batch_bytes = 0;
        while ((ch.n + kBatch <= max_reads && ch.bytes + last_batch_bytes + last_batch_bytes / 4 <= max_bytes) || ch.n == 0) {
            ch.batch_start.push_back(ch.n);
            const uint64_t bytes0 = ch.bytes;
            unsigned got = 0;
            while (got < kBatch) {
                const Rec a = s1_.peek();
                if (a.status < 0) {
                    s1_.consume();
                    break;
                }
                const std::shared_ptr<Block> &ba = s1_.block();
                hold(ch, ba);
                const bool nul_a = ba->has_nul;
                s1_.consume();
                Rec b;
                bool nul_b = false;
                if (s2_) {
                    b = s2_->peek();
                    if (b.status < 0) {  // mate 1 is dropped (FastqSplitter.hpp:61)
                        s2_->consume();
                        break;
                    }
                    const std::shared_ptr<Block> &bb = s2_->block();
                    hold(ch, bb);
                    nul_b = bb->has_nul;
                    s2_->consume();
                }
                append(ch, a, nul_a, b, nul_b);
                ++got;
            }
            if (got == 0) {
                ch.batch_start.pop_back();
                return false;  // empty batch: the reference's worker returns (main.cpp:70)
            }
            last_batch_bytes = ch.bytes - bytes0;
        }
        return true;
    }

private:
    static void hold(Chunk<Alloc> &ch, const std::shared_ptr<Block> &b)
    {
        // blocks arrive in order per stream: at most the last two entries can be this block
        const size_t k = ch.keep.size();
        if (k >= 1 && ch.keep[k - 1] == b) return;
        if (k >= 2 && ch.keep[k - 2] == b) return;
        ch.keep.push_back(b);
    }
    // C-string semantics of the reference (FastqSplitter.hpp:56: `seq1->seq.s` as const char*)
    static uint32_t clen(const char *s, uint32_t n, bool may_have_nul)
    {
        if (!may_have_nul) return n;
        const void *z = memchr(s, 0, n);
        return z ? (uint32_t)((const char *)z - s) : n;
    }
    void append(Chunk<Alloc> &ch, const Rec &a, bool nul_a, const Rec &b, bool nul_b)
    {
        const bool paired = (bool)s2_;
        const uint32_t l1 = clen(a.seq, a.seq_len, nul_a), l2 = paired ? clen(b.seq, b.seq_len, nul_b) : 0;
        const size_t total = paired ? (size_t)l1 + 1 + l2 : l1;
        if (ch.n == 0) {
            // size the staging buffers once from the first read (reads of a run have similar lengths)
            const uint64_t guess = std::min<uint64_t>(max_bytes_ + 64, (uint64_t)(total + 8) * max_reads_ * 9 / 8 + (1u << 16));
            ch.seq.reserve(guess, 0);
            if (with_qual_gpu_) ch.qual.reserve(guess, 0);
            ch.off.reserve(((size_t)max_reads_ + 2) * 4, 0);
            ((uint32_t *)ch.off.p)[0] = 0;
            ch.meta.reserve(max_reads_);
        }
        ch.off.reserve(((size_t)ch.n + 2) * 4, ((size_t)ch.n + 1) * 4);
        uint32_t *off = (uint32_t *)ch.off.p;
        ch.seq.reserve(ch.bytes + total + 8, ch.bytes);
        uint8_t *d = ch.seq.p + ch.bytes;
        memcpy(d, a.seq, l1);
        if (paired) {
            d[l1] = 'N';  // FastqSplitter.hpp:63,83
            memcpy(d + l1 + 1, b.seq, l2);
        }
        // qualities: mask_seq walks the QUAL string (FastqSplitter.hpp:104-108); positions it does
        // not reach are never masked -> pad with 0x7f, which is not < any mq
        const uint32_t ql1 = clen(a.qual, a.qual_len, nul_a), ql2 = paired ? clen(b.qual, b.qual_len, nul_b) : 0;
        if (with_qual_gpu_) {
            ch.qual.reserve(ch.bytes + total + 8, ch.bytes);
            uint8_t *q = ch.qual.p + ch.bytes;
            if (!paired) {
                const size_t c1 = ql1 < total ? ql1 : total;
                memcpy(q, a.qual, c1);
                if (c1 < total) memset(q + c1, 0x7f, total - c1);
            } else {
                // string(qual1) + "\33" + string(qual2), FastqSplitter.hpp:84
                size_t w = 0;
                const size_t c1 = ql1 < total ? ql1 : total;
                memcpy(q, a.qual, c1);
                w = c1;
                if (w < total) q[w++] = 0x1B;
                const size_t c2 = ql2 < total - w ? ql2 : total - w;
                memcpy(q + w, b.qual, c2);
                w += c2;
                if (w < total) memset(q + w, 0x7f, total - w);
            }
        }
        ReadMeta m;
        m.name1 = a.name;
        m.nlen1 = clen(a.name, a.name_len, nul_a);
        m.qual1 = a.qual;
        m.qlen1 = ql1;
        if (paired) {
            m.name2 = b.name;
            m.nlen2 = clen(b.name, b.name_len, nul_b);
            m.qual2 = b.qual;
            m.qlen2 = ql2;
        }
        m.len1 = l1;
        ch.meta.push_back(m);
        ch.bytes += total;
        ++ch.n;
        off[ch.n] = (uint32_t)ch.bytes;
    }

    RecordStream s1_;
    std::unique_ptr<RecordStream> s2_;
    bool with_qual_gpu_;
    unsigned max_reads_ = 0;
    uint64_t max_bytes_ = 0;
};

// Buffered writer over a file descriptor.
class OutBuf {
public:
    explicit OutBuf(int fd, size_t cap = 8u << 20) : fd_(fd), buf_(new char[cap]), cap_(cap) {}
    ~OutBuf() { flush(); }
    void put(const void *p, size_t n)
    {
        if (n > cap_ - len_) {
            flush();
            if (n >= cap_) {
                raw(p, n);
                return;
            }
        }
        memcpy(buf_.get() + len_, p, n);
        len_ += n;
    }
    void putc(char c)
    {
        if (len_ == cap_) flush();
        buf_[len_++] = c;
    }
    void flush()
    {
        if (len_) raw(buf_.get(), len_);
        len_ = 0;
    }

private:
    void raw(const void *p, size_t n)
    {
        const char *c = (const char *)p;
        while (fd_ >= 0 && n) {
            const ssize_t w = ::write(fd_, c, n);
            if (w <= 0) {
                if (w < 0 && errno == EINTR) continue;
                return;  // like the reference, output errors are not reported
            }
            c += w;
            n -= (size_t)w;
        }
    }
    int fd_;
    std::unique_ptr<char[]> buf_;
    size_t cap_, len_ = 0;
};

// ReadOutput::operator() (ReadOutput.hpp:37-50) for one chunk's associations.
template <class Alloc>
class Writer {
public:
    Writer(int fd_ssv, int fd_out1, int fd_out2, const std::vector<std::string> &legend, bool paired)
        : legend_(legend), paired_(paired), ssv_(fd_ssv), out1_(fd_out1), out2_(fd_out2), has1_(fd_out1 >= 0),
          has2_(fd_out2 >= 0)
    {
    }
    void flush()
    {
        ssv_.flush();
        out1_.flush();
        out2_.flush();
    }
    void write(const Chunk<Alloc> &ch)
    {
        const uint32_t *off = (const uint32_t *)ch.off.p;
        size_t next_batch = 0;
        const char *previd = "";  // `string previd = ""` per ReadOutput call = per batch
        uint32_t prevlen = 0;
        for (const AssocPair &as : ch.assoc) {
            const uint32_t r = as.read_idx, g = as.gene_idx;
            while (next_batch < ch.batch_start.size() && ch.batch_start[next_batch] <= r) {
                previd = "";
                prevlen = 0;
                ++next_batch;
            }
            const ReadMeta &m = ch.meta[r];
            ssv_.put(m.name1, m.nlen1);
            ssv_.putc(' ');
            if (g < legend_.size()) ssv_.put(legend_[g].data(), strlen(legend_[g].c_str()));
            ssv_.putc('\n');
            if (prevlen != m.nlen1 || memcmp(previd, m.name1, prevlen) != 0) {
                const char *s = (const char *)ch.seq.p + off[r];
                if (has1_) {
                    out1_.putc('@');
                    out1_.put(m.name1, m.nlen1);
                    out1_.putc('\n');
                    out1_.put(s, m.len1);
                    out1_.put("\n+\n", 3);
                    out1_.put(m.qual1, m.qlen1);
                    out1_.putc('\n');
                }
                if (has2_ && paired_) {
                    const uint32_t l2 = off[r + 1] - off[r] - m.len1 - 1;
                    out2_.putc('@');
                    out2_.put(m.name2, m.nlen2);
                    out2_.putc('\n');
                    out2_.put(s + m.len1 + 1, l2);
                    out2_.put("\n+\n", 3);
                    out2_.put(m.qual2, m.qlen2);
                    out2_.putc('\n');
                }
            }
            previd = m.name1;
            prevlen = m.nlen1;
        }
    }

private:
    const std::vector<std::string> &legend_;
    bool paired_;
    OutBuf ssv_, out1_, out2_;
    bool has1_, has2_;
};

}  // namespace shkhost

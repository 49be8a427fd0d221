// Block-wise FASTQ/FASTA ingest for the host side of shark-b200.
//
// FastqScanner yields exactly the sequence of kseq_read() outcomes that FastxReader (tests/host_tools/fastx.hpp,
// the record-at-a-time restatement of kseq.h:177-218) yields on the same bytes, but works on
// 8 MiB blocks: a record in the strict four-line shape
//        @name[ comment]\n  SEQ\n  +[anything]\n  QUAL\n        with |QUAL| == |SEQ|
// is recognised with four memchr calls and handed out as pointers into the block (no copy);
// everything else (FASTA, wrapped lines, '\r', empty lines, truncated records, garbage between
// records, the last bytes of a file) goes through the same character-level state machine as
// FastxReader, assembled into a side arena of the block.  One scanner per input file runs in its
// own thread; the outcome stream of a file does not depend on the other file, so pairing them
// afterwards (FastqSplitter.hpp:47-93) is exact.
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <cctype>
#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bgzf.hpp"

namespace shkhost {

// One kseq_read() outcome.  Pointers are not NUL-terminated and stay valid while the Block lives.
struct Rec {
    const char *name = nullptr, *seq = nullptr, *qual = nullptr;
    uint32_t name_len = 0, seq_len = 0, qual_len = 0;
    int32_t status = 0;  // >= 0: a record; -1 end of file; -2 truncated quality; -3 stream error
};

struct Block {
    std::vector<std::shared_ptr<char>> bufs;   // input buffers the records point into
    std::vector<std::unique_ptr<char[]>> arena;  // storage of records assembled by the general path
    size_t arena_used = 0, arena_cap = 0;
    std::vector<Rec> recs;
    bool has_nul = false;  // some buffer byte is 0: consumers apply C-string semantics per field
    bool all_ok = false;   // every outcome is a record (status >= 0) and no byte is 0: bulk consumers may take it whole

    char *arena_alloc(size_t n)
    {
        if (n > arena_cap - arena_used || arena.empty()) {
            const size_t cap = n > (1u << 16) ? n : (1u << 16);
            arena.emplace_back(new char[cap]);
            arena_cap = cap;
            arena_used = 0;
        }
        char *p = arena.back().get() + arena_used;
        arena_used += n;
        return p;
    }
};

// Recycles the input buffers (fresh 8 MiB allocations cost a page fault per 4 KiB).
class BufferPool {
public:
    static BufferPool &instance()
    {
        static BufferPool p;
        return p;
    }
    std::shared_ptr<char> get(size_t cap)
    {
        char *b = nullptr;
        if (cap <= kStd) {
            std::lock_guard<std::mutex> lk(mu_);
            if (!free_.empty()) {
                b = free_.back();
                free_.pop_back();
            }
        }
        if (!b) b = new char[cap <= kStd ? kStd : cap];
        const bool pooled = cap <= kStd;
        return std::shared_ptr<char>(b, [this, pooled](char *q) {
            if (pooled) {
                std::lock_guard<std::mutex> lk(mu_);
                if (free_.size() < 48) {
                    free_.push_back(q);
                    return;
                }
            }
            delete[] q;
        });
    }
    ~BufferPool()
    {
        for (char *q : free_) delete[] q;
    }
    static constexpr size_t kStd = (8u << 20) + (1u << 20);

private:
    std::mutex mu_;
    std::vector<char *> free_;
};

// One record in the strict four-line shape at base[pos..end):
//        @name[ comment]\n  SEQ\n  +[anything]\n  QUAL\n        with |QUAL| == |SEQ|
// kStrictOk: r filled (pointers into base), next = position after the record.  kStrictNeedMore: the bytes up
// to `end` do not hold the whole record.  kStrictFormat: not this shape - the character-level state machine
// decides.  Four memchr calls, no copy.
enum { kStrictOk = 0, kStrictNeedMore = 1, kStrictFormat = 2 };
inline int parse_strict(const char *base, size_t pos, size_t end, Rec &r, size_t &next)
{
    const char *p = base + pos, *e = base + end;
    if (p >= e) return kStrictNeedMore;
    if (*p != '@') return kStrictFormat;
    // header line: short, scanned once for both the end of the name and the end of the line
    const char *n0 = p + 1, *n1 = nullptr, *l1 = n0;
    for (;; ++l1) {
        if (l1 >= e) return kStrictNeedMore;
        const unsigned char c = (unsigned char)*l1;
        if (c > ' ') continue;  // isspace() bytes are all <= ' '
        if (c == '\n') break;
        if (!n1 && isspace(c)) {
            n1 = l1;
            const char *nl = (const char *)memchr(l1, '\n', (size_t)(e - l1));  // comment: skip it
            if (!nl) return kStrictNeedMore;
            l1 = nl;
            break;
        }
    }
    if (!n1) n1 = l1;
    const char *s = l1 + 1;
    if (s >= e) return kStrictNeedMore;
    if (*s == '\n' || *s == '>' || *s == '+' || *s == '@') return kStrictFormat;
    const char *l2 = (const char *)memchr(s, '\n', (size_t)(e - s));
    if (!l2) return kStrictNeedMore;
    if (l2[-1] == '\r') return kStrictFormat;
    const char *t = l2 + 1;
    if (e - t < 2) return kStrictNeedMore;
    if (*t != '+') return kStrictFormat;
    const char *l3 = t[1] == '\n' ? t + 1 : (const char *)memchr(t, '\n', (size_t)(e - t));
    if (!l3) return kStrictNeedMore;
    const char *q = l3 + 1;
    const size_t sl = (size_t)(l2 - s);
    if ((size_t)(e - q) <= sl) return kStrictNeedMore;  // the quality line and its '\n' must be here
    if (q[sl] != '\n') return kStrictFormat;            // shorter (a '\n' earlier is caught below) or longer
    if (memchr(q, '\n', sl) != nullptr || q[sl - 1] == '\r') return kStrictFormat;
    r.name = n0;
    r.name_len = (uint32_t)(n1 - n0);
    r.seq = s;
    r.seq_len = (uint32_t)sl;
    r.qual = q;
    r.qual_len = (uint32_t)sl;
    r.status = (int32_t)sl;
    next = (size_t)(q + sl + 1 - base);
    return kStrictOk;
}

class FastqScanner {
public:
    // start_offset: plain (uncompressed) files only - continue at that byte, at a record boundary (the parallel
    // scanner of fastpipe.hpp hands its tail over this way)
    explicit FastqScanner(const char *path, size_t block_bytes = 8u << 20, uint64_t start_offset = 0) : block_bytes_(block_bytes)
    {
        fd_ = open(path, O_RDONLY);
        if (fd_ < 0) return;
        if (start_offset) {
            if (lseek(fd_, (off_t)start_offset, SEEK_SET) < 0) {
                close(fd_);
                fd_ = -1;
            }
            return;
        }
        // gzip or plain, like gzopen/gzread; a plain file is then read with read(2) directly
        unsigned char magic[2] = {0, 0};
        const ssize_t m = pread(fd_, magic, 2, 0);
        if (m == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            // blocked gzip (bgzip): members located up front and inflated in parallel (bgzf.hpp); anything
            // else, and whatever follows the first member that is not such a block, goes through zlib
            if (bgzf_.open(fd_)) use_bgzf_ = true;
            else if (!open_zlib(0)) {
                close(fd_);
                fd_ = -1;
            }
        }
    }
    ~FastqScanner()
    {
        bgzf_.shutdown();     // its read-ahead thread uses fd_
        if (f_) gzclose(f_);  // closes fd_ too
        else if (fd_ >= 0) close(fd_);
    }
    FastqScanner(const FastqScanner &) = delete;
    FastqScanner &operator=(const FastqScanner &) = delete;
    bool ok() const { return fd_ >= 0; }

    // The next block of outcomes.  A block ends when its input buffer is used up or it holds
    // max_recs outcomes.  After end of file every block holds a single -1 (kseq_read keeps
    // returning -1); -2/-3 outcomes are part of the stream like any other.
    std::unique_ptr<Block> next(size_t max_recs = 1u << 20)
    {
        std::unique_ptr<Block> blk(new Block);
        blk_ = blk.get();
        attach_buffer();
        while (blk->recs.size() < max_recs) {
            if (last_char_ == 0) {
                const int f = fast_record();
                if (f == kOk) continue;
                if (f == kNeedMore) {
                    if (!blk->recs.empty() && !eof_ && !err_ && !err_pending_) break;  // hand the block out, refill on the next call
                    if (refill()) continue;
                    // no more data: the tail goes through the general path (exact end-of-file semantics)
                }
            }
            const long r = general_record();
            if (r == -1 || r == -3) break;  // sticky: do not spin at end of file
        }
        blk_ = nullptr;
        blk->all_ok = !blk->has_nul;
        for (const Rec &r : blk->recs)
            if (r.status < 0) {
                blk->all_ok = false;
                break;
            }
        return blk;
    }

private:
    enum { kOk = 0, kNeedMore = 1, kFormat = 2 };

    bool open_zlib(uint64_t at)
    {
        if (lseek(fd_, (off_t)at, SEEK_SET) < 0) return false;
        f_ = gzdopen(fd_, "r");
        if (f_) gzbuffer(f_, 1u << 20);
        return f_ != nullptr;
    }

    // ---- buffer management ---------------------------------------------------------------
    void attach_buffer()
    {
        if (buf_ && (blk_->bufs.empty() || blk_->bufs.back() != buf_)) {
            blk_->bufs.push_back(buf_);
            blk_->has_nul = blk_->has_nul || buf_has_nul_;
        }
    }
    // New buffer = unconsumed tail of the old one + fresh bytes.  Records already handed out keep
    // the old buffer alive through their block.
    bool refill()
    {
        if (err_pending_) err_ = true;  // bytes decoded before a stream error are parsed first, like kseq
        if (eof_ || err_) return false;
        const size_t tail = end_ - pos_;
        const size_t cap = tail + block_bytes_;
        std::shared_ptr<char> nb = BufferPool::instance().get(cap);
        if (tail) memcpy(nb.get(), buf_.get() + pos_, tail);
        size_t got = 0;
        while (got < block_bytes_) {  // gzread returns short counts at member boundaries
            long n;
            if (use_bgzf_) {
                n = bgzf_.read(nb.get() + tail + got, block_bytes_ - got);
                if (n == BgzfSource::kHandover) {
                    // the member at this offset is zlib's: a gzip member continues the stream exactly as under
                    // gzread; anything else after gzip data is trailing garbage, which gzread ignores (gz_look)
                    use_bgzf_ = false;
                    const uint64_t at = bgzf_.handover_offset();
                    unsigned char mg[2] = {0, 0};
                    if (pread(fd_, mg, 2, (off_t)at) == 2 && mg[0] == 0x1f && mg[1] == 0x8b) {
                        if (!open_zlib(at)) n = -1;
                        else continue;
                    } else {
                        n = 0;
                        garbage_tail_ = true;
                    }
                }
            } else if (garbage_tail_) {
                n = 0;
            } else if (f_) {
                n = gzread(f_, nb.get() + tail + got, (unsigned)(block_bytes_ - got));
            } else {
                do n = (long)read(fd_, nb.get() + tail + got, block_bytes_ - got);
                while (n < 0 && errno == EINTR);
            }
            if (n < 0) {
                if (got) err_pending_ = true;
                else err_ = true;
                break;
            }
            if (n == 0) {
                eof_ = true;
                break;
            }
            got += (size_t)n;
        }
        buf_ = nb;
        pos_ = 0;
        end_ = tail + got;
        buf_has_nul_ = end_ > 0 && memchr(buf_.get(), 0, end_) != nullptr;
        if (blk_) attach_buffer();
        return got > 0;
    }

    // ---- strict four-line records ----------------------------------------------------------
    int fast_record()
    {
        if (!buf_) return kNeedMore;
        Rec r;
        size_t next = 0;
        const int st = parse_strict(buf_.get(), pos_, end_, r, next);
        if (st != kStrictOk) return st == kStrictNeedMore ? kNeedMore : kFormat;
        blk_->recs.push_back(r);
        pos_ = next;
        return kOk;
    }

    // ---- general path: the state machine of FastxReader over this stream ---------------------
    int getc()
    {
        if (err_) return -3;
        if (pos_ >= end_ && !refill()) return err_ ? -3 : -1;
        return (int)(unsigned char)buf_.get()[pos_++];
    }
    enum Delim { kSpace, kLine };
    long get_until(Delim d, std::string &out, bool append, int *dret)
    {
        if (!append) out.clear();
        if (dret) *dret = 0;
        bool got_any = false;
        for (;;) {
            if (err_) return -3;
            if (pos_ >= end_ && !refill()) break;
            got_any = true;
            const unsigned char *p = (const unsigned char *)buf_.get() + pos_;
            const size_t avail = end_ - pos_;
            size_t i = 0;
            if (d == kLine) {
                const void *q = memchr(p, '\n', avail);
                i = q ? (size_t)((const unsigned char *)q - p) : avail;
            } else {
                while (i < avail && !isspace(p[i])) ++i;
            }
            out.append((const char *)p, i);
            pos_ += i;
            if (i < avail) {
                if (dret) *dret = p[i];
                ++pos_;
                break;
            }
        }
        if (!got_any) return -1;
        if (d == kLine && out.size() > 1 && out.back() == '\r') out.pop_back();
        return (long)out.size();
    }
    long read_general()
    {
        int c;
        if (last_char_ == 0) {
            while ((c = getc()) >= 0 && c != '>' && c != '@') {
            }
            if (c < 0) return c;
            last_char_ = c;
        }
        seq_.clear();
        qual_.clear();
        int delim = 0;
        long r = get_until(kSpace, name_, false, &delim);
        if (r < 0) return r;
        if (delim != '\n') {
            scratch_.clear();
            get_until(kLine, scratch_, false, nullptr);
        }
        while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq_.push_back((char)c);
            get_until(kLine, seq_, true, nullptr);
        }
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (long)seq_.size();
        while ((c = getc()) >= 0 && c != '\n') {
        }
        if (c == -1) return -2;
        while (get_until(kLine, qual_, true, nullptr) >= 0 && qual_.size() < seq_.size()) {
        }
        last_char_ = 0;
        if (seq_.size() != qual_.size()) return -2;
        return (long)seq_.size();
    }
    long general_record()
    {
        const long st = read_general();
        Rec r;
        r.status = (int32_t)st;
        if (st >= 0) {
            char *a = blk_->arena_alloc(name_.size() + seq_.size() + qual_.size() + 1);
            memcpy(a, name_.data(), name_.size());
            memcpy(a + name_.size(), seq_.data(), seq_.size());
            memcpy(a + name_.size() + seq_.size(), qual_.data(), qual_.size());
            r.name = a;
            r.name_len = (uint32_t)name_.size();
            r.seq = a + name_.size();
            r.seq_len = (uint32_t)seq_.size();
            r.qual = r.seq + seq_.size();
            r.qual_len = (uint32_t)qual_.size();
            if (memchr(a, 0, name_.size() + seq_.size() + qual_.size())) blk_->has_nul = true;
        }
        blk_->recs.push_back(r);
        return st;
    }

    int fd_ = -1;
    gzFile f_ = nullptr;  // set when the input is gzip read through zlib
    BgzfSource bgzf_;     // blocked gzip: parallel inflate, until a member that is not a BGZF block
    bool use_bgzf_ = false, garbage_tail_ = false;
    size_t block_bytes_;
    std::shared_ptr<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false, err_ = false, err_pending_ = false, buf_has_nul_ = false;
    int last_char_ = 0;
    Block *blk_ = nullptr;
    std::string name_, seq_, qual_, scratch_;
};

// Bounded blocking queue between pipeline stages.
template <class T>
class BoundedQueue {
public:
    explicit BoundedQueue(size_t cap = 4) : cap_(cap) {}
    void push(T v)
    {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < cap_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    T pop()
    {
        std::unique_lock<std::mutex> lk(mu_);
        not_empty_.wait(lk, [&] { return !q_.empty(); });
        T v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return v;
    }

private:
    std::mutex mu_;
    std::condition_variable not_empty_, not_full_;
    std::deque<T> q_;
    size_t cap_;
};

// A file's outcome stream, scanned ahead by its own thread.
class RecordStream {
public:
    explicit RecordStream(const char *path) : scanner_(path), q_(6) {}
    bool ok() const { return scanner_.ok(); }
    void start()
    {
        th_ = std::thread([this] {
            for (;;) {
                std::unique_ptr<Block> b = scanner_.next();
                const bool end = !b->recs.empty() && (b->recs.back().status == -1 || b->recs.back().status == -3);
                q_.push(std::shared_ptr<Block>(b.release()));
                if (end) break;  // the consumer repeats the final outcome itself
            }
        });
    }
    ~RecordStream()
    {
        if (th_.joinable()) {
            // drain so that a producer blocked on a full queue can finish
            while (!done_) advance();
            th_.join();
        }
    }
    // The current outcome; `block()` owns its bytes.
    const Rec &peek()
    {
        if (!cur_ || idx_ >= cur_->recs.size()) advance();
        return cur_->recs[idx_];
    }
    const std::shared_ptr<Block> &block() const { return cur_; }
    void consume()
    {
        const int32_t st = cur_->recs[idx_].status;
        if (done_ && idx_ + 1 >= cur_->recs.size() && (st == -1 || st == -3)) return;  // sticky end
        ++idx_;
    }

private:
    void advance()
    {
        if (done_) {
            idx_ = cur_->recs.size() - 1;
            return;
        }
        cur_ = q_.pop();
        idx_ = 0;
        const int32_t st = cur_->recs.empty() ? 0 : cur_->recs.back().status;
        if (st == -1 || st == -3) done_ = true;
    }
    FastqScanner scanner_;
    BoundedQueue<std::shared_ptr<Block>> q_;
    std::thread th_;
    std::shared_ptr<Block> cur_;
    size_t idx_ = 0;
    bool done_ = false;
};

}  // namespace shkhost

// Parallel decompression of blocked gzip (BGZF: what `bgzip` and most sequencing pipelines write) for the
// host ingest of shark-b200 ("multi-threaded/chunked FASTQ + gzip parsing").
//
// The reference reads every input through zlib's gzread (kseq.h: KSEQ_INIT(gzFile, gzread)), one inflate
// stream on one thread - 0.3-0.4 GB/s of FASTQ, two orders of magnitude below what the GPU path
// classifies.  A BGZF file is a series of independent gzip members of at most 64 KiB whose header carries
// the compressed size (extra subfield 'B','C') and whose trailer carries the uncompressed size, so the
// members that fill a buffer can be located up front and inflated in parallel, each straight to its final
// position.  Exactness: the bytes are those gzread yields.  Anything that is not a well-formed BGZF block
// (ordinary gzip members, a damaged or truncated block, a CRC or size mismatch, trailing bytes) ends the
// parallel path AT THAT MEMBER'S OFFSET and the caller continues there with zlib itself, so that errors and
// trailing garbage behave exactly as with gzread (tests/test_host_ingest.py compares outcome streams).
#pragma once
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <mutex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace shkhost {

// Threads used to inflate BGZF blocks (per input file).  SHK_INGEST_THREADS, else set_ingest_threads()
// (the CLI passes -t when it is > 1), else min(8, hardware threads / 2).
inline int &ingest_threads_setting()
{
    static int v = 0;
    return v;
}
inline void set_ingest_threads(int n) { ingest_threads_setting() = n; }
inline int ingest_threads()
{
    if (const char *ev = getenv("SHK_INGEST_THREADS")) {
        const int v = atoi(ev);
        if (v >= 1 && v <= 256) return v;
    }
    if (ingest_threads_setting() >= 1) return ingest_threads_setting() > 64 ? 64 : ingest_threads_setting();
    int hw = (int)std::thread::hardware_concurrency() / 2;
    return hw < 1 ? 1 : (hw > 8 ? 8 : hw);
}

class BgzfSource {
public:
    static constexpr long kHandover = -2;  // read(): continue with zlib at handover_offset()

    // True iff the file behind fd starts with a BGZF block.  fd is only pread, never moved or closed.
    bool open(int fd)
    {
        fd_ = fd;
        off_ = 0;
        unsigned char h[64];
        const ssize_t m = pread_full(h, sizeof h, 0);
        uint32_t hdr = 0, blen = 0;
        active_ = m >= 18 && parse_header(h, (size_t)m, hdr, blen) == kOk;
        return active_;
    }
    uint64_t handover_offset() const { return off_; }

    ~BgzfSource() { shutdown(); }
    // Stops the read-ahead thread (before the caller closes the file descriptor).
    void shutdown() { stop_ahead(); }
    BgzfSource() = default;
    BgzfSource(const BgzfSource &) = delete;
    BgzfSource &operator=(const BgzfSource &) = delete;

    // Up to n bytes of the decompressed stream -> dst.  > 0 bytes delivered; 0 end of file; -1 read error;
    // kHandover: the member at handover_offset() is not a block this class accepts.
    // With more than one ingest thread a read-ahead thread keeps the next 8 MiB inflated while the caller
    // parses the previous ones (two buffers); the bytes and the end conditions are those of read_sync().
    long read(char *dst, size_t n)
    {
        if (n == 0) return 0;
        if (!ahead_started_) {
            ahead_started_ = true;
            if (ingest_threads() > 1) {
                try {
                    for (auto &s : slot_) s.data.resize(kAheadBytes);
                    ahead_ = std::thread([this] { ahead_loop(); });
                    ahead_on_ = true;
                } catch (...) {
                    ahead_on_ = false;
                }
            }
        }
        if (!ahead_on_) return read_sync(dst, n);
        Slot &s = slot_[cons_ & 1];
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return s.full; });
        }
        if (s.status <= 0) return s.status;  // terminal and sticky: the producer has stopped
        const size_t k = s.len - s.pos < n ? s.len - s.pos : n;
        memcpy(dst, s.data.data() + s.pos, k);
        s.pos += k;
        if (s.pos == s.len) {
            {
                std::lock_guard<std::mutex> lk(mu_);
                s.full = false;
            }
            cv_.notify_all();
            ++cons_;
        }
        return (long)k;
    }

private:
    static constexpr size_t kAheadBytes = 8u << 20;
    struct Slot {
        std::vector<char> data;
        size_t len = 0, pos = 0;
        long status = 0;
        bool full = false;
    };
    void ahead_loop()
    {
        for (unsigned prod = 0;; ++prod) {
            Slot &s = slot_[prod & 1];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !s.full || quit_; });
                if (quit_) return;
            }
            const long r = read_sync(s.data.data(), kAheadBytes);
            {
                std::lock_guard<std::mutex> lk(mu_);
                s.status = r;
                s.len = r > 0 ? (size_t)r : 0;
                s.pos = 0;
                s.full = true;
            }
            cv_.notify_all();
            if (r <= 0) return;  // end of file, error or handover: nothing more to produce
        }
    }
    void stop_ahead()
    {
        if (!ahead_.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        ahead_.join();
    }
    std::thread ahead_;
    std::mutex mu_;
    std::condition_variable cv_;
    Slot slot_[2];
    unsigned cons_ = 0;
    bool ahead_started_ = false, ahead_on_ = false, quit_ = false;

    long read_sync(char *dst, size_t n)
    {
        if (n == 0) return 0;
        if (carry_pos_ < carry_len_) {
            const size_t k = carry_len_ - carry_pos_ < n ? carry_len_ - carry_pos_ : n;
            memcpy(dst, carry_.data() + carry_pos_, k);
            carry_pos_ += k;
            return (long)k;
        }
        for (;;) {
            if (pending_handover_) return kHandover;
            // compressed bytes for about n bytes of output (FASTQ deflates to a quarter or so); one block at least
            const size_t want = (n / 2 < (1u << 16) ? (1u << 16) : n / 2) + (1u << 16) + 64;
            if (cbuf_.size() < want) cbuf_.resize(want);
            const ssize_t got = pread_full(cbuf_.data(), want, off_);
            if (got < 0) return -1;
            if (got == 0) return 0;  // clean end of file at a block boundary
            plan_.clear();
            size_t cpos = 0, out = 0;
            bool stop_is_foreign = false;
            const bool at_eof = (size_t)got < want;  // the window reaches the end of the file
            while (cpos < (size_t)got) {
                uint32_t hdr = 0, blen = 0;
                const int ph = parse_header(cbuf_.data() + cpos, (size_t)got - cpos, hdr, blen);
                if (ph == kNo || (ph == kNeedMore && at_eof)) {  // a foreign member, garbage, or a header cut by the end of the file
                    stop_is_foreign = true;
                    break;
                }
                if (ph == kNeedMore) break;  // header cut by the window
                if (blen > (size_t)got - cpos) {  // block cut by the window - or by the end of a truncated file
                    stop_is_foreign = at_eof;
                    break;
                }
                const unsigned char *b = cbuf_.data() + cpos;
                const uint32_t isize = le32(b + blen - 4), crc = le32(b + blen - 8);
                if (isize > (1u << 16) || blen < hdr + 8) {
                    stop_is_foreign = true;
                    break;
                }
                if (out + isize > n) {
                    if (!plan_.empty() || out) break;  // the buffer is full enough
                    // a request smaller than one block: inflate into the carry buffer and serve it piecewise
                    if (carry_.size() < (1u << 16)) carry_.resize(1u << 16);
                    Item it{b + hdr, blen - hdr - 8, crc, isize, blen, 0};
                    if (!inflate_one(it, (char *)carry_.data())) {
                        pending_handover_ = true;
                        return kHandover;
                    }
                    off_ += blen;
                    carry_len_ = isize;
                    carry_pos_ = 0;
                    return read_sync(dst, n);
                }
                plan_.push_back(Item{b + hdr, blen - hdr - 8, crc, isize, blen, out});
                out += isize;
                cpos += blen;
            }
            if (plan_.empty()) {
                if (stop_is_foreign) {
                    pending_handover_ = true;
                    return kHandover;
                }
                return -1;  // cannot happen: a window always holds one whole block
            }
            // inflate the planned blocks, each to its final position
            std::atomic<size_t> next{0}, bad{plan_.size()};
            auto work = [&] {
                for (;;) {
                    const size_t i = next.fetch_add(1, std::memory_order_relaxed);
                    if (i >= plan_.size()) break;
                    if (!inflate_one(plan_[i], dst + plan_[i].out_off)) {
                        size_t cur = bad.load();
                        while (i < cur && !bad.compare_exchange_weak(cur, i)) {
                        }
                    }
                }
            };
            int nt = ingest_threads();
            if ((size_t)nt > plan_.size() / 4) nt = (int)(plan_.size() / 4);
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) {
                try {
                    th.emplace_back(work);
                } catch (...) {
                    break;
                }
            }
            work();
            for (auto &t : th) t.join();
            const size_t good = bad.load();
            size_t delivered = 0;
            uint64_t consumed = 0;
            for (size_t i = 0; i < good; ++i) {
                delivered += plan_[i].isize;
                consumed += plan_[i].block_len;
            }
            off_ += consumed;
            if (good < plan_.size() || stop_is_foreign) pending_handover_ = true;
            if (delivered) return (long)delivered;
            if (pending_handover_) return kHandover;
            // only empty blocks (the BGZF end-of-file marker): look further
        }
    }

    struct Item {
        const unsigned char *payload;  // raw deflate data
        uint32_t payload_len, crc, isize, block_len;
        size_t out_off;
    };
    static uint32_t le32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

    // A BGZF block header: gzip magic, deflate, FLG == FEXTRA exactly, an extra field holding 'B','C',2,BSIZE.
    // -> kOk with header length and total block length, kNo, or kNeedMore when `avail` bytes cannot tell.
    enum { kOk = 0, kNo = 1, kNeedMore = 2 };
    static int parse_header(const unsigned char *h, size_t avail, uint32_t &hdr_len, uint32_t &block_len)
    {
        static const unsigned char fixed[4] = {0x1f, 0x8b, 8, 4};
        for (size_t i = 0; i < 4 && i < avail; ++i)
            if (h[i] != fixed[i]) return kNo;
        if (avail < 12) return kNeedMore;
        const uint32_t xlen = (uint32_t)h[10] | ((uint32_t)h[11] << 8);
        if (xlen < 6 || xlen > 256) return kNo;
        if (12 + xlen > avail) return kNeedMore;
        uint32_t p = 12, bsize = 0;
        bool found = false;
        while (p + 4 <= 12 + xlen) {
            const uint32_t slen = (uint32_t)h[p + 2] | ((uint32_t)h[p + 3] << 8);
            if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) {
                bsize = (uint32_t)h[p + 4] | ((uint32_t)h[p + 5] << 8);
                found = true;
            }
            p += 4 + slen;
        }
        if (!found || p != 12 + xlen) return kNo;
        hdr_len = 12 + xlen;
        block_len = bsize + 1;
        return block_len >= hdr_len + 8 ? kOk : kNo;
    }

    static bool inflate_one(const Item &it, char *out)
    {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = const_cast<unsigned char *>(it.payload);
        zs.avail_in = it.payload_len;
        unsigned char dummy;  // an empty block (the end-of-file marker) still gets room to finish in
        zs.next_out = it.isize ? (unsigned char *)out : &dummy;
        zs.avail_out = it.isize ? it.isize : 1;
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.avail_in == 0 && zs.total_out == it.isize;
        inflateEnd(&zs);
        if (!ok) return false;
        return (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const unsigned char *)out, it.isize) == it.crc;
    }

    ssize_t pread_full(unsigned char *dst, size_t n, uint64_t off) const
    {
        size_t got = 0;
        while (got < n) {
            ssize_t r;
            do r = pread(fd_, dst + got, n - got, (off_t)(off + got));
            while (r < 0 && errno == EINTR);
            if (r < 0) return got ? (ssize_t)got : -1;
            if (r == 0) break;
            got += (size_t)r;
        }
        return (ssize_t)got;
    }

    int fd_ = -1;
    bool active_ = false, pending_handover_ = false;
    uint64_t off_ = 0;  // file offset of the first block not yet delivered
    std::vector<unsigned char> cbuf_, carry_;
    size_t carry_pos_ = 0, carry_len_ = 0;
    std::vector<Item> plan_;
};

}  // namespace shkhost

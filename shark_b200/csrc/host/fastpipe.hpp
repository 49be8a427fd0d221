// Parallel front end of the host pipeline: a worker pool, memory-mapped plain FASTQ
// files scanned by several threads at once, and RecordSource - one input file's kseq_read() outcome
// stream (kseq.h:177-218) with a bulk interface for the chunk builder.
//
// Exactness of the parallel scan.  The file is cut into segments at GUESSED record starts (a line that
// begins with '@' whose next-but-one line begins with '+').  Every worker parses strict four-line records
// (parse_strict, ingest.hpp) from its guess up to the next guess.  A segment's records are accepted only
// if the previous accepted segment ended EXACTLY on its start - by induction the accepted records are the
// ones a single sequential scan yields; a segment whose guess was wrong is parsed again from the true
// position.  The first byte that is not a strict record (FASTA, wrapped lines, '\r', a truncated tail,
// the end of the file) hands the rest of the file to the sequential FastqScanner at that offset, i.e. to
// the same character-level state machine as before.  Compressed inputs do not come here at all: they keep
// the streaming scanner (zlib / parallel BGZF inflate).
#pragma once
#include <sys/mman.h>
#include <sys/stat.h>

#include <atomic>
#include <functional>

#include "ingest.hpp"

namespace shkhost {

// Threads the host stages may use: SHK_HOST_THREADS, else the cores this process may run on (at most 32).
inline int host_threads()
{
    static const int n = [] {
        if (const char *ev = getenv("SHK_HOST_THREADS")) {
            const int v = atoi(ev);
            if (v >= 1 && v <= 256) return v;
        }
        int c = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) c = CPU_COUNT(&set);
        return c < 1 ? 1 : (c > 32 ? 32 : c);
    }();
    return n;
}

// run(n, fn): fn(i) for i in [0, n) on the pool's threads and the caller; returns when all are done.  Several
// threads may call run() at the same time (the scanners, the chunk builder and the writer share one pool).
class WorkPool {
public:
    static WorkPool &instance()
    {
        static WorkPool *p = new WorkPool(host_threads() - 1);  // lives for the process
        return *p;
    }
    explicit WorkPool(int n_workers)
    {
        for (int i = 0; i < n_workers; ++i) {
            try {
                workers_.emplace_back([this] { loop(); });
            } catch (...) {
                break;
            }
        }
    }
    int size() const { return (int)workers_.size() + 1; }
    void run(size_t n, const std::function<void(size_t)> &fn)
    {
        if (n == 0) return;
        if (n == 1 || workers_.empty()) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        job->n = n;
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(job);
        }
        cv_.notify_all();
        work_on(*job);
        std::unique_lock<std::mutex> lk(job->mu);
        job->cv.wait(lk, [&] { return job->done.load() == job->n; });
    }

private:
    struct Job {
        const std::function<void(size_t)> *fn = nullptr;
        size_t n = 0;
        std::atomic<size_t> next{0}, done{0};
        std::mutex mu;
        std::condition_variable cv;
    };
    void work_on(Job &j)
    {
        for (;;) {
            const size_t i = j.next.fetch_add(1);
            if (i >= j.n) return;
            (*j.fn)(i);
            if (j.done.fetch_add(1) + 1 == j.n) {
                std::lock_guard<std::mutex> lk(j.mu);
                j.cv.notify_all();
            }
        }
    }
    void loop()
    {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                for (;;) {
                    while (!jobs_.empty() && jobs_.front()->next.load() >= jobs_.front()->n) jobs_.pop_front();
                    if (!jobs_.empty()) break;
                    cv_.wait(lk);
                }
                job = jobs_.front();
            }
            work_on(*job);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> jobs_;
};

// A plain (not gzip) regular file mapped read-only.
class MappedFile {
public:
    explicit MappedFile(const char *path)
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        unsigned char magic[2] = {0, 0};
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
            !(pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                base_ = (const char *)m;
                size_ = (size_t)st.st_size;
                madvise(m, size_, MADV_SEQUENTIAL);
            }
        }
        close(fd);
    }
    ~MappedFile()
    {
        if (base_) munmap((void *)base_, size_);
    }
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
    bool ok() const { return base_ != nullptr; }
    const char *data() const { return base_; }
    size_t size() const { return size_; }

private:
    const char *base_ = nullptr;
    size_t size_ = 0;
};

struct Span {
    const Rec *recs;
    size_t n;
};

// One input file's outcome stream.  Plain regular files are mapped and scanned in parallel; everything else
// (gzip, BGZF, pipes) is scanned by one thread through FastqScanner.  Either way a producer thread fills a
// bounded queue of blocks; the consumer reads outcome by outcome (peek / consume, the exact path) or in bulk
// (clean_run / take).
class RecordSource {
public:
    explicit RecordSource(const char *path, size_t segment_bytes = 0) : path_(path), q_((size_t)std::max(8, 2 * host_threads() + 2))
    {
        if (const char *ev = getenv("SHK_SCAN_SEGMENT")) segment_bytes = (size_t)atol(ev);
        seg_ = segment_bytes ? segment_bytes : (size_t)(8u << 20);
        const bool want_map = !(getenv("SHK_NO_MMAP") && atoi(getenv("SHK_NO_MMAP")) != 0);
        if (want_map) map_.reset(new MappedFile(path));
        if (map_ && !map_->ok()) map_.reset();
        if (!map_) {
            stream_.reset(new FastqScanner(path));
            ok_ = stream_->ok();
        } else {
            ok_ = true;
        }
    }
    ~RecordSource()
    {
        if (th_.joinable()) {
            abort_.store(true);
            while (!done_) fetch();  // drain so that a producer blocked on a full queue can finish
            th_.join();
        }
    }
    RecordSource(const RecordSource &) = delete;
    RecordSource &operator=(const RecordSource &) = delete;
    bool ok() const { return ok_; }
    bool parallel() const { return (bool)map_; }
    void start()
    {
        th_ = std::thread([this] {
            if (map_) produce_mapped();
            else produce_stream(*stream_);
        });
    }

    // ---- outcome by outcome (RecordStream's interface) -----------------------------------------------
    const Rec &peek()
    {
        settle();
        return blocks_.front()->recs[idx_];
    }
    const std::shared_ptr<Block> &block()
    {
        settle();
        return blocks_.front();
    }
    void consume()
    {
        settle();
        const std::shared_ptr<Block> &b = blocks_.front();
        const int32_t st = b->recs[idx_].status;
        if (done_ && blocks_.size() == 1 && idx_ + 1 >= b->recs.size() && (st == -1 || st == -3)) return;  // sticky end
        ++idx_;
    }

    // ---- bulk --------------------------------------------------------------------------------------------
    // Number of outcomes from the cursor on, at most `want`, that are plain records (status >= 0) in blocks
    // without NUL bytes.  Blocks while the producer is behind.
    size_t clean_run(size_t want)
    {
        settle();
        size_t have = 0, b = 0;
        while (have < want) {
            if (b == blocks_.size()) {
                if (done_) break;
                fetch();
                continue;
            }
            const Block &blk = *blocks_[b];
            const size_t from = b == 0 ? idx_ : 0;
            ++b;
            if (blk.all_ok) {
                have += blk.recs.size() - from;
                continue;
            }
            if (blk.has_nul) break;
            size_t i = from;
            while (i < blk.recs.size() && blk.recs[i].status >= 0) ++i;
            have += i - from;
            if (i < blk.recs.size()) break;
        }
        return have < want ? have : want;
    }
    // The spans of the next n <= clean_run(...) records, without moving the cursor.
    void peek_run(size_t n, std::vector<Span> &spans, std::vector<std::shared_ptr<Block>> &keep)
    {
        settle();
        size_t from = idx_;
        for (size_t b = 0; n && b < blocks_.size(); ++b, from = 0) {
            const std::shared_ptr<Block> &blk = blocks_[b];
            const size_t m = std::min(n, blk->recs.size() - from);
            spans.push_back(Span{blk->recs.data() + from, m});
            keep.push_back(blk);
            n -= m;
        }
    }
    // Takes n <= clean_run(...) records: their spans and the blocks that own them.
    void take(size_t n, std::vector<Span> &spans, std::vector<std::shared_ptr<Block>> &keep)
    {
        while (n) {
            settle();
            const std::shared_ptr<Block> &b = blocks_.front();
            const size_t m = std::min(n, b->recs.size() - idx_);
            spans.push_back(Span{b->recs.data() + idx_, m});
            keep.push_back(b);
            idx_ += m;
            n -= m;
        }
    }

private:
    // cursor on a valid outcome: drops used-up blocks, waits for the next one
    void settle()
    {
        for (;;) {
            if (blocks_.empty()) {
                fetch();
                continue;
            }
            if (idx_ < blocks_.front()->recs.size()) return;
            if (done_ && blocks_.size() == 1) {  // sticky end: stay on the final outcome
                idx_ = blocks_.front()->recs.size() - 1;
                return;
            }
            blocks_.pop_front();
            idx_ = 0;
        }
    }
    void fetch()
    {
        if (done_) return;
        std::shared_ptr<Block> b = q_.pop();
        const int32_t st = b->recs.empty() ? 0 : b->recs.back().status;
        if (st == -1 || st == -3) done_ = true;
        if (!b->recs.empty()) blocks_.push_back(std::move(b));
    }

    // ---- producers ---------------------------------------------------------------------------------------
    void produce_stream(FastqScanner &sc)
    {
        for (;;) {
            std::unique_ptr<Block> b = sc.next();
            const bool end = !b->recs.empty() && (b->recs.back().status == -1 || b->recs.back().status == -3);
            q_.push(std::shared_ptr<Block>(b.release()));
            if (end) break;  // the consumer repeats the final outcome itself
        }
    }

    struct Segment {
        size_t start = 0, limit = 0;  // parse from start until pos >= limit
        std::vector<Rec> recs;
        size_t end = 0;      // position after the last record parsed
        bool stopped = false;  // a non-strict record (or the end of the file) at `end`
        bool has_nul = false;
    };
    // a line start at or after `from` that looks like the first line of a record
    size_t guess_start(size_t from) const
    {
        const char *base = map_->data();
        const size_t size = map_->size();
        if (from == 0) return 0;
        if (from >= size) return size;
        const char *nl = (const char *)memchr(base + from - 1, '\n', size - (from - 1));
        size_t p = nl ? (size_t)(nl - base) + 1 : size;
        for (int tries = 0; tries < 8 && p < size; ++tries) {
            const char *l1 = (const char *)memchr(base + p, '\n', size - p);
            if (!l1) return size;
            const char *l2 = (const char *)memchr(l1 + 1, '\n', size - (size_t)(l1 + 1 - base));
            if (base[p] == '@' && l2 && (size_t)(l2 + 1 - base) < size && l2[1] == '+') return p;
            p = (size_t)(l1 - base) + 1;
        }
        return p;
    }
    void parse_segment(Segment &s) const
    {
        const char *base = map_->data();
        const size_t size = map_->size();
        size_t pos = s.start;
        s.recs.clear();
        s.recs.reserve((s.limit - s.start) / 180 + 16);
        s.stopped = false;
        while (pos < s.limit) {
            Rec r;
            size_t next = 0;
            if (parse_strict(base, pos, size, r, next) != kStrictOk) {
                s.stopped = true;
                break;
            }
            s.recs.push_back(r);
            pos = next;
        }
        s.end = pos;
        s.has_nul = pos > s.start && memchr(base + s.start, 0, pos - s.start) != nullptr;
    }
    void emit(Segment &s)
    {
        if (s.recs.empty()) return;
        std::shared_ptr<Block> b(new Block);
        b->recs.swap(s.recs);
        b->has_nul = s.has_nul;
        b->all_ok = !s.has_nul;
        q_.push(std::move(b));
    }
    void produce_mapped()
    {
        const size_t size = map_->size();
        const int width = std::max(2, WorkPool::instance().size());  // segments scanned at once
        size_t true_pos = 0, next_boundary = 0;
        bool handed_over = false;
        std::vector<Segment> segs((size_t)width);
        while (true_pos < size && !handed_over && !abort_.load()) {
            // a wave of segments with guessed starts
            size_t n = 0;
            size_t g = true_pos;
            next_boundary = std::max(next_boundary, true_pos);
            for (; n < (size_t)width && g < size; ++n) {
                next_boundary += seg_;
                const size_t g_next = next_boundary >= size ? size : guess_start(next_boundary);
                segs[n].start = g;
                segs[n].limit = std::max(g_next, g);  // (a line longer than a segment: an empty segment, harmless)
                g = segs[n].limit;
            }
            WorkPool::instance().run(n, [&](size_t i) { parse_segment(segs[i]); });
            for (size_t i = 0; i < n && !handed_over; ++i) {
                Segment &s = segs[i];
                if (s.start != true_pos) {  // the guess was not a record start: parse again from the true position
                    if (true_pos >= s.limit) continue;  // the previous segment ran past this one entirely
                    s.start = true_pos;
                    parse_segment(s);
                }
                true_pos = s.end;
                const bool stop = s.stopped;
                emit(s);
                if (stop) handed_over = true;
            }
        }
        // the rest (possibly nothing but the end of the file) through the sequential scanner: exact kseq semantics
        FastqScanner tail(path_.c_str(), 8u << 20, true_pos);
        if (!tail.ok() || abort_.load()) {
            std::shared_ptr<Block> b(new Block);
            Rec r;
            r.status = -1;
            b->recs.push_back(r);
            q_.push(std::move(b));
            return;
        }
        produce_stream(tail);
    }

    std::string path_;
    std::unique_ptr<MappedFile> map_;
    std::unique_ptr<FastqScanner> stream_;
    bool ok_ = false;
    size_t seg_ = 8u << 20;
    BoundedQueue<std::shared_ptr<Block>> q_;
    std::thread th_;
    std::atomic<bool> abort_{false};
    std::deque<std::shared_ptr<Block>> blocks_;
    size_t idx_ = 0;
    bool done_ = false;
};

}  // namespace shkhost

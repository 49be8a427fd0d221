// Host-side test and measurement tool (no device code, no CUDA): used by tests/test_host_ingest.py
// and for tuning the ingest/output stages on a machine without a GPU.  Test infrastructure: fastx.hpp (next to
// this file) is the record-at-a-time restatement of kseq.h:177-218 that the product's scanners are checked against.
//   host_tools scan-check FILE [BLOCK_BYTES]   FastqScanner vs FastxReader, outcome by outcome
//   host_tools pscan-check FILE [SEGMENT_BYTES]  RecordSource (parallel scan of the mapped file + sequential tail)
//                                               vs FastxReader, outcome by outcome, through the bulk AND the
//                                               outcome-by-outcome interface
//   host_tools scan-dump FILE [BLOCK_BYTES [N]]  count / end status of the records the scanner yields, hash of the first N
//   host_tools ingest-bench FQ1 [FQ2] [--qual]  scan -> batcher (packing included) -> writer with every read kept
//                                               (results faked: this measures the host stages only)
//   host_tools pipe-check FQ1 [FQ2] [--qual Q] [--chunk N] [--bytes B]
//                                               batcher -> writer with every read reported for gene 0, chunks of N
//                                               reads: prints what the reference's ReadOutput would print for
//                                               such results (checked by the tests against a Python restatement)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "fastx.hpp"
#include "pipeline.hpp"

using namespace shkhost;

struct MallocAlloc {
    static void *alloc(size_t n) { return malloc(n); }
    static void free(void *p) { ::free(p); }
};

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// shk_host_pack's definition (include/shark_b200.h), restated: the library is not linked here
static void pack_scalar(const uint8_t *seq, const uint8_t *qual, int32_t min_quality, uint64_t n, uint64_t *codes, uint32_t *valid)
{
    const int mq = (int)(signed char)(unsigned char)((min_quality & 0xFF) + 33);
    const bool masking = (min_quality & 0xFF) != 0 && qual;
    for (uint64_t g = 0; g < (n + 31) / 32; ++g) {
        uint64_t c = 0;
        uint32_t v = 0;
        for (uint64_t i = g * 32; i < n && i < g * 32 + 32; ++i) {
            uint32_t ch = seq[i];
            if (masking && (int)(signed char)qual[i] < mq) ch = (ch - 64u) & 0xFFu;
            const uint32_t u = (ch | 0x20u) - 0x61u;
            if (u < 32u && ((0x00080045u >> u) & 1u)) {
                v |= 1u << (i - g * 32);
                c |= (uint64_t)((ch >> 1) & 3u) << (2 * (i - g * 32));
            }
        }
        codes[g] = c;
        valid[g] = v;
    }
}

static bool same_record(const Rec &r, long st, const std::string &n, const std::string &s, const std::string &q)
{
    if (st != r.status) return false;
    if (st < 0) return true;
    return n.size() == r.name_len && !memcmp(n.data(), r.name, r.name_len) && s.size() == r.seq_len &&
           !memcmp(s.data(), r.seq, r.seq_len) && q.size() == r.qual_len && !memcmp(q.data(), r.qual, r.qual_len);
}

// RecordSource against the record reader: `bulk` > 0 alternates bulk takes of up to that many records with
// single outcomes (the two interfaces share one cursor).
static int pscan_check(const char *path, size_t segment_bytes, size_t bulk)
{
    FastxReader ref(path, 1u << 16);
    RecordSource src(path, segment_bytes);
    if (!ref.ok() || !src.ok()) {
        printf("OPEN_FAILED\n");
        return ref.ok() == src.ok() ? 0 : 1;
    }
    src.start();
    std::string n, s, q;
    uint64_t count = 0, bulk_taken = 0;
    int ends = 0;
    while (ends < 3) {
        if (bulk) {
            const size_t m = src.clean_run(bulk);
            std::vector<Span> spans;
            std::vector<std::shared_ptr<Block>> keep;
            src.take(m, spans, keep);
            for (const Span &sp : spans)
                for (size_t i = 0; i < sp.n; ++i) {
                    const long st = ref.read(n, s, q);
                    if (!same_record(sp.recs[i], st, n, s, q)) {
                        printf("MISMATCH (bulk) outcome %llu\n", (unsigned long long)count);
                        return 1;
                    }
                    ++count;
                }
            bulk_taken += m;
        }
        const Rec r = src.peek();
        src.consume();
        const long st = ref.read(n, s, q);
        if (!same_record(r, st, n, s, q)) {
            printf("MISMATCH outcome %llu: status %ld vs %d\n", (unsigned long long)count, st, r.status);
            return 1;
        }
        ++count;
        if (st == -1 || st == -3) ++ends;
    }
    printf("OK %llu outcomes (%llu in bulk, %s)\n", (unsigned long long)count, (unsigned long long)bulk_taken,
           src.parallel() ? "mapped" : "stream");
    return 0;
}

static int scan_check(const char *path, size_t block_bytes)
{
    FastxReader ref(path, 1u << 16);
    FastqScanner sc(path, block_bytes);
    if (!ref.ok() || !sc.ok()) {
        printf("OPEN_FAILED\n");
        return ref.ok() == sc.ok() ? 0 : 1;
    }
    std::string n, s, q;
    uint64_t count = 0, fast = 0;
    int ends = 0;
    for (;;) {
        std::unique_ptr<Block> b = sc.next(1000);
        for (const Rec &r : b->recs) {
            const long st = ref.read(n, s, q);
            if (st != r.status) {
                printf("MISMATCH outcome %llu: status %ld vs %d\n", (unsigned long long)count, st, r.status);
                return 1;
            }
            if (st >= 0) {
                if (n.size() != r.name_len || memcmp(n.data(), r.name, r.name_len) || s.size() != r.seq_len ||
                    memcmp(s.data(), r.seq, r.seq_len) || q.size() != r.qual_len || memcmp(q.data(), r.qual, r.qual_len)) {
                    printf("MISMATCH outcome %llu: fields differ (name '%s')\n", (unsigned long long)count, n.c_str());
                    return 1;
                }
                bool in_buf = false;
                for (auto &bf : b->bufs) in_buf = in_buf || (r.seq >= bf.get());
                (void)in_buf;
            }
            ++count;
            if (st == -1 || st == -3) ++ends;
        }
        if (b->arena.empty()) fast += b->recs.size();
        if (ends >= 3) break;  // the end outcome is sticky on both sides
    }
    printf("OK %llu outcomes\n", (unsigned long long)count);
    return 0;
}

// Records up to the first non-record outcome: count, that outcome's status, FNV-1a over name \0 seq \0 qual \0 of
// the first `hash_first` records.
static int scan_dump(const char *path, size_t block_bytes, uint64_t hash_first)
{
    FastqScanner sc(path, block_bytes);
    if (!sc.ok()) {
        printf("OPEN_FAILED\n");
        return 0;
    }
    uint64_t h = 0xCBF29CE484222325ull, count = 0;
    auto eat = [&](const char *p, uint32_t n) {
        for (uint32_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 0x100000001B3ull;
        h = (h ^ 0) * 0x100000001B3ull;
    };
    for (;;) {
        std::unique_ptr<Block> b = sc.next(1000);
        for (const Rec &r : b->recs) {
            if (r.status < 0) {
                printf("DUMP count=%llu last=%d hash=%016llx\n", (unsigned long long)count, r.status, (unsigned long long)h);
                return 0;
            }
            if (count < hash_first) {
                eat(r.name, r.name_len);
                eat(r.seq, r.seq_len);
                eat(r.qual, r.qual_len);
            }
            ++count;
        }
    }
}

static int scan_bench(const char *path)
{
    const double t0 = now();
    FastqScanner sc(path);
    uint64_t n = 0, fastb = 0;
    for (;;) {
        std::unique_ptr<Block> b = sc.next();
        n += b->recs.size();
        if (b->arena.empty()) ++fastb;
        const int st = b->recs.back().status;
        if (st == -1 || st == -3) break;
    }
    const double t = now() - t0;
    printf("outcomes %llu in %.3fs (%.2f M/s), all-fast blocks %llu\n", (unsigned long long)n, t, n / t / 1e6, (unsigned long long)fastb);
    return 0;
}

static int ingest_bench(const char *f1, const char *f2, bool with_qual)
{
    const double t0 = now();
    Batcher<MallocAlloc> batcher(f1, f2, with_qual ? 20 : 0, pack_scalar);
    if (!batcher.files_ok()) return 1;
    batcher.start();
    std::vector<std::string> legend{"gene00000"};
    FILE *null = fopen("/dev/null", "w");
    const int fdn = fileno(null);
    const char *o1 = getenv("OUT1");
    int fd1 = o1 ? open(o1, O_RDWR | O_CREAT | O_TRUNC, 0666) : fdn;
    Writer<MallocAlloc> writer(fdn, fd1, f2 ? fdn : -1, legend, f2 != nullptr);
    Chunk<MallocAlloc> ch[2];
    uint64_t reads = 0, bytes = 0, bulk = 0;
    double t_fill = 0, t_write = 0;
    for (int i = 0;; i ^= 1) {
        double a = now();
        const bool more = batcher.fill(ch[i], 1000000, 640000000ull);
        double b = now();
        t_fill += b - a;
        ch[i].gene16_v.assign(ch[i].n, 0);
        ch[i].results_from_vectors();
        a = now();
        writer.write(ch[i]);
        t_write += now() - a;
        reads += ch[i].n;
        bytes += ch[i].bytes;
        bulk += ch[i].bulk ? ch[i].n : 0;
        if (!more) break;
    }
    writer.flush();
    const double t = now() - t0;
    printf("reads %llu (%llu in bulk chunks) bases %llu total %.3fs fill(wait+pack) %.3fs write %.3fs -> %.2f M reads/s, %d threads\n",
           (unsigned long long)reads, (unsigned long long)bulk, (unsigned long long)bytes, t, t_fill, t_write, reads / t / 1e6,
           host_threads());
    return 0;
}

// Batcher -> Writer with faked results: read i is reported for gene 0, except that every 7th read is dropped and
// every 5th goes through the multi list with genes 0 and 1.  stdout = ssv, OUT1 / OUT2 = the FASTQ files, stderr =
// the packed text (one line per chunk: n, bytes, FNV-1a of offsets, codes and validity words).
static int pipe_check(const char *f1, const char *f2, int min_quality, unsigned chunk, uint64_t max_bytes)
{
    Batcher<MallocAlloc> batcher(f1, f2, min_quality, pack_scalar);
    if (!batcher.files_ok()) return 1;
    batcher.start();
    std::vector<std::string> legend{"gA", "gB"};
    const char *o1 = getenv("OUT1"), *o2 = getenv("OUT2");
    const int fd1 = o1 ? open(o1, O_RDWR | O_CREAT | O_TRUNC, 0666) : -1;
    const int fd2 = (o2 && f2) ? open(o2, O_RDWR | O_CREAT | O_TRUNC, 0666) : -1;
    Writer<MallocAlloc> writer(STDOUT_FILENO, fd1, fd2, legend, f2 != nullptr);
    if (getenv("PREFAULT")) {  // the CLI's pre-faulting of mapped outputs, with the input sizes as bounds (+ slack: FASTA records grow)
        auto size_of = [](const char *p) -> uint64_t {
            struct stat st;
            return p && stat(p, &st) == 0 ? (uint64_t)st.st_size * 2 + 4096 : 0;
        };
        const uint64_t expect[3] = {0, fd1 >= 0 ? size_of(f1) : 0, fd2 >= 0 ? size_of(f2) : 0};
        writer.start_prefault(expect, 3);
    }
    Chunk<MallocAlloc> ch;
    uint64_t global = 0;
    for (;;) {
        const bool more = batcher.fill(ch, chunk, max_bytes);
        if (batcher.error()) {
            fprintf(stderr, "ERROR %s\n", batcher.error());
            return 3;
        }
        ch.gene16_v.assign(ch.n, 0);
        for (uint32_t r = 0; r < ch.n; ++r, ++global) {
            if (global % 7 == 3) ch.gene16_v[r] = kGeneNone;
            else if (global % 5 == 1) {
                ch.gene16_v[r] = kGeneMulti;
                ch.multi_v.push_back(AssocPair{r, 0});
                ch.multi_v.push_back(AssocPair{r, 1});
            }
        }
        ch.results_from_vectors();
        if (ch.n) {
            uint64_t h = 0xCBF29CE484222325ull;
            auto eat = [&](const void *p, size_t n) {
                for (size_t i = 0; i < n; ++i) h = (h ^ ((const unsigned char *)p)[i]) * 0x100000001B3ull;
            };
            eat(ch.off.p, ((size_t)ch.n + 1) * 4);
            eat(ch.codes.p, ch.groups() * 8);
            eat(ch.valid.p, ch.groups() * 4);
            std::string bs;
            for (uint32_t b : ch.batch_start) bs += std::to_string(b) + ",";
            fprintf(stderr, "CHUNK n=%u bytes=%llu bulk=%d batches=%s hash=%016llx\n", ch.n, (unsigned long long)ch.bytes, (int)ch.bulk,
                    bs.c_str(), (unsigned long long)h);
        }
        writer.write(ch);
        if (!more) break;
    }
    writer.flush();
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 3 && std::string(argv[1]) == "scan-check") return scan_check(argv[2], argc > 3 ? (size_t)atol(argv[3]) : (8u << 20));
    if (argc >= 3 && std::string(argv[1]) == "pscan-check")
        return pscan_check(argv[2], argc > 3 ? (size_t)atol(argv[3]) : (8u << 20), argc > 4 ? (size_t)atol(argv[4]) : 0);
    if (argc >= 3 && std::string(argv[1]) == "pipe-check") {
        const char *f2 = nullptr;
        int q = 0;
        unsigned chunk = 1000000;
        uint64_t max_bytes = 640000000ull;
        for (int i = 3; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--qual" && i + 1 < argc) q = atoi(argv[++i]);
            else if (a == "--chunk" && i + 1 < argc) chunk = (unsigned)atol(argv[++i]);
            else if (a == "--bytes" && i + 1 < argc) max_bytes = strtoull(argv[++i], nullptr, 10);
            else f2 = argv[i];
        }
        return pipe_check(argv[2], f2, q, chunk, max_bytes);
    }
    if (argc >= 3 && std::string(argv[1]) == "scan-dump") return scan_dump(argv[2], argc > 3 ? (size_t)atol(argv[3]) : (8u << 20), argc > 4 ? strtoull(argv[4], nullptr, 10) : ~0ull);
    if (argc >= 3 && std::string(argv[1]) == "scan-bench") return scan_bench(argv[2]);
    if (argc >= 3 && std::string(argv[1]) == "ingest-bench") {
        const char *f2 = nullptr;
        bool q = false;
        for (int i = 3; i < argc; ++i) {
            if (std::string(argv[i]) == "--qual") q = true;
            else f2 = argv[i];
        }
        return ingest_bench(argv[2], f2, q);
    }
    fprintf(stderr, "usage: host_tools scan-check FILE [BLOCK_BYTES] | ingest-bench FQ1 [FQ2] [--qual]\n");
    return 2;
}

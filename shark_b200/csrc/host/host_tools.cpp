// Host-side test and measurement tool (no device code, no CUDA): used by tests/test_host_ingest.py
// and for tuning the ingest/output stages on a machine without a GPU.
//   host_tools scan-check FILE [BLOCK_BYTES]   FastqScanner vs FastxReader, outcome by outcome
//   host_tools scan-dump FILE [BLOCK_BYTES [N]]  count / end status of the records the scanner yields, hash of the first N
//   host_tools ingest-bench FQ1 [FQ2] [--qual]  scanner -> batcher -> writer with every read kept
//                                               (results faked: this measures the host stages only)
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
    Batcher<MallocAlloc> batcher(f1, f2, with_qual);
    if (!batcher.files_ok()) return 1;
    batcher.start();
    std::vector<std::string> legend{"gene00000"};
    FILE *null = fopen("/dev/null", "w");
    const int fdn = fileno(null);
    const char *o1 = getenv("OUT1");
    int fd1 = o1 ? open(o1, O_WRONLY | O_CREAT | O_TRUNC, 0666) : fdn;
    Writer<MallocAlloc> writer(fdn, fd1, f2 ? fdn : -1, legend, f2 != nullptr);
    Chunk<MallocAlloc> ch[2];
    uint64_t reads = 0, bytes = 0;
    double t_fill = 0, t_write = 0;
    for (int i = 0;; i ^= 1) {
        double a = now();
        const bool more = batcher.fill(ch[i], 1000000, 640000000ull);
        double b = now();
        t_fill += b - a;
        ch[i].assoc.resize(ch[i].n);
        for (uint32_t r = 0; r < ch[i].n; ++r) ch[i].assoc[r] = AssocPair{r, 0};
        a = now();
        writer.write(ch[i]);
        t_write += now() - a;
        reads += ch[i].n;
        bytes += ch[i].bytes;
        if (!more) break;
    }
    writer.flush();
    const double t = now() - t0;
    printf("reads %llu bases %llu total %.3fs fill(wait+pack) %.3fs write %.3fs -> %.2f M reads/s\n", (unsigned long long)reads,
           (unsigned long long)bytes, t, t_fill, t_write, reads / t / 1e6);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 3 && std::string(argv[1]) == "scan-check") return scan_check(argv[2], argc > 3 ? (size_t)atol(argv[3]) : (8u << 20));
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

// shark-b200: command-line drop-in for the reference's `shark` (main.cpp:83-240) on top of the
// C ABI of libshark_b200.so.  Same options (argument_parser.hpp:29-174), same stdout ssv
// (ReadOutput.hpp:43), same filtered FASTQ files (ReadOutput.hpp:44-47), same stderr stage
// stamps.  The host does what the north star leaves on the host: FASTA/FASTQ parsing
// (fastx.hpp), batching into pinned SoA chunks (FastaSplitter.hpp / FastqSplitter.hpp) and
// output (ReadOutput.hpp); everything between goes through shk_index_build / shk_reads_submit /
// shk_reads_collect.  Extensions (not in the reference): --gpus N, --chunk-reads N.
//
// Pipeline: a parser thread fills chunk buffers; the main thread submits chunk i+1 to a free
// slot before it collects chunk i (double buffering per GPU, chunks round-robin over GPUs) and
// writes results in chunk order, which reproduces the reference's `-t 1` output order.
#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/shark_b200.h"
#include "fastx.hpp"

namespace {

// The reference's usage text is part of its observable behaviour (-h, argument errors).
const char *USAGE_MESSAGE =
    "Usage: shark -r <references> -1 <sample1> [OPTIONAL ARGUMENTS]\n"
    "\n"
    "Arguments:\n"
    "      -r, --reference                   reference sequences in FASTA format (can be gzipped)\n"
    "      -1, --sample1                     sample in FASTQ (can be gzipped)\n"
    "\n"
    "Optional arguments:\n"
    "      -h, --help                        display this help and exit\n"
    "      -2, --sample2                     second sample in FASTQ (optional, can be gzipped)\n"
    "      -o, --out1                        first output sample in FASTQ (default: sharked_sample.1)\n"
    "      -p, --out2                        second output sample in FASTQ (default: sharked_sample.2)\n"
    "      -k, --kmer-size                   size of the kmers to index (default:17, max:31)\n"
    "      -c, --confidence                  confidence for associating a read to a gene (default:0.6)\n"
    "      -b, --bf-size                     bloom filter size in GB (default:1)\n"
    "      -q, --min-base-quality            minimum base quality (assume FASTQ Illumina 1.8+ Phred scale, default:0, i.e., no filtering)\n"
    "      -s, --single                      report an association only if a single gene is found\n"
    "      -t, --threads                     number of threads (default:1)\n"
    "      -v, --verbose                     verbose mode\n";

struct Options {
    std::string fasta_path, sample1_path, sample2_path, out1_path, out2_path;
    bool paired = false;
    unsigned k = 17;
    double c = 0.6;
    uint64_t bf_size = 1ull << 33;
    char min_quality = 0;
    bool single = false, verbose = false;
    int n_threads = 1;
    int gpus = 1;                   // extension
    unsigned chunk_reads = 1000000; // extension; rounded to a multiple of the 50 000-read batch
};

// argument_parser.hpp:84-174, option by option (istringstream extraction included, so that e.g.
// "-k 17abc" parses like the reference does).
Options parse_arguments(int argc, char **argv)
{
    Options opt;
    static const char *shortopts = "t:r:1:2:o:p:k:c:b:q:svh";
    static const struct option longopts[] = {{"reference", required_argument, nullptr, 'r'},
                                             {"threads", required_argument, nullptr, 't'},
                                             {"sample1", required_argument, nullptr, '1'},
                                             {"sample2", required_argument, nullptr, '2'},
                                             {"out1", required_argument, nullptr, 'o'},
                                             {"out2", required_argument, nullptr, 'p'},
                                             {"kmer-size", required_argument, nullptr, 'k'},
                                             {"confidence", required_argument, nullptr, 'c'},
                                             {"bf-size", required_argument, nullptr, 'b'},
                                             {"min-base-quality", required_argument, nullptr, 'q'},
                                             {"single", no_argument, nullptr, 's'},
                                             {"verbose", no_argument, nullptr, 'v'},
                                             {"help", no_argument, nullptr, 'h'},
                                             {"gpus", required_argument, nullptr, 1000},
                                             {"chunk-reads", required_argument, nullptr, 1001},
                                             {nullptr, 0, nullptr, 0}};
    for (int ch; (ch = getopt_long(argc, argv, shortopts, longopts, nullptr)) != -1;) {
        std::istringstream arg(optarg != nullptr ? optarg : "");
        switch (ch) {
        case 'r': arg >> opt.fasta_path; break;
        case 't':
            arg >> opt.n_threads;
            if (opt.n_threads <= 0) {
                std::cerr << "USAGE_MESSAGE";  // sic: argument_parser.hpp:94
                std::cerr << "shark: at least 1 thread is required." << std::endl << "aborting..." << std::endl;
                exit(EXIT_FAILURE);
            }
            break;
        case '1': arg >> opt.sample1_path; break;
        case '2':
            arg >> opt.sample2_path;
            opt.paired = true;
            break;
        case 'o': arg >> opt.out1_path; break;
        case 'p': arg >> opt.out2_path; break;
        case 'k':
            arg >> opt.k;
            if (opt.k == 0 || opt.k > 31) {
                std::cerr << USAGE_MESSAGE;
                std::cerr << "shark: k must be in the range [1, 31]." << std::endl << "aborting..." << std::endl;
                exit(EXIT_FAILURE);
            }
            break;
        case 'c':
            arg >> opt.c;
            if (opt.c < 0 || opt.c > 1) {
                std::cerr << "shark: c must be in the range [0, 1]." << std::endl << "aborting..." << std::endl;
                exit(EXIT_FAILURE);
            }
            break;
        case 'b':
            arg >> opt.bf_size;
            opt.bf_size = opt.bf_size * (1ull << 33);  // "GB": argument_parser.hpp:130-134
            break;
        case 'q': {
            int mq = 0;
            arg >> mq;
            if (mq < 0) {
                std::cerr << USAGE_MESSAGE;
                std::cerr << "shark: q must be a positive value." << std::endl << "aborting..." << std::endl;
                exit(EXIT_FAILURE);
            }
            opt.min_quality = static_cast<char>(mq);
            break;
        }
        case 's': opt.single = true; break;
        case 'v': opt.verbose = true; break;
        case 'h': std::cerr << USAGE_MESSAGE; exit(EXIT_SUCCESS);
        case 1000: arg >> opt.gpus; break;
        case 1001: arg >> opt.chunk_reads; break;
        default:
            std::cerr << "shark : unknown argument" << std::endl;
            std::cerr << "\n" << USAGE_MESSAGE;
            exit(EXIT_FAILURE);
        }
    }
    if (opt.fasta_path.empty() || opt.sample1_path.empty()) {
        std::cerr << "shark : missing required files" << std::endl;
        std::cerr << "\n" << USAGE_MESSAGE;
        exit(EXIT_FAILURE);
    }
    if (opt.out1_path.empty()) opt.out1_path = "sharked_sample.1";
    if (opt.out2_path.empty() && !opt.sample2_path.empty()) opt.out2_path = "sharked_sample.2";
    if (opt.gpus < 1) opt.gpus = 1;
    return opt;
}

auto start_t = std::chrono::high_resolution_clock::now();
void pelapsed(const std::string &s)  // main.cpp:49-54
{
    auto now_t = std::chrono::high_resolution_clock::now();
    std::cerr << "[shark/" << s << "] Time elapsed "
              << std::chrono::duration_cast<std::chrono::milliseconds>(now_t - start_t).count() / 1000 << std::endl;
}

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "shark: " << msg << std::endl;
    exit(EXIT_FAILURE);
}

constexpr unsigned kBatch = 50000;  // FastqSplitter batch (main.cpp:215): ReadOutput's dedup resets per batch

// Pinned buffer from the library (falls back to nothing: allocation failure is fatal).
struct Pinned {
    uint8_t *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n)
    {
        if (n <= cap) return;
        size_t want = cap ? cap : (1u << 20);
        while (want < n) want *= 2;
        void *q = nullptr;
        if (shk_alloc_pinned(&q, want) != SHK_OK) die(std::string("cannot allocate pinned memory: ") + shk_last_error(nullptr));
        if (p) {
            memcpy(q, p, cap);
            shk_free_pinned(p);
        }
        p = (uint8_t *)q;
        cap = want;
    }
    ~Pinned()
    {
        if (p) shk_free_pinned(p);
    }
};

// One chunk of reads in the SoA layout of shk_reads_submit plus what ReadOutput needs.
struct Chunk {
    Pinned seq, qual, off;              // text (mate1 [+ 'N' + mate2]), qualities (+ 0x1B), uint32 offsets
    std::vector<char> host_qual;        // qualities when they are not sent to the GPU (min_quality == 0)
    std::vector<uint64_t> qual_pos;     // per read: start of its quality text in host_qual
    std::string names;                  // name1 '\0' [name2 '\0'] per read
    std::vector<uint64_t> name_pos;     // per read: start in names
    std::vector<uint32_t> len1;         // per read: length of mate 1 (mate 2 = total - len1 - 1)
    std::vector<uint32_t> batch_start;  // read indices where a 50 000-read batch begins
    uint32_t n = 0;
    uint64_t bytes = 0;
    bool last = false;
    uint64_t index = 0;
    void clear()
    {
        host_qual.clear();
        qual_pos.clear();
        names.clear();
        name_pos.clear();
        len1.clear();
        batch_start.clear();
        n = 0;
        bytes = 0;
        last = false;
    }
};

// FastqSplitter::operator() called until it returns an empty batch (main.cpp:66-77,
// FastqSplitter.hpp:47-93): a failed kseq_read ends the CURRENT batch only.  Fills one chunk with
// whole batches; returns false when the input is exhausted (the chunk may still hold reads).
class Batcher {
public:
    Batcher(const Options &o) : opt_(o), r1_(o.sample1_path.c_str()), with_qual_gpu_(o.min_quality != 0)
    {
        if (o.paired) r2_.reset(new shkhost::FastxReader(o.sample2_path.c_str()));
    }
    bool files_ok() const { return r1_.ok() && (!r2_ || r2_->ok()); }

    bool fill(Chunk &ch, unsigned max_reads, uint64_t max_bytes)
    {
        ch.clear();
        uint64_t last_batch_bytes = 0;
        while ((ch.n + kBatch <= max_reads && ch.bytes + last_batch_bytes + last_batch_bytes / 4 <= max_bytes) || ch.n == 0) {
            ch.batch_start.push_back(ch.n);
            const uint64_t bytes0 = ch.bytes;
            unsigned got = 0;
            while (got < kBatch) {
                if (r1_.read(n1_, s1_, q1_) < 0) break;
                if (r2_ && r2_->read(n2_, s2_, q2_) < 0) break;  // mate 1 is dropped (FastqSplitter.hpp:61)
                append(ch);
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
    // C-string semantics of the reference (FastqSplitter.hpp:56: `seq1->seq.s` as const char*)
    static size_t clen(const std::string &s)
    {
        size_t z = s.find('\0');
        return z == std::string::npos ? s.size() : z;
    }
    void append(Chunk &ch)
    {
        const size_t l1 = clen(s1_), l2 = r2_ ? clen(s2_) : 0;
        const size_t total = r2_ ? l1 + 1 + l2 : l1;
        ch.off.reserve(((size_t)ch.n + 2) * 4);
        uint32_t *off = (uint32_t *)ch.off.p;
        if (ch.n == 0) off[0] = 0;
        ch.seq.reserve(ch.bytes + total + 8);
        uint8_t *d = ch.seq.p + ch.bytes;
        memcpy(d, s1_.data(), l1);
        if (r2_) {
            d[l1] = 'N';  // FastqSplitter.hpp:63,83
            memcpy(d + l1 + 1, s2_.data(), l2);
        }
        // qualities: mask_seq walks the QUAL string (FastqSplitter.hpp:104-108); positions it does
        // not reach are never masked -> pad with 0x7f, which is not < any mq
        const size_t ql1 = clen(q1_), ql2 = r2_ ? clen(q2_) : 0;
        if (with_qual_gpu_) {
            ch.qual.reserve(ch.bytes + total + 8);
            uint8_t *q = ch.qual.p + ch.bytes;
            memset(q, 0x7f, total);
            if (!r2_) {
                memcpy(q, q1_.data(), ql1 < total ? ql1 : total);
            } else {
                // string(qual1) + "\33" + string(qual2), FastqSplitter.hpp:84
                std::string j;
                j.reserve(ql1 + 1 + ql2);
                j.append(q1_.data(), ql1).push_back('\33');
                j.append(q2_.data(), ql2);
                memcpy(q, j.data(), j.size() < total ? j.size() : total);
            }
        }
        // what ReadOutput prints: names, original sequences (from ch.seq) and quality strings
        ch.name_pos.push_back(ch.names.size());
        ch.names.append(n1_.c_str());
        ch.names.push_back('\0');
        if (r2_) {
            ch.names.append(n2_.c_str());
            ch.names.push_back('\0');
        }
        ch.qual_pos.push_back(ch.host_qual.size());
        ch.host_qual.insert(ch.host_qual.end(), q1_.data(), q1_.data() + ql1);
        ch.host_qual.push_back('\0');
        if (r2_) {
            ch.host_qual.insert(ch.host_qual.end(), q2_.data(), q2_.data() + ql2);
            ch.host_qual.push_back('\0');
        }
        ch.len1.push_back((uint32_t)l1);
        ch.bytes += total;
        ++ch.n;
        off[ch.n] = (uint32_t)ch.bytes;
    }

    const Options &opt_;
    shkhost::FastxReader r1_;
    std::unique_ptr<shkhost::FastxReader> r2_;
    bool with_qual_gpu_;
    std::string n1_, s1_, q1_, n2_, s2_, q2_;
};

// ReadOutput::operator() (ReadOutput.hpp:37-50) for one chunk's associations.
class Writer {
public:
    Writer(const Options &o, const std::vector<std::string> &legend) : legend_(legend), paired_(o.paired)
    {
        if (!o.out1_path.empty()) out1_ = fopen(o.out1_path.c_str(), "w");
        if (o.paired && !o.out2_path.empty()) out2_ = fopen(o.out2_path.c_str(), "w");
        if (out1_) setvbuf(out1_, nullptr, _IOFBF, 1 << 22);
        if (out2_) setvbuf(out2_, nullptr, _IOFBF, 1 << 22);
        setvbuf(stdout, nullptr, _IOFBF, 1 << 22);
    }
    ~Writer()
    {
        if (out1_) fclose(out1_);
        if (out2_) fclose(out2_);
        fflush(stdout);
    }
    void write(const Chunk &ch, const shk_chunk_result &res)
    {
        const uint32_t *off = (const uint32_t *)ch.off.p;
        size_t next_batch = 0;
        const char *previd = "";  // `string previd = ""` per ReadOutput call = per batch
        for (uint64_t i = 0; i < res.n_assoc; ++i) {
            const uint32_t r = res.assoc[i].read_idx, g = res.assoc[i].gene_idx;
            while (next_batch < ch.batch_start.size() && ch.batch_start[next_batch] <= r) {
                previd = "";
                ++next_batch;
            }
            const char *id1 = ch.names.data() + ch.name_pos[r];
            const char *gene = g < legend_.size() ? legend_[g].c_str() : "";
            fputs(id1, stdout);
            fputc(' ', stdout);
            fputs(gene, stdout);
            fputc('\n', stdout);
            if (strcmp(previd, id1) != 0) {
                const char *s = (const char *)ch.seq.p + off[r];
                const char *q1 = ch.host_qual.data() + ch.qual_pos[r];
                if (out1_) {
                    fputc('@', out1_);
                    fputs(id1, out1_);
                    fputc('\n', out1_);
                    fwrite(s, 1, ch.len1[r], out1_);
                    fputs("\n+\n", out1_);
                    fputs(q1, out1_);
                    fputc('\n', out1_);
                }
                if (out2_ && paired_) {
                    const char *id2 = id1 + strlen(id1) + 1;
                    const char *q2 = q1 + strlen(q1) + 1;
                    const uint32_t l2 = off[r + 1] - off[r] - ch.len1[r] - 1;
                    fputc('@', out2_);
                    fputs(id2, out2_);
                    fputc('\n', out2_);
                    fwrite(s + ch.len1[r] + 1, 1, l2, out2_);
                    fputs("\n+\n", out2_);
                    fputs(q2, out2_);
                    fputc('\n', out2_);
                }
            }
            previd = id1;
        }
    }

private:
    const std::vector<std::string> &legend_;
    bool paired_;
    FILE *out1_ = nullptr, *out2_ = nullptr;
};

// Simple blocking queue for the parser -> device hand-off.
template <class T>
class Channel {
public:
    void push(T v)
    {
        std::unique_lock<std::mutex> lk(mu_);
        q_.push_back(std::move(v));
        cv_.notify_all();
    }
    T pop()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        T v = std::move(q_.front());
        q_.pop_front();
        return v;
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

#define SHK_TRY(ctx, call)                                                              \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != SHK_OK) die(std::string(#call " failed: ") + shk_last_error(ctx));   \
    } while (0)

}  // namespace

int main(int argc, char *argv[])
{
    const Options opt = parse_arguments(argc, argv);

    if (opt.verbose) {  // main.cpp:113-123
        std::cerr << "Reference texts: " << opt.fasta_path << std::endl;
        std::cerr << "Sample 1: " << opt.sample1_path << std::endl;
        if (opt.paired) std::cerr << "Sample 2: " << opt.sample2_path << std::endl;
        std::cerr << "K-mer length: " << opt.k << std::endl;
        std::cerr << "Threshold value: " << opt.c << std::endl;
        std::cerr << "Only single associations: " << (opt.single ? "Yes" : "No") << std::endl;
        std::cerr << "Minimum base quality: " << static_cast<int>(opt.min_quality) << std::endl;
        std::cerr << std::endl;
    }

    // ---- reference: FastaSplitter (FastaSplitter.hpp:42-54) -> legend_ID + concatenated records.
    // Unlike the reference (which ignores open failures and later segfaults) we fail cleanly.
    std::vector<std::string> legend_ID;
    std::vector<uint8_t> ref_bases;
    std::vector<uint64_t> rec_off{0};
    {
        shkhost::FastxReader ref(opt.fasta_path.c_str());
        if (!ref.ok()) die("cannot open reference " + opt.fasta_path);
        std::string name, seq, qual;
        while (ref.read(name, seq, qual) >= 0) {
            legend_ID.emplace_back(name.c_str());
            const size_t l = strlen(seq.c_str());  // `seq->seq.s` as a C string (KmerBuilder input)
            ref_bases.insert(ref_bases.end(), seq.data(), seq.data() + l);
            rec_off.push_back(ref_bases.size());
        }
    }

    const unsigned chunk_reads = std::max(kBatch, opt.chunk_reads / kBatch * kBatch);
    // slot buffers are sized by this; a chunk is closed early when its text would not fit
    const uint64_t max_chunk_bytes = std::min<uint64_t>(0xF0000000ull, std::max<uint64_t>((uint64_t)chunk_reads * 640, 256ull << 20));
    std::vector<shk_ctx *> ctxs((size_t)opt.gpus, nullptr);
    for (int g = 0; g < opt.gpus; ++g) {
        shk_params p;
        memset(&p, 0, sizeof p);
        p.k = opt.k;
        p.c = opt.c;
        p.bf_bits = opt.bf_size;
        p.min_quality = (int32_t)(unsigned char)opt.min_quality;
        p.single = opt.single ? 1 : 0;
        p.device = g;
        p.n_slots = 2;
        p.max_reads_per_chunk = chunk_reads;
        p.max_bytes_per_chunk = max_chunk_bytes;
        if (shk_create(&p, &ctxs[g]) != SHK_OK) die(std::string("shk_create: ") + shk_last_error(nullptr));
    }
    shk_index_info info;
    SHK_TRY(ctxs[0], shk_index_build(ctxs[0], ref_bases.data(), rec_off.data(), (uint32_t)legend_ID.size(), &info));
    pelapsed("Transcript file processed");
    pelapsed("First switch performed");
    pelapsed("BF created from transcripts (" + std::to_string(info.n_genes) + " genes)");
    for (int g = 1; g < opt.gpus; ++g) SHK_TRY(ctxs[g], shk_index_replicate(ctxs[0], ctxs[g]));
    pelapsed("Second switch performed");
    std::vector<uint8_t>().swap(ref_bases);

    // ---- sample stage
    Batcher batcher(opt);
    if (!batcher.files_ok()) die("cannot open sample file(s)");
    Writer writer(opt, legend_ID);

    const size_t n_bufs = (size_t)opt.gpus * 2 + 2;
    std::vector<std::unique_ptr<Chunk>> pool;
    Channel<Chunk *> free_q, ready_q;
    for (size_t i = 0; i < n_bufs; ++i) {
        pool.emplace_back(new Chunk);
        free_q.push(pool.back().get());
    }
    std::thread parser([&] {
        uint64_t idx = 0;
        for (;;) {
            Chunk *ch = free_q.pop();
            const bool more = batcher.fill(*ch, chunk_reads, max_chunk_bytes);
            ch->index = idx++;
            ch->last = !more;
            ready_q.push(ch);
            if (!more) break;
        }
    });

    struct InFlight {
        Chunk *ch;
        int gpu;
        uint32_t slot;
    };
    std::deque<InFlight> inflight;
    const size_t max_inflight = (size_t)opt.gpus * 2;
    uint64_t submitted = 0;
    auto drain_one = [&] {
        InFlight f = inflight.front();
        inflight.pop_front();
        shk_chunk_result res;
        SHK_TRY(ctxs[f.gpu], shk_reads_collect(ctxs[f.gpu], f.slot, &res));
        writer.write(*f.ch, res);
        free_q.push(f.ch);
    };
    for (bool done = false; !done;) {
        Chunk *ch = ready_q.pop();
        done = ch->last;
        if (ch->n == 0) {
            free_q.push(ch);
            continue;
        }
        if (inflight.size() == max_inflight) drain_one();
        const int gpu = (int)(submitted % (uint64_t)opt.gpus);
        const uint32_t slot = (uint32_t)((submitted / (uint64_t)opt.gpus) % 2);
        SHK_TRY(ctxs[gpu], shk_reads_submit(ctxs[gpu], slot, ch->seq.p, opt.min_quality != 0 ? ch->qual.p : nullptr,
                                            (const uint32_t *)ch->off.p, ch->n));
        inflight.push_back({ch, gpu, slot});
        ++submitted;
    }
    while (!inflight.empty()) drain_one();
    parser.join();
    pelapsed("Sample completed");
    for (auto *c : ctxs) shk_destroy(c);
    pelapsed("Association done");
    return 0;
}

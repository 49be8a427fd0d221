// Host-side packing of read text (see shk_hostpack.h).  Plain C++ (g++), no CUDA.
#include "shk_hostpack.h"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

namespace shk {

namespace {

inline bool valid_base(uint32_t ch)
{
    const uint32_t u = (ch | 0x20u) - 0x61u;  // 'a' -> 0
    return u < 32u && ((0x00080045u >> u) & 1u);  // a, c, g, t
}

// groups [g0, g1) of 32 bytes; the last group of the text may be partial (n)
void pack_scalar(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t g0, uint64_t g1, uint64_t *codes,
                 uint32_t *valid)
{
    for (uint64_t g = g0; g < g1; ++g) {
        uint64_t c = 0;
        uint32_t v = 0;
        const uint64_t i0 = g * 32, i1 = i0 + 32 < n ? i0 + 32 : n;
        for (uint64_t i = i0; i < i1; ++i) {
            uint32_t ch = seq[i];
            if (qual && (int)(signed char)qual[i] < mq) ch = (ch - 64u) & 0xFFu;  // seq[i] = seq[i] - 64
            if (valid_base(ch)) {
                v |= 1u << (i - i0);
                c |= (uint64_t)((ch >> 1) & 3u) << (2 * (i - i0));
            }
        }
        codes[g] = c;
        valid[g] = v;
    }
}

#if defined(__x86_64__)
// AVX2, 64 bases (two groups) per iteration.  Validity: one pshufb maps the low nibble of a byte to the only
// value (ch | 0x20) may have for a base with that nibble ('a' 0x61, 'c' 0x63, 't' 0x74, 'g' 0x67), one compare
// settles it.  Codes: (ch >> 1) & 3, four per byte via maddubs (1, 4) and madd (1, 16), then the 16 dwords of
// the two vectors are narrowed to 16 bytes with two packs and one cross-lane permute.  Streaming stores: the
// output is read next by the DMA engine, not by this core.  In cache this runs at 15.5 GB/s per core
// (8.8 GB/s for the four-compares / extract version it replaced).
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t g0,
                                               uint64_t g1, uint64_t *codes, uint32_t *valid)
{
    const uint64_t full = n / 32;  // groups that are complete
    const uint64_t ge = g1 < full ? g1 : full;
    const __m256i lut = _mm256_setr_epi8(0, 0x61, 0, 0x63, 0x74, 0, 0, 0x67, 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 0x61, 0, 0x63, 0x74, 0, 0, 0x67, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i k20 = _mm256_set1_epi8(0x20), k40 = _mm256_set1_epi8(0x40), k3 = _mm256_set1_epi8(3), k0f = _mm256_set1_epi8(0x0F);
    const __m256i vmq = _mm256_set1_epi8((char)mq);
    const __m256i m14 = _mm256_set1_epi16(0x0401);       // bytes (1, 4): c0 + 4 c1
    const __m256i m116 = _mm256_set1_epi32(0x00100001);  // words (1, 16): t0 + 16 t1
    const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    // mq outside the signed-char range: the comparison is constant (mq is (signed char) in practice)
    const bool all_masked = mq > 127, none_masked = mq < -128, masking = qual && !none_masked;
    static const bool nt_env = !getenv("SHK_PACK_NO_NT");
    uint64_t g = g0;
    const bool nt = nt_env && ((reinterpret_cast<uintptr_t>(codes + g) & 15) == 0) && ((reinterpret_cast<uintptr_t>(valid + g) & 7) == 0);
    for (; g + 2 <= ge; g += 2) {
        __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(seq + g * 32));
        __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(seq + g * 32 + 32));
        if (masking) {  // seq[i] -= 64 where q < mq (signed)
            const __m256i q0 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(qual + g * 32));
            const __m256i q1 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(qual + g * 32 + 32));
            const __m256i lt0 = all_masked ? _mm256_set1_epi8(-1) : _mm256_cmpgt_epi8(vmq, q0);
            const __m256i lt1 = all_masked ? _mm256_set1_epi8(-1) : _mm256_cmpgt_epi8(vmq, q1);
            v0 = _mm256_sub_epi8(v0, _mm256_and_si256(lt0, k40));
            v1 = _mm256_sub_epi8(v1, _mm256_and_si256(lt1, k40));
        }
        const __m256i ok0 = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(v0, k0f)), _mm256_or_si256(v0, k20));
        const __m256i ok1 = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(v1, k0f)), _mm256_or_si256(v1, k20));
        const uint64_t vm = (uint64_t)(uint32_t)_mm256_movemask_epi8(ok0) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(ok1) << 32);
        __m256i x0 = _mm256_and_si256(_mm256_and_si256(_mm256_srli_epi16(v0, 1), k3), ok0);
        __m256i x1 = _mm256_and_si256(_mm256_and_si256(_mm256_srli_epi16(v1, 1), k3), ok1);
        x0 = _mm256_madd_epi16(_mm256_maddubs_epi16(x0, m14), m116);
        x1 = _mm256_madd_epi16(_mm256_maddubs_epi16(x1, m14), m116);
        __m256i y = _mm256_packus_epi32(x0, x1);   // 16-bit: lane 0 = x0[0..3] x1[0..3], lane 1 = x0[4..7] x1[4..7]
        y = _mm256_packus_epi16(y, y);             // bytes:  lane 0 low half = A B (x0.lo, x1.lo), lane 1 = C D (x0.hi, x1.hi)
        y = _mm256_permutevar8x32_epi32(y, perm);  // dwords A C B D = the 16 code bytes of the two groups, in order
        const __m128i out = _mm256_castsi256_si128(y);
        if (nt) {
            _mm_stream_si128(reinterpret_cast<__m128i *>(codes + g), out);
            _mm_stream_si64(reinterpret_cast<long long *>(valid + g), (long long)vm);
        } else {
            _mm_storeu_si128(reinterpret_cast<__m128i *>(codes + g), out);
            valid[g] = (uint32_t)vm;
            valid[g + 1] = (uint32_t)(vm >> 32);
        }
    }
    if (nt) _mm_sfence();
    if (g < g1) pack_scalar(seq, qual, mq, n, g, g1, codes, valid);  // an odd group and / or the partial last one
}

// 64 bases per iteration (two groups): byte compares straight into mask registers, vpmovdb compacts the codes.
__attribute__((target("avx512f,avx512bw,avx512vl"))) void pack_avx512(const uint8_t *seq, const uint8_t *qual, int mq,
                                                                      uint64_t n, uint64_t g0, uint64_t g1, uint64_t *codes,
                                                                      uint32_t *valid)
{
    const uint64_t full = n / 32;
    const uint64_t ge = g1 < full ? g1 : full;
    uint64_t g = g0;
    const __m512i k20 = _mm512_set1_epi8(0x20), k40 = _mm512_set1_epi8(0x40), k3 = _mm512_set1_epi8(3);
    const __m512i ca = _mm512_set1_epi8('a'), cc = _mm512_set1_epi8('c'), cg = _mm512_set1_epi8('g'), ct = _mm512_set1_epi8('t');
    const __m512i vmq = _mm512_set1_epi8((char)mq);
    const __m512i m14 = _mm512_set1_epi16(0x0401), m116 = _mm512_set1_epi32(0x00100001);
    const bool all_masked = mq > 127, none_masked = mq < -128;
    for (; g + 2 <= ge; g += 2) {
        __m512i v = _mm512_loadu_si512(seq + g * 32);
        if (qual && !none_masked) {
            const __m512i q = _mm512_loadu_si512(qual + g * 32);
            const __mmask64 lt = all_masked ? ~0ull : _mm512_cmplt_epi8_mask(q, vmq);  // q < mq, signed
            v = _mm512_mask_sub_epi8(v, lt, v, k40);
        }
        const __m512i lo = _mm512_or_si512(v, k20);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(lo, ca) | _mm512_cmpeq_epi8_mask(lo, cc) | _mm512_cmpeq_epi8_mask(lo, cg) |
                             _mm512_cmpeq_epi8_mask(lo, ct);
        __m512i x = _mm512_maskz_mov_epi8(ok, _mm512_and_si512(_mm512_srli_epi16(v, 1), k3));
        x = _mm512_madd_epi16(_mm512_maddubs_epi16(x, m14), m116);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(codes + g), _mm512_cvtepi32_epi8(x));
        valid[g] = (uint32_t)ok;
        valid[g + 1] = (uint32_t)(ok >> 32);
    }
    if (g < g1) pack_avx2(seq, qual, mq, n, g, g1, codes, valid);
}

bool have_avx512()
{
    // opt-in (SHK_PACK_AVX512=1): on the B200 host (16 threads, memory-bound) it measured no faster than AVX2 at
    // the packer (50-53 vs 50-51 Gbases/s with qualities, 70-75 vs 69-85 without) and a few percent slower end to
    // end on C2 (643-665 vs 665-691 M reads/s), profiles/hostpack_r1_v13.md
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl") && !getenv("SHK_PACK_SCALAR") && !getenv("SHK_PACK_AVX2") &&
                           getenv("SHK_PACK_AVX512") && atoi(getenv("SHK_PACK_AVX512")) != 0;
    return ok;
}
#endif

bool have_avx2()
{
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx2") && !getenv("SHK_PACK_SCALAR");
    return ok;
#else
    return false;
#endif
}

void pack_groups(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t g0, uint64_t g1, uint64_t *codes,
                 uint32_t *valid)
{
#if defined(__x86_64__)
    if (have_avx512()) {
        pack_avx512(seq, qual, mq, n, g0, g1, codes, valid);
        return;
    }
    if (have_avx2()) {
        pack_avx2(seq, qual, mq, n, g0, g1, codes, valid);
        return;
    }
#endif
    pack_scalar(seq, qual, mq, n, g0, g1, codes, valid);
}

// Persistent worker pool: run(fn) executes fn on every worker and on the caller, returns when all are done.
class Pool {
  public:
    explicit Pool(int n_workers)
    {
        for (int i = 0; i < n_workers; ++i) {
            try {
                workers_.emplace_back([this] { loop(); });
            } catch (...) {  // thread limit reached: work with the threads we have
                break;
            }
        }
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    void run(const std::function<void()> &fn)
    {
        std::lock_guard<std::mutex> serial(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &fn;
            pending_ = (int)workers_.size();
            ++gen_;
        }
        cv_work_.notify_all();
        fn();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void()> *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                job = job_;
            }
            (*job)();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void()> *job_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

Pool &pool()
{
    static Pool *p = [] {
        int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;  // the cores this process may run on (cgroup / taskset), not the machine's
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) n = CPU_COUNT(&set);
#endif
        if (n < 1) n = 1;
        if (n > 32) n = 32;
        if (const char *ev = getenv("SHK_PACK_THREADS")) {
            const int v = atoi(ev);
            if (v >= 1 && v <= 256) n = v;
        }
        return new Pool(n - 1);  // lives for the process: no destruction-order trouble at exit
    }();
    return *p;
}

}  // namespace

void host_pack(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t *codes, uint32_t *valid)
{
    pack_groups(seq, qual, mq, n, 0, (n + 31) / 32, codes, valid);
}

void host_pack_parallel(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t *codes, uint32_t *valid)
{
    const uint64_t groups = (n + 31) / 32;
    constexpr uint64_t kBlock = 4096;  // groups per grab = 128 KiB of text
    if (groups <= kBlock) {
        pack_groups(seq, qual, mq, n, 0, groups, codes, valid);
        return;
    }
    std::atomic<uint64_t> next{0};
    pool().run([&] {
        for (;;) {
            const uint64_t g0 = next.fetch_add(kBlock, std::memory_order_relaxed);
            if (g0 >= groups) break;
            pack_groups(seq, qual, mq, n, g0, g0 + kBlock < groups ? g0 + kBlock : groups, codes, valid);
        }
    });
}

int host_pack_threads() { return pool().size(); }

const char *host_pack_isa()
{
#if defined(__x86_64__)
    if (have_avx512()) return "avx512";
#endif
    return have_avx2() ? "avx2" : "scalar";
}

}  // namespace shk

// Index build on the device: replaces KmerBuilder + BloomfilterFiller (pass 1), the pass-2 loop of
// main.cpp, and both BF::switch_mode transitions.  Also the stand-alone probe (BF::get_index) and
// the random-sector microbenchmark.  Citations are reference file:line.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <unistd.h>
#include <thread>
#include <utility>
#include <vector>

#include "shk_internal.h"
#include "shk_scan.cuh"

namespace shk {

static constexpr uint64_t kInvalidPos = ~0ULL;

// =============================================================================================
// Device-wide exclusive scan of 32-bit values (hand-written, three phases: tile sums, scan of
// the tile sums by one CTA, tile scan + output).  Tile = 256 threads x 16 items, each warp owns
// 512 consecutive items and walks them in 16 coalesced rounds.
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// popcount of the 224 filter bits of sector i (the rank word w[7] is not counted)
struct SectorPopIn {
    const uint32_t *sectors;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(sectors + i * 8);
        uint4 a = p[0], b = p[1];
        return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z);
    }
};
struct SectorRankOut {
    uint32_t *sectors;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t prefix) const { sectors[i * 8 + 7] = prefix; }
};
struct U32In {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return p[i]; }
};
struct U32Out {
    uint32_t *p;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t prefix) const { p[i] = prefix; }
};

template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(In in, uint64_t n, uint32_t *tile_sums)
{
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)warp * (32 * kScanItems);
    uint32_t s = 0;
#pragma unroll 4
    for (int r = 0; r < kScanItems; ++r) {
        uint64_t i = base + (uint64_t)r * 32 + lane;
        if (i < n) s += in(i);
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if (lane == 0) warp_tot[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_tot[w];
        tile_sums[blockIdx.x] = t;
    }
}

template <class In, class Out>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(In in, Out out, uint64_t n, const uint32_t *tile_sums)
{
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)warp * (32 * kScanItems);
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        uint64_t i = base + (uint64_t)r * 32 + lane;
        v[r] = i < n ? in(i) : 0;
        s += v[r];
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if (lane == 0) warp_tot[warp] = s;
    __syncthreads();
    uint32_t run = tile_sums[blockIdx.x];
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        uint64_t i = base + (uint64_t)r * 32 + lane;
        uint32_t incl = warp_incl_scan(v[r], lane);
        if (i < n) out(i, run + incl - v[r]);
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
}

// Host driver.  tile_sums must hold ceil(n / kScanTile) words; d_total may be nullptr.
template <class In, class Out>
static int exclusive_scan(shk_ctx *ctx, cudaStream_t st, In in, Out out, uint64_t n, uint32_t *tile_sums,
                          uint32_t *d_total)
{
    uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    if (n_tiles == 0) {
        if (d_total) SHK_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st));
        return SHK_OK;
    }
    if (n_tiles > 0x7FFFFFFFull) return fail(ctx, SHK_E_LIMIT, "scan too large");
    scan_reduce_kernel<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, n, tile_sums);
    scan_tile_sums_kernel<<<1, 1024, 0, st>>>(tile_sums, (uint32_t)n_tiles, d_total);
    scan_apply_kernel<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, out, n, tile_sums);
    ctx->launches += 3;
    SHK_CUDA(ctx, cudaGetLastError());
    return SHK_OK;
}

// =============================================================================================
// K1: canonical k-mers of every reference position -> hash -> filter bit.
// Replaces KmerBuilder::operator() (KmerBuilder.hpp:40-72), BloomfilterFiller::operator()
// (BloomfilterFiller.hpp:38-46) and BF::add_at (bloomfilter.h:57-59).  One thread owns 8
// consecutive positions of the concatenated reference and rolls the forward k-mer over them
// (kmer_utils.hpp:73-75); a window is valid when the last k bytes are all ACGT/acgt and lie in
// one record.  The bit index of each window is kept (win_pos) for the second pass.
// =============================================================================================
constexpr int kPosPerThread = 8;

__device__ __forceinline__ uint32_t record_of(const uint64_t *rec_off, uint32_t n_rec, uint64_t x)
{
    // largest r with rec_off[r] <= x  (x < rec_off[n_rec])
    uint32_t lo = 0, hi = n_rec;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (rec_off[mid] <= x) lo = mid;
        else hi = mid;
    }
    return lo;
}

template <int MOD>
__global__ void __launch_bounds__(256)
enum_setbits_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ rec_off, uint32_t n_rec,
                    uint64_t lo, uint64_t hi, int k, FilterGeom g, uint32_t *sectors, uint64_t *win_pos,
                    uint32_t *rec_has_window, unsigned long long *n_windows)
{
    // windows ENDING in [lo, hi): the whole reference, or one shard of it (sharded build)
    const uint64_t x0 = lo + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kPosPerThread;
    uint32_t my_windows = 0;
    if (x0 < hi) {
        const uint64_t xend = min(x0 + (uint64_t)kPosPerThread, hi);
        uint64_t s = x0 >= (uint64_t)(k - 1) ? x0 - (k - 1) : 0;
        uint32_t r = record_of(rec_off, n_rec, s);
        uint64_t next_b = rec_off[r + 1];
        const uint64_t kmask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
        uint64_t fwd = 0;
        int run = 0;
        uint32_t marked = 0xFFFFFFFFu;
        for (uint64_t pos = s; pos < xend; ++pos) {
            while (pos >= next_b) {
                ++r;
                next_b = rec_off[r + 1];
                run = 0;
            }
            uint32_t ch = bases[pos];
            if (base_valid(ch)) {
                fwd = ((fwd << 2) | base_code(ch)) & kmask;
                ++run;
            } else {
                run = 0;
            }
            if (pos >= x0) {
                uint64_t p = kInvalidPos;
                if (run >= k) {
                    p = bit_index<MOD>(xxh64_u64(canonical(fwd, k)), g);
                    atomicOr(&sectors[phys_word(p)], 1u << (p & 31));  // `_bf[p % _size] = 1`
                    if (marked != r) {
                        rec_has_window[r] = 1;
                        marked = r;
                    }
                    ++my_windows;
                }
                win_pos[pos] = p;
            }
        }
    }
    my_windows = __reduce_add_sync(0xFFFFFFFFu, my_windows);
    if ((threadIdx.x & 31) == 0 && my_windows) atomicAdd(n_windows, (unsigned long long)my_windows);
}

// Gene index of every record = the reference's `nidx` (main.cpp:158-187): a record consumes an
// index unless its length is >= k and it has no valid window (`continue` at main.cpp:166 skips
// `++nidx` at 186).  One CTA; n_rec is small (<= a few 100k).
__global__ void __launch_bounds__(1024)
assign_nidx_kernel(const uint64_t *rec_off, const uint32_t *rec_has_window, uint32_t n_rec, int k, uint32_t *nidx,
                   uint32_t *n_genes)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_rec; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = 0;
        if (i < n_rec) {
            uint64_t len = rec_off[i + 1] - rec_off[i];
            v = (len >= (uint64_t)k && !rec_has_window[i]) ? 0u : 1u;
        }
        uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t wi = warp_incl_scan(w, lane);
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        uint32_t excl = carry_s + warp_tot[warp] + incl - v;
        if (i < n_rec) nidx[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_genes = carry_s;
}

// K3: second pass, counting.  Replaces the rank lookups of BF::add_to_kmer (bloomfilter.h:70):
// every window's bit index becomes the 0-based rank of its set bit, and cnt[rank] counts the
// occurrences (an upper bound of the list length before per-gene dedup).
__global__ void __launch_bounds__(256)
rank_count_kernel(uint64_t *win_pos, uint64_t total, const uint32_t *__restrict__ sectors, uint32_t *cnt)
{
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total) return;
    uint64_t p = win_pos[x];
    if (p == kInvalidPos) return;
    uint64_t q = p >> 5;
    uint64_t sec = q / kWordsPerSector;
    uint32_t slot = (uint32_t)(q - sec * kWordsPerSector);
    const uint4 *sp = reinterpret_cast<const uint4 *>(sectors + sec * 8);
    uint4 a = sp[0], b = sp[1];
    Sector s;
    s.w[0] = a.x, s.w[1] = a.y, s.w[2] = a.z, s.w[3] = a.w, s.w[4] = b.x, s.w[5] = b.y, s.w[6] = b.z, s.w[7] = b.w;
    uint32_t r = sector_rank(s, slot, (uint32_t)(p & 31));
    win_pos[x] = r;
    atomicAdd(&cnt[r], 1u);
}

// K4: second pass, filling.  Gene ids land unordered (and possibly repeated) in the slots that
// the scan of cnt reserved for each set bit.
template <class IdT>
__global__ void __launch_bounds__(256)
fill_kernel(const uint64_t *__restrict__ win_pos, uint64_t total, const uint64_t *__restrict__ rec_off, uint32_t n_rec,
            const uint32_t *__restrict__ nidx, const uint32_t *__restrict__ tmp_off, uint32_t *fill, IdT *tmp_ids)
{
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total) return;
    uint64_t v = win_pos[x];
    if (v == kInvalidPos) return;
    uint32_t r = (uint32_t)v;
    uint32_t g = nidx[record_of(rec_off, n_rec, x)];
    uint32_t slot = tmp_off[r] + atomicAdd(&fill[r], 1u);
    tmp_ids[slot] = (IdT)g;  // small_vector_t::push_back(uint16_t), small_vector.hpp:46 (uint32_t with SHK_F_WIDE_IDS)
}

// Sort + unique of each list (BF::add_to_kmer's `last() != input_idx` dedup with genes arriving
// in ascending order, bloomfilter.h:68-74 => every list is strictly ascending).  One thread
// per set bit for the common short lists; long lists are queued for the bitmap kernel.
constexpr uint32_t kShortList = 48;

template <class IdT>
__global__ void __launch_bounds__(256)
sort_unique_kernel(const uint32_t *__restrict__ tmp_off, uint32_t n_set, IdT *tmp_ids, uint32_t *len,
                   uint32_t *long_list, uint32_t *n_long)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_set) return;
    uint32_t b = tmp_off[r], n = tmp_off[r + 1] - b;
    if (n <= 1) {
        len[r] = n;
        return;
    }
    if (n > kShortList) {
        long_list[atomicAdd(n_long, 1u)] = r;
        return;
    }
    IdT *a = tmp_ids + b;
    for (uint32_t i = 1; i < n; ++i) {
        IdT key = a[i];
        uint32_t j = i;
        while (j > 0 && a[j - 1] > key) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = key;
    }
    uint32_t m = 1;
    for (uint32_t i = 1; i < n; ++i)
        if (a[i] != a[m - 1]) a[m++] = a[i];
    len[r] = m;
}

// Long lists: one CTA per list, a 65536-bit presence bitmap in shared memory gives the sorted
// unique ids directly (ids are 16-bit).
__global__ void __launch_bounds__(256)
sort_unique_long_kernel(const uint32_t *__restrict__ tmp_off, const uint32_t *__restrict__ long_list, uint16_t *tmp_ids,
                        uint32_t *len)
{
    __shared__ uint32_t bitmap[2048];
    __shared__ uint32_t warp_tot[8];
    const uint32_t r = long_list[blockIdx.x];
    const uint32_t b = tmp_off[r], n = tmp_off[r + 1] - b;
    for (int i = threadIdx.x; i < 2048; i += 256) bitmap[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
        uint32_t id = tmp_ids[b + i];
        atomicOr(&bitmap[id >> 5], 1u << (id & 31));
    }
    __syncthreads();
    // thread t owns words [8t, 8t+8)
    uint32_t c = 0;
    for (int w = 0; w < 8; ++w) c += __popc(bitmap[threadIdx.x * 8 + w]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(c, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t off = incl - c;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    for (int w = 0; w < 8; ++w) {
        uint32_t bits = bitmap[threadIdx.x * 8 + w];
        while (bits) {
            int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            tmp_ids[b + off++] = (uint16_t)((threadIdx.x * 8 + w) * 32 + bpos);
        }
    }
    if (threadIdx.x == 255) len[r] = off;
}

// SHK_F_WIDE_IDS: the same for 32-bit ids.  The presence bitmap (one bit per gene index) lives in global memory,
// one per CTA; a CTA takes long lists in turn: clear, mark, then compact the set bits in ascending order with a
// block-wide scan over 256 bitmap words at a time.
__global__ void __launch_bounds__(256)
sort_unique_long_wide_kernel(const uint32_t *__restrict__ tmp_off, const uint32_t *__restrict__ long_list, uint32_t n_long,
                             uint32_t *tmp_ids, uint32_t *len, uint32_t *bitmaps, uint32_t words)
{
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t base_s;
    uint32_t *bm = bitmaps + (uint64_t)blockIdx.x * words;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t li = blockIdx.x; li < n_long; li += gridDim.x) {
        const uint32_t r = long_list[li];
        const uint32_t b = tmp_off[r], n = tmp_off[r + 1] - b;
        for (uint32_t i = threadIdx.x; i < words; i += 256) bm[i] = 0;
        if (threadIdx.x == 0) base_s = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += 256) {
            const uint32_t id = tmp_ids[b + i];
            atomicOr(&bm[id >> 5], 1u << (id & 31));
        }
        __syncthreads();
        for (uint32_t w0 = 0; w0 < words; w0 += 256) {
            const uint32_t w = w0 + threadIdx.x;
            uint32_t bits = w < words ? bm[w] : 0u;
            const uint32_t c = __popc(bits);
            const uint32_t incl = warp_incl_scan(c, lane);
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            uint32_t off = base_s + incl - c;
            for (int ww = 0; ww < warp; ++ww) off += warp_tot[ww];
            while (bits) {
                const int bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                tmp_ids[b + off++] = w * 32u + (uint32_t)bpos;
            }
            __syncthreads();
            if (threadIdx.x == 255) base_s = off;  // thread 255 ends at base + all counts of this round
            __syncthreads();
        }
        if (threadIdx.x == 0) len[r] = base_s;
        __syncthreads();
    }
}

// Final layout (BF::switch_mode(2), bloomfilter.h:126-168): ids concatenated in rank order
// (`_index_kmer`), offsets (= select over `_bv`), and the 8-byte entry per set bit.
template <class IdT>
__global__ void __launch_bounds__(256)
finalize_lists_kernel(const uint32_t *__restrict__ tmp_off, const IdT *__restrict__ tmp_ids,
                      const uint32_t *__restrict__ csr_off, uint32_t n_set, IdT *csr_ids, uint64_t *entries)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_set) return;
    uint32_t o = csr_off[r], m = csr_off[r + 1] - o;
    const IdT *src = tmp_ids + tmp_off[r];
    if (m == 0) {
        entries[r] = 0;
        return;
    }
    for (uint32_t t = 0; t < m; ++t) csr_ids[o + t] = src[t];
    uint32_t id0 = src[0];
    if (sizeof(IdT) == 4) {  // wide entries: {length - 1, the id of a one-id list | CSR begin}
        entries[r] = make_wide_entry(m, m == 1 ? id0 : o);
        return;
    }
    uint32_t lo = m == 1 ? 0u : (m == 2 ? (uint32_t)src[1] : o);
    entries[r] = make_entry(id0, m, lo);
}

// Logical positions of the set bits in rank order (for shk_index_export).
__global__ void __launch_bounds__(256)
export_positions_kernel(const uint32_t *__restrict__ sectors, uint64_t n_sectors, uint64_t *pos)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    const uint4 *sp = reinterpret_cast<const uint4 *>(sectors + s * 8);
    uint4 a = sp[0], b = sp[1];
    uint32_t w[7] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z};
    uint32_t r = b.w;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        uint32_t bits = w[i];
        while (bits) {
            int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            pos[r++] = (s * kWordsPerSector + i) * 32 + bpos;
        }
    }
}

// Front table build (layout in shk_device.cuh).  Every set bit contributes L slots to its bucket
// (L = list length if <= 4, else one "long list" slot).  Three passes over the set bits / buckets:
//   front_count   slots needed per bucket
//   (scan)        overflow records needed per bucket -> first record index of each bucket
//   front_place   each set bit claims its logical slot range in the bucket with one atomicAdd and
//                 writes its slots; logical slot t lives in the bucket (t < 3 or the bucket fits)
//                 or in overflow record (t-3)/3, slot (t-3)%3
//   front_chain   chain pointers of the overflowing buckets
template <class F>
__device__ __forceinline__ void for_each_set_bit(const uint32_t *__restrict__ sectors, uint64_t s, F f)
{
    const uint4 *sp = reinterpret_cast<const uint4 *>(sectors + s * 8);
    uint4 a = sp[0], b = sp[1];
    uint32_t w[7] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z};
    uint32_t r = b.w;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        uint32_t bits = w[i];
        while (bits) {
            int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            f((s * kWordsPerSector + i) * 32 + bpos, r++);
        }
    }
}

__device__ __forceinline__ uint32_t front_slots_of(uint64_t e)
{
    const uint32_t ln = entry_len(e);
    return ln <= kFrontInlineMax ? ln : 1u;
}

// Per-bucket counter: low 16 bits = slots (at most 2^13 positions x 4), high 16 bits = number of
// 2-id lists.  In a chained bucket a 2-id list never straddles two records (the fast kernel
// relies on it: a plain id found in a record has its partner in the same record), which can
// waste one slot per such list - the record budget counts it.
__device__ __forceinline__ uint32_t front_cnt_slots(uint32_t c) { return c & 0xFFFFu; }
__device__ __forceinline__ uint32_t front_cnt_records(uint32_t c)
{
    const uint32_t slots = c & 0xFFFFu;
    return slots <= 4 ? 0u : (slots + (c >> 16) - 3u + 2u) / 3u;
}

__global__ void __launch_bounds__(256)
front_count_kernel(const uint32_t *__restrict__ sectors, uint64_t n_sectors, const uint64_t *__restrict__ entries,
                   FrontGeom fg, uint32_t *cnt)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    for_each_set_bit(sectors, s, [&](uint64_t p, uint32_t r) {
        const uint32_t L = front_slots_of(entries[r]);
        atomicAdd(&cnt[p >> fg.shift], L + (L == 2 ? 0x10000u : 0u));
    });
}

// overflow records needed by a bucket: 3 slots stay in the bucket, the rest in records of 3
struct FrontRecordsIn {
    const uint32_t *cnt;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return front_cnt_records(cnt[i]); }
};

// word index (in 32-bit words from the start of the table) of logical slot t of a bucket; an entry
// is 4 words (slots), or 8 with anchors: slots in words 0-3, anchors in words 4-7
__device__ __forceinline__ uint64_t front_slot_word(const FrontGeom &fg, uint64_t bucket, uint32_t total, uint32_t t,
                                                    uint32_t rec_base)
{
    const uint64_t wpe = 4ull * fg.stride;
    if (total <= 4 || t < 3) return bucket * wpe + t;
    const uint32_t u = t - 3;
    return (fg.n_buckets + rec_base + u / 3) * wpe + (u % 3);
}

__global__ void __launch_bounds__(256)
front_place_kernel(const uint32_t *__restrict__ sectors, uint64_t n_sectors, const uint64_t *__restrict__ entries,
                   const uint16_t *__restrict__ csr_ids, FrontGeom fg, const uint32_t *__restrict__ cnt,
                   const uint32_t *__restrict__ rec_base, uint32_t *cur, uint32_t *front,
                   const uint32_t *__restrict__ anchor_of_rank)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    for_each_set_bit(sectors, s, [&](uint64_t p, uint32_t r) {
        const uint64_t e = entries[r];
        const uint64_t bkt = p >> fg.shift;
        const uint32_t key = front_key((uint32_t)p & fg.off_mask);
        const uint32_t ln = entry_len(e), L = front_slots_of(e);
        const uint32_t total = front_cnt_slots(cnt[bkt]), rb = rec_base[bkt];
        uint32_t t0;
        if (total <= 4 || L != 2) {
            t0 = atomicAdd(&cur[bkt], L);
        } else {  // chained bucket, 2-id list: both slots in one record (logical slots 3i..3i+2)
            uint32_t seen = cur[bkt];
            for (;;) {
                t0 = seen + (seen % 3u == 2u ? 1u : 0u);
                const uint32_t prev = atomicCAS(&cur[bkt], seen, t0 + 2u);
                if (prev == seen) break;
                seen = prev;
            }
        }
        const uint32_t anchor = fg.stride == 2 ? anchor_of_rank[r] : 0u;
        if (ln > kFrontInlineMax) {
            front[front_slot_word(fg, bkt, total, t0, rb)] = key | kFrontLongFlag;
        } else {
            for (uint32_t i = 0; i < ln; ++i) {
                uint32_t g = i == 0 ? entry_id0(e) : (ln == 2 ? entry_lo(e) : (uint32_t)csr_ids[entry_lo(e) + i]);
                const uint64_t w = front_slot_word(fg, bkt, total, t0 + i, rb);
                front[w] = key | (ln >= 3 ? kFrontMultiFlag : 0u) | g;
                if (fg.stride == 2) front[w + 4] = anchor;
            }
        }
    });
}

__global__ void __launch_bounds__(256)
front_chain_kernel(FrontGeom fg, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ rec_base, uint32_t *front)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= fg.n_buckets) return;
    const uint32_t nrec = front_cnt_records(cnt[b]), rb = rec_base[b];
    if (nrec == 0) return;
    const uint64_t wpe = 4ull * fg.stride;
    front[b * wpe + 3] = kFrontChainBit | (uint32_t)(fg.n_buckets + rb);
    for (uint32_t i = 0; i + 1 < nrec; ++i)
        front[(fg.n_buckets + rb + i) * wpe + 3] = kFrontChainBit | (uint32_t)(fg.n_buckets + rb + i + 1);
}

// =============================================================================================
// Anchor-and-extend structures (layouts in shk_device.cuh).  `win` is pass 2's per-position array:
// win[e] = rank of the filter bit of the reference window ENDING at e, or kInvalidPos.
// =============================================================================================
__device__ __forceinline__ bool lists_equal(uint64_t e0, uint64_t e1, const uint16_t *__restrict__ csr_ids)
{
    const uint32_t n = entry_len(e0);
    if (n != entry_len(e1) || entry_id0(e0) != entry_id0(e1)) return false;
    if (n == 1) return true;
    if (n == 2) return entry_lo(e0) == entry_lo(e1);
    const uint16_t *a = csr_ids + entry_lo(e0), *b = csr_ids + entry_lo(e1);
    for (uint32_t i = 1; i < n; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

// E[e] (bit e of ebits) and the anchor of every set bit = the smallest window end that maps to it.
__global__ void __launch_bounds__(256)
ext_flags_kernel(const uint64_t *__restrict__ win, uint64_t total, const uint64_t *__restrict__ entries,
                 const uint16_t *__restrict__ csr_ids, uint32_t *ebits, uint32_t *anchor_of_rank)
{
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool E = false;
    if (x < total) {
        const uint64_t v = win[x];
        if (v != kInvalidPos) {
            const uint32_t r = (uint32_t)v;
            atomicMin(&anchor_of_rank[r], (uint32_t)x);
            if (x >= 1) {
                const uint64_t v0 = win[x - 1];
                if (v0 != kInvalidPos) E = (uint32_t)v0 == r || lists_equal(entries[(uint32_t)v0], entries[r], csr_ids);
            }
        }
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, E);
    if ((threadIdx.x & 31) == 0 && x < total) ebits[x >> 5] = m;
}

__device__ __forceinline__ uint32_t ebit(const uint32_t *__restrict__ ebits, uint64_t e, uint64_t total)
{
    return e < total ? (ebits[e >> 5] >> (e & 31)) & 1u : 0u;
}

// one thread per 16 reference positions -> one word of estream; per 32 positions -> one word of ref2
__global__ void __launch_bounds__(256)
ext_pack_kernel(const uint8_t *__restrict__ bases, uint64_t total, int k, const uint32_t *__restrict__ ebits,
                uint64_t *estream, uint64_t *ref2)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t t0 = w * 16;
    if (t0 >= total) return;
    uint64_t word = 0;
    uint32_t codes = 0;  // 16 bases, base i at bits 30-2i
    for (uint32_t i = 0; i < 16; ++i) {
        const uint64_t t = t0 + i;
        if (t >= total) break;
        const uint32_t ch = bases[t];
        const uint32_t code = base_valid(ch) ? base_code(ch) : 0u;
        const uint32_t nib = code | (ebit(ebits, t, total) << 2) | (ebit(ebits, t + (uint64_t)k, total) << 3);
        word |= (uint64_t)nib << (4 * i);
        codes |= code << (30 - 2 * i);
    }
    estream[w + 1] = word;
    // two threads share a ref2 word: the even one owns the high half
    uint32_t *r32 = reinterpret_cast<uint32_t *>(ref2 + (w >> 1) + 1);
    r32[(w & 1) ? 0 : 1] = codes;  // little-endian: word index 1 = bits 63..32
}

// coarse miss filter: bit (p >> coarse_shift) for every set position p
__global__ void __launch_bounds__(256)
coarse_fill_kernel(const uint32_t *__restrict__ sectors, uint64_t n_sectors, uint32_t coarse_shift, uint32_t *coarse)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    for_each_set_bit(sectors, s, [&](uint64_t p, uint32_t) {
        const uint64_t c = p >> coarse_shift;
        atomicOr(&coarse[c >> 5], 1u << (c & 31));
    });
}

// =============================================================================================
// Host orchestration of the build.
// =============================================================================================
template <class T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(uint64_t n) { return cudaMalloc((void **)&p, std::max<uint64_t>(n, 1) * sizeof(T)); }
    T *release()
    {
        T *q = p;
        p = nullptr;
        return q;
    }
};

// Device time of a build without the host gaps between its phases (cudaMalloc of GB-sized
// buffers, result read-backs): a segment opens after the allocations of a phase and closes right
// before the phase's stream synchronisation; build_ms is the sum of the segments.
struct SegTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> segs;
    bool open = false;
    void start(cudaStream_t st)
    {
        if (open) return;
        cudaEvent_t a = nullptr, b = nullptr;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        cudaEventRecord(a, st);
        segs.emplace_back(a, b);
        open = true;
    }
    void stop(cudaStream_t st)
    {
        if (!open) return;
        cudaEventRecord(segs.back().second, st);
        open = false;
    }
    // all segments must have completed (the caller has synchronised the stream)
    float total()
    {
        float sum = 0;
        for (auto &s : segs) {
            float t = 0;
            if (cudaEventElapsedTime(&t, s.first, s.second) == cudaSuccess) sum += t;
        }
        return sum;
    }
    void clear()
    {
        for (auto &s : segs) {
            cudaEventDestroy(s.first);
            cudaEventDestroy(s.second);
        }
        segs.clear();
        open = false;
    }
    ~SegTimer() { clear(); }
};
static inline void seg_start(SegTimer *t, cudaStream_t st)
{
    if (t) t->start(st);
}
static inline void seg_stop(SegTimer *t, cudaStream_t st)
{
    if (t) t->stop(st);
}

static void free_index_arrays(DeviceIndex &ix)
{
    if (ix.entries) cudaFree(ix.entries);
    if (ix.csr_off) cudaFree(ix.csr_off);
    if (ix.csr_ids) cudaFree(ix.csr_ids);
    if (ix.csr_ids32) cudaFree(ix.csr_ids32);
    if (ix.front) cudaFree(ix.front);
    if (ix.estream) cudaFree(ix.estream);
    if (ix.ref2) cudaFree(ix.ref2);
    if (ix.coarse) cudaFree(ix.coarse);
    if (ix.refr) cudaFree(ix.refr);
    if (ix.ebits) cudaFree(ix.ebits);
    if (ix.front_plain) cudaFree(ix.front_plain);
    ix.refr = nullptr;
    ix.ebits = nullptr;
    ix.front_plain = nullptr;
    ix.derived_bytes = 0;
    ix.front = nullptr;
    ix.entries = nullptr;
    ix.csr_off = nullptr;
    ix.csr_ids = nullptr;
    ix.csr_ids32 = nullptr;
    ix.estream = nullptr;
    ix.ref2 = nullptr;
    ix.coarse = nullptr;
    ix.egeom = ExtGeom{};
    ix.built = false;
}

// Sizes of the extension arrays for a reference of `total` bases (egeom.coarse_shift must be set).
static void ext_geometry(DeviceIndex &ix, uint64_t total)
{
    ix.egeom.total = total;
    ix.egeom.estream_words = total / 16 + 3;
    ix.egeom.ref2_words = total / 32 + 3;
    ix.egeom.coarse_words = (((ix.geom.bf_bits >> ix.egeom.coarse_shift) + 32) >> 5) + 1;
}

static int ext_alloc(shk_ctx *ctx)
{
    DeviceIndex &ix = ctx->index;
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.estream, ix.egeom.estream_words * 8));
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.ref2, ix.egeom.ref2_words * 8));
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.coarse, ix.egeom.coarse_words * 4));
    return SHK_OK;
}

// Is a path forced (shk_params.flags, or SHK_EXTEND=0/1 for tuning)?  -1 = no, else 0 / 1.
static int forced_extend(const shk_ctx *ctx)
{
    if (ctx->params.flags & SHK_F_EXTEND_ON) return 1;
    if (ctx->params.flags & SHK_F_EXTEND_OFF) return 0;
    if (const char *ev = getenv("SHK_EXTEND")) return atoi(ev) != 0;
    return -1;
}
// A plain front table (16 bytes per entry) of this many entries stays in L2 next to the streamed reads.
static bool front_fits_l2(uint64_t front_entries) { return front_entries * 16 <= (96ull << 20); }

// Extension structures: built unless forced off.  They serve the bulk kernel (packed reads) for any table size; the
// thread-per-read kernels use them only when the table is DRAM-sized or the flag forces it - for an L2-sized table
// they probe a slots-only copy of it (index_derive_bulk, info.plain_front) exactly as without the structures.
static bool decide_extend(const shk_ctx *ctx, uint64_t front_entries, uint64_t total)
{
    (void)front_entries;
    if (total == 0 || total >= 0xFFFFFF00ull) return false;
    const int f = forced_extend(ctx);
    return f != 0;
}

// Front table geometry: at most ~0.7 keys per 4-slot bucket (C2: 64 MB for 2.8 M keys; measured
// faster than 1.5 keys/bucket = 32 MB because chains are rarer, profiles/analyze_r1_v3.md), at
// least 32 positions per bucket (the table is then at most 4x the plain bit vector), at most 2^13
// (13-bit offsets).  SHK_FRONT_LOAD overrides the 0.7 (tuning).
static void front_geometry(DeviceIndex &ix)
{
    double load = 0.7;
    if (const char *ev = getenv("SHK_FRONT_LOAD")) {
        double v = atof(ev);
        if (v > 0.01 && v < 64) load = v;
    }
    const double per_key = load * (double)ix.geom.bf_bits / (double)std::max<uint64_t>(ix.info.n_set_bits, 1);
    uint32_t shift = 5;
    while (shift < kFrontMaxShift && (double)(1ull << (shift + 1)) <= per_key) ++shift;
    ix.fgeom.shift = shift;
    ix.fgeom.off_mask = (1u << shift) - 1u;
    ix.fgeom.n_buckets = (ix.geom.bf_bits + (1ull << shift) - 1) >> shift;
}

// Allocates the table (and the extension arrays) for a known geometry (info.front_shift /
// front_entries / extend / ref_bases / coarse_shift): used when the index is adopted from another GPU.
int index_alloc_front(shk_ctx *ctx)
{
    DeviceIndex &ix = ctx->index;
    if (ix.front) cudaFree(ix.front);
    if (ix.estream) cudaFree(ix.estream);
    if (ix.ref2) cudaFree(ix.ref2);
    if (ix.coarse) cudaFree(ix.coarse);
    if (ix.refr) cudaFree(ix.refr);
    if (ix.ebits) cudaFree(ix.ebits);
    if (ix.front_plain) cudaFree(ix.front_plain);
    ix.refr = nullptr, ix.ebits = nullptr, ix.front_plain = nullptr;
    ix.derived_bytes = 0;
    ix.front = nullptr;
    ix.estream = nullptr, ix.ref2 = nullptr, ix.coarse = nullptr;
    ix.egeom = ExtGeom{};
    if (ix.info.id_bits == 32) {  // SHK_F_WIDE_IDS: no front table
        ix.fgeom = FrontGeom{};
        return SHK_OK;
    }
    ix.fgeom.shift = ix.info.front_shift;
    ix.fgeom.off_mask = (1u << ix.fgeom.shift) - 1u;
    ix.fgeom.n_buckets = (ix.geom.bf_bits + (1ull << ix.fgeom.shift) - 1) >> ix.fgeom.shift;
    ix.fgeom.n_entries = ix.info.front_entries;
    ix.fgeom.stride = ix.info.extend ? 2u : 1u;
    if (ix.fgeom.n_entries < ix.fgeom.n_buckets || ix.fgeom.shift < 5 || ix.fgeom.shift > kFrontMaxShift ||
        (ix.info.extend && ix.info.coarse_shift > ix.fgeom.shift))
        return fail(ctx, SHK_E_ARG, "inconsistent front table geometry in index info");
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.front, ix.fgeom.n_entries * 16 * ix.fgeom.stride));
    if (ix.info.extend) {
        ix.egeom.enabled = 1;
        ix.egeom.coarse_shift = ix.info.coarse_shift;
        ext_geometry(ix, ix.info.ref_bases);
        return ext_alloc(ctx);
    }
    return SHK_OK;
}

// refr / ebits: the reference and the E flags in the form the bulk classification kernel reads them (shk_bulk.cu) -
// 32 positions per word like the packed reads, zero words on both sides.  Derived from estream, so every way an
// index becomes ready (build, staged protocol, adopt + finalize, load, replicate, sharded build) gets them.
__global__ void __launch_bounds__(256)
derive_bulk_kernel(const uint64_t *__restrict__ estream, uint64_t total, uint64_t *refr, uint32_t *ebits)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // positions 32w .. 32w+31
    if (w * 32 >= total) return;
    uint64_t codes = 0;
    uint32_t e = 0;
    for (uint32_t h = 0; h < 2; ++h) {
        const uint64_t t0 = w * 32 + h * 16;
        if (t0 >= total) break;
        const uint64_t word = estream[(t0 >> 4) + 1];
        for (uint32_t i = 0; i < 16; ++i) {
            const uint32_t nib = (uint32_t)(word >> (4 * i)) & 15u;
            const uint32_t code = nib & 3u, raw = code ^ (code >> 1);  // A0 C1 G2 T3 -> A0 C1 T2 G3
            codes |= (uint64_t)raw << (2 * (h * 16 + i));
            e |= ((nib >> 2) & 1u) << (h * 16 + i);
        }
    }
    refr[w + kDerivedPad] = codes;
    ebits[w + kDerivedPad] = e;
}

__global__ void __launch_bounds__(256)
derive_plain_kernel(const uint4 *__restrict__ front, uint64_t n_entries, uint4 *plain)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_entries) plain[i] = front[2 * i];  // the slots; chain pointers are entry indices, valid for any stride
}

int index_derive_bulk(shk_ctx *ctx)
{
    DeviceIndex &ix = ctx->index;
    if (ix.refr) cudaFree(ix.refr);
    if (ix.ebits) cudaFree(ix.ebits);
    if (ix.front_plain) cudaFree(ix.front_plain);
    ix.refr = nullptr, ix.ebits = nullptr, ix.front_plain = nullptr;
    ix.info.device_bytes -= std::min(ix.info.device_bytes, ix.derived_bytes);
    ix.derived_bytes = 0;
    ix.info.plain_front = 0;
    if (!ix.egeom.enabled || !ix.estream) return SHK_OK;
    const uint64_t words = ix.egeom.total / 32 + 1 + kDerivedPad + kDerivedTail;
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.refr, words * 8));
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.ebits, words * 4));
    ix.derived_bytes = words * 12;
    cudaStream_t st = ctx->build_stream;
    SHK_CUDA(ctx, cudaMemsetAsync(ix.refr, 0, words * 8, st));
    SHK_CUDA(ctx, cudaMemsetAsync(ix.ebits, 0, words * 4, st));
    const uint64_t n = (ix.egeom.total + 31) / 32;
    if (n) derive_bulk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ix.estream, ix.egeom.total, ix.refr, ix.ebits);
    SHK_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    if (forced_extend(ctx) < 0 && ix.front && ix.fgeom.stride == 2 && ix.fgeom.n_entries && front_fits_l2(ix.fgeom.n_entries)) {
        SHK_CUDA(ctx, cudaMalloc((void **)&ix.front_plain, ix.fgeom.n_entries * 16));
        derive_plain_kernel<<<(unsigned)((ix.fgeom.n_entries + 255) / 256), 256, 0, st>>>(ix.front, ix.fgeom.n_entries, ix.front_plain);
        SHK_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        ix.derived_bytes += ix.fgeom.n_entries * 16;
        ix.info.plain_front = 1;
    }
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    ix.info.device_bytes += ix.derived_bytes;
    return SHK_OK;
}

// Per-build inputs of the extension structures (alive only during index_build_device).
struct ExtBuildInputs {
    const uint8_t *d_bases;
    const uint64_t *d_win;  // per window end: rank of its filter bit, or kInvalidPos
    uint64_t total;
};

static int build_front(shk_ctx *ctx, cudaStream_t st, const uint64_t *d_entries, const uint16_t *d_csr_ids,
                       uint32_t *d_tiles, uint32_t *d_scalar, const ExtBuildInputs &xin, SegTimer *tm = nullptr)
{
    DeviceIndex &ix = ctx->index;
    front_geometry(ix);
    const uint64_t nb = ix.fgeom.n_buckets;
    if (nb >= 0x7FFFFFFFull) return fail(ctx, SHK_E_LIMIT, "front table too large");
    DevBuf<uint32_t> d_cnt, d_cur, d_rec, d_anchor, d_ebits;
    SHK_CUDA(ctx, d_cnt.alloc(nb));
    SHK_CUDA(ctx, d_cur.alloc(nb));
    SHK_CUDA(ctx, d_rec.alloc(nb + 1));
    seg_start(tm, st);
    SHK_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, nb * 4, st));
    SHK_CUDA(ctx, cudaMemsetAsync(d_cur.p, 0, nb * 4, st));
    const unsigned blocks_s = (unsigned)((ix.geom.n_sectors + 255) / 256);
    uint32_t n_rec = 0;
    if (ix.info.n_set_bits > 0) {
        front_count_kernel<<<blocks_s, 256, 0, st>>>(ix.sectors, ix.geom.n_sectors, d_entries, ix.fgeom, d_cnt.p);
        ctx->launches += 1;
        int rc = exclusive_scan(ctx, st, FrontRecordsIn{d_cnt.p}, U32Out{d_rec.p}, nb, d_tiles, d_scalar);
        if (rc) return rc;
        SHK_CUDA(ctx, cudaMemcpyAsync(&n_rec, d_scalar, 4, cudaMemcpyDeviceToHost, st));
        seg_stop(tm, st);
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
    }
    seg_stop(tm, st);
    ix.fgeom.n_entries = nb + n_rec;
    if (ix.fgeom.n_entries >= 0x7FFFFFFFull) return fail(ctx, SHK_E_LIMIT, "front table too large");
    const bool ext = ix.info.n_set_bits > 0 && decide_extend(ctx, ix.fgeom.n_entries, xin.total);
    ix.fgeom.stride = ext ? 2u : 1u;
    ix.egeom = ExtGeom{};
    if (ext) {
        // E flags + anchors, then the packed reference streams and the coarse filter
        ix.egeom.enabled = 1;
        uint32_t lg = 0;
        while (lg < 63 && (1ull << lg) < ix.geom.bf_bits) ++lg;
        // coarse filter: 2^28 bits (32 MB, L2-resident) unless the filter is smaller; SHK_COARSE_LOG2 (tuning) overrides
        uint32_t coarse_log2 = 28;
        if (const char *ev = getenv("SHK_COARSE_LOG2")) {
            const int v = atoi(ev);
            if (v >= 20 && v <= 32) coarse_log2 = (uint32_t)v;
        }
        ix.egeom.coarse_shift = std::min<uint32_t>(lg > coarse_log2 ? lg - coarse_log2 : 0, ix.fgeom.shift);
        ext_geometry(ix, xin.total);
        int rc = ext_alloc(ctx);
        if (rc) return rc;
        SHK_CUDA(ctx, d_anchor.alloc(ix.info.n_set_bits));
        SHK_CUDA(ctx, d_ebits.alloc((xin.total + 31) / 32 + 1));
        seg_start(tm, st);
        SHK_CUDA(ctx, cudaMemsetAsync(d_anchor.p, 0xFF, ix.info.n_set_bits * 4, st));
        SHK_CUDA(ctx, cudaMemsetAsync(ix.estream, 0, ix.egeom.estream_words * 8, st));
        SHK_CUDA(ctx, cudaMemsetAsync(ix.ref2, 0, ix.egeom.ref2_words * 8, st));
        SHK_CUDA(ctx, cudaMemsetAsync(ix.coarse, 0, ix.egeom.coarse_words * 4, st));
        ext_flags_kernel<<<(unsigned)((xin.total + 255) / 256), 256, 0, st>>>(xin.d_win, xin.total, d_entries, d_csr_ids,
                                                                              d_ebits.p, d_anchor.p);
        ext_pack_kernel<<<(unsigned)(((xin.total + 15) / 16 + 255) / 256), 256, 0, st>>>(
            xin.d_bases, xin.total, (int)ctx->params.k, d_ebits.p, ix.estream, ix.ref2);
        coarse_fill_kernel<<<blocks_s, 256, 0, st>>>(ix.sectors, ix.geom.n_sectors, ix.egeom.coarse_shift, ix.coarse);
        ctx->launches += 3;
        SHK_CUDA(ctx, cudaGetLastError());
    }
    if (ix.front) cudaFree(ix.front);
    ix.front = nullptr;
    const uint64_t front_bytes = ix.fgeom.n_entries * 16 * ix.fgeom.stride;
    seg_stop(tm, st);  // cudaFree / cudaMalloc below wait for the device
    SHK_CUDA(ctx, cudaMalloc((void **)&ix.front, front_bytes));
    seg_start(tm, st);
    SHK_CUDA(ctx, cudaMemsetAsync(ix.front, 0xFF, front_bytes, st));
    if (ix.info.n_set_bits > 0) {
        uint32_t *f32 = reinterpret_cast<uint32_t *>(ix.front);
        front_place_kernel<<<blocks_s, 256, 0, st>>>(ix.sectors, ix.geom.n_sectors, d_entries, d_csr_ids, ix.fgeom,
                                                     d_cnt.p, d_rec.p, d_cur.p, f32, d_anchor.p);
        front_chain_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(ix.fgeom, d_cnt.p, d_rec.p, f32);
        ctx->launches += 2;
        SHK_CUDA(ctx, cudaGetLastError());
    }
    seg_stop(tm, st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));  // the temporaries die with this scope
    ix.info.front_shift = ix.fgeom.shift;
    ix.info.front_entries = ix.fgeom.n_entries;
    ix.info.extend = ext ? 1u : 0u;
    ix.info.ref_bases = xin.total;
    ix.info.coarse_shift = ix.egeom.coarse_shift;
    return SHK_OK;
}

// Second half of the list build, shared by the one-call build and the staged protocol: lists that
// sit unordered (and repeated) at tmp_ids[tmp_off[r] ..] become the sorted unique CSR + entries.
// d_len (n_set + 1 words) receives the final lengths, d_long / d_n_long queue the long lists.
template <class IdT>
static int finish_lists(shk_ctx *ctx, cudaStream_t st, uint32_t n_set, const uint32_t *d_tmp_off, IdT *d_tmp_ids,
                        uint32_t *d_len, uint32_t *d_long, uint32_t *d_n_long, uint32_t *d_tiles, DevBuf<uint32_t> &d_csr_off,
                        DevBuf<IdT> &d_csr_ids, DevBuf<uint64_t> &d_entries, uint64_t &tot_ids, SegTimer *tm = nullptr,
                        uint32_t n_genes = 0)
{
    const unsigned blocks_r = (unsigned)(((uint64_t)n_set + 255) / 256);
    SHK_CUDA(ctx, cudaMemsetAsync(d_n_long, 0, 4, st));
    sort_unique_kernel<IdT><<<blocks_r, 256, 0, st>>>(d_tmp_off, n_set, d_tmp_ids, d_len, d_long, d_n_long);
    ctx->launches += 1;
    uint32_t h_long = 0;
    SHK_CUDA(ctx, cudaMemcpyAsync(&h_long, d_n_long, 4, cudaMemcpyDeviceToHost, st));
    seg_stop(tm, st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    seg_start(tm, st);
    DevBuf<uint32_t> d_bitmaps;
    if (h_long > 0) {
        if constexpr (sizeof(IdT) == 2) {
            sort_unique_long_kernel<<<h_long, 256, 0, st>>>(d_tmp_off, d_long, d_tmp_ids, d_len);
        } else {
            const uint32_t words = (std::max<uint32_t>(n_genes, 1) + 31) / 32;
            const uint32_t ctas = std::min<uint32_t>(h_long, 296);
            seg_stop(tm, st);
            SHK_CUDA(ctx, d_bitmaps.alloc((uint64_t)ctas * words));
            seg_start(tm, st);
            sort_unique_long_wide_kernel<<<ctas, 256, 0, st>>>(d_tmp_off, d_long, h_long, d_tmp_ids, d_len, d_bitmaps.p, words);
        }
        ctx->launches += 1;
    }
    int rc = exclusive_scan(ctx, st, U32In{d_len}, U32Out{d_csr_off.p}, (uint64_t)n_set, d_tiles, d_csr_off.p + n_set);
    if (rc) return rc;
    uint32_t h_tot = 0;
    SHK_CUDA(ctx, cudaMemcpyAsync(&h_tot, d_csr_off.p + n_set, 4, cudaMemcpyDeviceToHost, st));
    seg_stop(tm, st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    tot_ids = h_tot;
    // the reference counts ids in an int (bloomfilter.h:130); SHK_F_WIDE_IDS lifts that to the 32 bits of our offsets
    if (tot_ids >= (sizeof(IdT) == 2 ? 0x7FFFFFFFull : 0xFFFFFFFFull))
        return fail(ctx, SHK_E_LIMIT, "total id count overflows the reference's int (bloomfilter.h:130)");
    SHK_CUDA(ctx, d_csr_ids.alloc(tot_ids));
    seg_start(tm, st);
    finalize_lists_kernel<IdT><<<blocks_r, 256, 0, st>>>(d_tmp_off, d_tmp_ids, d_csr_off.p, n_set, d_csr_ids.p, d_entries.p);
    ctx->launches += 1;
    SHK_CUDA(ctx, cudaGetLastError());
    return SHK_OK;
}

template <int MOD>
static void launch_enum(shk_ctx *ctx, cudaStream_t st, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec,
                        uint64_t lo, uint64_t hi, uint64_t *win_pos, uint32_t *has_window, unsigned long long *n_windows)
{
    uint64_t threads = (hi - lo + kPosPerThread - 1) / kPosPerThread;
    unsigned blocks = (unsigned)((threads + 255) / 256);
    enum_setbits_kernel<MOD><<<blocks, 256, 0, st>>>(bases, rec_off, n_rec, lo, hi, (int)ctx->params.k, ctx->index.geom,
                                                      ctx->index.sectors, win_pos, has_window, n_windows);
    ctx->launches += 1;
}

// Device state of one build.  It lives on the stack of index_build_device for the one-call build
// and in shk_ctx::shard between the calls of the sharded protocol.
struct BuildState {
    DevBuf<uint8_t> bases;
    DevBuf<uint64_t> rec_off, win;  // win[e]: bit index (pass 1) / rank (pass 2) of the window ending at e
    DevBuf<uint32_t> has, nidx, scalars, tiles;
    DevBuf<unsigned long long> nwin;
    uint64_t total = 0;
    uint32_t n_rec = 0;
    uint64_t lo = 0, hi = 0;  // window ends this context enumerates itself
    uint32_t n_genes = 0, n_set = 0;
    unsigned long long n_windows = 0;
    SegTimer tm;
    double t_wall0 = 0;
};

static double wall_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Pass 1 over the window ends [lo, hi) of the reference (KmerBuilder + BloomfilterFiller + BF::add_at).
static int build_pass1(shk_ctx *ctx, BuildState &bs, const uint8_t *ref_bases, const uint64_t *rec_off, uint32_t n_rec,
                       uint64_t lo, uint64_t hi)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    const uint64_t total = rec_off[n_rec];
    if (total >= (1ULL << 32)) return fail(ctx, SHK_E_LIMIT, "reference larger than 4 Gbases is not supported");
    for (uint32_t i = 0; i < n_rec; ++i)
        if (rec_off[i + 1] < rec_off[i]) return fail(ctx, SHK_E_ARG, "rec_offsets must be non-decreasing");
    free_index_arrays(ix);
    bs.t_wall0 = wall_ms();
    bs.total = total, bs.n_rec = n_rec, bs.lo = lo, bs.hi = hi;
    const uint64_t sector_bytes = ix.geom.n_sectors * 32;
    SHK_CUDA(ctx, bs.bases.alloc(total));
    SHK_CUDA(ctx, bs.rec_off.alloc((uint64_t)n_rec + 1));
    SHK_CUDA(ctx, bs.win.alloc(total));
    SHK_CUDA(ctx, bs.has.alloc(n_rec));
    SHK_CUDA(ctx, bs.nidx.alloc(n_rec));
    SHK_CUDA(ctx, bs.scalars.alloc(8));  // 0 n_genes, 1 n_set, 2 n_occ, 3 tot_ids, 4 n_long
    SHK_CUDA(ctx, bs.nwin.alloc(1));
    uint64_t max_scan_n = std::max<uint64_t>(ix.geom.n_sectors, total + 1);
    SHK_CUDA(ctx, bs.tiles.alloc(max_scan_n / kScanTile + 2));
    bs.tm.start(st);
    SHK_CUDA(ctx, cudaMemcpyAsync(bs.bases.p, ref_bases, total, cudaMemcpyHostToDevice, st));
    SHK_CUDA(ctx, cudaMemcpyAsync(bs.rec_off.p, rec_off, ((uint64_t)n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    SHK_CUDA(ctx, cudaMemsetAsync(ix.sectors, 0, sector_bytes, st));
    SHK_CUDA(ctx, cudaMemsetAsync(bs.has.p, 0, std::max<uint64_t>(n_rec, 1) * 4, st));
    SHK_CUDA(ctx, cudaMemsetAsync(bs.scalars.p, 0, 8 * 4, st));
    SHK_CUDA(ctx, cudaMemsetAsync(bs.nwin.p, 0, 8, st));
    if (hi > lo) {
        switch (ix.geom.mod_kind) {
        case MOD_POW2: launch_enum<MOD_POW2>(ctx, st, bs.bases.p, bs.rec_off.p, n_rec, lo, hi, bs.win.p, bs.has.p, bs.nwin.p); break;
        case MOD_B33: launch_enum<MOD_B33>(ctx, st, bs.bases.p, bs.rec_off.p, n_rec, lo, hi, bs.win.p, bs.has.p, bs.nwin.p); break;
        default: launch_enum<MOD_GENERIC>(ctx, st, bs.bases.p, bs.rec_off.p, n_rec, lo, hi, bs.win.p, bs.has.p, bs.nwin.p); break;
        }
        SHK_CUDA(ctx, cudaGetLastError());
    }
    return SHK_OK;
}

// `nidx` of every record (main.cpp:158-187) and BF::switch_mode(1): the rank directory and num_kmer
// (bloomfilter.h:121-122).  Needs the complete filter and the complete window flags.
static int build_rank(shk_ctx *ctx, BuildState &bs)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    bs.tm.start(st);
    assign_nidx_kernel<<<1, 1024, 0, st>>>(bs.rec_off.p, bs.has.p, bs.n_rec, (int)ctx->params.k, bs.nidx.p, bs.scalars.p + 0);
    ctx->launches += 1;
    int rc = exclusive_scan(ctx, st, SectorPopIn{ix.sectors}, SectorRankOut{ix.sectors}, ix.geom.n_sectors, bs.tiles.p,
                            bs.scalars.p + 1);
    if (rc) return rc;
    uint32_t h_scalars[8];
    SHK_CUDA(ctx, cudaMemcpyAsync(h_scalars, bs.scalars.p, sizeof h_scalars, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaMemcpyAsync(&bs.n_windows, bs.nwin.p, 8, cudaMemcpyDeviceToHost, st));
    bs.tm.stop(st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    bs.n_genes = h_scalars[0];
    bs.n_set = h_scalars[1];
    if (bs.n_genes > 65536 && !ctx->wide_ids)
        return fail(ctx, SHK_E_LIMIT,
                    "%u gene indices: the reference stores gene ids in 16 bits (small_vector.hpp:46), more than 65536 "
                    "need SHK_F_WIDE_IDS (shark-b200 --wide-ids)",
                    bs.n_genes);
    if (bs.n_set >= 0x7FFFFFFFu || bs.n_windows >= 0xFFFFFFFFull)
        return fail(ctx, SHK_E_LIMIT, "too many set bits / windows");
    return SHK_OK;
}

// Sharded build: bit index -> rank for this context's own window ends only (no counting yet).
__global__ void __launch_bounds__(256)
rank_convert_kernel(uint64_t *win_pos, uint64_t lo, uint64_t hi, const uint32_t *__restrict__ sectors)
{
    uint64_t x = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= hi) return;
    uint64_t p = win_pos[x];
    if (p == kInvalidPos) return;
    uint64_t q = p >> 5;
    uint64_t sec = q / kWordsPerSector;
    uint32_t slot = (uint32_t)(q - sec * kWordsPerSector);
    const uint4 *sp = reinterpret_cast<const uint4 *>(sectors + sec * 8);
    uint4 a = sp[0], b = sp[1];
    Sector s;
    s.w[0] = a.x, s.w[1] = a.y, s.w[2] = a.z, s.w[3] = a.w, s.w[4] = b.x, s.w[5] = b.y, s.w[6] = b.z, s.w[7] = b.w;
    win_pos[x] = sector_rank(s, slot, (uint32_t)(p & 31));
}

// Sharded build: occurrences per rank over the gathered window array (+ the number of windows).
__global__ void __launch_bounds__(256)
count_ranks_kernel(const uint64_t *__restrict__ win, uint64_t total, uint32_t *cnt, unsigned long long *n_windows)
{
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t v = x < total ? win[x] : kInvalidPos;
    const bool valid = v != kInvalidPos;
    if (valid) atomicAdd(&cnt[(uint32_t)v], 1u);
    uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(n_windows, (unsigned long long)n);
}

// Pass 2 (main.cpp:154-189 -> BF::add_to_kmer), BF::switch_mode(2), front table, info.
// win_is_rank: the window array already holds ranks (sharded build, gathered from all shards).
template <class IdT>
static int build_lists_t(shk_ctx *ctx, BuildState &bs, bool win_is_rank)
{
    constexpr bool kWide = sizeof(IdT) == 4;
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    const uint64_t total = bs.total;
    const uint32_t n_set = bs.n_set;
    int rc;
    DevBuf<uint32_t> d_cnt, d_fill, d_tmp_off, d_long;
    DevBuf<IdT> d_tmp_ids;
    DevBuf<uint32_t> d_csr_off;
    DevBuf<IdT> d_csr_ids;
    DevBuf<uint64_t> d_entries;
    SHK_CUDA(ctx, d_cnt.alloc((uint64_t)n_set + 1));
    SHK_CUDA(ctx, d_fill.alloc((uint64_t)n_set + 1));
    SHK_CUDA(ctx, d_tmp_off.alloc((uint64_t)n_set + 1));
    SHK_CUDA(ctx, d_long.alloc((uint64_t)n_set + 1));
    // the sharded build learns the total number of windows only while counting: one id per position bounds it
    SHK_CUDA(ctx, d_tmp_ids.alloc((win_is_rank ? total : bs.n_windows) + 1));
    SHK_CUDA(ctx, d_csr_off.alloc((uint64_t)n_set + 1));
    SHK_CUDA(ctx, d_entries.alloc((uint64_t)n_set + 1));
    bs.tm.start(st);
    SHK_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, ((uint64_t)n_set + 1) * 4, st));
    SHK_CUDA(ctx, cudaMemsetAsync(d_fill.p, 0, ((uint64_t)n_set + 1) * 4, st));
    uint64_t tot_ids = 0;
    if (n_set > 0) {
        unsigned blocks_x = (unsigned)((total + 255) / 256);
        if (win_is_rank) {
            SHK_CUDA(ctx, cudaMemsetAsync(bs.nwin.p, 0, 8, st));
            count_ranks_kernel<<<blocks_x, 256, 0, st>>>(bs.win.p, total, d_cnt.p, bs.nwin.p);
            SHK_CUDA(ctx, cudaMemcpyAsync(&bs.n_windows, bs.nwin.p, 8, cudaMemcpyDeviceToHost, st));
        } else {
            rank_count_kernel<<<blocks_x, 256, 0, st>>>(bs.win.p, total, ix.sectors, d_cnt.p);
        }
        ctx->launches += 1;
        rc = exclusive_scan(ctx, st, U32In{d_cnt.p}, U32Out{d_tmp_off.p}, (uint64_t)n_set, bs.tiles.p, d_tmp_off.p + n_set);
        if (rc) return rc;
        fill_kernel<IdT><<<blocks_x, 256, 0, st>>>(bs.win.p, total, bs.rec_off.p, bs.n_rec, bs.nidx.p, d_tmp_off.p, d_fill.p,
                                                   d_tmp_ids.p);
        ctx->launches += 1;
        rc = finish_lists<IdT>(ctx, st, n_set, d_tmp_off.p, d_tmp_ids.p, d_fill.p, d_long.p, bs.scalars.p + 4, bs.tiles.p,
                               d_csr_off, d_csr_ids, d_entries, tot_ids, &bs.tm, bs.n_genes);
        if (rc) return rc;
    } else {
        SHK_CUDA(ctx, cudaMemsetAsync(d_csr_off.p, 0, 4, st));
        bs.tm.stop(st);
        SHK_CUDA(ctx, d_csr_ids.alloc(1));
    }
    bs.tm.stop(st);
    // front table over the finished entries (16-bit ids only: with SHK_F_WIDE_IDS every read takes the
    // reference-shaped path filter word -> sector rank -> entry -> CSR)
    ix.info.n_set_bits = n_set;
    if constexpr (kWide) {
        ix.fgeom = FrontGeom{};
        ix.egeom = ExtGeom{};
        ix.info.front_shift = 0, ix.info.front_entries = 0, ix.info.extend = 0, ix.info.ref_bases = total, ix.info.coarse_shift = 0;
    } else {
        uint64_t need_tiles = (((ix.geom.bf_bits + 31) >> 5) + kScanTile - 1) / kScanTile + 2;
        DevBuf<uint32_t> d_tiles2;
        SHK_CUDA(ctx, d_tiles2.alloc(need_tiles));
        rc = build_front(ctx, st, d_entries.p, reinterpret_cast<const uint16_t *>(d_csr_ids.p), d_tiles2.p, bs.scalars.p + 5,
                         ExtBuildInputs{bs.bases.p, bs.win.p, total}, &bs.tm);
        if (rc) return rc;
    }
    bs.tm.stop(st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    if (bs.n_windows >= 0xFFFFFFFFull) return fail(ctx, SHK_E_LIMIT, "too many windows");

    ix.entries = d_entries.release();
    ix.csr_off = d_csr_off.release();
    if constexpr (kWide) ix.csr_ids32 = reinterpret_cast<uint32_t *>(d_csr_ids.release());
    else ix.csr_ids = reinterpret_cast<uint16_t *>(d_csr_ids.release());
    ix.info.id_bits = kWide ? 32u : 16u;
    ix.info.n_records = bs.n_rec;
    ix.info.n_genes = bs.n_genes;
    ix.info.n_set_bits = n_set;
    ix.info.tot_ids = tot_ids;
    ix.info.n_windows = bs.n_windows;
    ix.info.bf_bits = ix.geom.bf_bits;
    ix.info.device_bytes = ix.geom.n_sectors * 32 + ((uint64_t)n_set + 1) * (8 + 4) + tot_ids * sizeof(IdT) +
                           ix.fgeom.n_entries * 16 * ix.fgeom.stride +
                           (ix.egeom.enabled ? ix.egeom.estream_words * 8 + ix.egeom.ref2_words * 8 + ix.egeom.coarse_words * 4 : 0);
    ix.info.build_ms = bs.tm.total();
    ix.info.build_wall_ms = (float)(wall_ms() - bs.t_wall0);
    ix.built = true;
    return SHK_OK;
}

static int build_lists(shk_ctx *ctx, BuildState &bs, bool win_is_rank)
{
    return ctx->wide_ids ? build_lists_t<uint32_t>(ctx, bs, win_is_rank) : build_lists_t<uint16_t>(ctx, bs, win_is_rank);
}

int index_build_device(shk_ctx *ctx, const uint8_t *ref_bases, const uint64_t *rec_off, uint32_t n_rec)
{
    BuildState bs;
    int rc = build_pass1(ctx, bs, ref_bases, rec_off, n_rec, 0, rec_off[n_rec]);
    if (rc) return rc;
    if ((rc = build_rank(ctx, bs)) != 0) return rc;
    if ((rc = build_lists(ctx, bs, false)) != 0) return rc;
    ctx->index.info.n_shards = 1;
    return SHK_OK;
}

// =============================================================================================
// Sharded build (second mode): every context enumerates the windows of one shard of
// the reference records into its own filter; the filters are OR-merged with a P2P kernel (slice
// reduce, then slice gather: each context reads 2 (n-1)/n of a filter over NVLink instead of n-1
// filters); every context then builds the same rank directory, converts its own windows to ranks,
// gathers the other shards' ranks (disjoint position ranges) and finishes pass 2 locally.  The
// result is the index of the one-call build, on every context.  The caller provides the barriers
// between the steps (every call returns with its device work finished).
// =============================================================================================
struct ShardBuild {
    BuildState bs;
    uint32_t shard = 0, n_shards = 0;
    std::vector<uint64_t> cut;  // n_shards + 1 position cuts (record boundaries)
    int step = 0;               // 1 begun, 2 merged (phase 1), 3 merged (phase 2), 4 ranked
};

constexpr int kMaxShards = 16;
struct PeerVecs {
    const uint4 *p[kMaxShards];
    int n;
};

// local[i] |= peer[i] for every peer, i in [v0, v1) (128-bit words).  Peers do not write this
// slice while it runs (they reduce their own slices).
__global__ void __launch_bounds__(256)
or_merge_kernel(uint4 *__restrict__ local, PeerVecs peers, uint64_t v0, uint64_t v1)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += stride) {
        uint4 v = local[i];
        for (int j = 0; j < peers.n; ++j) {
            const uint4 t = peers.p[j][i];
            v.x |= t.x, v.y |= t.y, v.z |= t.z, v.w |= t.w;
        }
        local[i] = v;
    }
}

// dst[i] = src[i] over NVLink (or within the device), i in [v0, v1)
__global__ void __launch_bounds__(256)
p2p_gather_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint64_t v0, uint64_t v1)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += stride) dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
p2p_gather_u64_kernel(uint64_t *__restrict__ dst, const uint64_t *__restrict__ src, uint64_t lo, uint64_t hi)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) dst[i] = src[i];
}

struct PeerFlags {
    const uint32_t *p[kMaxShards];
    int n;
};
__global__ void __launch_bounds__(256) or_flags_kernel(uint32_t *local, PeerFlags peers, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = local[i];
    for (int j = 0; j < peers.n; ++j) v |= peers.p[j][i];
    local[i] = v;
}

void shard_free(shk_ctx *ctx)
{
    delete ctx->shard;
    ctx->shard = nullptr;
}

// Position cuts at record boundaries, balanced by bases: cut[s] = the record boundary nearest to
// s * total / n.  A pure function of the offsets: identical on every rank.
static void shard_cuts(const uint64_t *rec_off, uint32_t n_rec, uint32_t n_shards, std::vector<uint64_t> &cut)
{
    const uint64_t total = rec_off[n_rec];
    cut.assign(n_shards + 1, total);
    cut[0] = 0;
    for (uint32_t s = 1; s < n_shards; ++s) {
        const uint64_t want = (uint64_t)((unsigned __int128)total * s / n_shards);
        const uint64_t *it = std::lower_bound(rec_off, rec_off + n_rec + 1, want);  // *it >= want
        uint64_t c = *it;
        if (it != rec_off && want - it[-1] < c - want) c = it[-1];  // the nearer boundary
        cut[s] = std::max(c, cut[s - 1]);
    }
}

void shard_cuts_host(const uint64_t *rec_off, uint32_t n_rec, uint32_t n_shards, uint64_t *cuts)
{
    std::vector<uint64_t> c;
    shard_cuts(rec_off, n_rec, n_shards, c);
    std::copy(c.begin(), c.end(), cuts);
}

static void enable_peer(int from_dev, int to_dev)
{
    if (from_dev == to_dev) return;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, from_dev, to_dev) == cudaSuccess && can) cudaDeviceEnablePeerAccess(to_dev, 0);
    cudaGetLastError();  // "already enabled" is fine
}

int shard_begin(shk_ctx *ctx, const uint8_t *ref_bases, const uint64_t *rec_off, uint32_t n_rec, uint32_t shard,
                uint32_t n_shards, shk_shard_mem *mine)
{
    if (n_shards == 0 || n_shards > (uint32_t)kMaxShards || shard >= n_shards)
        return fail(ctx, SHK_E_ARG, "shard %u of %u: at most %d shards", shard, n_shards, kMaxShards);
    shard_free(ctx);
    ctx->shard = new (std::nothrow) ShardBuild;
    if (!ctx->shard) return fail(ctx, SHK_E_NOMEM, "out of host memory");
    ShardBuild &sb = *ctx->shard;
    sb.shard = shard, sb.n_shards = n_shards;
    shard_cuts(rec_off, n_rec, n_shards, sb.cut);
    int rc = build_pass1(ctx, sb.bs, ref_bases, rec_off, n_rec, sb.cut[shard], sb.cut[shard + 1]);
    if (rc) return rc;
    cudaStream_t st = ctx->build_stream;
    sb.bs.tm.stop(st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    memset(mine, 0, sizeof *mine);
    mine->dev_ptr[0] = ctx->index.sectors;
    mine->dev_ptr[1] = sb.bs.win.p;
    mine->dev_ptr[2] = sb.bs.has.p;
    mine->device = ctx->device;
    mine->shard = shard;
    mine->pid = (int64_t)getpid();
    mine->ipc_ok = 1;
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, mine->dev_ptr[i]) != cudaSuccess) {
            cudaGetLastError();
            mine->ipc_ok = 0;  // same-process peers still work through dev_ptr
            break;
        }
        static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(mine->ipc_handle[i], &h, 64);
    }
    sb.step = 1;
    return SHK_OK;
}

int shard_open(shk_ctx *ctx, const shk_shard_mem *peer, shk_shard_mem *opened)
{
    *opened = *peer;
    if (peer->pid == (int64_t)getpid()) {  // same process: the pointers are usable as they are
        enable_peer(ctx->device, peer->device);
        return SHK_OK;
    }
    if (!peer->ipc_ok) return fail(ctx, SHK_E_CUDA, "shard %u exported no CUDA IPC handles", peer->shard);
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t h;
        memcpy(&h, peer->ipc_handle[i], 64);
        void *p = nullptr;
        SHK_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        opened->dev_ptr[i] = p;
    }
    opened->pid = -(int64_t)getpid();  // marks pointers that shk_shard_close must release
    return SHK_OK;
}

int shard_close(shk_ctx *ctx, shk_shard_mem *opened)
{
    if (opened->pid != -(int64_t)getpid()) return SHK_OK;
    for (int i = 0; i < 3; ++i)
        if (opened->dev_ptr[i]) {
            cudaIpcCloseMemHandle(opened->dev_ptr[i]);
            opened->dev_ptr[i] = nullptr;
        }
    cudaGetLastError();
    opened->pid = 0;
    return SHK_OK;
}

int shard_merge(shk_ctx *ctx, int phase, const shk_shard_mem *all)
{
    ShardBuild *sbp = ctx->shard;
    if (!sbp || sbp->step != phase) return fail(ctx, SHK_E_STATE, "shk_shard_merge(%d) out of order", phase);
    ShardBuild &sb = *sbp;
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    const uint32_t n = sb.n_shards, me = sb.shard;
    const uint64_t n_vec = ix.geom.n_sectors * 2;  // 128-bit words of the filter
    auto slice = [&](uint32_t s) { return (uint64_t)((unsigned __int128)n_vec * s / n); };
    const unsigned blocks = (unsigned)ctx->sm_count * 8;
    uint4 *local = reinterpret_cast<uint4 *>(ix.sectors);
    sb.bs.tm.start(st);
    if (phase == 1) {
        PeerVecs pv{};
        PeerFlags pf{};
        for (uint32_t s = 0; s < n; ++s) {
            if (s == me) continue;
            pv.p[pv.n++] = reinterpret_cast<const uint4 *>(all[s].dev_ptr[0]);
            pf.p[pf.n++] = reinterpret_cast<const uint32_t *>(all[s].dev_ptr[2]);
        }
        if (pv.n && slice(me + 1) > slice(me)) {
            or_merge_kernel<<<blocks, 256, 0, st>>>(local, pv, slice(me), slice(me + 1));
            ctx->launches += 1;
        }
        // window flags: records belong to exactly one shard, the OR over all shards is the full array
        if (pf.n && sb.bs.n_rec) {
            or_flags_kernel<<<(sb.bs.n_rec + 255) / 256, 256, 0, st>>>(sb.bs.has.p, pf, sb.bs.n_rec);
            ctx->launches += 1;
        }
    } else {
        for (uint32_t s = 0; s < n; ++s) {
            if (s == me || slice(s + 1) <= slice(s)) continue;
            p2p_gather_kernel<<<blocks, 256, 0, st>>>(local, reinterpret_cast<const uint4 *>(all[s].dev_ptr[0]), slice(s),
                                                      slice(s + 1));
            ctx->launches += 1;
        }
    }
    SHK_CUDA(ctx, cudaGetLastError());
    sb.bs.tm.stop(st);
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    sb.step = phase + 1;
    return SHK_OK;
}

int shard_rank(shk_ctx *ctx)
{
    ShardBuild *sbp = ctx->shard;
    if (!sbp || sbp->step != 3) return fail(ctx, SHK_E_STATE, "shk_shard_rank out of order");
    ShardBuild &sb = *sbp;
    int rc = build_rank(ctx, sb.bs);
    if (rc) return rc;
    cudaStream_t st = ctx->build_stream;
    if (sb.bs.hi > sb.bs.lo && sb.bs.n_set > 0) {
        sb.bs.tm.start(st);
        rank_convert_kernel<<<(unsigned)((sb.bs.hi - sb.bs.lo + 255) / 256), 256, 0, st>>>(sb.bs.win.p, sb.bs.lo, sb.bs.hi,
                                                                                          ctx->index.sectors);
        ctx->launches += 1;
        SHK_CUDA(ctx, cudaGetLastError());
        sb.bs.tm.stop(st);
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
    }
    sb.step = 4;
    return SHK_OK;
}

int shard_finish(shk_ctx *ctx, const shk_shard_mem *all)
{
    ShardBuild *sbp = ctx->shard;
    if (!sbp || sbp->step != 4) return fail(ctx, SHK_E_STATE, "shk_shard_finish out of order");
    ShardBuild &sb = *sbp;
    cudaStream_t st = ctx->build_stream;
    const unsigned blocks = (unsigned)ctx->sm_count * 8;
    sb.bs.tm.start(st);
    for (uint32_t s = 0; s < sb.n_shards; ++s) {
        if (s == sb.shard || sb.cut[s + 1] <= sb.cut[s]) continue;
        p2p_gather_u64_kernel<<<blocks, 256, 0, st>>>(sb.bs.win.p, reinterpret_cast<const uint64_t *>(all[s].dev_ptr[1]),
                                                      sb.cut[s], sb.cut[s + 1]);
        ctx->launches += 1;
    }
    SHK_CUDA(ctx, cudaGetLastError());
    sb.bs.tm.stop(st);  // build_lists allocates first; the gathers run meanwhile
    int rc = build_lists(ctx, sb.bs, true);
    if (rc) return rc;
    ctx->index.info.n_shards = sb.n_shards;
    sb.step = 5;
    return SHK_OK;
}

__global__ void __launch_bounds__(256) widen_ids_kernel(const uint16_t *__restrict__ in, uint64_t n, uint32_t *out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

int index_export_device(shk_ctx *ctx, uint64_t *pos, uint32_t *off, uint16_t *ids, uint32_t *ids32)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    const uint64_t n_set = ix.info.n_set_bits;
    if (pos && n_set) {
        DevBuf<uint64_t> d_pos;
        SHK_CUDA(ctx, d_pos.alloc(n_set));
        unsigned blocks = (unsigned)((ix.geom.n_sectors + 255) / 256);
        export_positions_kernel<<<blocks, 256, 0, st>>>(ix.sectors, ix.geom.n_sectors, d_pos.p);
        ctx->launches += 1;
        SHK_CUDA(ctx, cudaGetLastError());
        SHK_CUDA(ctx, cudaMemcpyAsync(pos, d_pos.p, n_set * 8, cudaMemcpyDeviceToHost, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (off) SHK_CUDA(ctx, cudaMemcpyAsync(off, ix.csr_off, (n_set + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (ids && ix.info.tot_ids) {
        if (!ix.csr_ids) return fail(ctx, SHK_E_STATE, "the index holds 32-bit ids (SHK_F_WIDE_IDS): use shk_index_export_wide");
        SHK_CUDA(ctx, cudaMemcpyAsync(ids, ix.csr_ids, ix.info.tot_ids * 2, cudaMemcpyDeviceToHost, st));
    }
    if (ids32 && ix.info.tot_ids) {
        if (ix.csr_ids32) {
            SHK_CUDA(ctx, cudaMemcpyAsync(ids32, ix.csr_ids32, ix.info.tot_ids * 4, cudaMemcpyDeviceToHost, st));
        } else {
            DevBuf<uint32_t> d_w;
            SHK_CUDA(ctx, d_w.alloc(ix.info.tot_ids));
            widen_ids_kernel<<<(unsigned)((ix.info.tot_ids + 255) / 256), 256, 0, st>>>(ix.csr_ids, ix.info.tot_ids, d_w.p);
            ctx->launches += 1;
            SHK_CUDA(ctx, cudaGetLastError());
            SHK_CUDA(ctx, cudaMemcpyAsync(ids32, d_w.p, ix.info.tot_ids * 4, cudaMemcpyDeviceToHost, st));
            SHK_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    return SHK_OK;
}

// =============================================================================================
// K5: stand-alone probe = BF::get_index (bloomfilter.h:78-102), one thread per k-mer.
// =============================================================================================
template <int MOD>
__global__ void __launch_bounds__(256)
probe_kernel(const uint64_t *__restrict__ kmers, uint64_t n, FilterGeom g, const uint32_t *__restrict__ sectors,
             const uint32_t *__restrict__ csr_off, long long *rank, uint32_t *begin, uint32_t *len)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t p = bit_index<MOD>(xxh64_u64(kmers[i]), g);
    uint64_t q = p >> 5;
    uint64_t sec = q / kWordsPerSector;
    uint32_t slot = (uint32_t)(q - sec * kWordsPerSector), bit = (uint32_t)(p & 31);
    Sector s = ld_sector(reinterpret_cast<const Sector *>(sectors) + sec);
    if ((s.w[slot] >> bit) & 1u) {  // `_bf[bf_idx]`
        uint32_t r = sector_rank(s, slot, bit);  // `_brank(bf_idx + 1) - 1`
        rank[i] = r;
        uint32_t b = csr_off[r];  // `_select_bv(rank - 1) + 1`
        begin[i] = b;
        len[i] = csr_off[r + 1] - b;  // `_select_bv(rank)` inclusive end
    } else {
        rank[i] = -1;
        begin[i] = 0;
        len[i] = 0;
    }
}

int probe_device(shk_ctx *ctx, const uint64_t *kmers, uint64_t n, int64_t *rank, uint32_t *begin, uint32_t *len)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    if (n == 0) return SHK_OK;
    DevBuf<uint64_t> d_k;
    DevBuf<long long> d_rank;
    DevBuf<uint32_t> d_begin, d_len;
    SHK_CUDA(ctx, d_k.alloc(n));
    SHK_CUDA(ctx, d_rank.alloc(n));
    SHK_CUDA(ctx, d_begin.alloc(n));
    SHK_CUDA(ctx, d_len.alloc(n));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_k.p, kmers, n * 8, cudaMemcpyHostToDevice, st));
    unsigned blocks = (unsigned)((n + 255) / 256);
    switch (ix.geom.mod_kind) {
    case MOD_POW2: probe_kernel<MOD_POW2><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.csr_off, d_rank.p, d_begin.p, d_len.p); break;
    case MOD_B33: probe_kernel<MOD_B33><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.csr_off, d_rank.p, d_begin.p, d_len.p); break;
    default: probe_kernel<MOD_GENERIC><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.csr_off, d_rank.p, d_begin.p, d_len.p); break;
    }
    ctx->launches += 1;
    SHK_CUDA(ctx, cudaGetLastError());
    SHK_CUDA(ctx, cudaMemcpyAsync(rank, d_rank.p, n * 8, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaMemcpyAsync(begin, d_begin.p, n * 4, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaMemcpyAsync(len, d_len.p, n * 4, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    return SHK_OK;
}

// Probe throughput on resident k-mers, through the hot path's own access sequence:
// one filter word (32 B sector) -> on a hit the whole sector (rank) -> the 8-byte entry.
template <int MOD>
__global__ void __launch_bounds__(256)
probe_bench_kernel(const uint64_t *__restrict__ kmers, uint64_t n, FilterGeom g, const uint32_t *__restrict__ sectors,
                   const uint64_t *__restrict__ entries, unsigned long long *hits, unsigned long long *checksum)
{
    constexpr int ILP = 4;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t my_hits = 0;
    uint64_t acc = 0;
    for (uint64_t i = i0; i < n; i += stride * ILP) {
        uint64_t p[ILP];
        uint32_t w[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            uint64_t idx = i + (uint64_t)j * stride;
            p[j] = idx < n ? bit_index<MOD>(xxh64_u64(kmers[idx]), g) : kInvalidPos;
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) w[j] = p[j] != kInvalidPos ? ld_filter_word(sectors + phys_word(p[j]), pol_first) : 0u;
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (p[j] != kInvalidPos && ((w[j] >> (p[j] & 31)) & 1u)) {
                uint64_t q = p[j] >> 5, sec = q / kWordsPerSector;
                Sector s = ld_sector(reinterpret_cast<const Sector *>(sectors) + sec);
                uint32_t r = sector_rank(s, (uint32_t)(q - sec * kWordsPerSector), (uint32_t)(p[j] & 31));
                acc += ld_u64_hint(entries + r, pol_last);
                ++my_hits;
            }
        }
    }
    my_hits = __reduce_add_sync(0xFFFFFFFFu, my_hits);
    if ((threadIdx.x & 31) == 0) {
        if (my_hits) atomicAdd(hits, (unsigned long long)my_hits);
    }
    if (acc == 0x1234567ULL) atomicAdd(checksum, acc);  // keeps the entry loads alive
}

int probe_bench_device(shk_ctx *ctx, const uint64_t *kmers, uint64_t n, uint32_t reps, float *ms, uint64_t *hits)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    DevBuf<uint64_t> d_k;
    DevBuf<unsigned long long> d_c;
    SHK_CUDA(ctx, d_k.alloc(n));
    SHK_CUDA(ctx, d_c.alloc(2));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_k.p, kmers, n * 8, cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    SHK_CUDA(ctx, cudaEventCreate(&e0));
    SHK_CUDA(ctx, cudaEventCreate(&e1));
    unsigned blocks = (unsigned)ctx->sm_count * 8;
    float best = 1e30f;
    for (uint32_t r = 0; r < reps + 1; ++r) {  // first repetition is the warm-up
        SHK_CUDA(ctx, cudaMemsetAsync(d_c.p, 0, 16, st));
        SHK_CUDA(ctx, cudaEventRecord(e0, st));
        switch (ix.geom.mod_kind) {
        case MOD_POW2: probe_bench_kernel<MOD_POW2><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.entries, d_c.p, d_c.p + 1); break;
        case MOD_B33: probe_bench_kernel<MOD_B33><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.entries, d_c.p, d_c.p + 1); break;
        default: probe_bench_kernel<MOD_GENERIC><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, ix.entries, d_c.p, d_c.p + 1); break;
        }
        ctx->launches += 1;
        SHK_CUDA(ctx, cudaEventRecord(e1, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        if (r > 0 && t < best) best = t;
    }
    unsigned long long h[2] = {0, 0};
    SHK_CUDA(ctx, cudaMemcpy(h, d_c.p, 16, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = best;
    if (hits) *hits = h[0];
    return SHK_OK;
}

// =============================================================================================
// B0: random 32-byte-sector loads over the filter allocation = the measured ceiling that the
// probe throughput is normalised by.  Independent addresses (splitmix64 of a
// counter), 8 loads in flight per thread, whole sector consumed.
// =============================================================================================
__global__ void __launch_bounds__(256)
random_sector_kernel(const Sector *__restrict__ base, uint64_t n_sectors, uint64_t n_loads, uint64_t seed,
                     unsigned long long *sink)
{
    constexpr int ILP = 8;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = i0; i < n_loads; i += stride * ILP) {
        Sector s[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            uint64_t idx = i + (uint64_t)j * stride;
            uint64_t a = __umul64hi(splitmix64(seed + idx), n_sectors);  // uniform in [0, n_sectors)
            if (idx < n_loads) s[j] = ld_sector(base + a);
            else s[j].w[0] = s[j].w[7] = 0;
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc ^= s[j].w[0] ^ s[j].w[7];
    }
    if (acc == 0x9E3779B9u) atomicAdd(sink, 1ULL);
}

int random_sector_bench_device(shk_ctx *ctx, uint64_t n_loads, uint64_t span_bytes, uint64_t seed, float *ms)
{
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    uint64_t n_sectors = ix.geom.n_sectors;
    if (span_bytes && span_bytes / 32 < n_sectors) n_sectors = span_bytes / 32;
    if (n_sectors == 0) return fail(ctx, SHK_E_ARG, "empty span");
    DevBuf<unsigned long long> d_sink;
    SHK_CUDA(ctx, d_sink.alloc(1));
    SHK_CUDA(ctx, cudaMemsetAsync(d_sink.p, 0, 8, st));
    cudaEvent_t e0, e1;
    SHK_CUDA(ctx, cudaEventCreate(&e0));
    SHK_CUDA(ctx, cudaEventCreate(&e1));
    unsigned blocks = (unsigned)ctx->sm_count * 8;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        SHK_CUDA(ctx, cudaEventRecord(e0, st));
        random_sector_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const Sector *>(ix.sectors), n_sectors, n_loads,
                                                     seed + (uint64_t)r * n_loads, d_sink.p);
        ctx->launches += 1;
        SHK_CUDA(ctx, cudaEventRecord(e1, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        if (r > 0 && t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = best;
    return SHK_OK;
}

// =============================================================================================
// Staged build = the reference's own functor protocol (seam 2), for callers that keep
// the reference's main.cpp and replace KmerBuilder / BloomfilterFiller / class BF one for one
// (include/shark_b200_functors.hpp).  Same device structures, same results as index_build_device.
// =============================================================================================

// KmerBuilder::operator() (KmerBuilder.hpp:40-72): the hash of every valid window, by window end.
// Same walk as K1 (8 positions per thread, rolling forward k-mer), but the full 64-bit hash is kept
// and no bit is set.
__global__ void __launch_bounds__(256)
enum_hashes_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ rec_off, uint32_t n_rec, uint64_t total,
                   int k, uint64_t *win_hash, uint32_t *win_valid)
{
    const uint64_t x0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kPosPerThread;
    if (x0 >= total) return;
    const uint64_t xend = min(x0 + (uint64_t)kPosPerThread, total);
    uint64_t s = x0 >= (uint64_t)(k - 1) ? x0 - (k - 1) : 0;
    uint32_t r = record_of(rec_off, n_rec, s);
    uint64_t next_b = rec_off[r + 1];
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    uint64_t fwd = 0;
    int run = 0;
    for (uint64_t pos = s; pos < xend; ++pos) {
        while (pos >= next_b) {
            ++r;
            next_b = rec_off[r + 1];
            run = 0;
        }
        const uint32_t ch = bases[pos];
        if (base_valid(ch)) {
            fwd = ((fwd << 2) | base_code(ch)) & kmask;
            ++run;
        } else {
            run = 0;
        }
        if (pos >= x0) {
            const bool v = run >= k;
            win_valid[pos] = v ? 1u : 0u;
            win_hash[pos] = v ? xxh64_u64(canonical(fwd, k)) : 0ULL;  // _get_hash(min(kmer, rckmer))
        }
    }
}

__global__ void __launch_bounds__(256)
compact_hashes_kernel(const uint64_t *__restrict__ win_hash, const uint32_t *__restrict__ win_valid,
                      const uint32_t *__restrict__ win_off, uint64_t total, uint64_t *out)
{
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x < total && win_valid[x]) out[win_off[x]] = win_hash[x];
}

int kmer_hashes_device(shk_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint64_t *hashes,
                       uint64_t cap, uint64_t *n_hashes)
{
    cudaStream_t st = ctx->build_stream;
    const uint64_t total = rec_off[n_rec];
    *n_hashes = 0;
    if (total >= (1ULL << 32)) return fail(ctx, SHK_E_LIMIT, "more than 4 Gbases in one call");
    for (uint32_t i = 0; i < n_rec; ++i)
        if (rec_off[i + 1] < rec_off[i]) return fail(ctx, SHK_E_ARG, "rec_offsets must be non-decreasing");
    if (total == 0) return SHK_OK;
    DevBuf<uint8_t> d_bases;
    DevBuf<uint64_t> d_rec_off, d_hash, d_out;
    DevBuf<uint32_t> d_valid, d_off, d_tiles;
    SHK_CUDA(ctx, d_bases.alloc(total));
    SHK_CUDA(ctx, d_rec_off.alloc((uint64_t)n_rec + 1));
    SHK_CUDA(ctx, d_hash.alloc(total));
    SHK_CUDA(ctx, d_valid.alloc(total));
    SHK_CUDA(ctx, d_off.alloc(total + 1));
    SHK_CUDA(ctx, d_tiles.alloc(total / kScanTile + 2));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_bases.p, bases, total, cudaMemcpyHostToDevice, st));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_rec_off.p, rec_off, ((uint64_t)n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint64_t threads = (total + kPosPerThread - 1) / kPosPerThread;
    enum_hashes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_bases.p, d_rec_off.p, n_rec, total,
                                                                          (int)ctx->params.k, d_hash.p, d_valid.p);
    ctx->launches += 1;
    int rc = exclusive_scan(ctx, st, U32In{d_valid.p}, U32Out{d_off.p}, total, d_tiles.p, d_off.p + total);
    if (rc) return rc;
    uint32_t h_n = 0;
    SHK_CUDA(ctx, cudaMemcpyAsync(&h_n, d_off.p + total, 4, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    *n_hashes = h_n;
    ctx->staged.hashed = true;
    if (h_n == 0) return SHK_OK;
    if (h_n > cap || !hashes) return fail(ctx, SHK_E_CAPACITY, "%u hashes do not fit into a buffer of %llu", h_n,
                                          (unsigned long long)cap);
    SHK_CUDA(ctx, d_out.alloc(h_n));
    compact_hashes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_hash.p, d_valid.p, d_off.p, total, d_out.p);
    ctx->launches += 1;
    SHK_CUDA(ctx, cudaGetLastError());
    SHK_CUDA(ctx, cudaMemcpyAsync(hashes, d_out.p, (uint64_t)h_n * 8, cudaMemcpyDeviceToHost, st));
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    return SHK_OK;
}

// BF::add_at (bloomfilter.h:57-59): `_bf[p % _size] = 1`
template <int MOD>
__global__ void __launch_bounds__(256)
add_at_kernel(const uint64_t *__restrict__ positions, uint64_t n, FilterGeom g, uint32_t *sectors)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t p = bit_index<MOD>(positions[i], g);
    atomicOr(&sectors[phys_word(p)], 1u << (p & 31));
}

int staged_add_at(shk_ctx *ctx, const uint64_t *positions, uint64_t n)
{
    if (ctx->staged.mode != 0) return fail(ctx, SHK_E_STATE, "add_at is only possible in mode 0 (mode is %d)", ctx->staged.mode);
    if (n == 0) return SHK_OK;
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    DevBuf<uint64_t> d_pos;
    SHK_CUDA(ctx, d_pos.alloc(n));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_pos.p, positions, n * 8, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    switch (ix.geom.mod_kind) {
    case MOD_POW2: add_at_kernel<MOD_POW2><<<blocks, 256, 0, st>>>(d_pos.p, n, ix.geom, ix.sectors); break;
    case MOD_B33: add_at_kernel<MOD_B33><<<blocks, 256, 0, st>>>(d_pos.p, n, ix.geom, ix.sectors); break;
    default: add_at_kernel<MOD_GENERIC><<<blocks, 256, 0, st>>>(d_pos.p, n, ix.geom, ix.sectors); break;
    }
    ctx->launches += 1;
    SHK_CUDA(ctx, cudaGetLastError());
    SHK_CUDA(ctx, cudaStreamSynchronize(st));  // positions may be freed by the caller
    return SHK_OK;
}

// BF::add_to_kmer (bloomfilter.h:64-70): kmer -> hash % size -> `_brank(bf_idx)` = set bits strictly
// below the position (whether or not the position itself is set, exactly like the reference).
template <int MOD>
__global__ void __launch_bounds__(256)
kmer_rank_kernel(const uint64_t *__restrict__ kmers, uint64_t n, FilterGeom g, const uint32_t *__restrict__ sectors,
                 uint32_t n_set, uint16_t gene, uint32_t *pair_rank, uint16_t *pair_gene, uint32_t *cnt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t p = bit_index<MOD>(xxh64_u64(kmers[i]), g);
    const uint64_t q = p >> 5;
    const uint64_t sec = q / kWordsPerSector;
    const uint32_t slot = (uint32_t)(q - sec * kWordsPerSector);
    const uint4 *sp = reinterpret_cast<const uint4 *>(sectors + sec * 8);
    const uint4 a = sp[0], b = sp[1];
    Sector s;
    s.w[0] = a.x, s.w[1] = a.y, s.w[2] = a.z, s.w[3] = a.w, s.w[4] = b.x, s.w[5] = b.y, s.w[6] = b.z, s.w[7] = b.w;
    uint32_t r = sector_rank(s, slot, (uint32_t)(p & 31));
    if (r >= n_set) r = 0xFFFFFFFFu;  // `_set_index[num_kmer]`: out of bounds in the reference, dropped here
    else atomicAdd(&cnt[r], 1u);
    pair_rank[i] = r;
    pair_gene[i] = gene;
}

__global__ void __launch_bounds__(256)
pairs_fill_kernel(const uint32_t *__restrict__ pair_rank, const uint16_t *__restrict__ pair_gene, uint64_t n,
                  const uint32_t *__restrict__ tmp_off, uint32_t *fill, uint16_t *tmp_ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = pair_rank[i];
    if (r == 0xFFFFFFFFu) return;
    tmp_ids[tmp_off[r] + atomicAdd(&fill[r], 1u)] = pair_gene[i];
}

void staged_free(shk_ctx *ctx)
{
    StagedBuild &sb = ctx->staged;
    if (sb.cnt) cudaFree(sb.cnt);
    if (sb.pair_rank) cudaFree(sb.pair_rank);
    if (sb.pair_gene) cudaFree(sb.pair_gene);
    sb.cnt = nullptr, sb.pair_rank = nullptr, sb.pair_gene = nullptr;
    sb.n_pairs = sb.cap = 0;
}

int staged_add_to_kmer(shk_ctx *ctx, const uint64_t *kmers, uint64_t n, int32_t input_idx)
{
    StagedBuild &sb = ctx->staged;
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    if (sb.mode != 1) return SHK_OK;  // `if (_mode != 1) return;` (bloomfilter.h:62-63)
    if (input_idx < 0 || input_idx > 65535)
        return fail(ctx, SHK_E_LIMIT, "gene index %d: the reference stores gene ids in 16 bits (small_vector.hpp:46)", input_idx);
    if (input_idx < sb.last_idx)
        return fail(ctx, SHK_E_STATE, "add_to_kmer: input_idx %d after %d; indices must not decrease (main.cpp:159-187)",
                    input_idx, sb.last_idx);
    sb.last_idx = input_idx;
    if (n == 0) return SHK_OK;
    if (sb.n_pairs + n >= 0xFFFFFFFFull) return fail(ctx, SHK_E_LIMIT, "too many k-mer occurrences");
    if (sb.n_pairs + n > sb.cap) {  // grow geometrically, keep what is there
        uint64_t cap = std::max<uint64_t>(sb.cap * 2, std::max<uint64_t>(sb.n_pairs + n, 1u << 20));
        uint32_t *nr = nullptr;
        uint16_t *ng = nullptr;
        SHK_CUDA(ctx, cudaMalloc((void **)&nr, cap * 4));
        cudaError_t e = cudaMalloc((void **)&ng, cap * 2);
        if (e != cudaSuccess) {
            cudaFree(nr);
            SHK_CUDA(ctx, e);
        }
        if (sb.n_pairs) {
            cudaMemcpyAsync(nr, sb.pair_rank, sb.n_pairs * 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(ng, sb.pair_gene, sb.n_pairs * 2, cudaMemcpyDeviceToDevice, st);
            cudaStreamSynchronize(st);
        }
        if (sb.pair_rank) cudaFree(sb.pair_rank);
        if (sb.pair_gene) cudaFree(sb.pair_gene);
        sb.pair_rank = nr, sb.pair_gene = ng, sb.cap = cap;
    }
    DevBuf<uint64_t> d_k;
    SHK_CUDA(ctx, d_k.alloc(n));
    SHK_CUDA(ctx, cudaMemcpyAsync(d_k.p, kmers, n * 8, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    uint32_t *pr = sb.pair_rank + sb.n_pairs;
    uint16_t *pg = sb.pair_gene + sb.n_pairs;
    const uint16_t gene = (uint16_t)input_idx;
    switch (ix.geom.mod_kind) {
    case MOD_POW2: kmer_rank_kernel<MOD_POW2><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, sb.n_set, gene, pr, pg, sb.cnt); break;
    case MOD_B33: kmer_rank_kernel<MOD_B33><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, sb.n_set, gene, pr, pg, sb.cnt); break;
    default: kmer_rank_kernel<MOD_GENERIC><<<blocks, 256, 0, st>>>(d_k.p, n, ix.geom, ix.sectors, sb.n_set, gene, pr, pg, sb.cnt); break;
    }
    ctx->launches += 1;
    SHK_CUDA(ctx, cudaGetLastError());
    SHK_CUDA(ctx, cudaStreamSynchronize(st));
    sb.n_pairs += n;
    return SHK_OK;
}

int staged_switch_mode(shk_ctx *ctx, int new_mode, uint64_t *n_set_bits)
{
    StagedBuild &sb = ctx->staged;
    DeviceIndex &ix = ctx->index;
    cudaStream_t st = ctx->build_stream;
    if (sb.mode == 0 && new_mode == 1) {
        // rank directory + num_kmer (bloomfilter.h:121-122); `_set_index.resize(num_kmer)` = cnt
        DevBuf<uint32_t> d_tiles, d_total;
        SHK_CUDA(ctx, d_tiles.alloc(ix.geom.n_sectors / kScanTile + 2));
        SHK_CUDA(ctx, d_total.alloc(1));
        int rc = exclusive_scan(ctx, st, SectorPopIn{ix.sectors}, SectorRankOut{ix.sectors}, ix.geom.n_sectors, d_tiles.p,
                                d_total.p);
        if (rc) return rc;
        uint32_t n_set = 0;
        SHK_CUDA(ctx, cudaMemcpyAsync(&n_set, d_total.p, 4, cudaMemcpyDeviceToHost, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
        if (n_set >= 0x7FFFFFFFu) return fail(ctx, SHK_E_LIMIT, "too many set bits");
        staged_free(ctx);
        SHK_CUDA(ctx, cudaMalloc((void **)&sb.cnt, ((uint64_t)n_set + 1) * 4));
        SHK_CUDA(ctx, cudaMemsetAsync(sb.cnt, 0, ((uint64_t)n_set + 1) * 4, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
        sb.n_set = n_set;
        sb.last_idx = -1;
        sb.mode = 1;
        if (n_set_bits) *n_set_bits = n_set;
        return SHK_OK;
    }
    if (sb.mode == 1 && new_mode == 2) {
        free_index_arrays(ix);
        cudaEvent_t e0, e1;
        SHK_CUDA(ctx, cudaEventCreate(&e0));
        SHK_CUDA(ctx, cudaEventCreate(&e1));
        SHK_CUDA(ctx, cudaEventRecord(e0, st));
        const uint32_t n_set = sb.n_set;
        DevBuf<uint32_t> d_fill, d_tmp_off, d_long, d_scalars, d_tiles, d_csr_off;
        DevBuf<uint16_t> d_tmp_ids, d_csr_ids;
        DevBuf<uint64_t> d_entries;
        SHK_CUDA(ctx, d_fill.alloc((uint64_t)n_set + 1));
        SHK_CUDA(ctx, d_tmp_off.alloc((uint64_t)n_set + 1));
        SHK_CUDA(ctx, d_long.alloc((uint64_t)n_set + 1));
        SHK_CUDA(ctx, d_scalars.alloc(8));
        SHK_CUDA(ctx, d_tiles.alloc((uint64_t)n_set / kScanTile + 2));
        SHK_CUDA(ctx, d_tmp_ids.alloc(sb.n_pairs + 1));
        SHK_CUDA(ctx, d_csr_off.alloc((uint64_t)n_set + 1));
        SHK_CUDA(ctx, d_entries.alloc((uint64_t)n_set + 1));
        SHK_CUDA(ctx, cudaMemsetAsync(d_fill.p, 0, ((uint64_t)n_set + 1) * 4, st));
        SHK_CUDA(ctx, cudaMemsetAsync(d_scalars.p, 0, 32, st));
        uint64_t tot_ids = 0;
        if (n_set > 0) {
            int rc = exclusive_scan(ctx, st, U32In{sb.cnt}, U32Out{d_tmp_off.p}, (uint64_t)n_set, d_tiles.p, d_tmp_off.p + n_set);
            if (rc) return rc;
            if (sb.n_pairs) {
                pairs_fill_kernel<<<(unsigned)((sb.n_pairs + 255) / 256), 256, 0, st>>>(sb.pair_rank, sb.pair_gene, sb.n_pairs,
                                                                                        d_tmp_off.p, d_fill.p, d_tmp_ids.p);
                ctx->launches += 1;
            }
            rc = finish_lists<uint16_t>(ctx, st, n_set, d_tmp_off.p, d_tmp_ids.p, d_fill.p, d_long.p, d_scalars.p + 4, d_tiles.p,
                                        d_csr_off, d_csr_ids, d_entries, tot_ids);
            if (rc) return rc;
        } else {
            SHK_CUDA(ctx, cudaMemsetAsync(d_csr_off.p, 0, 4, st));
            SHK_CUDA(ctx, d_csr_ids.alloc(1));
        }
        ix.info = shk_index_info{};
        ix.info.n_set_bits = n_set;
        ix.info.id_bits = 16;
        {
            const uint64_t need_tiles = (((ix.geom.bf_bits + 31) >> 5) + kScanTile - 1) / kScanTile + 2;
            DevBuf<uint32_t> d_tiles2;
            SHK_CUDA(ctx, d_tiles2.alloc(need_tiles));
            // no reference text at this seam: the extension structures are not built
            int rc = build_front(ctx, st, d_entries.p, d_csr_ids.p, d_tiles2.p, d_scalars.p + 5, ExtBuildInputs{nullptr, nullptr, 0});
            if (rc) return rc;
        }
        SHK_CUDA(ctx, cudaEventRecord(e1, st));
        SHK_CUDA(ctx, cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        ix.entries = d_entries.release();
        ix.csr_off = d_csr_off.release();
        ix.csr_ids = d_csr_ids.release();
        ix.info.n_genes = (uint32_t)(sb.last_idx + 1);
        ix.info.n_records = ix.info.n_genes;
        ix.info.tot_ids = tot_ids;
        ix.info.n_windows = sb.n_pairs;
        ix.info.bf_bits = ix.geom.bf_bits;
        ix.info.device_bytes = ix.geom.n_sectors * 32 + ((uint64_t)n_set + 1) * (8 + 4) + tot_ids * 2 +
                               ix.fgeom.n_entries * 16 * ix.fgeom.stride;
        ix.info.build_ms = ms;
        ix.built = true;
        staged_free(ctx);
        sb.mode = 2;
        if (n_set_bits) *n_set_bits = n_set;
        return SHK_OK;
    }
    return fail(ctx, SHK_E_STATE, "switch_mode(%d) in mode %d: only 0 -> 1 and 1 -> 2 exist (bloomfilter.h:112-184)", new_mode,
                sb.mode);
}

}  // namespace shk

// Read classification on the device: replaces FastqSplitter's masking rule, ReadAnalyzer::operator()
// and the ordering contract of ReadOutput.  Citations are reference file:line.
//
// Kernels of one chunk (all on the slot's stream):
//   analyze_reads_kernel   one warp per read: mask, 2-bit pack, canonical k-mers, front-table probe
//                          (one 16-byte load per k-mer), per-gene coverage/hits in a
//                          register-resident 8-gene table, argmax, threshold
//   analyze_slow_kernel    exact path for reads the fast path gives up on (more than 8 genes, a
//                          list longer than 8, or a text longer than 1024 bytes)
//   scan_tile_sums_kernel  exclusive scan of the per-tile association counts
//   scatter_assoc_kernel   ordered (read_idx, gene_idx) list + keep flags
#include "shk_internal.h"
#include "shk_scan.cuh"

namespace shk {

constexpr int kSlots = 8;        // genes tracked per read on the fast path
constexpr int kGroup = 3;        // rounds of 32 windows whose probes are in flight together
constexpr int kWarpsPerCta = 8;
constexpr int kReadsPerWarp = kReadsPerTile / kWarpsPerCta;
constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    uint32_t lo = __shfl_sync(kFull, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(kFull, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// Union of the intervals [e-k+1, e] over the set bits e of M (64 positions, bit x = position x):
// position x is covered iff some e in [x, x+k-1] is set.
__device__ __forceinline__ uint64_t dilate_down(uint64_t m, int k)
{
    int covered = 1;
    while (covered * 2 <= k) {
        m |= m >> covered;
        covered *= 2;
    }
    if (covered < k) m |= m >> (k - covered);
    return m;
}

// Per-read gene table of the fast path, held in registers of the whole warp:
//   slot s (0..kSlots-1): gene id lives in lane s of `gene`; its hit bitmask over window end
//   positions is spread over the lanes - lane c holds the word for positions [32c, 32c+32).
struct WarpTable {
    uint32_t gene;
    uint32_t mask[kSlots];
    int nslots;
    bool overflow;

    __device__ __forceinline__ void init()
    {
        gene = 0xFFFFFFFFu;
        nslots = 0;
        overflow = false;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) mask[s] = 0;
    }
    // all arguments are warp-uniform: gene g was hit by the windows ending at positions
    // base + i for every set bit i of m
    __device__ __forceinline__ void update(uint32_t g, uint32_t base, uint32_t m, int lane)
    {
        uint32_t found = __ballot_sync(kFull, lane < nslots && gene == g);
        int idx;
        if (found) {
            idx = __ffs(found) - 1;
        } else {
            if (nslots == kSlots) {
                overflow = true;
                return;
            }
            idx = nslots++;
            if (lane == idx) gene = g;
        }
        const uint32_t w0 = base >> 5, sh = base & 31u;
        uint32_t add = 0;
        if ((uint32_t)lane == w0) add = m << sh;
        else if ((uint32_t)lane == w0 + 1 && sh) add = m >> (32u - sh);
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if (s == idx) mask[s] |= add;
    }
};

// One 32-position chunk of a read text: load, mask (FastqSplitter.hpp:104-109), validity and
// 2-bit codes (kmer_utils.hpp:29-41), pack to 64 bits, and for the lane's window
// [pos-k+1, pos] the hashed filter position.  Returns the lane's window validity.
struct ChunkState {
    uint64_t prevP;  // packed codes of the previous chunk (base i at bits 63-2i..62-2i)
    uint32_t prevV;  // validity mask of the previous chunk
};

template <bool HAS_QUAL, int MOD>
__device__ __forceinline__ bool chunk_window(const ReadKernelArgs &a, uint32_t off0, uint32_t n, int chunk, int lane,
                                             ChunkState &cs, uint32_t &len, uint32_t &pw, uint32_t &bit)
{
    const uint32_t pos = (uint32_t)chunk * 32u + (uint32_t)lane;
    uint32_t ch = 0;
    if (pos < n) {
        ch = a.seq[off0 + pos];
        if (HAS_QUAL) {
            int q = (int)(signed char)a.qual[off0 + pos];
            if (q < a.mq) ch = (ch - 64u) & 0xFFu;  // seq[i] = seq[i] - 64
        }
    }
    const bool valid = base_valid(ch);
    const uint32_t V = __ballot_sync(kFull, valid);
    len += __popc(V);  // ReadAnalyzer.hpp:46-49
    const uint32_t val = valid ? base_code(ch) << (30 - 2 * (lane & 15)) : 0u;
    const uint32_t hiw = __reduce_or_sync(kFull, lane < 16 ? val : 0u);
    const uint32_t low = __reduce_or_sync(kFull, lane >= 16 ? val : 0u);
    const uint64_t P = ((uint64_t)hiw << 32) | low;
    // window validity: bits [32+lane-k+1, 32+lane] of (prevV : V)
    const int k = a.k;
    const uint64_t VV = (uint64_t)cs.prevV | ((uint64_t)V << 32);
    const uint32_t kbits = (1u << k) - 1u;  // k <= 31
    const bool wv = (((uint32_t)(VV >> (33 + lane - k))) & kbits) == kbits;
    // forward k-mer: low 2k bits of (prevP : P) >> 2*(31-lane)
    const int s = 2 * (31 - lane);
    uint64_t fwd = P >> s;
    if (s) fwd |= cs.prevP << (64 - s);
    fwd &= (1ULL << (2 * k)) - 1ULL;
    pw = 0;
    bit = 0;
    if (wv) {
        uint64_t p = bit_index<MOD>(xxh64_u64(canonical(fwd, k)), a.geom);
        pw = (uint32_t)phys_word(p);
        bit = (uint32_t)(p & 31);
    }
    cs.prevP = P;
    cs.prevV = V;
    return wv;
}

// Adds the hits of one round (windows ending at base + lane) to the table.  H = ballot of hit
// lanes, e = the lane's entry.
__device__ __forceinline__ void accumulate_round(const ReadKernelArgs &a, WarpTable &tab, uint32_t base, uint32_t H,
                                                 bool hit, uint64_t e, int lane)
{
    uint32_t rem = H;
    while (rem && !tab.overflow) {
        const int leader = __ffs(rem) - 1;
        const uint64_t el = shfl64(e, leader);
        const uint32_t grp = __ballot_sync(kFull, hit && e == el);  // lanes with the identical list
        rem &= ~grp;
        const uint32_t ln = entry_len(el);
        if (ln > (uint32_t)kSlots) {
            tab.overflow = true;
            break;
        }
        tab.update(entry_id0(el), base, grp, lane);
        if (ln == 2) {
            tab.update(entry_lo(el), base, grp, lane);
        } else if (ln >= 3) {
            const uint32_t b = entry_lo(el);
            for (uint32_t t = 1; t < ln; ++t) tab.update(a.csr_ids[b + t], base, grp, lane);
        }
    }
}

// Writes the winners (ascending gene id) of a read with >= 2 associations into the tie pool.
// Returns the pool offset (warp-uniform); sets the overflow flag when the pool is exhausted.
__device__ __forceinline__ uint32_t pool_reserve(const ReadKernelArgs &a, uint32_t count, int lane)
{
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&a.counters->pool_used, count);
    base = __shfl_sync(kFull, base, 0);
    if ((uint64_t)base + count > a.pool_cap) {
        if (lane == 0) a.counters->pool_overflow = 1;
        return 0xFFFFFFFFu;
    }
    return base;
}

// Full path for one position p: bit vector word -> sector rank -> entry (bloomfilter.h:87-101).
// Returns false when the bit is clear.
__device__ __forceinline__ bool full_probe(const ReadKernelArgs &a, uint64_t p, uint64_t pol_first, uint64_t pol_last,
                                           uint64_t &e)
{
    const uint32_t pw = (uint32_t)phys_word(p), bit = (uint32_t)(p & 31);
    const uint32_t w = ld_filter_word(a.sectors + pw, pol_first);
    if (!((w >> bit) & 1u)) return false;
    Sector s = ld_sector(reinterpret_cast<const Sector *>(a.sectors) + (pw >> 3));
    e = ld_u64_hint(a.entries + sector_rank(s, pw & 7u, bit), pol_last);
    return true;
}

template <bool HAS_QUAL, int MOD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) analyze_reads_kernel(const ReadKernelArgs a)
{
    // per warp: packed 2-bit codes and validity masks of the read's 32-byte chunks; index c+1
    // holds chunk c, index 0 is the all-invalid chunk "-1"
    __shared__ uint64_t sP[kWarpsPerCta][kMaxFastLen / 32 + 2];
    __shared__ uint32_t sV[kWarpsPerCta][kMaxFastLen / 32 + 2];
    __shared__ uint32_t s_assoc[kWarpsPerCta];
    __shared__ uint32_t s_probes[kWarpsPerCta];
    __shared__ uint32_t s_hits[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    uint32_t warp_assoc = 0, warp_probes = 0, warp_hits = 0;
    const int k = a.k;
    const uint32_t kbits = (1u << k) - 1u;
    const uint64_t kmask2 = (1ULL << (2 * k)) - 1ULL;
    if (lane == 0) {
        sP[warp][0] = 0;
        sV[warp][0] = 0;
    }

    for (int it = 0; it < kReadsPerWarp; ++it) {
        const uint32_t r = blockIdx.x * kReadsPerTile + (uint32_t)it * kWarpsPerCta + (uint32_t)warp;
        if (r >= a.n_reads) break;
        const uint32_t off0 = a.off[r];
        const uint32_t n = a.off[r + 1] - off0;
        if (n > kMaxFastLen) {
            if (lane == 0) {
                a.rec[r] = make_uint2(0u, 0u);
                a.slow_list[atomicAdd(&a.counters->n_slow, 1u)] = r;
            }
            continue;
        }
        // ---- phase 1: text -> validity masks + packed codes (FastqSplitter.hpp:104-109,
        //      kmer_utils.hpp:29-41), number of valid bases (ReadAnalyzer.hpp:46-49)
        const int nch = (int)((n + 31u) >> 5);
        uint32_t len = 0;
        __syncwarp();
        for (int c = 0; c < nch; ++c) {
            const uint32_t pos = (uint32_t)c * 32u + (uint32_t)lane;
            uint32_t ch = 0;
            if (pos < n) {
                ch = a.seq[off0 + pos];
                if (HAS_QUAL) {
                    int q = (int)(signed char)a.qual[off0 + pos];
                    if (q < a.mq) ch = (ch - 64u) & 0xFFu;  // seq[i] = seq[i] - 64
                }
            }
            const bool valid = base_valid(ch);
            const uint32_t V = __ballot_sync(kFull, valid);
            len += __popc(V);
            const uint32_t val = valid ? base_code(ch) << (30 - 2 * (lane & 15)) : 0u;
            const uint32_t hiw = __reduce_or_sync(kFull, lane < 16 ? val : 0u);
            const uint32_t low = __reduce_or_sync(kFull, lane >= 16 ? val : 0u);
            if (lane == 0) {
                sP[warp][c + 1] = ((uint64_t)hiw << 32) | low;
                sV[warp][c + 1] = V;
            }
        }
        __syncwarp();
        // ---- phase 2: every candidate window end e = k-1 .. n-1, 32 per round, kGroup rounds of
        //      front-table loads in flight (ReadAnalyzer.hpp:50-87 without the sequential roll)
        WarpTable tab;
        tab.init();
        const int n_rounds = n >= (uint32_t)k && len >= (uint32_t)k ? (int)((n - (uint32_t)k + 32u) >> 5) : 0;
        for (int t0 = 0; t0 < n_rounds && !tab.overflow; t0 += kGroup) {
            uint4 q[kGroup];
            uint32_t bucket[kGroup], offk[kGroup];
            bool wv[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                wv[j] = false;
                bucket[j] = offk[j] = 0;
                const uint32_t e = (uint32_t)(k - 1) + 32u * (uint32_t)(t0 + j) + (uint32_t)lane;
                if (t0 + j < n_rounds && e < n) {
                    const uint32_t c = e >> 5, i = e & 31u;
                    const uint64_t VV = (uint64_t)sV[warp][c] | ((uint64_t)sV[warp][c + 1] << 32);
                    wv[j] = (((uint32_t)(VV >> (33u + i - (uint32_t)k))) & kbits) == kbits;
                    if (wv[j]) {
                        const int sh = 2 * (31 - (int)i);
                        uint64_t fwd = sP[warp][c + 1] >> sh;
                        if (sh) fwd |= sP[warp][c] << (64 - sh);
                        fwd &= kmask2;
                        const uint64_t p = bit_index<MOD>(xxh64_u64(canonical(fwd, k)), a.geom);
                        bucket[j] = (uint32_t)(p >> a.fgeom.shift);
                        offk[j] = (uint32_t)p & a.fgeom.off_mask;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                q[j] = wv[j] ? ld_front(a.front + bucket[j], pol_last) : make_uint4(kFrontEmpty, kFrontEmpty, kFrontEmpty, kFrontEmpty);
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                if (t0 + j >= n_rounds) break;
                uint32_t g = 0;
                uint64_t e = 0;
                bool hit = false;
                if (wv[j]) {
                    const int res = front_lookup(q[j], offk[j], g);
                    if (res == FRONT_SINGLE) {
                        hit = true;
                        e = make_entry(g, 1u, 0u);
                    } else if (res == FRONT_FULL) {
                        const uint64_t p = ((uint64_t)bucket[j] << a.fgeom.shift) | offk[j];
                        hit = full_probe(a, p, pol_first, pol_last, e);
                    }
                }
                warp_probes += __popc(__ballot_sync(kFull, wv[j]));
                const uint32_t H = __ballot_sync(kFull, hit);
                if (H) {
                    warp_hits += __popc(H);
                    accumulate_round(a, tab, (uint32_t)(k - 1) + 32u * (uint32_t)(t0 + j), H, hit, e, lane);
                }
            }
        }
        if (tab.overflow) {
            if (lane == 0) {
                a.rec[r] = make_uint2(0u, 0u);
                a.slow_list[atomicAdd(&a.counters->n_slow, 1u)] = r;
            }
            continue;
        }
        // per-gene coverage and hit count (ReadAnalyzer.hpp:58-60,81-83 in closed form)
        uint32_t mycov = 0, myhits = 0;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            if (s < tab.nslots) {
                const uint32_t W = tab.mask[s];
                uint32_t Wn = __shfl_down_sync(kFull, W, 1);
                if (lane == 31) Wn = 0;
                const uint64_t D = dilate_down((uint64_t)W | ((uint64_t)Wn << 32), k);
                const uint32_t cov = __reduce_add_sync(kFull, (uint32_t)__popc((uint32_t)D));
                const uint32_t hits = __reduce_add_sync(kFull, (uint32_t)__popc(W));
                if (lane == s) {
                    mycov = cov;
                    myhits = hits;
                }
            }
        }
        // argmax with ties (ReadAnalyzer.hpp:90-102), threshold and -s (ReadAnalyzer.hpp:104)
        const bool live = lane < tab.nslots;
        const uint32_t maxc = __reduce_max_sync(kFull, live ? mycov : 0u);
        const uint32_t maxh = __reduce_max_sync(kFull, (live && mycov == maxc) ? myhits : 0u);
        uint32_t win = __ballot_sync(kFull, live && mycov == maxc && myhits == maxh);
        uint32_t count = (uint32_t)__popc(win);
        const bool pass = count > 0 && (double)maxc >= __dmul_rn(a.c, (double)len) && (!a.single || count == 1);
        if (!pass) count = 0;
        uint32_t payload = 0;
        if (count == 1) {
            payload = __shfl_sync(kFull, tab.gene, __ffs(win) - 1);
        } else if (count >= 2) {
            payload = pool_reserve(a, count, lane);
            if (payload != 0xFFFFFFFFu) {
                for (uint32_t t = 0; t < count; ++t) {  // ascending gene order = std::map order
                    const bool in = (win >> lane) & 1u;
                    const uint32_t gmin = __reduce_min_sync(kFull, in ? tab.gene : 0xFFFFFFFFu);
                    if (lane == 0) a.pool[payload + t] = gmin;
                    win &= ~__ballot_sync(kFull, in && tab.gene == gmin);
                }
            }
        }
        if (lane == 0) a.rec[r] = make_uint2(count, payload);
        warp_assoc += count;
    }
    if (lane == 0) {
        s_assoc[warp] = warp_assoc;
        s_probes[warp] = warp_probes;
        s_hits[warp] = warp_hits;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ta = 0, tp = 0, th = 0;
        for (int w = 0; w < kWarpsPerCta; ++w) ta += s_assoc[w], tp += s_probes[w], th += s_hits[w];
        a.tile_sums[blockIdx.x] = ta;
        if (tp) atomicAdd(&a.counters->n_probes, (unsigned long long)tp);
        if (th) atomicAdd(&a.counters->n_hits, (unsigned long long)th);
    }
}

// ---------------------------------------------------------------------------------------------
// Exact path: any number of genes, any read length.  One warp per read; a dense per-warp table
// indexed by gene id (ids are 16-bit, small_vector.hpp:46) holds {stamp, cov, hits, last} and is
// updated window by window in read order exactly as ReadAnalyzer.hpp:56-62,79-86 does, the
// lanes sharing the (distinct) ids of one list.  Stamps make clearing unnecessary.
// ---------------------------------------------------------------------------------------------
template <bool HAS_QUAL, int MOD>
__global__ void __launch_bounds__(128) analyze_slow_kernel(const ReadKernelArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slab = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slab >= a.n_slow_slabs) return;
    const uint32_t n_slow = a.counters->n_slow;
    if (slab >= n_slow) return;
    const uint64_t pol_first = make_policy_evict_first(), pol_last = make_policy_evict_last();
    const Sector *sector_base = reinterpret_cast<const Sector *>(a.sectors);
    uint4 *table = a.slow_table + (uint64_t)slab * a.n_genes;
    uint32_t stamp = a.slow_stamp[slab];
    const uint32_t k = (uint32_t)a.k;
    unsigned long long probes = 0, hits_total = 0;

    for (uint32_t i = slab; i < n_slow; i += a.n_slow_slabs) {
        const uint32_t r = a.slow_list[i];
        const uint32_t off0 = a.off[r];
        const uint32_t n = a.off[r + 1] - off0;
        if (++stamp == 0) {  // stamp wrapped: clear the slab once
            for (uint32_t g = lane; g < a.n_genes; g += 32) table[g] = make_uint4(0, 0, 0, 0);
            stamp = 1;
            __syncwarp();
        }
        const int nch = (int)((n + 31u) >> 5);
        ChunkState cs{0ULL, 0u};
        uint32_t len = 0;
        for (int c = 0; c < nch; ++c) {
            uint32_t pw, bit;
            const bool wv = chunk_window<HAS_QUAL, MOD>(a, off0, n, c, lane, cs, len, pw, bit);
            const uint32_t w = wv ? ld_filter_word(a.sectors + pw, pol_first) : 0u;
            const bool hit = wv && ((w >> bit) & 1u);
            probes += __popc(__ballot_sync(kFull, wv));
            uint32_t H = __ballot_sync(kFull, hit);
            uint64_t e = 0;
            if (hit) {
                Sector s = ld_sector(sector_base + (pw >> 3));
                e = ld_u64_hint(a.entries + sector_rank(s, pw & 7u, bit), pol_last);
            }
            hits_total += __popc(H);
            while (H) {  // windows in read order
                const int src = __ffs(H) - 1;
                H &= H - 1;
                const uint64_t el = shfl64(e, src);
                const uint32_t epos = (uint32_t)c * 32u + (uint32_t)src;
                const uint32_t ln = entry_len(el);
                for (uint32_t t = lane; t < ln; t += 32) {
                    uint32_t g;
                    if (t == 0) g = entry_id0(el);
                    else if (ln == 2) g = entry_lo(el);
                    else g = a.csr_ids[entry_lo(el) + t];
                    uint4 ent = table[g];
                    if (ent.x != stamp) {
                        // fresh map entry: `pos - 0` >= k for every window, so cov = k
                        ent = make_uint4(stamp, k, 1u, epos);
                    } else {
                        ent.y += min(k, epos - ent.w);
                        ent.z += 1u;
                        ent.w = epos;
                    }
                    table[g] = ent;
                }
                __syncwarp();
            }
        }
        // argmax over the touched genes in ascending id order
        uint32_t maxc = 0, maxh = 0;
        for (uint32_t g = lane; g < a.n_genes; g += 32) {
            const uint4 ent = table[g];
            if (ent.x == stamp && (ent.y > maxc || (ent.y == maxc && ent.z > maxh))) {
                maxc = ent.y;
                maxh = ent.z;
            }
        }
        const uint32_t wmaxc = __reduce_max_sync(kFull, maxc);
        const uint32_t wmaxh = __reduce_max_sync(kFull, maxc == wmaxc ? maxh : 0u);
        uint32_t count = 0;
        uint32_t first_gene = 0xFFFFFFFFu;
        for (uint32_t g0 = 0; g0 < a.n_genes; g0 += 32) {
            const uint32_t g = g0 + lane;
            bool is = false;
            if (g < a.n_genes) {
                const uint4 ent = table[g];
                is = ent.x == stamp && ent.y == wmaxc && ent.z == wmaxh;
            }
            const uint32_t b = __ballot_sync(kFull, is);
            if (b && first_gene == 0xFFFFFFFFu) first_gene = g0 + (uint32_t)(__ffs(b) - 1);
            count += __popc(b);
        }
        const bool pass = count > 0 && wmaxc > 0 && (double)wmaxc >= __dmul_rn(a.c, (double)len) &&
                          (!a.single || count == 1);
        if (!pass) count = 0;
        uint32_t payload = first_gene;
        if (count >= 2) {
            payload = pool_reserve(a, count, lane);
            if (payload != 0xFFFFFFFFu) {
                uint32_t o = payload;
                for (uint32_t g0 = 0; g0 < a.n_genes; g0 += 32) {
                    const uint32_t g = g0 + lane;
                    bool is = false;
                    if (g < a.n_genes) {
                        const uint4 ent = table[g];
                        is = ent.x == stamp && ent.y == wmaxc && ent.z == wmaxh;
                    }
                    const uint32_t b = __ballot_sync(kFull, is);
                    if (is) a.pool[o + __popc(b & ((1u << lane) - 1u))] = g;
                    o += __popc(b);
                }
            }
        }
        if (lane == 0) {
            a.rec[r] = make_uint2(count, payload);
            if (count) atomicAdd(&a.tile_sums[r / kReadsPerTile], count);
        }
        __syncwarp();
    }
    if (lane == 0) {
        a.slow_stamp[slab] = stamp;
        if (probes) atomicAdd(&a.counters->n_probes, probes);
        if (hits_total) atomicAdd(&a.counters->n_hits, hits_total);
    }
}

// ---------------------------------------------------------------------------------------------
// K7: ordered output.  tile_base = exclusive scan of the per-tile counts (scan_tile_sums_kernel);
// every tile then scans its 64 reads and writes its associations in read order, genes ascending
// (the order ReadOutput prints them, ReadOutput.hpp:40-49), plus the per-read keep flag.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kReadsPerTile)
scatter_assoc_kernel(const ReadKernelArgs a, uint64_t assoc_cap, const uint32_t *total)
{
    __shared__ uint32_t warp_tot[kReadsPerTile / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kReadsPerTile + threadIdx.x;
    uint2 rc = make_uint2(0u, 0u);
    if (r < a.n_reads) rc = a.rec[r];
    const uint32_t incl = warp_incl_scan(rc.x, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t o = a.tile_base[blockIdx.x] + incl - rc.x;
    for (int w = 0; w < warp; ++w) o += warp_tot[w];
    if (r < a.n_reads) {
        a.keep[r] = rc.x ? 1 : 0;
        if ((uint64_t)o + rc.x <= assoc_cap) {
            if (rc.x == 1) {
                a.assoc[o] = shk_assoc{r, rc.y};
            } else if ((uint64_t)rc.y + rc.x <= a.pool_cap) {  // pool overflow: the host re-runs the chunk
                for (uint32_t t = 0; t < rc.x; ++t) a.assoc[o + t] = shk_assoc{r, a.pool[rc.y + t]};
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.counters->n_assoc = *total;
}

template <bool HAS_QUAL, int MOD>
static void launch_typed(const ReadKernelArgs &a, cudaStream_t st, unsigned tiles, unsigned slow_blocks, cudaEvent_t ev_ka)
{
    analyze_reads_kernel<HAS_QUAL, MOD><<<tiles, kWarpsPerCta * 32, 0, st>>>(a);
    if (ev_ka) cudaEventRecord(ev_ka, st);
    analyze_slow_kernel<HAS_QUAL, MOD><<<slow_blocks, 128, 0, st>>>(a);
}

int launch_read_kernels(shk_ctx *ctx, const ReadKernelArgs &a, uint64_t assoc_cap, cudaStream_t st, cudaEvent_t ev_k0,
                        cudaEvent_t ev_ka, cudaEvent_t ev_k1)
{
    if (ev_k0) cudaEventRecord(ev_k0, st);
    int launched = 0;
    if (a.n_reads > 0) {
        const unsigned tiles = (a.n_reads + kReadsPerTile - 1) / kReadsPerTile;
        const unsigned slow_blocks = (a.n_slow_slabs + 3) / 4;
        const bool q = a.qual != nullptr;
        switch (a.geom.mod_kind) {
        case MOD_POW2: q ? launch_typed<true, MOD_POW2>(a, st, tiles, slow_blocks, ev_ka) : launch_typed<false, MOD_POW2>(a, st, tiles, slow_blocks, ev_ka); break;
        case MOD_B33: q ? launch_typed<true, MOD_B33>(a, st, tiles, slow_blocks, ev_ka) : launch_typed<false, MOD_B33>(a, st, tiles, slow_blocks, ev_ka); break;
        default: q ? launch_typed<true, MOD_GENERIC>(a, st, tiles, slow_blocks, ev_ka) : launch_typed<false, MOD_GENERIC>(a, st, tiles, slow_blocks, ev_ka); break;
        }
        // tile_base <- exclusive scan(tile_sums); the grand total lands in tile_base[tiles]
        cudaMemcpyAsync(a.tile_base, a.tile_sums, (size_t)tiles * 4, cudaMemcpyDeviceToDevice, st);
        scan_tile_sums_kernel<<<1, 1024, 0, st>>>(a.tile_base, tiles, a.tile_base + tiles);
        scatter_assoc_kernel<<<tiles, kReadsPerTile, 0, st>>>(a, assoc_cap, a.tile_base + tiles);
        launched = 4;
        ctx->launches += 4;
    }
    else if (ev_ka) cudaEventRecord(ev_ka, st);
    if (ev_k1) cudaEventRecord(ev_k1, st);
    return launched;
}

// Re-runs only the scatter (after the host grew the association buffer).
int launch_scatter(shk_ctx *ctx, const ReadKernelArgs &a, uint64_t assoc_cap, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    const unsigned tiles = (a.n_reads + kReadsPerTile - 1) / kReadsPerTile;
    scatter_assoc_kernel<<<tiles, kReadsPerTile, 0, st>>>(a, assoc_cap, a.tile_base + tiles);
    ctx->launches += 1;
    return 1;
}

}  // namespace shk

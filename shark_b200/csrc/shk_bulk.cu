// Bulk classification of packed reads over the extension structures (K6b).  Replaces the same reference code as
// analyze_reads_kernel - ReadAnalyzer::operator() (ReadAnalyzer.hpp:39-110) behind FastqSplitter's masking
// (FastqSplitter.hpp:104-109, already folded into the validity bits of the packed form) - for indexes whose front
// table is DRAM-sized.  Citations are reference file:line.
//
// Why another kernel.  analyze_reads_kernel<EXT> walks every read base by base and hashes every window, although most
// windows of a read that follows the reference are known without a hash; in its thread-per-read layout a warp pays for
// the union of its lanes' paths, so skipping hashes per lane buys nothing (profiles/analyze_r2.md).  Here the two halves
// of the work are separated and each runs in the layout that suits it:
//
//   * per read, 32 positions at a time (one lane = one read, one round = one 32-base word of every read of the warp):
//     validity runs -> the mask WV of valid windows (build_kmer's restart rule, kmer_utils.hpp:57-71); under the read's
//     current DIAGONAL (a hypothesis "read position q is reference position base +- q") the packed read word is XORed
//     with the packed reference word -> match mask -> D = windows that are, base for base, the reference window on the
//     diagonal -> S = D[p-1] & D[p] & E' = windows whose gene-id list is that of the window before (E: shk_device.cuh).
//     S needs no proof beyond itself: whatever the diagonal is, D says the two read windows ARE the two reference
//     windows, and E says those carry the same list.  The ids of a run of S windows are applied in O(1)
//     (cov += min(k, pos - last) + L - 1, hits += L; the reference's per-window update summed over the run).
//   * all windows that are valid but not S (WV & ~S: the windows covering a mismatch, the first window of a run,
//     every window of a read that follows no reference) are LOOKUPS, and the warp shares them evenly: the lookups of a
//     round form one queue (prefix sum over the lanes' counts), every lane takes an equal slice whoever owns the
//     windows, builds the k-mers from the owners' code words in shared memory (no rolling: a window is a funnel shift),
//     and does hash -> coarse filter -> front table exactly like analyze_reads_kernel.  Hits go back to the owner
//     through shared memory; a plain hit of a read without a working diagonal is verified against the reference
//     (ref2) and becomes the read's new diagonal.
//   * the owner applies hits and runs in window order to the same 4-gene register table and ends the read with the
//     same code as analyze_reads_kernel (shk_reads.cuh).  Reads it cannot hold (more than 4 genes, a list of more than
//     2 ids, more than kMaxFastLen bases) go to the warp-per-read kernels as before.
//
// Results are those of the reference by construction: the ids of every valid window are either looked up in the exact
// front table or copied from the previous window under a condition that implies equality.
#include "shk_internal.h"
#include "shk_reads.cuh"
#include "shk_scan.cuh"

#include <cstdlib>

namespace shk {

#ifndef SHK_BULK_MIN_BLOCKS
#define SHK_BULK_MIN_BLOCKS 4
#endif
#ifndef SHK_BULK_ILP
#define SHK_BULK_ILP 4
#endif
constexpr int kBulkWarps = 4;
constexpr uint32_t kBulkComplex = 0xFFFFFFFEu;  // "ids of the previous window": a list of more than 2 ids
constexpr int kBulkThreads = kBulkWarps * 32;
static_assert(kBulkThreads == (int)kReadsPerTile, "one CTA of the bulk kernel = one scan tile");

struct BulkWarp {
    uint64_t cprev[32];            // per lane: code words of the previous and of the current 32 positions of its read
    uint64_t ccur[32];
    unsigned long long cand[32];   // per lane: best diagonal candidate of the round {pos in word:31, strand:1, e:32}
    uint32_t need[32];             // per lane: windows of the round that must be looked up
    uint32_t pre[33];              // exclusive prefix sum of popc(need)
    uint32_t hit[32];              // per lane: looked-up windows that are set in the filter (plain lists)
    uint32_t res[32][33];          // [lane][pos in word]: ids of a hit, A | B << 16 (B == A: one id)
    uint32_t cplx[32];             // per lane: hits whose list has more than 2 ids (the owner walks the bucket itself)
};

// reverse the order of the 32 2-bit groups of a word
__device__ __forceinline__ uint64_t pair_reverse(uint64_t x)
{
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}
// bits at even positions of x -> 16 contiguous bits
__device__ __forceinline__ uint32_t squeeze_even(uint32_t x)
{
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
// 32 2-bit groups -> one bit per group: group != 0
__device__ __forceinline__ uint32_t nonzero_pairs(uint64_t x)
{
    const uint64_t nz = x | (x >> 1);
    return squeeze_even((uint32_t)nz) | (squeeze_even((uint32_t)(nz >> 32)) << 16);
}
// bit i of the result: bits [32 + i - k + 1, 32 + i] of (prev : cur) are all set (k <= 32)
__device__ __forceinline__ uint32_t runs_ge(uint32_t prev, uint32_t cur, uint32_t k)
{
    uint64_t r = (uint64_t)prev | ((uint64_t)cur << 32);
    uint32_t have = 1;
    while (2u * have <= k) {  // runs of `have` -> runs of 2 * have
        r &= r << have;
        have *= 2u;
    }
    if (k > have) r &= r << (k - have);  // k - have <= have: the two runs overlap or touch
    return (uint32_t)(r >> 32);
}
// 32 positions t0 .. t0 + 31 of the derived reference arrays (t0 >= -32 * kDerivedPad)
__device__ __forceinline__ uint64_t ref_codes(const uint64_t *refr, int64_t t0, uint64_t pol)
{
    const uint64_t u = (uint64_t)(t0 + 32 * (int64_t)kDerivedPad);
    const uint32_t sh = 2u * ((uint32_t)u & 31u);
    const uint64_t lo = ld_u64_hint(refr + (u >> 5), pol), hi = ld_u64_hint(refr + (u >> 5) + 1, pol);
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}
__device__ __forceinline__ uint32_t ref_flags(const uint32_t *ebits, int64_t t0, uint64_t pol)
{
    const uint64_t u = (uint64_t)(t0 + 32 * (int64_t)kDerivedPad);
    const uint32_t lo = ld_u32_hint(ebits + (u >> 5), pol), hi = ld_u32_hint(ebits + (u >> 5) + 1, pol);
    return __funnelshift_r(lo, hi, (uint32_t)u & 31u);
}

// A read's diagonal: read position q <-> reference position base + q (forward) or base - q (reverse strand).
struct Diagonal {
    int64_t base;
    uint32_t dir;
    bool on;
};

// Word w of a read under a diagonal: M = positions whose base is valid and equals the reference base (its complement
// on the reverse strand), E = the same-list flag of the window ENDING at each position (shk_device.cuh: forward E[e];
// reverse strand: the read window ending at q is the reference window ending at e = base - q + k - 1 and the window
// before it ends at e + 1, so the flag is E[e + 1]).
__device__ __forceinline__ void match_word(const ReadKernelArgs &a, const Diagonal &dg, uint32_t w, uint64_t C, uint32_t V,
                                           uint32_t k, uint64_t pol, uint32_t &M, uint32_t &E)
{
    M = 0u;
    E = 0u;
    if (!dg.on || V == 0u) return;
    const int64_t t0 = dg.dir ? dg.base - 32 * (int64_t)w - 31 : dg.base + 32 * (int64_t)w;
    if (t0 < -32 * (int64_t)kDerivedPad || t0 > (int64_t)a.ref_total) return;  // the arrays have zero words around them
    uint64_t R = ref_codes(a.refr, t0, pol);
    if (dg.dir) {
        R = pair_reverse(R) ^ 0xAAAAAAAAAAAAAAAAULL;  // complement of A0 C1 T2 G3 = code ^ 2
        E = __brev(ref_flags(a.ebits, t0 + (int64_t)k, pol));
    } else {
        E = ref_flags(a.ebits, t0, pol);
    }
    M = V & ~nonzero_pairs(C ^ R);
}

template <int MOD>
__global__ void __launch_bounds__(kBulkThreads, SHK_BULK_MIN_BLOCKS)
analyze_bulk_kernel(const ReadKernelArgs a)
{
    __shared__ BulkWarp shared[kBulkWarps];
    const int lane = threadIdx.x & 31;
    BulkWarp &sh = shared[threadIdx.x >> 5];
    const uint32_t r = a.r0 + blockIdx.x * kBulkThreads + threadIdx.x;
    const uint32_t tile = a.r0 / kReadsPerTile + blockIdx.x;
    const uint64_t pol_first = a.pol_first, pol_last = a.pol_last;
    const uint64_t pol_front = pol_first;  // a DRAM-sized table must not displace the L2-resident arrays
    const uint32_t k = (uint32_t)a.k;
    const uint64_t kmask2 = (1ULL << (2 * k)) - 1ULL;
    const uint32_t fshift = a.fgeom.shift, fmask = a.fgeom.off_mask;

    uint32_t count = 0, payload = 0, my_probes = 0, my_hits = 0, my_ext = 0, my_loads = 0;
    bool slow = false;
    uint32_t n = 0, src0 = 0;
    if (r < a.r1) {
        const uint32_t off0 = a.off[r];
        n = a.off[r + 1] - off0;
        src0 = off0 - a.pack_base;
        if (n > kMaxFastLen) {
            slow = true;
            n = 0;
        }
    }
    const uint32_t my_words = (n + 31u) >> 5;
    const uint32_t rounds = __reduce_max_sync(kFull, my_words);

    Mru4 tab;
    tab.init();
    Diagonal dg{0, 0u, false};
    uint64_t prevC = 0;
    uint32_t prevV = 0, prevM = 0, prevDtop = 0, len = 0;
    uint32_t prevA = kFrontEmpty, prevB = kFrontEmpty, last_ev = 0xFFFFFFF0u;  // ids and end of the last window with ids

    for (uint32_t w = 0; w < rounds; ++w) {
        // ---- the owner's word: validity runs, match runs under the diagonal, what has to be looked up ----
        const bool mine = w < my_words && !tab.overflow;
        uint64_t C = 0;
        uint32_t V = 0;
        if (mine) {
            const uint32_t bp = src0 + 32u * w, gi = bp >> 5, sb = bp & 31u;
            const uint64_t c_lo = ld_u64_hint(a.pcodes + gi, pol_first), c_hi = ld_u64_hint(a.pcodes + gi + 1, pol_first);
            const uint32_t v_lo = ld_text_word(a.pvalid + gi, pol_first), v_hi = ld_text_word(a.pvalid + gi + 1, pol_first);
            C = sb ? (c_lo >> (2u * sb)) | (c_hi << (64u - 2u * sb)) : c_lo;
            V = __funnelshift_r(v_lo, v_hi, sb);
            const uint32_t rem = n - 32u * w;  // positions past the read belong to the next one
            if (rem < 32u) V &= (1u << rem) - 1u;
        }
        const uint32_t WV = runs_ge(prevV, V, k);
        len += __popc(V);        // ReadAnalyzer.hpp:46-49
        my_probes += __popc(WV);
        uint32_t M, E;
        match_word(a, dg, w, C, V, k, pol_last, M, E);
        if (!mine) M = 0u;
        const uint32_t D = runs_ge(prevM, M, k);
        const uint32_t S = D & ((D << 1) | prevDtop) & E;
        const uint32_t need = WV & ~S;

        sh.cprev[lane] = prevC;
        sh.ccur[lane] = C;
        sh.need[lane] = need;
        sh.hit[lane] = 0u;
        sh.cplx[lane] = 0u;
        sh.cand[lane] = ~0ULL;
        const uint32_t incl = warp_incl_scan((uint32_t)__popc(need), lane);
        sh.pre[lane + 1] = incl;
        if (lane == 0) sh.pre[0] = 0u;
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        // reads whose diagonal explains nothing in this round take the anchor of their first plain hit as the next one
        const uint32_t want = __ballot_sync(kFull, WV != 0u && S == 0u);
        __syncwarp();

        // ---- the warp's lookups, an equal slice per lane ----
        if (total) {
            const uint32_t per = (total + 31u) >> 5;
            const uint32_t i0 = min(total, (uint32_t)lane * per), i1 = min(total, i0 + per);
            uint32_t left = i1 - i0, o = 0, m = 0;
            if (left) {
#pragma unroll
                for (uint32_t s = 16; s; s >>= 1)  // the largest o with pre[o] <= i0: the owner of item i0
                    if (sh.pre[o + s] <= i0) o += s;
                m = sh.need[o];
                uint32_t j = i0 - sh.pre[o], at = 0, mm = m;  // drop the j lowest set bits of m
#pragma unroll
                for (uint32_t s = 16; s; s >>= 1) {
                    const uint32_t c = __popc(mm & ((1u << s) - 1u));
                    if (j >= c) {
                        j -= c;
                        mm >>= s;
                        at += s;
                    }
                }
                m &= ~0u << at;
            }
            while (left) {
                uint32_t oo[SHK_BULK_ILP], pp[SHK_BULK_ILP], bucket[SHK_BULK_ILP], key[SHK_BULK_ILP], cidx[SHK_BULK_ILP],
                    cwd[SHK_BULK_ILP];
                uint64_t fwd[SHK_BULK_ILP], rcm[SHK_BULK_ILP];
                bool ok[SHK_BULK_ILP];
#pragma unroll
                for (int u = 0; u < SHK_BULK_ILP; ++u) {
                    ok[u] = left != 0u;
                    oo[u] = 0u, pp[u] = 0u;
                    if (ok[u]) {
                        while (m == 0u) m = sh.need[++o];
                        pp[u] = (uint32_t)__ffs(m) - 1u;
                        m &= m - 1u;
                        oo[u] = o;
                        --left;
                    }
                }
#pragma unroll
                for (int u = 0; u < SHK_BULK_ILP; ++u) {
                    // the window ending at position 32 + p of (previous word : current word), LSB first
                    const uint64_t lo = sh.cprev[oo[u]], hi = sh.ccur[oo[u]];
                    const uint32_t s2 = 2u * (33u + pp[u] - k);  // >= 4
                    uint64_t X = s2 < 64u ? (lo >> s2) | (hi << (64u - s2)) : hi >> (s2 - 64u);
                    X &= kmask2;
                    X ^= (X >> 1) & 0x5555555555555555ULL;           // A0 C1 T2 G3 -> A0 C1 G2 T3 (kmer_utils.hpp:29-41)
                    rcm[u] = ~X & kmask2;                             // revcompl, kmer_utils.hpp:47-55: LSB-first complement
                    fwd[u] = pair_reverse(X) >> (64u - 2u * k);       // first base in the most significant group
                    const uint64_t pb = bit_index<MOD>(xxh64_u64(fwd[u] < rcm[u] ? fwd[u] : rcm[u]), a.geom);
                    bucket[u] = (uint32_t)(pb >> fshift);
                    key[u] = front_key((uint32_t)pb & fmask);
                    cidx[u] = (bucket[u] << a.coarse_rel) | (key[u] >> a.coarse_key_shift);
                    cwd[u] = 0u;
                    if (ok[u]) cwd[u] = ld_u32_hint(a.coarse + (cidx[u] >> 5), pol_last);
                }
                uint4 q[SHK_BULK_ILP];
#pragma unroll
                for (int u = 0; u < SHK_BULK_ILP; ++u) {
                    q[u] = make_uint4(kFrontEmpty, kFrontEmpty, kFrontEmpty, kFrontEmpty);
                    if ((cwd[u] >> (cidx[u] & 31u)) & 1u) {
                        q[u] = ld_front(a.front + (uint64_t)bucket[u] * 2u, pol_front);
                        ++my_loads;
                    }
                }
#pragma unroll
                for (int u = 0; u < SHK_BULK_ILP; ++u) {
                    uint4 qq = q[u];
                    const uint32_t kb = key[u];
                    uint32_t A, B, cur = bucket[u];
                    for (;;) {  // the position's two smallest ids (slot - key; shk_device.cuh), along the bucket's chain
                        const uint32_t d0 = qq.x - kb, d1 = qq.y - kb, d2 = qq.z - kb, d3 = qq.w - kb;
                        const uint32_t lo01 = min(d0, d1), hi01 = max(d0, d1), lo23 = min(d2, d3), hi23 = max(d2, d3);
                        A = min(lo01, lo23);
                        B = min(max(lo01, lo23), min(hi01, hi23));
                        if (A < kFrontLim || (int32_t)qq.w >= -1) break;
                        cur = qq.w & 0x7FFFFFFFu;
                        qq = ld_front(a.front + (uint64_t)cur * 2u, pol_front);
                    }
                    if (A < kFrontLim) {
                        const uint32_t ow = oo[u], p = pp[u];
                        if (A < 0x10000u && (B < 0x10000u || B >= kFrontLim)) {  // a list of one or two ids
                            sh.res[ow][p] = A | ((B < 0x10000u ? B : A) << 16);
                            atomicOr(&sh.hit[ow], 1u << p);
                            if ((want >> ow) & 1u) {
                                // the slot's anchor names a reference window with the same filter bit; it gives a
                                // diagonal only if that window IS the read's window (either strand)
                                const uint4 an = ld_front(a.front + (uint64_t)cur * 2u + 1u, pol_front);
                                const uint32_t sa = A + kb;
                                const uint32_t e = qq.x == sa ? an.x : (qq.y == sa ? an.y : (qq.z == sa ? an.z : an.w));
                                const uint64_t rk = ref2_window(a.ref2, e, kmask2, pol_last);
                                if (rk == fwd[u]) atomicMin(&sh.cand[ow], ((unsigned long long)p << 33) | e);
                                else if (rk == rcm[u]) atomicMin(&sh.cand[ow], ((unsigned long long)p << 33) | (1ULL << 32) | e);
                            }
                        } else {
                            atomicOr(&sh.cplx[ow], 1u << p);  // 3 ids or more
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- the owner applies hits and runs in window order (ReadAnalyzer.hpp:56-62, 79-86) ----
        bool rediag = false;
        if (mine) {
            const uint32_t cplx = sh.cplx[lane];
            uint32_t ev = sh.hit[lane] | cplx | S;
            while (ev && !tab.overflow) {
                const uint32_t p = (uint32_t)__ffs(ev) - 1u, pos = 32u * w + p;
                uint32_t A, B, L;
                if ((S >> p) & 1u) {  // a run of windows with the ids of the window before
                    const uint32_t rest = ~(S >> p);
                    L = rest ? (uint32_t)__ffs(rest) - 1u : 32u;
                    const bool cont = last_ev + 1u == pos;  // (a window without ids before it: none here either)
                    A = cont ? prevA : kFrontEmpty;
                    B = cont ? prevB : kFrontEmpty;
                    my_ext += L;
                    if (cont && prevA == kBulkComplex) {  // a run of lists of 3 ids or more: not held here
                        tab.overflow = true;
                        break;
                    }
                } else if ((cplx >> p) & 1u) {
                    // a list of 3 or 4 ids (or a longer one: exact path): every slot of the bucket and of its chain, in
                    // list order, as analyze_reads_kernel's generic path does
                    const uint32_t s2 = 2u * (33u + p - k);
                    uint64_t X = s2 < 64u ? (prevC >> s2) | (C << (64u - s2)) : C >> (s2 - 64u);
                    X &= kmask2;
                    X ^= (X >> 1) & 0x5555555555555555ULL;
                    const uint64_t rcx = ~X & kmask2, fwx = pair_reverse(X) >> (64u - 2u * k);
                    const uint64_t pb = bit_index<MOD>(xxh64_u64(fwx < rcx ? fwx : rcx), a.geom);
                    const uint32_t kb = front_key((uint32_t)pb & fmask);
                    uint4 qq = ld_front(a.front + (pb >> fshift) * 2u, pol_front);
                    for (;;) {
#pragma unroll 1
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t sl = i == 0 ? qq.x : (i == 1 ? qq.y : (i == 2 ? qq.z : qq.w));
                            const uint32_t d = sl ^ kb;
                            if (d < kFrontLim) {
                                if (d & kFrontLongFlag) tab.overflow = true;
                                else tab.hit(d & 0xFFFFu, pos, k);
                            }
                        }
                        if (!front_is_chain(qq.w)) break;
                        qq = ld_front(a.front + (uint64_t)(qq.w & 0x7FFFFFFFu) * 2u, pol_front);
                    }
                    my_hits += 1u;
                    prevA = kBulkComplex, prevB = kFrontEmpty, last_ev = pos;
                    ev &= ev - 1u;
                    continue;
                } else {
                    const uint32_t rr = sh.res[lane][p];
                    A = rr & 0xFFFFu;
                    B = rr >> 16;
                    if (B == A) B = kFrontEmpty;
                    L = 1u;
                }
                ev &= ~(((L < 32u ? (1u << L) : 0u) - 1u) << p);
                if (A != kFrontEmpty) {
                    my_hits += L;
                    const bool hasB = B != kFrontEmpty;
                    const bool a0 = A == tab.g0, a1 = A == tab.g1, b0 = hasB && B == tab.g0, b1 = hasB && B == tab.g1;
                    if ((a0 | a1) & (!hasB | b0 | b1)) {
                        if (a0 | b0) {
                            tab.c0 += min(k, pos - tab.l0) + (L - 1u);
                            tab.h0 += L;
                            tab.l0 = pos + L - 1u;
                        }
                        if (a1 | b1) {
                            tab.c1 += min(k, pos - tab.l1) + (L - 1u);
                            tab.h1 += L;
                            tab.l1 = pos + L - 1u;
                        }
                    } else {
                        tab.hit(A, pos, k);
                        if (!tab.overflow) tab.more(A, L - 1u);
                        if (hasB && !tab.overflow) {
                            tab.hit(B, pos, k);
                            if (!tab.overflow) tab.more(B, L - 1u);
                        }
                    }
                    prevA = A;
                    prevB = B;
                    last_ev = pos + L - 1u;
                }
            }
            const unsigned long long cd = sh.cand[lane];
            if (cd != ~0ULL) {
                const uint32_t e = (uint32_t)cd, pos = 32u * w + (uint32_t)(cd >> 33);
                dg.on = true;
                dg.dir = (uint32_t)(cd >> 32) & 1u;
                dg.base = dg.dir ? (int64_t)e - (int64_t)k + 1 + (int64_t)pos : (int64_t)e - (int64_t)pos;
                rediag = true;
            }
        }
        if (rediag) {  // the next word's windows reach into this one: its matches under the new diagonal
            uint32_t E2;
            match_word(a, dg, w, C, V, k, pol_last, M, E2);
        }
        prevC = C;
        prevV = V;
        prevM = M;
        prevDtop = rediag ? 0u : D >> 31;
    }

    if (r < a.r1) {
        if (tab.overflow) slow = true;
        if (slow) {
            count = 0, payload = 0, my_probes = 0, my_hits = 0;  // counted by the kernel that classifies the read
            a.slow_list[atomicAdd(&a.counters->n_slow, 1u)] = r;
        } else {
            finish_read(a, tab, len, count, payload);
        }
        a.rec[r] = make_uint2(count, payload);
    }
    const uint32_t wa = __reduce_add_sync(kFull, count), wp = __reduce_add_sync(kFull, my_probes),
                   wh = __reduce_add_sync(kFull, my_hits), wm = __reduce_add_sync(kFull, multi_entries(count, payload)),
                   wk = __popc(__ballot_sync(kFull, count != 0u)), we = __reduce_add_sync(kFull, my_ext),
                   wl = __reduce_add_sync(kFull, my_loads);
    if (lane == 0) {
        if (wm) atomicAdd(&a.tile_sums[tile], wm);
        if (wa) atomicAdd(&a.counters->n_assoc, (unsigned long long)wa);
        if (wk) atomicAdd(&a.counters->n_kept, (unsigned long long)wk);
        if (wp) atomicAdd(&a.counters->n_probes, (unsigned long long)wp);
        if (wh) atomicAdd(&a.counters->n_hits, (unsigned long long)wh);
        if (we) atomicAdd(&a.counters->n_extended, (unsigned long long)we);
        if (wl) atomicAdd(&a.counters->n_table_loads, (unsigned long long)wl);
    }
}

// SHK_BULK=0 keeps packed reads on analyze_reads_kernel<EXT, PACKED> (A/B measurements, tests of both kernels).
bool bulk_enabled()
{
    const char *e = std::getenv("SHK_BULK");
    return !(e && e[0] == '0');
}

void launch_bulk_kernel(const ReadKernelArgs &a, cudaStream_t st, unsigned blocks)
{
    switch (a.geom.mod_kind) {
    case MOD_POW2: analyze_bulk_kernel<MOD_POW2><<<blocks, kBulkThreads, 0, st>>>(a); break;
    case MOD_B33: analyze_bulk_kernel<MOD_B33><<<blocks, kBulkThreads, 0, st>>>(a); break;
    default: analyze_bulk_kernel<MOD_GENERIC><<<blocks, kBulkThreads, 0, st>>>(a); break;
    }
}

}  // namespace shk

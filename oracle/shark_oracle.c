/*
 * shark_oracle.c - CPU restatement of Shark's k-mer Bloom-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (shark_b200/, include/, the C-ABI
 * library, the shark-b200 CLI) links, imports or executes this file.  It is used by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the CHECKER.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the known answers of SURVEY.md App. B.3 (taken from the reference's kmer_utils.hpp
 *       compiled as-is),
 *   (2) the reference's own fixtures example/ENSG00000277117.truth.ssv and
 *       example/sharked.sample_{1,2}.truth.fq (copied as tests/golden/example/),
 *   (3) golden outputs produced by the unmodified reference binary (oracle/_ref/shark,
 *       built by oracle/Makefile from /root/reference) on generated edge-case inputs
 *       (tests/golden/make_golden.py, committed with its outputs).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * The restatement is written from the algorithm's description; data
 * structures differ from the reference on purpose (sparse sorted bit positions instead of
 * a 1 GiB sdsl bit vector; CSR instead of small_vector + select) - only semantics match.
 *
 * shko_index_build_wide is NOT a restatement of something the reference computes: it is this
 * restatement with the reference's two narrow types widened (16-bit gene ids, `int` id total),
 * the checker of the opt-in extension SHK_F_WIDE_IDS.  Inside the reference's
 * limits it equals shko_index_build (tests/test_gpu_wide.py checks that too); beyond them its
 * parity is "unpinned" by construction - the reference has no defined behaviour there.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Base codes.  kmer_utils.hpp:29-41 (to_int: A/a=1 C/c=2 G/g=3 T/t=4, else 0); the k-mer
 * uses to_int-1 (kmer_utils.hpp:68).  Bytes >= 128 index out of bounds in the reference
 * (UB); we define them as invalid.
 * ---------------------------------------------------------------------------------- */
static inline int base_code(uint8_t ch) /* -1 = invalid, else 0..3 */
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

int shko_base_code(uint8_t ch) { return base_code(ch); }

/* XXH64 of the 8 little-endian bytes of v, seed 0.
 * kmer_utils.hpp:81-83 -> xxhash.hpp:495-500 -> endian_align<64> 459-492 (len<32 branch:
 * h = seed + PRIME5 (487), h += len (489)) -> endian_align_sub_ending<64> 425-456 (one
 * 8-byte lane 427-433, avalanche 449-453).  Primes: xxhash.hpp:349. */
#define P1 11400714785074694791ULL
#define P2 14029467366897019727ULL
#define P3 1609587929392839161ULL
#define P4 9650029242287828579ULL
#define P5 2870177450012600261ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

uint64_t shko_xxh64_u64(uint64_t v)
{
    uint64_t h = P5 + 8;
    uint64_t k1 = rotl64(v * P2, 31) * P1; /* round(0, v): xxhash.hpp:367-373 */
    h ^= k1;
    h = rotl64(h, 27) * P1 + P4;
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

/* kmer_utils.hpp:47-55 */
uint64_t shko_revcompl(uint64_t kmer, int k)
{
    uint64_t rc = 0;
    kmer = ~kmer;
    for (int i = 0; i < k; ++i) {
        rc = (rc << 2) | (kmer & 3);
        kmer >>= 2;
    }
    return rc;
}

/* kmer_utils.hpp:73-79 */
uint64_t shko_lsappend(uint64_t kmer, uint64_t c, int k) { return ((kmer << 2) | c) & ((1ULL << (2 * k)) - 1); }
uint64_t shko_rsprepend(uint64_t kmer, uint64_t c, int k) { return (kmer >> 2) | (c << (2 * k - 2)); }

/* kmer_utils.hpp:57-71.  Finds the first window of k valid chars at or after *p, returns
 * the forward k-mer and leaves *p one past the window; returns -1 and *p = n if none. */
int64_t shko_build_kmer(const uint8_t *s, int64_t n, int64_t *p, int k)
{
    for (int64_t q = *p; q < n && q < *p + k; ++q)
        if (base_code(s[q]) < 0) *p = q + 1;
    if (*p + k > n) {
        *p = n;
        return -1;
    }
    uint64_t kmer = 0;
    for (int64_t end = *p + k; *p < end; ++*p) kmer = (kmer << 2) | (uint64_t)base_code(s[*p]);
    return (int64_t)kmer;
}

/* Canonical k-mer enumeration of a string - the loop shared by KmerBuilder.hpp:43-68,
 * main.cpp:163-183 and ReadAnalyzer.hpp:50-87 (rolling with lsappend/rsprepend, rebuild
 * after an invalid char).  Writes canonical k-mers and the index of each window's last
 * char.  Returns the count, or -1 when the reference would `continue` (first build_kmer
 * fails: KmerBuilder.hpp:46, main.cpp:166, ReadAnalyzer.hpp:53).  A string shorter than k
 * returns 0 (the reference never enters the block). */
int64_t shko_enumerate(const uint8_t *s, int64_t n, int k, uint64_t *canon, int64_t *endpos)
{
    if (n < k) return 0;
    int64_t pos = 0, cnt = 0;
    int64_t b = shko_build_kmer(s, n, &pos, k);
    if (b < 0) return -1;
    uint64_t kmer = (uint64_t)b, rc = shko_revcompl(kmer, k);
    if (canon) canon[cnt] = kmer < rc ? kmer : rc;
    if (endpos) endpos[cnt] = pos - 1;
    ++cnt;
    for (; pos < n; ++pos) {
        int c = base_code(s[pos]);
        if (c < 0) {
            ++pos;
            b = shko_build_kmer(s, n, &pos, k);
            if (b < 0) break;
            kmer = (uint64_t)b;
            rc = shko_revcompl(kmer, k);
            --pos;
        } else {
            kmer = shko_lsappend(kmer, (uint64_t)c, k);
            rc = shko_rsprepend(rc, (uint64_t)((~c) & 3), k); /* reverse_char: kmer_utils.hpp:43-45 */
        }
        if (canon) canon[cnt] = kmer < rc ? kmer : rc;
        if (endpos) endpos[cnt] = pos;
        ++cnt;
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------
 * Index (class BF, bloomfilter.h:36-203) in sparse form:
 *   pos[r]      sorted distinct set-bit positions   (== the 1s of _bf; rank(pos[r]) == r)
 *   off[r..r+1] list boundaries                     (== select over _bv, bloomfilter.h:142-148)
 *   ids[]       concatenated ascending gene ids     (== _index_kmer, bloomfilter.h:156-167)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int k;
    uint64_t bf_bits;
    uint64_t n_set;     /* num_kmer, bloomfilter.h:122 */
    uint64_t tot_ids;   /* tot_idx, bloomfilter.h:130 */
    uint32_t n_records; /* legend_ID.size() */
    uint32_t n_genes;   /* final nidx, main.cpp:186 */
    uint64_t *pos;
    uint32_t *off;
    uint16_t *ids;
    uint32_t *ids32;    /* the same ids untruncated: what SHK_F_WIDE_IDS indexes hold (see shko_index_build_wide) */
    int wide;
} shko_index;

typedef struct { uint64_t pos; uint32_t gene; } occ_t;
static int cmp_occ(const void *a, const void *b)
{
    const occ_t *x = a, *y = b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->gene < y->gene ? -1 : x->gene > y->gene;
}

/* Builds the index from the concatenated record sequences (exactly as parsed: case and
 * non-ACGT bytes preserved).  rec_off has n_records+1 entries.
 * Pass 1 = KmerBuilder.hpp:40-72 + BloomfilterFiller.hpp:38-46 + bloomfilter.h:57-59;
 * switch 0->1 = bloomfilter.h:112-125; pass 2 = main.cpp:154-189 + bloomfilter.h:61-75
 * (including the nidx rule: `continue` at main.cpp:166 skips `++nidx` at 186);
 * switch 1->2 = bloomfilter.h:126-184.  Gene ids are stored as uint16_t
 * (small_vector.hpp:46): callers must keep n_records <= 65536. */
static shko_index *index_build(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_records, int k, uint64_t bf_bits,
                               int wide);
shko_index *shko_index_build(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_records, int k,
                             uint64_t bf_bits)
{
    return index_build(bases, rec_off, n_records, k, bf_bits, 0);
}
/* The same algorithm with the two widenings of SHK_F_WIDE_IDS: gene ids keep 32 bits where the
 * reference's small_vector_t::push_back(uint16_t) / vector<uint16_t> (small_vector.hpp:46, bloomfilter.h:45) keep
 * 16, and nothing is counted in an `int`.  This is the reference's algorithm with those two types changed - what
 * the reference would compute if its id type were wide enough - not something the reference itself can run. */
shko_index *shko_index_build_wide(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_records, int k,
                                  uint64_t bf_bits)
{
    return index_build(bases, rec_off, n_records, k, bf_bits, 1);
}
static shko_index *index_build(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_records, int k, uint64_t bf_bits,
                               int wide)
{
    shko_index *ix = calloc(1, sizeof *ix);
    ix->wide = wide;
    ix->k = k;
    ix->bf_bits = bf_bits;
    ix->n_records = n_records;
    uint64_t total = rec_off[n_records];
    occ_t *occ = malloc((total + 1) * sizeof *occ);
    uint64_t n_occ = 0, maxlen = 0;
    for (uint32_t i = 0; i < n_records; ++i)
        if (rec_off[i + 1] - rec_off[i] > maxlen) maxlen = rec_off[i + 1] - rec_off[i];
    uint64_t *canon = malloc((maxlen + 1) * sizeof *canon);
    uint32_t nidx = 0;
    for (uint32_t i = 0; i < n_records; ++i) {
        const uint8_t *s = bases + rec_off[i];
        int64_t n = (int64_t)(rec_off[i + 1] - rec_off[i]);
        if (n >= k) {
            int64_t c = shko_enumerate(s, n, k, canon, NULL);
            if (c < 0) continue; /* main.cpp:166: no ++nidx */
            for (int64_t j = 0; j < c; ++j) {
                occ[n_occ].pos = shko_xxh64_u64(canon[j]) % bf_bits; /* bloomfilter.h:58,66 */
                occ[n_occ].gene = nidx;
                ++n_occ;
            }
        }
        ++nidx;
    }
    ix->n_genes = nidx;
    free(canon);
    qsort(occ, n_occ, sizeof *occ, cmp_occ);
    /* distinct positions */
    uint64_t n_set = 0, tot = 0;
    for (uint64_t j = 0; j < n_occ; ++j) {
        if (j == 0 || occ[j].pos != occ[j - 1].pos) ++n_set, ++tot;
        else if (occ[j].gene != occ[j - 1].gene) ++tot; /* bloomfilter.h:72 dedup on last() */
    }
    ix->n_set = n_set;
    ix->tot_ids = tot;
    ix->pos = malloc((n_set + 1) * sizeof *ix->pos);
    ix->off = malloc((n_set + 1) * sizeof *ix->off);
    ix->ids = malloc((tot + 1) * sizeof *ix->ids);
    ix->ids32 = malloc((tot + 1) * sizeof *ix->ids32);
    uint64_t r = 0, t = 0;
    for (uint64_t j = 0; j < n_occ; ++j) {
        if (j == 0 || occ[j].pos != occ[j - 1].pos) {
            ix->pos[r] = occ[j].pos;
            ix->off[r] = (uint32_t)t;
            ++r;
            ix->ids32[t] = occ[j].gene;
            ix->ids[t++] = (uint16_t)occ[j].gene;
        } else if (occ[j].gene != occ[j - 1].gene) {
            ix->ids32[t] = occ[j].gene;
            ix->ids[t++] = (uint16_t)occ[j].gene;
        }
    }
    ix->off[n_set] = (uint32_t)t;
    free(occ);
    return ix;
}

void shko_index_free(shko_index *ix)
{
    if (!ix) return;
    free(ix->pos);
    free(ix->off);
    free(ix->ids);
    free(ix->ids32);
    free(ix);
}

uint64_t shko_index_n_set(const shko_index *ix) { return ix->n_set; }
uint64_t shko_index_tot_ids(const shko_index *ix) { return ix->tot_ids; }
uint32_t shko_index_n_genes(const shko_index *ix) { return ix->n_genes; }
const uint64_t *shko_index_pos(const shko_index *ix) { return ix->pos; }
const uint32_t *shko_index_off(const shko_index *ix) { return ix->off; }
const uint16_t *shko_index_ids(const shko_index *ix) { return ix->ids; }
const uint32_t *shko_index_ids32(const shko_index *ix) { return ix->ids32; }

/* rank of a set position, or -1 (bit clear).  bloomfilter.h:87-90: `_bf[bf_idx]` then
 * `_brank(bf_idx+1)` = r+1 for the r-th (0-based) set bit. */
static int64_t find_rank(const shko_index *ix, uint64_t p)
{
    uint64_t lo = 0, hi = ix->n_set;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (ix->pos[mid] < p) lo = mid + 1;
        else hi = mid;
    }
    return (lo < ix->n_set && ix->pos[lo] == p) ? (int64_t)lo : -1;
}

/* BF::get_index, bloomfilter.h:78-102, for n canonical k-mers.  Miss -> rank -1, len 0
 * (the reference's inclusive (begin, begin-1) pair, Q10).  begin = offset into ids. */
void shko_probe(const shko_index *ix, const uint64_t *kmers, uint64_t n, int64_t *rank, uint32_t *begin,
                uint32_t *len)
{
    for (uint64_t i = 0; i < n; ++i) {
        int64_t r = find_rank(ix, shko_xxh64_u64(kmers[i]) % ix->bf_bits);
        rank[i] = r;
        begin[i] = r < 0 ? 0 : ix->off[r];
        len[i] = r < 0 ? 0 : ix->off[r + 1] - ix->off[r];
    }
}

/* ------------------------------------------------------------------------------------
 * Read text.  FastqSplitter.hpp:104-113: seq[i] -= 64 where qual[i] < min_quality+33, both
 * as (signed) char (`const char mq = min_quality + 33`, FastqSplitter.hpp:75).  Applied in
 * place on a copy.  The paired joiner ('N' / 0x1B, FastqSplitter.hpp:83-84) is supplied by
 * the caller, who passes already-joined text.
 * ---------------------------------------------------------------------------------- */
void shko_mask(uint8_t *seq, const uint8_t *qual, uint64_t n, int min_quality)
{
    if (min_quality == 0) return; /* FastqSplitter.hpp:51 */
    int8_t mq = (int8_t)(uint8_t)((uint8_t)min_quality + 33);
    for (uint64_t i = 0; i < n; ++i)
        if ((int8_t)qual[i] < mq) seq[i] = (uint8_t)(seq[i] - 64);
}

/* ------------------------------------------------------------------------------------
 * ReadAnalyzer::operator(), ReadAnalyzer.hpp:39-110, for a batch in SoA form.
 * Per read i (text = seq[off[i]..off[i+1]) after masking):
 *   out_count[i]  number of associations emitted (0 = read dropped)
 * and the associations themselves appended to (a_read, a_gene) in emission order
 * (ascending gene index per read, ReadAnalyzer.hpp:93-107).  Returns the total number of
 * associations; a_read/a_gene may be NULL to only count; cap = capacity of a_* arrays.
 * ---------------------------------------------------------------------------------- */
typedef struct { int gene; uint32_t cov, hits, last; } gcov_t;

static gcov_t *map_get(gcov_t **tab, int *n, int *cap, int gene)
{
    /* std::map<int, gene_cov_t>::operator[] : sorted by key, value-initialised on insert */
    int lo = 0, hi = *n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((*tab)[mid].gene < gene) lo = mid + 1;
        else hi = mid;
    }
    if (lo < *n && (*tab)[lo].gene == gene) return &(*tab)[lo];
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 16;
        *tab = realloc(*tab, (size_t)*cap * sizeof **tab);
    }
    memmove(&(*tab)[lo + 1], &(*tab)[lo], (size_t)(*n - lo) * sizeof **tab);
    (*tab)[lo].gene = gene;
    (*tab)[lo].cov = (*tab)[lo].hits = (*tab)[lo].last = 0;
    ++*n;
    return &(*tab)[lo];
}

uint64_t shko_analyze(const shko_index *ix, const uint8_t *seq, const uint8_t *qual, const uint64_t *off,
                      uint64_t n_reads, double c, int min_quality, int single, uint32_t *out_count,
                      uint64_t *a_read, uint32_t *a_gene, uint64_t cap)
{
    const uint32_t k = (uint32_t)ix->k;
    uint64_t n_assoc = 0, maxlen = 0;
    for (uint64_t i = 0; i < n_reads; ++i)
        if (off[i + 1] - off[i] > maxlen) maxlen = off[i + 1] - off[i];
    uint8_t *text = malloc(maxlen + 1);
    uint64_t *canon = malloc((maxlen + 1) * sizeof *canon);
    int64_t *endpos = malloc((maxlen + 1) * sizeof *endpos);
    gcov_t *tab = NULL;
    int ntab = 0, captab = 0;
    int *winners = NULL;
    int nwin = 0, capwin = 0;

    for (uint64_t i = 0; i < n_reads; ++i) {
        int64_t n = (int64_t)(off[i + 1] - off[i]);
        memcpy(text, seq + off[i], (size_t)n);
        if (qual) shko_mask(text, qual + off[i], (uint64_t)n, min_quality);
        ntab = 0;
        uint32_t len = 0; /* ReadAnalyzer.hpp:46-49 */
        for (int64_t p = 0; p < n; ++p) len += base_code(text[p]) >= 0;
        if (out_count) out_count[i] = 0;
        if (len >= k) {
            int64_t cnt = shko_enumerate(text, n, (int)k, canon, endpos);
            if (cnt < 0) continue; /* ReadAnalyzer.hpp:53 */
            for (int64_t j = 0; j < cnt; ++j) {
                int64_t r = find_rank(ix, shko_xxh64_u64(canon[j]) % ix->bf_bits);
                if (r < 0) continue;
                for (uint32_t t = ix->off[r]; t < ix->off[r + 1]; ++t) {
                    gcov_t *g = map_get(&tab, &ntab, &captab, ix->wide ? (int)ix->ids32[t] : (int)ix->ids[t]);
                    if (j == 0) {
                        /* ReadAnalyzer.hpp:57-61: pos is one past the window, unsigned math */
                        uint32_t pos = (uint32_t)(endpos[0] + 1);
                        uint32_t d = pos - g->last;
                        g->cov += k < d ? k : d;
                        g->hits = 1;
                        g->last = pos - 1;
                    } else {
                        /* ReadAnalyzer.hpp:80-85 */
                        uint32_t pos = (uint32_t)endpos[j];
                        uint32_t d = pos - g->last;
                        g->cov += k < d ? k : d;
                        g->hits += 1;
                        g->last = pos;
                    }
                }
            }
        }
        /* ReadAnalyzer.hpp:90-102: lexicographic max of (cov, hits), ties in map order */
        uint32_t max = 0, maxk = 0;
        nwin = 0;
        for (int t = 0; t < ntab; ++t) {
            if (tab[t].cov == max && tab[t].hits == maxk) {
                if (nwin == capwin) { capwin = capwin ? capwin * 2 : 16; winners = realloc(winners, (size_t)capwin * sizeof *winners); }
                winners[nwin++] = tab[t].gene;
            } else if (tab[t].cov > max || (tab[t].cov == max && tab[t].hits > maxk)) {
                nwin = 0;
                max = tab[t].cov;
                maxk = tab[t].hits;
                if (nwin == capwin) { capwin = capwin ? capwin * 2 : 16; winners = realloc(winners, (size_t)capwin * sizeof *winners); }
                winners[nwin++] = tab[t].gene;
            }
        }
        /* ReadAnalyzer.hpp:104: unsigned -> double compare against c*len in double */
        volatile double thr = c * (double)len;
        if ((double)max >= thr && (!single || nwin == 1)) {
            for (int t = 0; t < nwin; ++t) {
                if (a_read && n_assoc < cap) {
                    a_read[n_assoc] = i;
                    a_gene[n_assoc] = (uint32_t)winners[t];
                }
                ++n_assoc;
            }
            if (out_count) out_count[i] = (uint32_t)nwin;
        }
    }
    free(text);
    free(canon);
    free(endpos);
    free(tab);
    free(winners);
    return n_assoc;
}

/* Per-gene (cov, hits) table of ONE read after masking - the intermediate state of
 * ReadAnalyzer.hpp:43-87 that the reference binary cannot show.  Returns the number of
 * genes (ascending gene order), writing at most cap entries. */
int shko_read_table(const shko_index *ix, const uint8_t *text, int64_t n, int *genes, uint32_t *cov,
                    uint32_t *hits, int cap)
{
    const uint32_t k = (uint32_t)ix->k;
    uint64_t *canon = malloc((size_t)(n + 1) * sizeof *canon);
    int64_t *endpos = malloc((size_t)(n + 1) * sizeof *endpos);
    gcov_t *tab = NULL;
    int ntab = 0, captab = 0;
    uint32_t len = 0;
    for (int64_t p = 0; p < n; ++p) len += base_code(text[p]) >= 0;
    if (len >= k) {
        int64_t cnt = shko_enumerate(text, n, (int)k, canon, endpos);
        for (int64_t j = 0; j < cnt; ++j) {
            int64_t r = find_rank(ix, shko_xxh64_u64(canon[j]) % ix->bf_bits);
            if (r < 0) continue;
            for (uint32_t t = ix->off[r]; t < ix->off[r + 1]; ++t) {
                gcov_t *g = map_get(&tab, &ntab, &captab, ix->wide ? (int)ix->ids32[t] : (int)ix->ids[t]);
                uint32_t pos = (uint32_t)(j == 0 ? endpos[0] + 1 : endpos[j]);
                uint32_t d = pos - g->last;
                g->cov += k < d ? k : d;
                g->hits = j == 0 ? 1 : g->hits + 1;
                g->last = j == 0 ? pos - 1 : pos;
            }
        }
    }
    for (int t = 0; t < ntab && t < cap; ++t) {
        genes[t] = tab[t].gene;
        cov[t] = tab[t].cov;
        hits[t] = tab[t].hits;
    }
    free(canon);
    free(endpos);
    free(tab);
    return ntab;
}

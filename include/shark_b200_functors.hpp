/*
 * shark_b200_functors.hpp - the reference's own functor interface on top of the C ABI.
 *
 * AlgoLab/shark has no plugin API; its hot path is reached through `class BF` (bloomfilter.h:36-203),
 * `KmerBuilder` (KmerBuilder.hpp:34-76), `BloomfilterFiller` (BloomfilterFiller.hpp:34-52) and
 * `ReadAnalyzer` (ReadAnalyzer.hpp:32-119), wired together in main.cpp.  This header declares classes
 * with the SAME names, constructors, call operators, argument meaning and ownership rules, implemented
 * with libshark_b200.so (include/shark_b200.h), so that the reference's main.cpp compiles UNCHANGED
 * against the B200 path:
 *
 *     g++ -O3 -std=c++14 -DNDEBUG -I<repo>/include -include shark_b200_functors.hpp \
 *         <reference>/main.cpp -L<repo>/shark_b200 -lshark_b200 -lz -pthread
 *
 * `-include` puts this file first; it defines the include guards of the four headers it replaces, so
 * the reference's `#include "bloomfilter.h"` etc. become no-ops, while kseq.h, argument_parser.hpp,
 * FastaSplitter.hpp, FastqSplitter.hpp, ReadOutput.hpp and kmer_utils.hpp (pass 2 of main.cpp
 * enumerates k-mers on the host) stay the reference's own.  No sdsl-lite is needed any more.
 *
 * Error behaviour: the reference signals nothing (no exceptions, no codes; failures are UB).  Here a
 * failing library call prints shk_last_error() to stderr and exits with EXIT_FAILURE - there is no
 * CPU fallback to continue with.
 */
#ifndef SHARK_B200_FUNCTORS_HPP
#define SHARK_B200_FUNCTORS_HPP

/* the reference headers replaced by this file */
#define _BLOOM_FILTER_HPP
#define KMER_BUILDER_HPP
#define BF_FILLER_HPP
#define READANALYZER_HPP

#include <algorithm>
#include <array>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "shark_b200.h"

using namespace std;  // the replaced headers do this at global scope and main.cpp relies on it

/* common.hpp:30-36 (interface types of the sample stage).  Same guard, so that the reference's
 * own common.hpp is used when it comes first and skipped when this file comes first. */
#ifndef SHARK_COMMON_HPP
#define SHARK_COMMON_HPP
struct sharseq_t {
  string id, seq, qual;
};
typedef std::pair<string, std::pair<sharseq_t, sharseq_t>> elem_t;
typedef std::pair<string, std::pair<sharseq_t, sharseq_t>> assoc_t;
#endif

namespace shark_b200 {

inline void die(const shk_ctx *ctx, const char *what) {
  fprintf(stderr, "[shark-b200] %s failed: %s\n", what, shk_last_error(ctx));
  exit(EXIT_FAILURE);
}
#define SHK_B200_CHECK(ctx, call)                     \
  do {                                                \
    if ((call) != SHK_OK) ::shark_b200::die(ctx, #call); \
  } while (0)

inline int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

}  // namespace shark_b200

/* ---------------------------------------------------------------------------------------------
 * class BF (bloomfilter.h:36-203): the 3-mode index, device-resident.
 *   mode 0  add_at(p)                 set bit p % size
 *   mode 1  add_to_kmer(kmers, idx)   attach gene idx to the lists of the k-mers' bits
 *   mode 2  get_index(kmer)           inclusive iterator pair into the id array (Q10: a miss is
 *                                     (begin, begin - 1))
 * ------------------------------------------------------------------------------------------- */
class BF {
public:
  typedef uint64_t kmer_t;
  typedef uint64_t hash_t;
  typedef vector<uint16_t> index_kmer_t;

  /* `BF bloom(opt::bf_size)` (main.cpp:108).  k, c, -s reach the device later, with KmerBuilder's
   * and ReadAnalyzer's constructors, exactly as in the reference. */
  BF(const size_t size) : _size(size) {
    shk_params p;
    memset(&p, 0, sizeof p);
    p.k = 31;  // placeholder until a functor that knows k is constructed (shk_set_options)
    p.c = 0.0;
    p.bf_bits = size;
    p.device = shark_b200::env_int("SHARK_B200_DEVICE", 0);
    p.n_slots = (uint32_t)max(1, shark_b200::env_int("SHARK_B200_SLOTS", 4));
    p.max_reads_per_chunk = 1u << 16;  // FastqSplitter hands out 50 000 reads per call (main.cpp:215)
    p.max_bytes_per_chunk = (uint64_t)p.max_reads_per_chunk * 640;
    _n_slots = p.n_slots;
    _max_reads = p.max_reads_per_chunk;
    _max_bytes = p.max_bytes_per_chunk;
    if (shk_create(&p, &_ctx) != SHK_OK) shark_b200::die(nullptr, "shk_create");
    _pending.reserve(kFlush);
    for (uint32_t s = 0; s < _n_slots; ++s) _free_slots.push_back(s);
  }

  ~BF() { shk_destroy(_ctx); }

  /* bloomfilter.h:57-59.  Single positions are batched; BloomfilterFiller passes whole vectors. */
  void add_at(const uint64_t p) {
    _pending.push_back(p);
    if (_pending.size() >= kFlush) flush();
  }
  void add_at(const uint64_t *positions, size_t n) {
    flush();
    SHK_B200_CHECK(_ctx, shk_bf_add_at(_ctx, positions, n));
  }

  /* bloomfilter.h:61-75: hashes the k-mers (the vector is overwritten with `hash % size`, sorted, as
   * the reference leaves it - callers clear it anyway), ranks them, appends input_idx. */
  void add_to_kmer(vector<uint64_t> &kmers, const int input_idx) {
    if (shk_bf_mode(_ctx) != 1) return;
    SHK_B200_CHECK(_ctx, shk_bf_add_to_kmer(_ctx, kmers.data(), kmers.size(), input_idx));
  }

  /* bloomfilter.h:78-102.  Host-side view for callers other than ReadAnalyzer (which runs on the
   * device): one shk_probe per call into a lazily downloaded copy of `_index_kmer`. */
  pair<index_kmer_t::const_iterator, index_kmer_t::const_iterator> get_index(const kmer_t &kmer) const {
    int start_pos = 0;
    int end_pos = -1;
    {
      lock_guard<mutex> lock(_mtx);
      if (!_exported) {
        shk_index_info info;
        SHK_B200_CHECK(_ctx, shk_index_info_get(_ctx, &info));
        _index_kmer.resize(info.tot_ids);
        SHK_B200_CHECK(_ctx, shk_index_export(_ctx, nullptr, nullptr, _index_kmer.data()));
        _exported = true;
      }
      int64_t rank = -1;
      uint32_t begin = 0, len = 0;
      SHK_B200_CHECK(_ctx, shk_probe(_ctx, &kmer, 1, &rank, &begin, &len));
      if (rank >= 0) {
        start_pos = (int)begin;
        end_pos = (int)begin + (int)len - 1;
      }
    }
    return make_pair(_index_kmer.cbegin() + start_pos, _index_kmer.cbegin() + end_pos);
  }

  /* bloomfilter.h:112-184: 0 -> 1 and 1 -> 2 only; anything else returns false. */
  bool switch_mode(const int new_mode) {
    const int mode = shk_bf_mode(_ctx);
    if ((mode == 0 && new_mode == 1) || (mode == 1 && new_mode == 2)) {
      flush();
      SHK_B200_CHECK(_ctx, shk_bf_switch_mode(_ctx, new_mode, nullptr));
      return true;
    }
    return false;
  }

  /* ---- used by the functors of this header (not part of the reference's interface) ---- */
  shk_ctx *ctx() const { return _ctx; }
  size_t size() const { return _size; }
  uint32_t n_slots() const { return _n_slots; }
  uint32_t max_reads() const { return _max_reads; }
  uint64_t max_bytes() const { return _max_bytes; }
  uint32_t acquire_slot() const {
    unique_lock<mutex> lock(_slot_mtx);
    _slot_cv.wait(lock, [&] { return !_free_slots.empty(); });
    const uint32_t s = _free_slots.back();
    _free_slots.pop_back();
    return s;
  }
  void release_slot(uint32_t s) const {
    {
      lock_guard<mutex> lock(_slot_mtx);
      _free_slots.push_back(s);
    }
    _slot_cv.notify_one();
  }

private:
  BF() = delete;
  const BF &operator=(const BF &) = delete;
  const BF &operator=(const BF &&) = delete;

  void flush() {
    if (_pending.empty()) return;
    SHK_B200_CHECK(_ctx, shk_bf_add_at(_ctx, _pending.data(), _pending.size()));
    _pending.clear();
  }

  static constexpr size_t kFlush = 1u << 20;
  const size_t _size;
  shk_ctx *_ctx = nullptr;
  uint32_t _n_slots = 0, _max_reads = 0;
  uint64_t _max_bytes = 0;
  vector<uint64_t> _pending;
  mutable mutex _mtx;
  mutable bool _exported = false;
  mutable index_kmer_t _index_kmer;
  mutable mutex _slot_mtx;
  mutable condition_variable _slot_cv;
  mutable vector<uint32_t> _free_slots;
};

/* ---------------------------------------------------------------------------------------------
 * KmerBuilder (KmerBuilder.hpp:34-76): hashes of all canonical k-mers of a batch of records.
 * Takes ownership of `texts` (deletes it) and returns a heap vector the caller deletes, like the
 * reference.  Re-entrant (main.cpp:136-140 calls it from -t threads): calls are serialised on the
 * builder's device context.
 * ------------------------------------------------------------------------------------------- */
class KmerBuilder {
public:
  KmerBuilder(size_t _k) : k(_k) {
    shk_params p;
    memset(&p, 0, sizeof p);
    p.k = (uint32_t)_k;
    p.c = 0.0;
    p.bf_bits = 64;  // this context only hashes; it owns no filter worth the name
    p.device = shark_b200::env_int("SHARK_B200_DEVICE", 0);
    p.n_slots = 1;
    p.max_reads_per_chunk = 1;
    p.max_bytes_per_chunk = 64;
    if (shk_create(&p, &ctx) != SHK_OK) shark_b200::die(nullptr, "shk_create");
  }
  ~KmerBuilder() { shk_destroy(ctx); }
  KmerBuilder(const KmerBuilder &) = delete;
  KmerBuilder &operator=(const KmerBuilder &) = delete;

  vector<uint64_t> *operator()(vector<pair<string, string>> *texts) const {
    vector<uint64_t> *kmer_pos = new vector<uint64_t>();
    vector<uint8_t> bases;
    vector<uint64_t> rec_off(1, 0);
    size_t total = 0;
    for (const auto &p : *texts) total += p.second.size();
    bases.reserve(total);
    for (const auto &p : *texts) {
      bases.insert(bases.end(), p.second.begin(), p.second.end());
      rec_off.push_back(bases.size());
    }
    kmer_pos->resize(total);
    uint64_t n = 0;
    {
      lock_guard<mutex> lock(mtx);
      SHK_B200_CHECK(ctx, shk_kmer_hashes(ctx, bases.data(), rec_off.data(), (uint32_t)texts->size(), kmer_pos->data(),
                                         total, &n));
    }
    kmer_pos->resize(n);
    delete texts;
    return kmer_pos;
  }

private:
  const size_t k;
  shk_ctx *ctx = nullptr;
  mutable mutex mtx;
};

/* ---------------------------------------------------------------------------------------------
 * BloomfilterFiller (BloomfilterFiller.hpp:34-52): `bf->add_at(p)` for every hash, serialised
 * under one mutex; takes ownership of `positions`.
 * ------------------------------------------------------------------------------------------- */
class BloomfilterFiller {
public:
  BloomfilterFiller(BF *_bf) : bf(_bf) {}

  void operator()(vector<uint64_t> *positions) {
    {
      std::lock_guard<std::mutex> lock(mtx);
      bf->add_at(positions->data(), positions->size());
    }
    delete positions;
  }

private:
  BF *const bf;
  std::mutex mtx;
};

/* ---------------------------------------------------------------------------------------------
 * ReadAnalyzer (ReadAnalyzer.hpp:32-119).  `operator()` is const and is called concurrently by the
 * -t worker threads (main.cpp:72,219-223); every call takes one chunk slot of the BF's context,
 * packs the batch's texts (p.first: mate1 [+ 'N' + mate2], already masked by FastqSplitter) into
 * the slot's pinned staging buffers, runs the device classification and appends
 * `{ legend_ID[gene], p.second }` in read order, genes ascending - the order the reference's
 * std::map iteration produces (ReadAnalyzer.hpp:90-108).
 * ------------------------------------------------------------------------------------------- */
class ReadAnalyzer {
public:
  typedef vector<assoc_t> output_t;

  ReadAnalyzer(BF *_bf, const vector<string> &_legend_ID, unsigned int _k, double _c, bool _only_single = false)
      : bf(_bf), legend_ID(_legend_ID), k(_k), c(_c), only_single(_only_single) {
    // masking already happened on the host (FastqSplitter.hpp:70,104-109) -> min_quality 0 here
    SHK_B200_CHECK(bf->ctx(), shk_set_options(bf->ctx(), k, c, 0, only_single ? 1 : 0));
    staging.resize(bf->n_slots());
  }
  ~ReadAnalyzer() {
    for (auto &st : staging) {
      if (st.seq) shk_free_pinned(st.seq);
      if (st.off) shk_free_pinned(st.off);
    }
  }

  void operator()(const vector<elem_t> &reads, output_t &associations) const {
    size_t i = 0;
    while (i < reads.size()) {
      // one chunk: as many reads as fit the slot (a 50 000-read batch always does)
      size_t j = i;
      uint64_t bytes = 0;
      while (j < reads.size() && j - i < bf->max_reads() && bytes + reads[j].first.size() <= bf->max_bytes()) {
        bytes += reads[j].first.size();
        ++j;
      }
      if (j == i) {
        fprintf(stderr, "[shark-b200] a read of %zu bytes exceeds the chunk capacity of %llu bytes\n",
                reads[i].first.size(), (unsigned long long)bf->max_bytes());
        exit(EXIT_FAILURE);
      }
      const uint32_t slot = bf->acquire_slot();
      const Staging st = staging_of(slot);
      uint64_t o = 0;
      st.off[0] = 0;
      for (size_t r = i; r < j; ++r) {
        const string &t = reads[r].first;
        memcpy(st.seq + o, t.data(), t.size());
        o += t.size();
        st.off[r - i + 1] = (uint32_t)o;
      }
      shk_chunk_result res;
      SHK_B200_CHECK(bf->ctx(), shk_reads_submit(bf->ctx(), slot, st.seq, nullptr, st.off, (uint32_t)(j - i)));
      SHK_B200_CHECK(bf->ctx(), shk_reads_collect(bf->ctx(), slot, &res));
      for (uint64_t a = 0; a < res.n_assoc; ++a) {
        const elem_t &p = reads[i + res.assoc[a].read_idx];
        associations.push_back({legend_ID[res.assoc[a].gene_idx], get<1>(p)});  // ReadAnalyzer.hpp:106
      }
      bf->release_slot(slot);
      i = j;
    }
  }

private:
  struct Staging {
    uint8_t *seq = nullptr;
    uint32_t *off = nullptr;
  };
  Staging staging_of(uint32_t slot) const {  // pinned buffers of a slot, allocated on first use
    lock_guard<mutex> lock(mtx);
    Staging &st = staging[slot];
    if (!st.seq) {
      void *a = nullptr, *b = nullptr;
      if (shk_alloc_pinned(&a, bf->max_bytes()) != SHK_OK || shk_alloc_pinned(&b, ((size_t)bf->max_reads() + 1) * 4) != SHK_OK)
        shark_b200::die(nullptr, "shk_alloc_pinned");
      st.seq = (uint8_t *)a;
      st.off = (uint32_t *)b;
    }
    return st;
  }

  BF *const bf;
  const vector<string> &legend_ID;
  const unsigned int k;
  const double c;
  const bool only_single;
  mutable mutex mtx;
  mutable vector<Staging> staging;
};

#endif /* SHARK_B200_FUNCTORS_HPP */

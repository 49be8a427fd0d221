// Host-side packing of read text for the H2D link (shk_reads_submit with SHK_F_HOST_PACK).
//
// The end-to-end rate of the read path is bound by PCIe: the reference-facing call takes the reads
// as the reference holds them - ASCII text, one byte per base, plus one quality byte per base under
// -q (FastqSplitter.hpp:47-93) - and the kernels are 3-5 x faster than the link.  Everything the
// kernels need from a text byte is 3 bits: is it a valid base AFTER the quality masking
// (`seq[i] -= 64` where `qual[i] < q+33`, FastqSplitter.hpp:104-109; to_int, kmer_utils.hpp:29-41),
// and which of the four.  These functions compute exactly those bits on the host cores (AVX-512 or AVX2, a
// scalar fallback elsewhere), so that 0.375 bytes per base cross the link instead of 1 (2 under -q);
// a small kernel expands them back to text in HBM and the classification kernels run unchanged.
#pragma once
#include <cstdint>

namespace shk {

// Per text byte i: ch = seq[i], minus 64 (mod 256) when qual != nullptr and (signed char)qual[i] < mq;
// valid bit i = ch is one of ACGTacgt; code i = (ch >> 1) & 3 (A 0, C 1, T 2, G 3) or 0 when invalid.
// codes[i / 32] holds 32 codes (base i % 32 at bits 2 (i % 32)), valid[i / 32] 32 bits; both arrays have
// ceil(n / 32) words and the bits past n are zero.  Single thread, bytes [0, n).
void host_pack(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t *codes, uint32_t *valid);

// The same over all pack threads of the process (a persistent pool, created on first use;
// SHK_PACK_THREADS overrides its size = min(hardware threads, 32)).  Calls are serialised.
void host_pack_parallel(const uint8_t *seq, const uint8_t *qual, int mq, uint64_t n, uint64_t *codes, uint32_t *valid);
int host_pack_threads();
const char *host_pack_isa();  // "avx512", "avx2" or "scalar"

}  // namespace shk

// RFC1951 bit emission shared by the sm_100a emission kernels (finish.cuh) and by host code
// (test mock, fixed-block path): what the reference produces in
//   ZopfliLengthsToSymbols   /root/reference/src/zopfli/tree.c:30-69      -> canonical_codes()
//   EncodeTree / AddDynamicTree               deflate.c:105-272            -> tree_tokens(), write_tree_header()
//   AddLZ77Data                               deflate.c:297-333            -> symbol_bits()
//   AddNonCompressedBlock                     deflate.c:625-663            -> stored_layout()
// restated for bulk emission: every quantity is a pure function of (code lengths, symbol), tokens of
// the code-length RLE come from closed forms per run instead of counting loops, and all writers go
// through a BitSink so the same code runs on one device thread, a device CTA or the host.
//
// Bit order (deflate.c:45-72): values are appended LSB first; Huffman codes are appended MSB first,
// i.e. bit-reversed -- canonical_codes() returns them already reversed.
#pragma once
#include <stdint.h>

#include "deflate_size.hpp"

namespace zb {

// Per-block emission plan, produced on the device by k_block_plan (finish.cuh) or by the host
// estimators; lives in device memory on the product path (the host only sees the three costs).
struct BlockPlan {
  uint8_t ll_len[kNumLL];   // dynamic code lengths chosen by GetDynamicLengths (deflate.c:569-582)
  uint8_t d_len[kNumD];
  uint32_t tree_flags;      // use_16 | use_17 << 1 | use_18 << 2: first minimum of deflate.c:251-272
  uint32_t tree_bits;       // size of the encoded tree with those flags
  uint64_t dyn_bits;        // ZopfliCalculateBlockSize btype 2 (3 header bits included)
  uint64_t fixed_bits;      // btype 1
  uint64_t unc_bits;        // btype 0
  uint64_t nbytes;          // input bytes the symbols cover (ZopfliLZ77GetByteRange lz77.c:160-166)
};

ZB_HD uint32_t reverse_code(uint32_t v, int n) {
#if defined(__CUDA_ARCH__)
  return n ? __brev(v) >> (32 - n) : 0u;
#else
  uint32_t r = 0;
  for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
#endif
}

// Canonical prefix codes (RFC1951 3.2.2 / tree.c:30-69) for n symbols with lengths <= 15; codes come
// back bit-reversed (ready for an LSB-first sink).  Symbols with length 0 get code 0.
template <typename LenT, typename CodeT>
ZB_HD void canonical_codes(const LenT* len, int n, CodeT* code) {
  uint32_t cnt[16];
  for (int l = 0; l < 16; l++) cnt[l] = 0;
  for (int i = 0; i < n; i++) cnt[len[i]]++;
  cnt[0] = 0;
  uint32_t next[16];
  uint32_t c = 0;
  for (int l = 1; l <= 15; l++) { c = (c + cnt[l - 1]) << 1; next[l] = c; }
  for (int i = 0; i < n; i++) {
    const int l = len[i];
    code[i] = l ? (CodeT)reverse_code(next[l]++, l) : (CodeT)0;
  }
}

// Writes HLIT/HDIST/HCLEN, the code-length code and the RLE tokens; returns the number of bits.
// With a counting sink this is CalculateTreeSize's per-flag evaluation (deflate.c:277-290).
template <typename LenT, typename Sink>
ZB_HD uint32_t write_tree_header(const LenT* ll_len, const LenT* d_len, unsigned flags, Sink& sink) {
  const TreeShape sh = tree_shape(ll_len, d_len);
  uint32_t clcount[19];
  for (int k = 0; k < 19; k++) clcount[k] = 0;
  tree_tokens(ll_len, d_len, sh, flags, [&](unsigned sym, unsigned) { clcount[sym]++; });
  uint8_t clcl[19];
  uint16_t clcode[19];
  PmScratch<19, 7> pm;
  length_limited<19, 7>(clcount, 19, 7, clcl, pm);
  canonical_codes(clcl, 19, clcode);
  unsigned hclen = 15;   // trailing unused code-length codes (in transmission order) are dropped
  while (hclen > 0 && clcount[clcl_rank_symbol(hclen + 3)] == 0) hclen--;
  uint32_t bits = 14 + (hclen + 4) * 3;
  sink.put(sh.hlit, 5);
  sink.put(sh.hdist, 5);
  sink.put(hclen, 4);
  for (unsigned k = 0; k < hclen + 4; k++) sink.put(clcl[clcl_rank_symbol(k)], 3);
  tree_tokens(ll_len, d_len, sh, flags, [&](unsigned sym, unsigned extra) {
    const int xb = sym == 16 ? 2 : (sym == 17 ? 3 : (sym == 18 ? 7 : 0));
    sink.put((uint32_t)clcode[sym] | (extra << clcl[sym]), clcl[sym] + xb);
    bits += clcl[sym] + xb;
  });
  return bits;
}

struct CountSink {  // size-only evaluation
  ZB_HD void put(uint32_t, int) {}
};

// One LZ77 symbol as (litlen part, dist part): each part is code + extra bits, at most 20 / 28 bits
// (deflate.c:297-333).  dist == 0 => literal, the dist part is empty.
struct SymBits { uint32_t v0, v1; uint8_t n0, n1; };

template <typename LenT, typename CodeT>
ZB_HD SymBits symbol_bits_of(unsigned litlen, unsigned dist, const LenT* ll_len, const CodeT* ll_code,
                             const LenT* d_len, const CodeT* d_code) {
  SymBits r;
  if (dist == 0) {
    r.v0 = ll_code[litlen];
    r.n0 = ll_len[litlen];
    r.v1 = 0;
    r.n1 = 0;
    return r;
  }
  const int ls = length_symbol((int)litlen), ds = dist_symbol((int)dist);
  const int lx = length_extra_bits((int)litlen), dx = dist_extra_bits((int)dist);
  r.v0 = (uint32_t)ll_code[ls] | ((uint32_t)length_extra_bits_value((int)litlen) << ll_len[ls]);
  r.n0 = (uint8_t)(ll_len[ls] + lx);
  r.v1 = (uint32_t)d_code[ds] | ((uint32_t)dist_extra_bits_value((int)dist) << d_len[ds]);
  r.n1 = (uint8_t)(d_len[ds] + dx);
  return r;
}

// Stored blocks (deflate.c:625-663): [3 header bits][pad to byte][LEN][NLEN][<= 65535 bytes], repeated.
// Bit position after writing `nbytes` input bytes as stored blocks starting at bit position `pos`.
ZB_HD uint64_t stored_end(uint64_t pos, uint64_t nbytes) {
  uint64_t left = nbytes;
  do {
    const uint64_t bs = left > 65535 ? 65535 : left;
    pos = ((pos + 3 + 7) & ~(uint64_t)7) + 32 + bs * 8;
    left -= bs;
  } while (left);
  return pos;
}

}  // namespace zb

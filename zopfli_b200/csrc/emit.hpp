// RFC1951 bit emission for one deflate block: what AddLZ77Block / AddDynamicTree / EncodeTree /
// AddLZ77Data / ZopfliLengthsToSymbols produce (/root/reference/src/zopfli/deflate.c:38-72,
// 105-272, 297-333, 682-745; tree.c:30-69), restated over a 64-bit accumulator instead of the
// reference's bit-at-a-time appends.  Every compressed block is emitted at bit offset 0 into its
// own BitString so blocks can be produced concurrently and spliced by a bit-offset scan
// (SURVEY 8(e)); stored blocks depend on the true bit offset (deflate.c:643-649) and are
// written at splice time.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "lz77_store.hpp"

namespace zb {

struct BitString {
  std::vector<uint8_t> bytes;
  uint64_t acc = 0;
  int nacc = 0;       // bits pending in acc
  uint64_t nbits = 0;  // total bits written

  void add_bits(uint32_t value, int n) {  // LSB first (deflate.c:45-56)
    acc |= (uint64_t)value << nacc;
    nacc += n;
    nbits += (uint64_t)n;
    while (nacc >= 8) { bytes.push_back((uint8_t)acc); acc >>= 8; nacc -= 8; }
  }
  void flush() {
    if (nacc > 0) { bytes.push_back((uint8_t)acc); acc = 0; nacc = 0; }
  }
};

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
  return r;
}

// tree.c:30-69, returning codes already bit-reversed so they can go through add_bits
// (AddHuffmanBits deflate.c:62-72 writes the code MSB first).
template <typename LenT>
inline void lengths_to_reversed_codes(const LenT* lengths, int n, int maxbits, uint32_t* codes) {
  uint32_t bl_count[16] = {0}, next_code[16] = {0};
  for (int i = 0; i < n; i++) bl_count[lengths[i]]++;
  bl_count[0] = 0;
  uint32_t code = 0;
  for (int bits = 1; bits <= maxbits; bits++) {
    code = (code + bl_count[bits - 1]) << 1;
    next_code[bits] = code;
  }
  for (int i = 0; i < n; i++) {
    int len = lengths[i];
    codes[i] = len ? reverse_bits(next_code[len]++, len) : 0;
  }
}

// EncodeTree with output (deflate.c:105-249)
inline void emit_tree(const uint8_t* ll_lengths, const uint8_t* d_lengths, bool use_16, bool use_17,
                      bool use_18, BitString& out) {
  static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  unsigned hlit = 29, hdist = 29;
  while (hlit > 0 && ll_lengths[257 + hlit - 1] == 0) hlit--;
  while (hdist > 0 && d_lengths[1 + hdist - 1] == 0) hdist--;
  const unsigned hlit2 = hlit + 257, total = hlit2 + hdist + 1;
  uint32_t clcounts[19] = {0};
  std::vector<uint8_t> rle, rle_bits;
  for (unsigned i = 0; i < total; i++) {
    unsigned symbol = i < hlit2 ? ll_lengths[i] : d_lengths[i - hlit2];
    unsigned count = 1;
    if (use_16 || (symbol == 0 && (use_17 || use_18))) {
      for (unsigned j = i + 1; j < total && symbol == (unsigned)(j < hlit2 ? ll_lengths[j] : d_lengths[j - hlit2]); j++)
        count++;
    }
    i += count - 1;
    if (symbol == 0 && count >= 3) {
      if (use_18) while (count >= 11) {
        unsigned c2 = count > 138 ? 138 : count;
        rle.push_back(18); rle_bits.push_back((uint8_t)(c2 - 11)); clcounts[18]++; count -= c2;
      }
      if (use_17) while (count >= 3) {
        unsigned c2 = count > 10 ? 10 : count;
        rle.push_back(17); rle_bits.push_back((uint8_t)(c2 - 3)); clcounts[17]++; count -= c2;
      }
    }
    if (use_16 && count >= 4) {
      count--;
      clcounts[symbol]++;
      rle.push_back((uint8_t)symbol); rle_bits.push_back(0);
      while (count >= 3) {
        unsigned c2 = count > 6 ? 6 : count;
        rle.push_back(16); rle_bits.push_back((uint8_t)(c2 - 3)); clcounts[16]++; count -= c2;
      }
    }
    clcounts[symbol] += count;
    while (count > 0) { rle.push_back((uint8_t)symbol); rle_bits.push_back(0); count--; }
  }
  uint8_t clcl[19];
  uint32_t clcodes[19];
  PmScratch<19, 7> s;
  length_limited<19, 7>(clcounts, 19, 7, clcl, s);
  lengths_to_reversed_codes(clcl, 19, 7, clcodes);
  unsigned hclen = 15;
  while (hclen > 0 && clcounts[order[hclen + 4 - 1]] == 0) hclen--;
  out.add_bits(hlit, 5);
  out.add_bits(hdist, 5);
  out.add_bits(hclen, 4);
  for (unsigned i = 0; i < hclen + 4; i++) out.add_bits(clcl[order[i]], 3);
  for (size_t i = 0; i < rle.size(); i++) {
    out.add_bits(clcodes[rle[i]], clcl[rle[i]]);
    if (rle[i] == 16) out.add_bits(rle_bits[i], 2);
    else if (rle[i] == 17) out.add_bits(rle_bits[i], 3);
    else if (rle[i] == 18) out.add_bits(rle_bits[i], 7);
  }
}

// AddLZ77Block for btype 1 / 2 (deflate.c:682-745) at bit offset 0.
inline void emit_compressed_block(int btype, bool final, const Lz77Store& st, size_t lstart,
                                  size_t lend, BitString& out) {
  uint8_t ll_lengths[kNumLL], d_lengths[kNumD];
  uint32_t ll_codes[kNumLL], d_codes[kNumD];
  out.add_bits(final ? 1 : 0, 1);
  out.add_bits(btype & 1, 1);
  out.add_bits((btype & 2) >> 1, 1);
  if (btype == 1) {
    for (int i = 0; i < kNumLL; i++) ll_lengths[i] = (uint8_t)fixed_ll_length(i);
    for (int i = 0; i < kNumD; i++) d_lengths[i] = 5;
  } else {
    uint32_t h[320];
    DynScratch s;
    st.range_hist(lstart, lend, h);
    dynamic_block_bits(h, ll_lengths, d_lengths, s);
    int flags = 0;
    best_tree_size(ll_lengths, d_lengths, &flags);  // AddDynamicTree deflate.c:251-272
    emit_tree(ll_lengths, d_lengths, (flags & 1) != 0, (flags & 2) != 0, (flags & 4) != 0, out);
  }
  lengths_to_reversed_codes(ll_lengths, kNumLL, 15, ll_codes);
  lengths_to_reversed_codes(d_lengths, kNumD, 15, d_codes);
  for (size_t i = lstart; i < lend; i++) {  // AddLZ77Data deflate.c:297-333
    unsigned dist = st.dists[i], litlen = st.litlens[i];
    if (dist == 0) {
      out.add_bits(ll_codes[litlen], ll_lengths[litlen]);
    } else {
      unsigned lls = st.llsym[i], ds = st.dsym[i];
      out.add_bits(ll_codes[lls], ll_lengths[lls]);
      out.add_bits((uint32_t)length_extra_bits_value((int)litlen), length_extra_bits((int)litlen));
      out.add_bits(d_codes[ds], d_lengths[ds]);
      out.add_bits((uint32_t)dist_extra_bits_value((int)dist), dist_extra_bits((int)dist));
    }
  }
  out.add_bits(ll_codes[256], ll_lengths[256]);  // end symbol
  out.flush();
}

}  // namespace zb

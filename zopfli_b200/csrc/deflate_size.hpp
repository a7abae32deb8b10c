// Exact DEFLATE block-size arithmetic, shared by host code (block splitter, emitter) and by the
// sm_100a kernels (per-iteration block cost inside the iterate loop, split-cost evaluation).
//
// What the reference computes here (citations relative to /root/reference/src/zopfli/):
//   ZopfliLengthLimitedCodeLengths  katajainen.c:172-262   -> length_limited()
//   PatchDistanceCodesForBuggyDecoders deflate.c:86-99     -> patch_distance_codes()
//   EncodeTree(size_only)           deflate.c:105-249      -> encode_tree_size()
//   CalculateTreeSize               deflate.c:277-290      -> best_tree_size()
//   OptimizeHuffmanForRle           deflate.c:434-518      -> optimize_for_rle()
//   CalculateBlockSymbolSize*       deflate.c:348-420      -> symbol_bits()
//   GetDynamicLengths / TryOptimizeHuffmanForRle deflate.c:525-582 -> dynamic_lengths()
//
// The reference's boundary package-merge (a recursive node-pool algorithm) is replaced by the
// classic level-by-level package-merge over flat arrays -- no pointers, no recursion, O(n*L)
// words of scratch -- with the reference's tie rule (a leaf precedes a package only when it is
// STRICTLY lighter, katajainen.c:93-104).  Bit-identical lengths; checked against the compiled
// reference in tests/test_host_logic.py.
#pragma once
#include <stdint.h>

#include "symbols.hpp"

namespace zb {

template <int NMAX, int LMAX>
struct PmScratch {
  uint32_t key[NMAX];                     // leaves sorted by (weight << 9 | symbol)
  uint32_t row[2][2 * NMAX - 2];          // item weights of the previous / current level
  uint32_t mask[LMAX][(2 * NMAX + 29) / 32];  // bit = item is a leaf
};

// freq: n counts (< 2^22, as katajainen.c:168-170's int comparator also requires);
// out: n code lengths (0 for unused symbols).
template <int NMAX, int LMAX, typename FreqT, typename OutT>
ZB_HD void length_limited(const FreqT* freq, int n, int maxbits, OutT* out,
                          PmScratch<NMAX, LMAX>& s) {
  int ns = 0;
  for (int i = 0; i < n; i++) {
    out[i] = 0;
    if (freq[i]) s.key[ns++] = ((uint32_t)freq[i] << 9) | (uint32_t)i;
  }
  if (ns == 0) return;                                     // katajainen.c:208-211
  if (ns == 1) { out[s.key[0] & 511] = 1; return; }        // :212-216
  if (ns == 2) { out[s.key[0] & 511] = 1; out[s.key[1] & 511] = 1; return; }  // :217-222
  // katajainen.c:224-235: sort by weight, symbol index breaking ties (keys are unique).
  // Shell sort: no recursion, in place.
  for (int gap = ns >> 1; gap > 0; gap = (gap == 2) ? 1 : (gap * 5) / 11) {
    for (int i = gap; i < ns; i++) {
      uint32_t k = s.key[i];
      int j = i;
      while (j >= gap && s.key[j - gap] > k) { s.key[j] = s.key[j - gap]; j -= gap; }
      s.key[j] = k;
    }
  }
  if (ns - 1 < maxbits) maxbits = ns - 1;                  // :238-240
  const int maxitems = 2 * ns - 2;  // only the first 2n-2 items of any level can be selected
  const int mwords = (maxitems + 31) >> 5;
  // level 0: the leaves
  int prevlen = ns;
  for (int i = 0; i < ns; i++) s.row[0][i] = s.key[i] >> 9;
  for (int w = 0; w < mwords; w++) s.mask[0][w] = 0;
  for (int i = 0; i < ns; i++) s.mask[0][i >> 5] |= 1u << (i & 31);
  int cur = 1;
  for (int lev = 1; lev < maxbits; lev++, cur ^= 1) {
    const uint32_t* prev = s.row[cur ^ 1];
    uint32_t* row = s.row[cur];
    uint32_t* mk = s.mask[lev];
    for (int w = 0; w < mwords; w++) mk[w] = 0;
    const int npk = prevlen >> 1;
    int li = 0, pi = 0, len = 0;
    while (len < maxitems && (li < ns || pi < npk)) {
      bool leaf;
      uint32_t sum = 0;
      if (pi < npk) {
        sum = prev[2 * pi] + prev[2 * pi + 1];
        leaf = li < ns && sum > (s.key[li] >> 9);
      } else {
        leaf = true;
      }
      if (leaf) { row[len] = s.key[li] >> 9; mk[len >> 5] |= 1u << (len & 31); li++; }
      else { row[len] = sum; pi++; }
      len++;
    }
    prevlen = len;
  }
  // Selection (ExtractBitLengths katajainen.c:145-163): take the first 2n-2 items of the last
  // level; p selected packages pull in the first 2p items of the level below.  A leaf of
  // sorted rank r gets one bit per level whose selected-leaf count exceeds r.
  int need = maxitems;
  for (int lev = maxbits - 1; lev >= 0; lev--) {
    const uint32_t* mk = s.mask[lev];
    int c = 0, full = need >> 5, rem = need & 31;
    for (int w = 0; w < full; w++) {
#if defined(__CUDA_ARCH__)
      c += __popc(mk[w]);
#else
      c += __builtin_popcount(mk[w]);
#endif
    }
    if (rem) {
      uint32_t v = mk[full] & ((1u << rem) - 1);
#if defined(__CUDA_ARCH__)
      c += __popc(v);
#else
      c += __builtin_popcount(v);
#endif
    }
    for (int i = 0; i < c; i++) out[s.key[i] & 511]++;
    need = 2 * (need - c);
  }
}

template <typename LenT>
ZB_HD void patch_distance_codes(LenT* d_lengths) {  // deflate.c:86-99
  int num = 0;
  for (int i = 0; i < 30; i++) {
    if (d_lengths[i]) num++;
    if (num >= 2) return;
  }
  if (num == 0) d_lengths[0] = d_lengths[1] = 1;
  else if (num == 1) d_lengths[d_lengths[0] ? 1 : 0] = 1;
}

// Run-length statistics of the code-length sequence for one (use_16,use_17,use_18) choice:
// the 19 code-length-code counts of deflate.c:137-196 without materialising the RLE stream.
template <typename LenT>
ZB_HD void tree_rle_counts(const LenT* ll_lengths, const LenT* d_lengths, bool use_16, bool use_17,
                           bool use_18, uint32_t* clcounts, unsigned* hlit_out,
                           unsigned* hdist_out) {
  unsigned hlit = 29, hdist = 29;
  for (int i = 0; i < 19; i++) clcounts[i] = 0;
  while (hlit > 0 && ll_lengths[257 + hlit - 1] == 0) hlit--;    // deflate.c:133-134
  while (hdist > 0 && d_lengths[1 + hdist - 1] == 0) hdist--;
  const unsigned hlit2 = hlit + 257;
  const unsigned total = hlit2 + hdist + 1;
  for (unsigned i = 0; i < total; i++) {
    unsigned symbol = i < hlit2 ? ll_lengths[i] : d_lengths[i - hlit2];
    unsigned count = 1;
    if (use_16 || (symbol == 0 && (use_17 || use_18))) {
      for (unsigned j = i + 1;
           j < total && symbol == (unsigned)(j < hlit2 ? ll_lengths[j] : d_lengths[j - hlit2]); j++)
        count++;
    }
    i += count - 1;
    if (symbol == 0 && count >= 3) {
      if (use_18) while (count >= 11) { unsigned c2 = count > 138 ? 138 : count; clcounts[18]++; count -= c2; }
      if (use_17) while (count >= 3) { unsigned c2 = count > 10 ? 10 : count; clcounts[17]++; count -= c2; }
    }
    if (use_16 && count >= 4) {
      count--;
      clcounts[symbol]++;
      while (count >= 3) { unsigned c2 = count > 6 ? 6 : count; clcounts[16]++; count -= c2; }
    }
    clcounts[symbol] += count;
  }
  *hlit_out = hlit;
  *hdist_out = hdist;
}

// deflate.c:105-249 with out == NULL
template <typename LenT>
ZB_HD uint32_t encode_tree_size(const LenT* ll_lengths, const LenT* d_lengths, bool use_16,
                                bool use_17, bool use_18) {
  uint32_t clcounts[19];
  uint8_t clcl[19];
  unsigned hlit, hdist;
  tree_rle_counts(ll_lengths, d_lengths, use_16, use_17, use_18, clcounts, &hlit, &hdist);
  PmScratch<19, 7> s;
  length_limited<19, 7>(clcounts, 19, 7, clcl, s);
  const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  unsigned hclen = 15;
  while (hclen > 0 && clcounts[order[hclen + 4 - 1]] == 0) hclen--;
  uint32_t r = 14 + (hclen + 4) * 3;
  for (int i = 0; i < 19; i++) r += clcl[i] * clcounts[i];
  r += clcounts[16] * 2 + clcounts[17] * 3 + clcounts[18] * 7;
  return r;
}

// deflate.c:277-290 / 251-272: first minimum over the 8 flag combinations.
template <typename LenT>
ZB_HD uint32_t best_tree_size(const LenT* ll_lengths, const LenT* d_lengths, int* best_flags) {
  uint32_t best = 0;
  int bf = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t size = encode_tree_size(ll_lengths, d_lengths, (i & 1) != 0, (i & 2) != 0, (i & 4) != 0);
    if (best == 0 || size < best) { best = size; bf = i; }
  }
  if (best_flags) *best_flags = bf;
  return best;
}

// deflate.c:379-401 (the `Small` variant :348-374 adds up the same integers symbol by symbol)
template <typename CntT, typename LenT>
ZB_HD uint64_t symbol_bits(const CntT* llc, const CntT* dc, const LenT* ll, const LenT* d) {
  uint64_t r = 0;
  for (int i = 0; i < 256; i++) r += (uint64_t)ll[i] * llc[i];
  for (int i = 257; i < 286; i++) r += (uint64_t)(ll[i] + length_symbol_extra_bits(i)) * llc[i];
  for (int i = 0; i < 30; i++) r += (uint64_t)(d[i] + dist_symbol_extra_bits(i)) * dc[i];
  return r + ll[256];
}

// deflate.c:434-518.  `good` is caller scratch of `length` bytes.
template <typename CntT>
ZB_HD void optimize_for_rle(int length, CntT* counts, uint8_t* good) {
  for (; length >= 0; --length) {
    if (length == 0) return;
    if (counts[length - 1] != 0) break;
  }
  for (int i = 0; i < length; ++i) good[i] = 0;
  CntT symbol = counts[0];
  int stride = 0;
  for (int i = 0; i < length + 1; ++i) {
    if (i == length || counts[i] != symbol) {
      if ((symbol == 0 && stride >= 5) || (symbol != 0 && stride >= 7))
        for (int k = 0; k < stride; ++k) good[i - k - 1] = 1;
      stride = 1;
      if (i != length) symbol = counts[i];
    } else {
      ++stride;
    }
  }
  stride = 0;
  uint64_t limit = counts[0], sum = 0;
  for (int i = 0; i < length + 1; ++i) {
    bool brk = (i == length) || good[i];
    if (!brk) {
      uint64_t c = counts[i];
      brk = (c > limit ? c - limit : limit - c) >= 4;
    }
    if (brk) {
      if (stride >= 4 || (stride >= 3 && sum == 0)) {
        int count = (int)((sum + (uint64_t)(stride / 2)) / (uint64_t)stride);
        if (count < 1) count = 1;
        if (sum == 0) count = 0;
        for (int k = 0; k < stride; ++k) counts[i - k - 1] = (CntT)count;
      }
      stride = 0;
      sum = 0;
      if (i < length - 3)
        limit = ((uint64_t)counts[i] + counts[i + 1] + counts[i + 2] + counts[i + 3] + 2) / 4;
      else if (i < length) limit = counts[i];
      else limit = 0;
    }
    ++stride;
    if (i != length) sum += counts[i];
  }
}

struct DynScratch {
  PmScratch<kNumLL, 15> pm;
  uint32_t llc2[kNumLL];
  uint32_t dc2[kNumD];
  uint8_t ll2[kNumLL];
  uint8_t d2[kNumD];
  uint8_t good[kNumLL];
};

// GetDynamicLengths (deflate.c:569-582): code lengths minimising tree + data bits between the
// plain and the RLE-smoothed histograms.  llc[256] must already be 1.  Returns the bit count
// without the 3 block-header bits.
ZB_HD uint64_t dynamic_lengths(const uint32_t* llc, const uint32_t* dc, uint8_t* ll, uint8_t* d,
                               DynScratch& s) {
  length_limited<kNumLL, 15>(llc, kNumLL, 15, ll, s.pm);
  length_limited<kNumLL, 15>(dc, kNumD, 15, d, s.pm);
  patch_distance_codes(d);
  uint64_t size1 = best_tree_size(ll, d, nullptr) + symbol_bits(llc, dc, ll, d);
  for (int i = 0; i < kNumLL; i++) s.llc2[i] = llc[i];
  for (int i = 0; i < kNumD; i++) s.dc2[i] = dc[i];
  optimize_for_rle(kNumLL, s.llc2, s.good);
  optimize_for_rle(kNumD, s.dc2, s.good);
  length_limited<kNumLL, 15>(s.llc2, kNumLL, 15, s.ll2, s.pm);
  length_limited<kNumLL, 15>(s.dc2, kNumD, 15, s.d2, s.pm);
  patch_distance_codes(s.d2);
  uint64_t size2 = best_tree_size(s.ll2, s.d2, nullptr) + symbol_bits(llc, dc, s.ll2, s.d2);
  if (size2 < size1) {  // deflate.c:553-557
    for (int i = 0; i < kNumLL; i++) ll[i] = s.ll2[i];
    for (int i = 0; i < kNumD; i++) d[i] = s.d2[i];
    return size2;
  }
  return size1;
}

// GetFixedTree deflate.c:335-342
ZB_HD int fixed_ll_length(int i) { return i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)); }

template <typename CntT>
ZB_HD uint64_t fixed_symbol_bits(const CntT* llc, const CntT* dc) {
  uint64_t r = 0;
  for (int i = 0; i < 256; i++) r += (uint64_t)fixed_ll_length(i) * llc[i];
  for (int i = 257; i < 286; i++) r += (uint64_t)(fixed_ll_length(i) + length_symbol_extra_bits(i)) * llc[i];
  for (int i = 0; i < 30; i++) r += (uint64_t)(5 + dist_symbol_extra_bits(i)) * dc[i];
  return r + 7;  // end symbol
}

// ZopfliCalculateBlockSize btype 0 (deflate.c:590-597)
ZB_HD uint64_t stored_bits(uint64_t nbytes) {
  uint64_t rem = nbytes % 65535;
  uint64_t blocks = nbytes / 65535 + (rem ? 1 : 0);
  return blocks * 5 * 8 + nbytes * 8;
}

}  // namespace zb

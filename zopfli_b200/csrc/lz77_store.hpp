// Host-side LZ77 symbol list with sampled cumulative histograms, and the block-cost estimators
// the splitter and the block-type choice are built on.
//
// Plays the role of the reference's ZopfliLZ77Store (/root/reference/src/zopfli/lz77.h:44-62,
// lz77.c:98-217) but is laid out for bulk transfer from the GPU: plain SoA vectors, one
// histogram snapshot every kSnap symbols (the reference keeps one running counter per symbol
// slot, lz77.c:107-124, which cannot be filled in parallel).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "deflate_size.hpp"

namespace zb {

struct Lz77Store {
  static constexpr size_t kSnap = 256;
  std::vector<uint16_t> litlens;  // literal byte or match length
  std::vector<uint16_t> dists;    // 0 => literal
  std::vector<size_t> pos;        // byte position of the symbol in the input
  std::vector<uint16_t> llsym;    // literal/length alphabet symbol
  std::vector<uint8_t> dsym;      // distance alphabet symbol (undefined for literals)
  std::vector<uint32_t> snaps;    // [(size/kSnap)+1][320] histogram of symbols [0, k*kSnap)

  size_t size() const { return litlens.size(); }
  void clear() { litlens.clear(); dists.clear(); pos.clear(); llsym.clear(); dsym.clear(); snaps.clear(); }

  // Appends n symbols starting at byte position `start`; call finalize() when done.
  void append(const uint16_t* ll, const uint16_t* dd, size_t n, size_t start) {
    size_t p = start, base = litlens.size();
    litlens.insert(litlens.end(), ll, ll + n);
    dists.insert(dists.end(), dd, dd + n);
    pos.resize(base + n);
    llsym.resize(base + n);
    dsym.resize(base + n);
    for (size_t i = 0; i < n; i++) {
      pos[base + i] = p;
      if (dd[i] == 0) { llsym[base + i] = ll[i]; dsym[base + i] = 0; p += 1; }
      else { llsym[base + i] = (uint16_t)length_symbol(ll[i]); dsym[base + i] = (uint8_t)dist_symbol(dd[i]); p += ll[i]; }
    }
  }
  void append(const Lz77Store& o) {  // ZopfliAppendLZ77Store lz77.c:151-158
    litlens.insert(litlens.end(), o.litlens.begin(), o.litlens.end());
    dists.insert(dists.end(), o.dists.begin(), o.dists.end());
    pos.insert(pos.end(), o.pos.begin(), o.pos.end());
    llsym.insert(llsym.end(), o.llsym.begin(), o.llsym.end());
    dsym.insert(dsym.end(), o.dsym.begin(), o.dsym.end());
  }
  void finalize() {
    size_t n = size(), ns = n / kSnap + 1;
    snaps.assign(ns * 320, 0);
    uint32_t run[320];
    memset(run, 0, sizeof(run));
    for (size_t i = 0; i < n; i++) {
      if (i % kSnap == 0) memcpy(&snaps[(i / kSnap) * 320], run, sizeof(run));
      run[llsym[i]]++;
      if (dists[i]) run[288 + dsym[i]]++;
    }
    if (n % kSnap == 0) memcpy(&snaps[(n / kSnap) * 320], run, sizeof(run));
  }
  // histogram of symbols [0, i)
  void prefix_hist(size_t i, uint32_t* h) const {
    size_t k = i / kSnap;
    memcpy(h, &snaps[k * 320], 320 * sizeof(uint32_t));
    for (size_t j = k * kSnap; j < i; j++) {
      h[llsym[j]]++;
      if (dists[j]) h[288 + dsym[j]]++;
    }
  }
  // ZopfliLZ77GetHistogram lz77.c:189-217: h[0..288) literal/length, h[288..320) distance
  void range_hist(size_t lstart, size_t lend, uint32_t* h) const {
    if (lend - lstart < 2 * kSnap) {
      memset(h, 0, 320 * sizeof(uint32_t));
      for (size_t j = lstart; j < lend; j++) {
        h[llsym[j]]++;
        if (dists[j]) h[288 + dsym[j]]++;
      }
      return;
    }
    uint32_t a[320];
    prefix_hist(lend, h);
    prefix_hist(lstart, a);
    for (int i = 0; i < 320; i++) h[i] -= a[i];
  }
  // ZopfliLZ77GetByteRange lz77.c:160-166
  size_t byte_range(size_t lstart, size_t lend) const {
    if (lstart == lend) return 0;
    size_t l = lend - 1;
    return pos[l] + (dists[l] == 0 ? 1 : litlens[l]) - pos[lstart];
  }
};

// ZopfliCalculateBlockSize (deflate.c:584-608) from a histogram. Integer bit counts; the
// reference carries them in doubles, which represent them exactly.
struct BlockCosts {
  uint64_t stored, fixed, dynamic;
};

inline uint64_t dynamic_block_bits(const uint32_t* hist, uint8_t* ll_out, uint8_t* d_out,
                                   DynScratch& s) {
  uint32_t llc[kNumLL];
  memcpy(llc, hist, sizeof(llc));
  llc[256] = 1;  // deflate.c:575
  uint8_t ll[kNumLL], d[kNumD];
  uint64_t r = 3 + dynamic_lengths(llc, hist + 288, ll_out ? ll_out : ll, d_out ? d_out : d, s);
  return r;
}

inline uint64_t fixed_block_bits(const uint32_t* hist) { return 3 + fixed_symbol_bits(hist, hist + 288); }

// ZopfliCalculateBlockSizeAutoType (deflate.c:610-621). `whole_size` is the size of the store
// the range lives in: the reference tests lz77->size, not the range.
inline uint64_t auto_type_bits(const Lz77Store& st, size_t lstart, size_t lend, DynScratch& s) {
  uint32_t h[320];
  st.range_hist(lstart, lend, h);
  uint64_t unc = stored_bits(st.byte_range(lstart, lend));
  uint64_t fixed = st.size() > 1000 ? unc : fixed_block_bits(h);
  uint64_t dyn = dynamic_block_bits(h, nullptr, nullptr, s);
  return (unc < fixed && unc < dyn) ? unc : (fixed < dyn ? fixed : dyn);
}

}  // namespace zb

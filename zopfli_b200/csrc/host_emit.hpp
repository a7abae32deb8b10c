// Host-side twins of k_block_plan / k_emit (finish.cuh), built from the SAME shared arithmetic
// (deflate_size.hpp, emit_bits.hpp).  The product emits on the device; these serve the C-ABI test seam
// ZopfliB200HostEmitBlock and the CPU test mock (tests/hostmock), so that the emission code the kernels
// run is also exercised on a box without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#include "emit_bits.hpp"
#include "lz77_store.hpp"

namespace zb {

struct HostBitSink {  // LSB-first writer into a zero-initialised byte buffer
  uint8_t* bytes;
  uint64_t pos;
  void put(uint32_t v, int n) {
    if (n == 0) return;
    uint64_t val = (uint64_t)v << (pos & 7);
    for (size_t b = (size_t)(pos >> 3); val; b++, val >>= 8) bytes[b] |= (uint8_t)val;
    pos += (uint64_t)n;
  }
};

// what k_block_plan computes for symbols [0, n) of (ll, d)
inline void host_block_plan(const uint16_t* ll, const uint16_t* d, size_t n, BlockPlan& p) {
  uint32_t h[320];
  memset(h, 0, sizeof(h));
  uint64_t nbytes = 0;
  for (size_t i = 0; i < n; i++) {
    if (d[i] == 0) { h[ll[i]]++; nbytes += 1; }
    else { h[length_symbol(ll[i])]++; h[288 + dist_symbol(d[i])]++; nbytes += ll[i]; }
  }
  DynScratch s;
  memset(&p, 0, sizeof(p));
  p.nbytes = nbytes;
  p.unc_bits = stored_bits(nbytes);
  p.fixed_bits = fixed_block_bits(h);
  p.dyn_bits = dynamic_block_bits(h, p.ll_len, p.d_len, s);
  int flags = 0;
  p.tree_bits = best_tree_size(p.ll_len, p.d_len, &flags);
  p.tree_flags = (uint32_t)flags;
}

// what k_emit writes for one compressed piece (btype 1 or 2) at sink.pos; returns the bits written
inline uint64_t host_emit_block(int btype, bool final, const uint16_t* ll, const uint16_t* d, size_t n,
                                const BlockPlan* plan, HostBitSink& sink) {
  const uint64_t start = sink.pos;
  uint8_t ll_len[kNumLL], d_len[kNumD];
  uint16_t ll_code[kNumLL], d_code[kNumD];
  sink.put((final ? 1u : 0u) | (btype == 1 ? 2u : 4u), 3);
  if (btype == 2) {
    memcpy(ll_len, plan->ll_len, kNumLL);
    memcpy(d_len, plan->d_len, kNumD);
    write_tree_header(ll_len, d_len, plan->tree_flags, sink);
  } else {
    for (int i = 0; i < kNumLL; i++) ll_len[i] = (uint8_t)fixed_ll_length(i);
    for (int i = 0; i < kNumD; i++) d_len[i] = 5;
  }
  canonical_codes(ll_len, kNumLL, ll_code);
  canonical_codes(d_len, kNumD, d_code);
  for (size_t i = 0; i < n; i++) {
    const SymBits sb = symbol_bits_of(ll[i], d[i], ll_len, ll_code, d_len, d_code);
    sink.put(sb.v0, sb.n0);
    sink.put(sb.v1, sb.n1);
  }
  sink.put(ll_code[256], ll_len[256]);
  return sink.pos - start;
}

// stored piece (AddNonCompressedBlock deflate.c:625-663) at sink.pos
inline void host_emit_stored(bool final, const uint8_t* data, uint64_t nbytes, HostBitSink& sink) {
  uint64_t left = nbytes;
  do {
    const uint64_t bs = left > 65535 ? 65535 : left;
    const bool last = left == bs;
    sink.put((final && last) ? 1u : 0u, 3);
    sink.pos = (sink.pos + 7) & ~(uint64_t)7;
    sink.put((uint32_t)bs, 16);
    sink.put((~(uint32_t)bs) & 0xffffu, 16);
    memcpy(sink.bytes + (sink.pos >> 3), data, (size_t)bs);
    sink.pos += bs * 8;
    data += bs;
    left -= bs;
  } while (left);
}

}  // namespace zb

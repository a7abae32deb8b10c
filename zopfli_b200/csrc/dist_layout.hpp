// Pure arithmetic of the multi-GPU path (dist.cpp), kept apart so it can be tested without GPUs or NCCL
// (tests/test_sharding_gloo.py drives it over gloo): which bytes a rank owns, and where every rank's bits
// land in the one output stream.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "symbols.hpp"

namespace zb {

inline size_t dist_num_master_blocks(size_t insize) {  // deflate.c:912-924 do/while
  return insize == 0 ? 1 : (insize + kMasterBlock - 1) / kMasterBlock;
}

// Rank `rank` of `world` owns the master blocks [rank*nm/world, (rank+1)*nm/world): bytes [a, b).  Its
// device copy starts at `base` <= a: the 32 KiB dictionary in front of the shard (squeeze.c:229-230),
// rounded down to 16 bytes for the vector loads of the kernels.  Ranks without a master block get a == b.
inline void dist_shard(size_t insize, int world, int rank, size_t* a, size_t* b, size_t* base) {
  const size_t nm = dist_num_master_blocks(insize);
  const size_t lo = (size_t)rank * nm / (size_t)world, hi = (size_t)(rank + 1) * nm / (size_t)world;
  *a = lo * (size_t)kMasterBlock < insize ? lo * (size_t)kMasterBlock : insize;
  *b = hi * (size_t)kMasterBlock < insize ? hi * (size_t)kMasterBlock : insize;
  *base = *a > (size_t)kWindow ? (*a - kWindow) & ~(size_t)15 : 0;
  if (lo == hi) *base = *a = *b;
}

// len8[r*stride + p] = bits rank r's blocks occupy when its first bit sits at bit phase p (stored blocks pad
// to a byte boundary, deflate.c:643-649, so the length depends on the phase).  start[0] = phase0 and
// start[r+1] = start[r] + len8[r][start[r] & 7]: absolute bit offsets counted from bit 0 of the byte that
// holds the caller's phase0 bits.
inline void dist_placement(const uint64_t* len8, size_t stride, int world, unsigned phase0, uint64_t* start) {
  start[0] = phase0 & 7u;
  for (int r = 0; r < world; r++) start[r + 1] = start[r] + len8[(size_t)r * stride + (start[r] & 7)];
}

// bytes of the stream rank r's bits touch, counted from the byte of its first bit
inline size_t dist_bytes_touched(const uint64_t* start, int r) {
  if (start[r + 1] == start[r]) return 0;
  return (size_t)(((start[r] & 7) + (start[r + 1] - start[r]) + 7) / 8);
}

}  // namespace zb

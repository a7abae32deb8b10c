// Split-point search over an LZ77 symbol list: what ZopfliBlockSplitLZ77 computes
// (/root/reference/src/zopfli/blocksplitter.c:43-96 FindMinimum, :125-128 SplitCost,
// :195-213 FindLargestSplittableBlock, :215-273 ZopfliBlockSplitLZ77), restated around a
// pluggable batched cost evaluator so the probes of one search round can be priced together
// (on host threads today; the same interface feeds the device evaluator).
#pragma once
#include <algorithm>
#include <functional>
#include <vector>

#include "lz77_store.hpp"

namespace zb {

// cost(lstart, lend) = ZopfliCalculateBlockSizeAutoType over that symbol range.
using RangeCostFn = std::function<uint64_t(size_t, size_t)>;

// blocksplitter.c:43-96.  f(i) = cost(start0, i) + cost(i, end0).
inline size_t find_minimum(const RangeCostFn& cost, size_t cstart, size_t cend, size_t start,
                           size_t end, uint64_t* smallest) {
  auto f = [&](size_t i) { return cost(cstart, i) + cost(i, cend); };
  const uint64_t kLarge = ~(uint64_t)0;  // ZOPFLI_LARGE_FLOAT: larger than any real cost
  if (end - start < 1024) {
    uint64_t best = kLarge;
    size_t result = start;
    for (size_t i = start; i < end; i++) {
      uint64_t v = f(i);
      if (v < best) { best = v; result = i; }
    }
    *smallest = best;
    return result;
  }
  constexpr int NUM = 9;
  size_t p[NUM];
  uint64_t vp[NUM];
  uint64_t lastbest = kLarge;
  size_t pos = start;
  for (;;) {
    if (end - start <= (size_t)NUM) break;
    for (int i = 0; i < NUM; i++) {
      p[i] = start + (size_t)(i + 1) * ((end - start) / (NUM + 1));
      vp[i] = f(p[i]);
    }
    int besti = 0;
    uint64_t best = vp[0];
    for (int i = 1; i < NUM; i++)
      if (vp[i] < best) { best = vp[i]; besti = i; }
    if (best > lastbest) break;
    start = besti == 0 ? start : p[besti - 1];
    end = besti == NUM - 1 ? end : p[besti + 1];
    pos = p[besti];
    lastbest = best;
  }
  *smallest = lastbest;
  return pos;
}

// blocksplitter.c:215-273.  Returns the split points as LZ77 symbol indices (sorted).
inline std::vector<size_t> block_split_lz77(const RangeCostFn& cost, size_t lz77size,
                                            size_t maxblocks) {
  std::vector<size_t> points;
  if (lz77size < 10) return points;  // :225
  std::vector<unsigned char> done(lz77size, 0);
  size_t lstart = 0, lend = lz77size, numblocks = 1;
  for (;;) {
    if (maxblocks > 0 && numblocks >= maxblocks) break;
    uint64_t splitcost;
    size_t llpos = find_minimum(cost, lstart, lend, lstart + 1, lend, &splitcost);
    uint64_t origcost = cost(lstart, lend);
    if (splitcost > origcost || llpos == lstart + 1 || llpos == lend) {
      done[lstart] = 1;
    } else {
      points.insert(std::upper_bound(points.begin(), points.end(), llpos), llpos);  // AddSorted :130-143
      numblocks++;
    }
    // FindLargestSplittableBlock :195-213
    size_t longest = 0;
    bool found = false;
    for (size_t i = 0; i <= points.size(); i++) {
      size_t s = i == 0 ? 0 : points[i - 1];
      size_t e = i == points.size() ? lz77size - 1 : points[i];
      if (!done[s] && e - s > longest) { lstart = s; lend = e; found = true; longest = e - s; }
    }
    if (!found) break;
    if (lend - lstart < 10) break;
  }
  return points;
}

}  // namespace zb

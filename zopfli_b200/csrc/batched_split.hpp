// Batched, speculative form of ZopfliBlockSplitLZ77 (/root/reference/src/zopfli/blocksplitter.c:
// 43-96 FindMinimum, :195-273) for many LZ77 stores at once.
//
// The reference runs one FindMinimum at a time and one probe at a time.  Both are pure functions of
// the symbol range, so here (a) every probe of a FindMinimum round is priced in one batch,
// (b) FindMinimum is started for EVERY current block of every store as soon as the block exists
// (whichever the "largest first" rule picks next, its answer is already there), and (c) all stores
// advance in lockstep rounds, and (d) while few searches are active a round prices two or three
// levels of the nine-point recursion at once (every interval the next level could narrow to),
// which divides the number of rounds -- the latency that is left -- by the depth.  One round =
// one call of the batch evaluator (a kernel launch on the device, k_split_eval).  The sequential decision logic (largest splittable block first, `done`
// marks, maxblocks, the `lz77size - 1` end quirk of :203) is replayed unchanged on the cached
// answers, so the split points are identical to the reference's.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <map>
#include <utility>
#include <vector>

namespace zb {

struct EvalReq { uint32_t store, lstart, lend; };
using BatchEvalFn = std::function<void(const std::vector<EvalReq>&, std::vector<uint64_t>&)>;

namespace bsplit {

constexpr uint64_t kLarge = ~(uint64_t)0;

struct FmTask {  // one FindMinimum(SplitCost) over the block [cstart, cend)
  uint32_t store;
  size_t cstart, cend;
  size_t start, end;      // current search interval (start = cstart+1, end = cend initially)
  bool exhaustive;
  size_t pos;
  uint64_t lastbest = kLarge;
  bool have_orig = false;
  uint64_t origcost = 0;
  bool finished = false;
  size_t llpos = 0;
  uint64_t splitcost = 0;
  size_t req_base = 0;   // the orig-cost request of the current batch, if any
  int root = -1;         // node of the current interval in this round's speculation tree
};

struct SpecNode {        // one (start, end) interval of the nine-point recursion and its probes
  size_t start, end;
  size_t p[9];
  size_t req_base;       // 18 requests: (cstart, p[i]), (p[i], cend)
  bool has;              // end - start > 9
  int child[9];
};

struct StoreState {
  size_t n = 0;
  std::vector<size_t> points;
  std::map<size_t, bool> done;  // block start -> not splittable
  size_t numblocks = 1;
  size_t lstart = 0, lend = 0;
  bool finished = false;
  std::map<std::pair<size_t, size_t>, size_t> fm;  // (cstart,cend) -> task index
};

}  // namespace bsplit

// budget: probes per round up to which deeper speculation is worth its kernel time
inline size_t default_split_budget() {
  static size_t v = [] { const char* e = getenv("ZOPFLI_B200_SPLIT_BUDGET"); return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)6000; }();
  return v;
}

inline std::vector<std::vector<size_t>> batched_block_split(const std::vector<size_t>& sizes, size_t maxblocks,
                                                             const BatchEvalFn& eval, size_t budget = default_split_budget()) {
  using namespace bsplit;
  const size_t ns = sizes.size();
  std::vector<StoreState> S(ns);
  std::vector<FmTask> T;
  auto spawn = [&](uint32_t st, size_t a, size_t b) {
    auto key = std::make_pair(a, b);
    if (S[st].fm.count(key)) return;
    FmTask t;
    t.store = st; t.cstart = a; t.cend = b;
    t.start = a + 1; t.end = b;
    t.exhaustive = (t.end - t.start) < 1024;  // blocksplitter.c:45
    t.pos = t.start;
    S[st].fm[key] = T.size();
    T.push_back(t);
  };
  auto speculate = [&](uint32_t st) {  // FindMinimum for every current block that could be picked
    StoreState& s = S[st];
    if (s.finished) return;
    if (maxblocks > 0 && s.numblocks >= maxblocks) return;
    for (size_t i = 0; i <= s.points.size(); i++) {
      size_t a = i == 0 ? 0 : s.points[i - 1];
      size_t b = i == s.points.size() ? s.n - 1 : s.points[i];  // blocksplitter.c:203
      if (s.done.count(a) || b < a || b - a < 10) continue;
      spawn(st, a, b);
    }
  };
  // replay of the reference loop on cached answers; returns when it needs an unfinished task
  auto advance = [&](uint32_t st) {
    StoreState& s = S[st];
    while (!s.finished) {
      if (maxblocks > 0 && s.numblocks >= maxblocks) { s.finished = true; break; }
      auto it = s.fm.find(std::make_pair(s.lstart, s.lend));
      if (it == s.fm.end()) { spawn(st, s.lstart, s.lend); break; }
      const FmTask& t = T[it->second];
      if (!t.finished) break;
      if (t.splitcost > t.origcost || t.llpos == s.lstart + 1 || t.llpos == s.lend) {  // :251
        s.done[s.lstart] = true;
      } else {
        s.points.insert(std::upper_bound(s.points.begin(), s.points.end(), t.llpos), t.llpos);
        s.numblocks++;
      }
      size_t longest = 0;  // FindLargestSplittableBlock :195-213
      bool found = false;
      for (size_t i = 0; i <= s.points.size(); i++) {
        size_t a = i == 0 ? 0 : s.points[i - 1];
        size_t b = i == s.points.size() ? s.n - 1 : s.points[i];
        if (!s.done.count(a) && b - a > longest) { s.lstart = a; s.lend = b; found = true; longest = b - a; }
      }
      if (!found) { s.finished = true; break; }
      if (s.lend - s.lstart < 10) { s.finished = true; break; }
    }
    speculate(st);
  };
  for (size_t st = 0; st < ns; st++) {
    S[st].n = sizes[st];
    if (sizes[st] < 10) { S[st].finished = true; continue; }  // :225
    S[st].lstart = 0;
    S[st].lend = sizes[st];
    advance((uint32_t)st);
  }
  std::vector<EvalReq> reqs;
  std::vector<uint64_t> costs;
  std::vector<SpecNode> nodes;
  const size_t kBudget = budget;
  for (;;) {
    reqs.clear();
    nodes.clear();
    bool any = false;
    size_t narrowing = 0;
    for (FmTask& t : T)
      if (!t.finished && !S[t.store].finished && !t.exhaustive && t.end - t.start > 9) narrowing++;
    const int depth = narrowing * 18 * 91 <= kBudget ? 3 : (narrowing * 18 * 10 <= kBudget ? 2 : 1);
    for (FmTask& t : T) {
      t.root = -1;
      if (t.finished || S[t.store].finished) continue;
      any = true;
      t.req_base = reqs.size();
      if (!t.have_orig) reqs.push_back({t.store, (uint32_t)t.cstart, (uint32_t)t.cend});
      if (t.exhaustive) {
        for (size_t i = t.start; i < t.end; i++) {
          reqs.push_back({t.store, (uint32_t)t.cstart, (uint32_t)i});
          reqs.push_back({t.store, (uint32_t)i, (uint32_t)t.cend});
        }
      } else {
        // speculation tree over the intervals the recursion can reach within `depth` levels
        struct Builder {
          std::vector<SpecNode>& nodes; std::vector<EvalReq>& reqs; const FmTask& t;
          int build(size_t start, size_t end, int d) {
            const int id = (int)nodes.size();
            nodes.push_back(SpecNode());
            SpecNode n;
            n.start = start; n.end = end; n.has = end - start > 9;  // blocksplitter.c:73
            n.req_base = reqs.size();
            for (int i = 0; i < 9; i++) n.child[i] = -1;
            if (n.has) {
              for (int i = 0; i < 9; i++) {
                n.p[i] = start + (size_t)(i + 1) * ((end - start) / 10);
                reqs.push_back({t.store, (uint32_t)t.cstart, (uint32_t)n.p[i]});
                reqs.push_back({t.store, (uint32_t)n.p[i], (uint32_t)t.cend});
              }
              if (d > 1)
                for (int i = 0; i < 9; i++)
                  n.child[i] = build(i == 0 ? start : n.p[i - 1], i == 8 ? end : n.p[i + 1], d - 1);
            }
            nodes[id] = n;
            return id;
          }
        } bld{nodes, reqs, t};
        t.root = bld.build(t.start, t.end, depth);
      }
    }
    if (!any) break;
    costs.assign(reqs.size(), 0);
    if (!reqs.empty()) eval(reqs, costs);
    for (FmTask& t : T) {
      if (t.finished || S[t.store].finished) continue;
      size_t k = t.req_base;
      if (!t.have_orig) { t.origcost = costs[k++]; t.have_orig = true; }
      if (t.exhaustive) {  // blocksplitter.c:45-58
        uint64_t best = kLarge;
        size_t result = t.start;
        for (size_t i = t.start; i < t.end; i++) {
          uint64_t v = costs[k] + costs[k + 1];
          k += 2;
          if (v < best) { best = v; result = i; }
        }
        t.llpos = result; t.splitcost = best; t.finished = true;
        continue;
      }
      for (int cur = t.root; cur >= 0;) {  // blocksplitter.c:71-91, one loop iteration per tree level
        const SpecNode& n = nodes[cur];
        if (!n.has) { t.llpos = t.pos; t.splitcost = t.lastbest; t.finished = true; break; }
        const size_t kb = n.req_base;
        int besti = 0;
        uint64_t best = costs[kb] + costs[kb + 1];
        for (int i = 1; i < 9; i++) {
          uint64_t v = costs[kb + 2 * i] + costs[kb + 2 * i + 1];
          if (v < best) { best = v; besti = i; }
        }
        if (best > t.lastbest) { t.llpos = t.pos; t.splitcost = t.lastbest; t.finished = true; break; }
        t.start = besti == 0 ? n.start : n.p[besti - 1];
        t.end = besti == 8 ? n.end : n.p[besti + 1];
        t.pos = n.p[besti];
        t.lastbest = best;
        cur = n.child[besti];
      }
      if (!t.finished && t.end - t.start <= 9) { t.llpos = t.pos; t.splitcost = t.lastbest; t.finished = true; }
    }
    for (size_t st = 0; st < ns; st++) advance((uint32_t)st);
  }
  std::vector<std::vector<size_t>> out(ns);
  for (size_t st = 0; st < ns; st++) out[st] = S[st].points;
  return out;
}

}  // namespace zb

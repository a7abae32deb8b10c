/*
 * dp_int_model.c -- TEST INFRASTRUCTURE ONLY (see zopfli_oracle.h).
 *
 * Sequential model of the INTEGER formulation of the forward DP that k_iterate uses while all costs
 * of a group of 32 positions provably stay inside one float binade (zopfli_b200/csrc/iterate.cuh,
 * "integer window").  GetBestLengths (squeeze.c:217-309) stores costs as float and computes every
 * relaxation as  (double)costs[j] + cost  compared in double against (double)costs[j+k]  and stored
 * rounded to float.  Inside one binade [2^e, 2^(e+1)) with float ulp u = 2^(e-23):
 *
 *   fl64(x + c)            = x + c'      c' = c rounded to the double grid 2^(e-52) of the binade
 *   (float)fl64(x + c)     = x + k1 u    k1 = round-to-nearest(c' / u), provided c'/u is no exact tie
 *   newCost < costs[t]    <=> x + k1 u < costs[t], or equal and the float rounding went UP (`up`)
 *
 * so a cost is the integer (float bits - bits(2^e)), an edge is "+ k1", and the strict double compare
 * plus the sequential "first strict improvement wins" order of the reference become ONE unsigned
 * minimum over packed words  (value << 7) | tiebreak,  tiebreak = up ? 63 - o : 64 + o,  o = order of
 * the edge's source among the sources of the target (0: length >= 35, 1..32: lengths 34..3, 33: the
 * literal).  This file runs that model next to the oracle's reference DP on every pass of
 * zo_lz77_optimal / zo_lz77_optimal_fixed and counts positions whose cost or length_array entry
 * differs.  It includes zopfli_oracle.c to reach its static helpers; nothing in the product links it.
 */
#include "zopfli_oracle.c"

#include <stdio.h>

typedef struct {
  uint64_t steps_total, steps_int, groups_int, groups_int_ring, mismatches, passes, tie_binades, builds;
} ZoIntDpStats;
static ZoIntDpStats g_ids;

#define INFP 0xffffffffu
#define NOEDGE (1u << 30)

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

typedef struct {
  int e, tie;
  uint32_t base;           /* float bits of 2^e */
  uint32_t te[31][35];     /* [dsym][len]; row 30: no edge */
  uint32_t lit[256];
  uint32_t tl[29][30];     /* long edges: (k1 << 1) | up per (length symbol - 257, distance symbol) */
} BinadeTab;

/* k1 / up / tie of one edge cost c in binade e; order o */
static uint32_t pack_edge(double c, int e, int o, int* tie) {
  const double B = ldexp(1.0, e);
  const double s0 = B + c;          /* rounds c to the binade's double grid */
  const float f = (float)s0;
  const double df = (double)f;
  const uint32_t k1 = fbits(f) - fbits((float)B);
  const int up = df > s0;
  if (fabs(df - s0) == ldexp(1.0, e - 24)) *tie = 1;
  return (k1 << 7) | (uint32_t)(up ? 63 - o : 64 + o);
}

static void build_tab(BinadeTab* bt, int e, const ZoStats* st) {
  int ds, len, i;
  bt->e = e;
  bt->tie = 0;
  bt->base = fbits((float)ldexp(1.0, e));
  for (ds = 0; ds < 30; ds++)
    for (len = 3; len <= 34; len++) {
      const double c = length_extra_bits(len) + dist_symbol_extra_bits(ds) + st->ll_symbols[length_symbol(len)] + st->d_symbols[ds];
      bt->te[ds][len] = pack_edge(c, e, 35 - len, &bt->tie);
    }
  for (len = 3; len <= 34; len++) bt->te[30][len] = NOEDGE | (uint32_t)(64 + 35 - len);
  for (i = 0; i < 256; i++) bt->lit[i] = pack_edge(st->ll_symbols[i], e, 33, &bt->tie);
  for (i = 0; i < 29; i++)
    for (ds = 0; ds < 30; ds++) {
      /* the cost of a length depends on it only through its symbol (extra bits are a function of the symbol) */
      const double c = length_symbol_extra_bits(257 + i) + dist_symbol_extra_bits(ds) + st->ll_symbols[257 + i] + st->d_symbols[ds];
      const uint32_t w = pack_edge(c, e, 0, &bt->tie); /* o = 0: tiebreak 63 (up) or 64 */
      bt->tl[i][ds] = ((w >> 7) << 1) | ((w & 127u) == 63u ? 1u : 0u);
    }
  g_ids.builds++;
}

static uint16_t decode_tb(uint32_t p, uint16_t keep) {
  const uint32_t tb = p & 127u, o = tb < 64 ? 63 - tb : tb - 64;
  if (o == 33) return 1;
  if (o == 0) return keep;
  return (uint16_t)(35 - o);
}

static void int_model(const ZoSegment* seg, const ZoTable* t, const ZoStats* st,
                      const uint16_t* la_ref, const float* costs_ref) {
  const uint8_t* in = seg->in;
  const size_t nb = seg->inend - seg->instart;
  size_t j, k, g;
  const double mincost = model_min_cost(st);
  const double cost258 = cost_stat(ZO_MAX_MATCH, 1, st);
  float* costs = (float*)malloc((nb + 1 + 300) * sizeof(float));
  uint16_t* la = (uint16_t*)calloc(nb + 1 + 300, sizeof(uint16_t));
  uint32_t* P = (uint32_t*)malloc((nb + 1 + 300) * sizeof(uint32_t));
  double mn = 1e300, mx = 0, ml = 0, margin_up;
  int skip_noop, imode = 0, ds, len, i;
  BinadeTab bt;
  size_t dirty_until = 0, guard_until = 0;
  bt.e = -1; bt.tie = 1; bt.base = 0;
  for (len = 3; len <= 258; len++)
    for (ds = 0; ds < 30; ds++) {
      const double c = length_extra_bits(len) + dist_symbol_extra_bits(ds) + st->ll_symbols[length_symbol(len)] + st->d_symbols[ds];
      if (c < mn) mn = c;
      if (c > mx) mx = c;
    }
  for (i = 0; i < 256; i++) if (st->ll_symbols[i] > ml) ml = st->ll_symbols[i];
  skip_noop = mincost <= mn;
  margin_up = 36.0 * ml + mx + 1.0;
  for (j = 1; j < nb + 1 + 300; j++) costs[j] = (float)ZO_LARGE;
  costs[0] = 0;
  g_ids.passes++;

  for (g = 0; g * 32 < nb; g++) {
    const size_t j0 = g * 32, jend = j0 + 32 < nb ? j0 + 32 : nb;
    int flagged = 0, want_int = 0, ring_dirty;
    for (j = j0; j < jend; j++)
      if (t->skipped[j] || (j > 0 && t->skipped[j - 1])) flagged = 1;   /* long-run shortcut: fp64 general path only */
    ring_dirty = j0 + 35 <= dirty_until;
    if (skip_noop && !flagged && j0 + 32 <= nb && j0 >= guard_until) {
      const double cj = costs[j0];
      int e;
      (void)frexp(cj, &e);
      e -= 1; /* cj in [2^e, 2^(e+1)) */
      if (cj > 0 && e >= 1 && e <= 100) {
        if (e != bt.e) {
          if (imode) { /* binade change: leave the integer representation with the old base first */
            for (k = j0 + 1; k <= j0 + 34; k++) {
              if (P[k] >= NOEDGE) costs[k] = (float)ZO_LARGE;
              else { costs[k] = bitsf((P[k] >> 7) + bt.base); la[k] = decode_tb(P[k], la[k]); }
            }
            imode = 0;
          }
          build_tab(&bt, e, st);
          if (bt.tie) g_ids.tie_binades++;
        }
        /* no lower margin: words only grow by positive edges, so it is enough that every pending value
         * carried into the window lies in the binade (checked below when the mode is entered) */
        if (!bt.tie && ldexp(1.0, e) > margin_up && cj < ldexp(1.0, e + 1) - margin_up)
          want_int = 1;
        if (want_int && !imode) {
          for (k = j0 + 1; k <= j0 + 34; k++)
            if (costs[k] < (float)ldexp(1.0, e)) want_int = 0;
          /* every finite ring entry (targets beyond the window) must lie in the binade: it joins as an integer */
          if (ring_dirty)
            for (k = j0 + 35; k <= dirty_until; k++)
              if (costs[k] != (float)ZO_LARGE && !(costs[k] >= (float)ldexp(1.0, e) && costs[k] < (float)ldexp(1.0, e + 1))) want_int = 0;
        }
      }
    }
    if (want_int) {
      uint32_t CJ;
      const double lo = ldexp(1.0, bt.e), hi = ldexp(1.0, bt.e + 1);
      if (!imode) { /* double -> int: pending targets j0+1 .. j0+34 */
        for (k = j0 + 1; k <= j0 + 34; k++) {
          const float v = costs[k];
          if (!(v < hi)) P[k] = INFP;
          else {
            if (v < lo) { g_ids.mismatches++; fprintf(stderr, "int model: pending below the binade at %zu\n", k); }
            P[k] = ((fbits(v) - bt.base) << 7) | (uint32_t)(64 + (la[k] >= 35 ? 0 : 35 - la[k]));
          }
        }
        imode = 1;
      }
      CJ = (fbits(costs[j0]) - bt.base) << 7;
      for (j = j0; j < j0 + 32; j++) {
        const size_t tj = j + 35;
        uint32_t X;
        /* literal */
        X = CJ + bt.lit[in[seg->instart + j]];
        if (P[j + 1] < X) X = P[j + 1];
        /* lengths 3..34: every lane relaxes, rows beyond the match length hold "no edge" */
        {
          const unsigned leng = t->length[j];
          uint32_t r = t->runoff[j];
          for (k = 3; k <= 34; k++) {
            uint32_t cand;
            if (k <= leng && j + k <= nb) {
              while ((t->runs[r] >> 16) < k) r++;
              cand = CJ + bt.te[dist_symbol((int)(t->runs[r] & 0xffff))][k];
            } else {
              cand = CJ + bt.te[30][k];
            }
            if (cand < P[j + k]) P[j + k] = cand;
          }
        }
        /* lengths 35..: straight into the ring (absolute float bits), first strict improvement wins */
        {
          const unsigned leng = t->length[j];
          const size_t kend = leng < nb - j ? leng : nb - j;
          const uint32_t cjb = bt.base + (CJ >> 7);
          uint32_t r = t->runoff[j];
          for (k = 35; k <= kend; k++) {
            uint32_t e2, rl;
            while ((t->runs[r] >> 16) < k) r++;
            e2 = bt.tl[length_symbol((int)k) - 257][dist_symbol((int)(t->runs[r] & 0xffff))];
            rl = cjb + (e2 >> 1);
            if (rl >= fbits((float)hi)) { g_ids.mismatches++; fprintf(stderr, "int model: long edge beyond the binade at %zu\n", j); }
            if (rl < fbits(costs[j + k]) || (rl == fbits(costs[j + k]) && (e2 & 1u))) { costs[j + k] = bitsf(rl); la[j + k] = (uint16_t)k; }
          }
          if (kend > 34 && j + kend > dirty_until) dirty_until = j + kend;
        }
        /* the lane that just handed over target j+3 takes on target j+35: the ring entry joins
         * (after this step's own length-35 edge has been pushed) */
        if (tj <= dirty_until) {
          const float v = costs[tj];
          if (v == (float)ZO_LARGE) P[tj] = INFP;
          else {
            if (!(v >= lo && v < hi)) { g_ids.mismatches++; fprintf(stderr, "int model: ring value outside the binade at %zu\n", tj); }
            P[tj] = ((fbits(v) - bt.base) << 7) | 64u;
          }
        } else {
          if (costs[tj] != (float)ZO_LARGE) { g_ids.mismatches++; fprintf(stderr, "int model: clean ring expected at %zu\n", tj); }
          P[tj] = INFP;
        }
        costs[j + 1] = bitsf((X >> 7) + bt.base);
        la[j + 1] = decode_tb(X, la[j + 1]);
        CJ = X & ~127u;
        g_ids.steps_int++;
      }
      g_ids.groups_int++;
      if (ring_dirty) g_ids.groups_int_ring++;
      continue;
    }
    if (imode) { /* int -> double: pending targets j0+1 .. j0+34 */
      for (k = j0 + 1; k <= j0 + 34; k++) {
        if (P[k] >= NOEDGE) costs[k] = (float)ZO_LARGE;
        else { costs[k] = bitsf((P[k] >> 7) + bt.base); la[k] = decode_tb(P[k], la[k]); }
      }
      imode = 0;
    }
    for (j = j0; j < jend; j++) { /* reference arithmetic (what the fp64 paths of the kernel do) */
      const size_t ii = seg->instart + j;
      size_t kend;
      unsigned leng;
      uint32_t r;
      double mc;
      if (t->skipped[j]) {
        costs[j + ZO_MAX_MATCH] = (float)(costs[j] + cost258);
        la[j + ZO_MAX_MATCH] = ZO_MAX_MATCH;
        if (j + ZO_MAX_MATCH > dirty_until) dirty_until = j + ZO_MAX_MATCH;
        guard_until = j + 600;
        continue;
      }
      {
        double nc = cost_stat(in[ii], 0, st) + costs[j];
        if (nc < costs[j + 1]) { costs[j + 1] = (float)nc; la[j + 1] = 1; }
      }
      leng = t->length[j];
      kend = leng < nb - j ? leng : nb - j;
      mc = mincost + costs[j];
      r = t->runoff[j];
      for (k = 3; k <= kend; k++) {
        double nc;
        while ((t->runs[r] >> 16) < k) r++;
        if (costs[j + k] <= mc) continue;
        nc = cost_stat((unsigned)k, t->runs[r] & 0xffff, st) + costs[j];
        if (nc < costs[j + k]) { costs[j + k] = (float)nc; la[j + k] = (uint16_t)k; }
      }
      if (kend > 34 && j + kend > dirty_until) dirty_until = j + kend;
    }
  }
  g_ids.steps_total += nb;
  for (j = 1; j <= nb; j++)
    if (costs[j] != costs_ref[j] || la[j] != la_ref[j]) {
      if (g_ids.mismatches < 5)
        fprintf(stderr, "int model mismatch at %zu/%zu: cost %.9g vs %.9g, length %u vs %u\n", j, nb,
                (double)costs[j], (double)costs_ref[j], la[j], la_ref[j]);
      g_ids.mismatches++;
    }
  free(costs); free(la); free(P);
}

/* Runs ZopfliLZ77Optimal (numiterations > 0) or ZopfliLZ77OptimalFixed (0) with the model checking
 * every DP pass; returns the number of mismatching positions, fills the counters. */
uint64_t zo_dp_int_check(const uint8_t* in, size_t instart, size_t inend, int numiterations, uint64_t* out8) {
  ZoStore s;
  memset(&g_ids, 0, sizeof(g_ids));
  zo_store_init(&s);
  zo_dp_observer = int_model;
  if (numiterations > 0) zo_lz77_optimal(in, instart, inend, numiterations, &s);
  else zo_lz77_optimal_fixed(in, instart, inend, &s);
  zo_dp_observer = NULL;
  zo_store_free(&s);
  if (out8) memcpy(out8, &g_ids, sizeof(g_ids));
  return g_ids.mismatches;
}

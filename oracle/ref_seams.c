/*
 * ref_seams.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin flat-argument wrappers around the UNMODIFIED reference (linked from
 * oracle/_ref/libzopfli_ref.so, headers included from /root/reference at build time) so that
 * Python/ctypes tests can read the three parity seams of SURVEY.md section 4 without restating
 * the reference's struct layouts in Python:
 *   seam 3  per-position ZopfliFindLongestMatch output (length, distance, sublen[3..length])
 *   seam 2  ZopfliLZ77Store contents out of ZopfliLZ77Optimal / ...Fixed / ...Greedy
 *   seam 1  is reached directly (ZopfliCompress / ZopfliDeflate / ZopfliDeflatePart are flat already)
 * No reference source is copied here; every function below only CALLS the reference.
 * Built only where /root/reference exists; the resulting .so travels to the GPU box.
 */
#include <stdlib.h>
#include <string.h>

#include "blocksplitter.h"
#include "deflate.h"
#include "hash.h"
#include "katajainen.h"
#include "lz77.h"
#include "squeeze.h"
#include "tree.h"
#include "zopfli.h"

void OptimizeHuffmanForRle(int length, size_t* counts); /* deflate.c:434, non-static, undeclared */

/* Runs the reference's hash update loop (squeeze.c:237-248) over [instart, inend) and records
 * ZopfliFindLongestMatch(limit=258, sublen) with the cache disabled (lmc == NULL).
 * out_sublen is [n][259]; entries outside 3..length are left 0. Also dumps the hash state. */
void ref_match_table(const unsigned char* in, size_t instart, size_t inend,
                     unsigned short* out_len, unsigned short* out_dist,
                     unsigned short* out_sublen, unsigned short* out_same, int* out_hv,
                     int* out_hv2) {
  ZopfliOptions opt;
  ZopfliBlockState s;
  ZopfliHash hash;
  ZopfliHash* h = &hash;
  size_t i, windowstart = instart > ZOPFLI_WINDOW_SIZE ? instart - ZOPFLI_WINDOW_SIZE : 0;
  unsigned short sublen[259];
  ZopfliInitOptions(&opt);
  ZopfliInitBlockState(&opt, instart, inend, 0, &s);
  ZopfliAllocHash(ZOPFLI_WINDOW_SIZE, h);
  ZopfliResetHash(ZOPFLI_WINDOW_SIZE, h);
  ZopfliWarmupHash(in, windowstart, inend, h);
  for (i = windowstart; i < instart; i++) ZopfliUpdateHash(in, i, inend, h);
  for (i = instart; i < inend; i++) {
    size_t j = i - instart, k;
    unsigned short dist, leng;
    ZopfliUpdateHash(in, i, inend, h);
    memset(sublen, 0, sizeof(sublen));
    ZopfliFindLongestMatch(&s, h, in, i, inend, ZOPFLI_MAX_MATCH, sublen, &dist, &leng);
    out_len[j] = leng;
    out_dist[j] = dist;
    if (out_sublen) {
      for (k = 0; k < 259; k++) out_sublen[j * 259 + k] = (k >= 3 && k <= leng) ? sublen[k] : 0;
    }
    if (out_same) out_same[j] = h->same[i & ZOPFLI_WINDOW_MASK];
    if (out_hv) out_hv[j] = h->hashval[i & ZOPFLI_WINDOW_MASK];
    if (out_hv2) out_hv2[j] = h->hashval2[i & ZOPFLI_WINDOW_MASK];
  }
  ZopfliCleanHash(h);
  ZopfliCleanBlockState(&s);
}

/* Limited walk, sublen == NULL (FollowPath's call shape, squeeze.c:367). */
void ref_limited_match(const unsigned char* in, size_t instart, size_t inend,
                       const unsigned short* limits, unsigned short* out_len,
                       unsigned short* out_dist) {
  ZopfliOptions opt;
  ZopfliBlockState s;
  ZopfliHash hash;
  ZopfliHash* h = &hash;
  size_t i, windowstart = instart > ZOPFLI_WINDOW_SIZE ? instart - ZOPFLI_WINDOW_SIZE : 0;
  ZopfliInitOptions(&opt);
  ZopfliInitBlockState(&opt, instart, inend, 0, &s);
  ZopfliAllocHash(ZOPFLI_WINDOW_SIZE, h);
  ZopfliResetHash(ZOPFLI_WINDOW_SIZE, h);
  ZopfliWarmupHash(in, windowstart, inend, h);
  for (i = windowstart; i < instart; i++) ZopfliUpdateHash(in, i, inend, h);
  for (i = instart; i < inend; i++) {
    size_t j = i - instart;
    ZopfliUpdateHash(in, i, inend, h);
    if (limits[j] >= 3) {
      ZopfliFindLongestMatch(&s, h, in, i, inend, limits[j], 0, &out_dist[j], &out_len[j]);
    } else {
      out_len[j] = 0;
      out_dist[j] = 0;
    }
  }
  ZopfliCleanHash(h);
  ZopfliCleanBlockState(&s);
}

static size_t export_store(const ZopfliLZ77Store* st, unsigned short* litlens,
                           unsigned short* dists, size_t cap) {
  size_t i;
  for (i = 0; i < st->size && i < cap; i++) {
    litlens[i] = st->litlens[i];
    dists[i] = st->dists[i];
  }
  return st->size;
}

/* mode 0: ZopfliLZ77Optimal(numiterations), 1: ZopfliLZ77OptimalFixed, 2: ZopfliLZ77Greedy
 * with the block-level cache (as squeeze.c:481), 3: greedy without cache (blocksplitter.c:296). */
size_t ref_lz77(const unsigned char* in, size_t instart, size_t inend, int mode, int numiterations,
                unsigned short* litlens, unsigned short* dists, size_t cap) {
  ZopfliOptions opt;
  ZopfliBlockState s;
  ZopfliLZ77Store store;
  size_t n;
  ZopfliInitOptions(&opt);
  ZopfliInitLZ77Store(in, &store);
  ZopfliInitBlockState(&opt, instart, inend, mode != 3, &s);
  if (mode == 0) {
    ZopfliLZ77Optimal(&s, in, instart, inend, numiterations, &store);
  } else if (mode == 1) {
    ZopfliLZ77OptimalFixed(&s, in, instart, inend, &store);
  } else {
    ZopfliHash hash;
    ZopfliAllocHash(ZOPFLI_WINDOW_SIZE, &hash);
    ZopfliLZ77Greedy(&s, in, instart, inend, &store, &hash);
    ZopfliCleanHash(&hash);
  }
  n = export_store(&store, litlens, dists, cap);
  ZopfliCleanBlockState(&s);
  ZopfliCleanLZ77Store(&store);
  return n;
}

static void import_store(const unsigned char* in, const unsigned short* litlens,
                         const unsigned short* dists, size_t n, size_t instart,
                         ZopfliLZ77Store* store) {
  size_t i, pos = instart;
  ZopfliInitLZ77Store(in, store);
  for (i = 0; i < n; i++) {
    ZopfliStoreLitLenDist(litlens[i], dists[i], pos, store);
    pos += dists[i] == 0 ? 1 : litlens[i];
  }
}

/* ZopfliCalculateBlockSize on an explicit symbol list. */
double ref_block_size(const unsigned char* in, const unsigned short* litlens,
                      const unsigned short* dists, size_t n, size_t lstart, size_t lend,
                      int btype) {
  ZopfliLZ77Store store;
  double r;
  import_store(in, litlens, dists, n, 0, &store);
  r = btype < 0 ? ZopfliCalculateBlockSizeAutoType(&store, lstart, lend)
                : ZopfliCalculateBlockSize(&store, lstart, lend, btype);
  ZopfliCleanLZ77Store(&store);
  return r;
}

/* ZopfliBlockSplitLZ77 on an explicit symbol list; returns npoints. */
size_t ref_block_split_lz77(const unsigned char* in, const unsigned short* litlens,
                            const unsigned short* dists, size_t n, size_t maxblocks,
                            size_t* out_points, size_t cap) {
  ZopfliOptions opt;
  ZopfliLZ77Store store;
  size_t* pts = 0;
  size_t np = 0, i;
  ZopfliInitOptions(&opt);
  import_store(in, litlens, dists, n, 0, &store);
  ZopfliBlockSplitLZ77(&opt, &store, maxblocks, &pts, &np);
  for (i = 0; i < np && i < cap; i++) out_points[i] = pts[i];
  free(pts);
  ZopfliCleanLZ77Store(&store);
  return np;
}

/* ZopfliBlockSplit (byte positions); returns npoints. */
size_t ref_block_split(const unsigned char* in, size_t instart, size_t inend, size_t maxblocks,
                       size_t* out_points, size_t cap) {
  ZopfliOptions opt;
  size_t* pts = 0;
  size_t np = 0, i;
  ZopfliInitOptions(&opt);
  ZopfliBlockSplit(&opt, in, instart, inend, maxblocks, &pts, &np);
  for (i = 0; i < np && i < cap; i++) out_points[i] = pts[i];
  free(pts);
  return np;
}

int ref_length_limited(const size_t* freq, int n, int maxbits, unsigned* out) {
  return ZopfliLengthLimitedCodeLengths(freq, n, maxbits, out);
}

void ref_entropy(const size_t* count, size_t n, double* out) { ZopfliCalculateEntropy(count, n, out); }

void ref_optimize_rle(int length, size_t* counts) { OptimizeHuffmanForRle(length, counts); }

/*
 * zopfli.h -- public C API of zopfli-b200, ABI-compatible with google/zopfli's libzopfli.so.1.
 *
 * Every declaration below replaces the reference interface cited next to it (paths relative to
 * /root/reference/src/zopfli/).  Same names, argument order, struct layout and ownership rules,
 * so existing callers (the zopfli CLI zopfli_bin.c:112, zopflipng's CustomPNGDeflate
 * zopflipng_lib.cc:60, the cgo wrapper go/zopfli/zopfli.go:46) relink unchanged.
 *
 * Ownership (util.h:124-155, zopfli.h:82-83): results are APPENDED to *out; *out is grown with
 * malloc/realloc so that its capacity is always the smallest power of two >= *outsize, the
 * caller frees it with free(), and may keep appending with the same growth rule.
 * Errors: none are returned (all functions are void, like the reference); allocation or CUDA
 * failures terminate the process with a message.  There is no CPU fallback.
 */
#ifndef ZOPFLI_B200_ZOPFLI_H_
#define ZOPFLI_B200_ZOPFLI_H_

#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* zopfli.h:33-64 -- six ints, 24 bytes */
typedef struct ZopfliOptions {
  int verbose;            /* print progress to stderr */
  int verbose_more;       /* print more */
  int numiterations;      /* optimal-parse iterations per block (default 15) */
  int blocksplitting;     /* 1: choose block boundaries by cost (default) */
  int blocksplittinglast; /* unused, kept for layout compatibility */
  int blocksplittingmax;  /* max blocks per 1,000,000-byte master block, 0 = unlimited (default 15) */
} ZopfliOptions;

/* zopfli.h:67, util.c:28-35 */
void ZopfliInitOptions(ZopfliOptions* options);

/* zopfli.h:70-74 */
typedef enum {
  ZOPFLI_FORMAT_GZIP,
  ZOPFLI_FORMAT_ZLIB,
  ZOPFLI_FORMAT_DEFLATE
} ZopfliFormat;

/* zopfli.h:86-88, zopfli_lib.c:28-42 */
void ZopfliCompress(const ZopfliOptions* options, ZopfliFormat output_type,
                    const unsigned char* in, size_t insize,
                    unsigned char** out, size_t* outsize);

/* deflate.h:58-60, deflate.c:908-931.  btype 0 stored, 1 fixed, 2 best of all; *bp is the
 * number of bits already used in the last output byte (0 on the first call). */
void ZopfliDeflate(const ZopfliOptions* options, int btype, int final,
                   const unsigned char* in, size_t insize,
                   unsigned char* bp, unsigned char** out, size_t* outsize);

/* deflate.h:67-70, deflate.c:811-906.  Bytes before instart serve as the LZ77 dictionary. */
void ZopfliDeflatePart(const ZopfliOptions* options, int btype, int final,
                       const unsigned char* in, size_t instart, size_t inend,
                       unsigned char* bp, unsigned char** out, size_t* outsize);

/* lz77.h:44-62 -- the output type of the hot path, as callers of deflate.h see it: symbol i is a
 * literal litlens[i] when dists[i] == 0, else a match of length litlens[i] at distance dists[i].
 * Only litlens, dists and size are read by the two functions below; the remaining members keep
 * the reference's layout (pointers it maintains for its own bookkeeping). */
typedef struct ZopfliLZ77Store {
  unsigned short* litlens;
  unsigned short* dists;
  size_t size;
  const unsigned char* data;
  size_t* pos;
  unsigned short* ll_symbol;
  unsigned short* d_symbol;
  size_t* ll_counts;
  size_t* d_counts;
} ZopfliLZ77Store;

/* deflate.h:79-80, deflate.c:584-608.  Exact size in bits of symbols [lstart, lend) as one block of
 * type btype (0 stored, 1 fixed, 2 dynamic).  Host arithmetic (integers); no GPU involved. */
double ZopfliCalculateBlockSize(const ZopfliLZ77Store* lz77, size_t lstart, size_t lend, int btype);

/* deflate.h:85-86, deflate.c:610-621.  Minimum over the three block types (fixed only considered
 * for stores of at most 1000 symbols, as in the reference). */
double ZopfliCalculateBlockSizeAutoType(const ZopfliLZ77Store* lz77, size_t lstart, size_t lend);

/* gzip_container.h:42-44, gzip_container.c:84-124 */
void ZopfliGzipCompress(const ZopfliOptions* options,
                        const unsigned char* in, size_t insize,
                        unsigned char** out, size_t* outsize);

/* zlib_container.h:42-44, zlib_container.c:50-79 */
void ZopfliZlibCompress(const ZopfliOptions* options,
                        const unsigned char* in, size_t insize,
                        unsigned char** out, size_t* outsize);

#ifdef __cplusplus
}
#endif
#endif /* ZOPFLI_B200_ZOPFLI_H_ */

/*
 * zopfli_b200.h -- additional C-ABI entry points of libzopfli.so.1 (zopfli-b200 build).
 *
 * These expose the hot-path seams the reference keeps internal, for parity tests, the bench
 * harness and multi-GPU sharding.  Plain pointers and sizes only.  Each cites the reference
 * function whose result it returns (paths relative to /root/reference/src/zopfli/).
 * All functions return 0 on success; CUDA failures abort (no CPU fallback exists).
 */
#ifndef ZOPFLI_B200_EXTRA_H_
#define ZOPFLI_B200_EXTRA_H_

#include <stddef.h>
#include <stdint.h>

#include "zopfli.h"

#ifdef __cplusplus
extern "C" {
#endif

/* LZ77 parse of in[instart, inend) with in[max(0,instart-32768), instart) as dictionary.
 * mode 0: ZopfliLZ77Optimal (squeeze.h:41-44) with `numiterations`;
 * mode 1: ZopfliLZ77OptimalFixed (squeeze.h:56-59);
 * mode 2: ZopfliLZ77Greedy (lz77.h:136-138).
 * Writes up to cap symbols (dists[i]==0 => literal litlens[i]) and the symbol count. */
int ZopfliB200LZ77(const unsigned char* in, size_t insize, size_t instart, size_t inend, int mode,
                   int numiterations, unsigned short* litlens, unsigned short* dists, size_t cap,
                   size_t* size);

/* Batched form: n ranges of one input parsed in one device pass.  off[i]/cnt[i] locate the
 * symbols of range i inside litlens/dists (total capacity cap); cost[i] (may be NULL) receives
 * the exact dynamic-block bit size of the returned parse (ZopfliCalculateBlockSize btype 2,
 * deflate.h:79-80) for mode 0. */
int ZopfliB200LZ77Batch(const unsigned char* in, size_t insize, size_t n, const size_t* instart,
                        const size_t* inend, int mode, int numiterations, unsigned short* litlens,
                        unsigned short* dists, size_t cap, size_t* off, size_t* cnt,
                        uint64_t* cost);

/* ZopfliFindLongestMatch (lz77.h:112-115) for every position of [instart, inend), limit 258,
 * cache disabled: length[j], distance[j], sublen[j*259 + k] for k in 3..length[j] (0 elsewhere),
 * plus the per-position hash state same / hashval / hashval2 (hash.h:29-47). Any output may be NULL. */
int ZopfliB200MatchTable(const unsigned char* in, size_t insize, size_t instart, size_t inend,
                         unsigned short* length, unsigned short* distance, unsigned short* sublen,
                         unsigned short* same, unsigned short* hashval, unsigned short* hashval2);

/* ZopfliCalculateBlockSize(btype 2) as a function of the 288+32 symbol histogram
 * (deflate.c:569-608), evaluated on the device (where=1) or by the host code (where=0). */
uint64_t ZopfliB200DynamicBlockBits(const uint32_t* hist320, int where);

/* ZopfliCalculateBlockSizeAutoType (deflate.h:85-86) of nreq symbol ranges of one LZ77 store,
 * evaluated on the device (k_split_eval) -- the splitter's cost oracle. */
int ZopfliB200DeviceAutoTypeBits(const unsigned short* litlens, const unsigned short* dists, size_t n,
                                 size_t nreq, const size_t* lstart, const size_t* lend, uint64_t* out);

/* Host logic seams (no GPU needed): ZopfliBlockSplitLZ77 (blocksplitter.h:42-45) and
 * ZopfliCalculateBlockSize[AutoType] (deflate.h:79-86) over an explicit symbol list whose first
 * symbol starts at byte 0 of `in`. */
size_t ZopfliB200HostBlockSplitLZ77(const unsigned char* in, const unsigned short* litlens,
                                    const unsigned short* dists, size_t n, size_t maxblocks,
                                    size_t* points, size_t cap);
/* The same search as the product runs it (batched_split.hpp: every FindMinimum of every current
 * block in lockstep rounds, `budget` probes per round bound the speculation depth), costs from the
 * host estimators, for several stores at once: store s = symbols [off[s], off[s]+size[s]).
 * points[s*cap ..] receives store s's split points, npoints[s] their number. */
void ZopfliB200HostBatchedSplit(const unsigned short* litlens, const unsigned short* dists, size_t nstores,
                                const size_t* off, const size_t* size, size_t maxblocks, size_t budget,
                                size_t* points, size_t cap, size_t* npoints);
double ZopfliB200HostBlockSize(const unsigned char* in, const unsigned short* litlens,
                               const unsigned short* dists, size_t n, size_t lstart, size_t lend,
                               int btype /* 0,1,2 or -1 for AutoType */);
/* AddLZ77Block (deflate.c:682-745) for btype 1/2 at bit offset 0: emits one block into out
 * (capacity cap bytes), returns the number of bits. */
uint64_t ZopfliB200HostEmitBlock(const unsigned char* in, const unsigned short* litlens,
                                 const unsigned short* dists, size_t n, size_t lstart, size_t lend,
                                 int btype, int final, unsigned char* out, size_t cap);
/* OptimizeHuffmanForRle (deflate.c:434-518) as restated for host and device, in place, n <= 288. */
void ZopfliB200HostOptimizeRle(uint32_t* counts, int n);
/* ZopfliLengthLimitedCodeLengths (katajainen.h:35-36) as restated for host and device. */
int ZopfliB200HostLengthLimited(const uint32_t* freq, int n, int maxbits, unsigned* bitlengths);

/* Sharding: compress in[start, end) -- cut into master blocks of 1,000,000 bytes
 * counted from `start` (util.h:60; deflate.c:912-924), bytes before `start` being the LZ77
 * dictionary (the 32 KiB halo of a shard) -- into a position-independent SPAN: a sequence of
 *   u8 kind (0 bits, 1 stored) | u8 final | u64 nbits-or-nbytes | payload bytes
 * Compressed blocks are encoded at bit offset 0; stored blocks carry their raw bytes because
 * their padding depends on the final bit offset (deflate.c:643-649).  `final` marks the last
 * block of the range.  If dev_in is non-NULL it is a device pointer holding the same bytes as
 * `in` (16-byte aligned, readable 16 bytes past insize) and no host-to-device copy is made. */
int ZopfliB200DeflateSpan(const ZopfliOptions* options, const unsigned char* in, size_t insize,
                          const unsigned char* dev_in, size_t start, size_t end, int final,
                          unsigned char** span, size_t* spansize);
/* Splices a span onto a stream whose last byte has *bp bits in use (bit-offset scan).  The span is
 * validated first (record sizes against spansize); returns 1 and leaves the output untouched if
 * it is truncated or malformed. */
int ZopfliB200AppendSpan(const unsigned char* span, size_t spansize, unsigned char* bp,
                         unsigned char** out, size_t* outsize);
/* Bit offset, relative to the first bit the call wrote, at which each master block (deflate.c:
 * 908-931) of the most recent ZopfliDeflate / ZopfliCompress call starts, plus the end offset as last
 * entry.  Returns the number of entries (master blocks + 1).  Lets a checker compare sampled master
 * blocks of one big stream with the reference's ZopfliDeflatePart of the same range. */
size_t ZopfliB200LastMasterBitOffsets(uint64_t* offsets, size_t cap);
/* ZopfliCompress with the input already resident on the device (bench `value` leg). */
void ZopfliB200CompressDevice(const ZopfliOptions* options, ZopfliFormat output_type,
                              const unsigned char* in, size_t insize, const unsigned char* dev_in,
                              unsigned char** out, size_t* outsize);

/* ---- several GPUs, one stream ----
 * The master blocks of one input (the reference's independent unit, deflate.c:908-931) are sharded
 * over the GPUs of one box; NCCL scatters the byte ranges (+ 32 KiB dictionary) from rank 0's GPU and
 * gathers the compressed bits straight into their final bit positions on rank 0.
 *  - one process, N GPUs: set ZOPFLI_B200_GPUS=N; ZopfliCompress / ZopfliDeflate(btype 2) use it
 *    transparently (ncclCommInitAll, one host thread per GPU);
 *  - one process per GPU (e.g. torchrun): rank 0 calls ZopfliB200DistUniqueId, the id is given to every
 *    rank out of band, every rank calls ZopfliB200DistInit once (device = ZOPFLI_B200_DEVICE / LOCAL_RANK)
 *    and then ZopfliB200DistCompress COLLECTIVELY with the same options, format and insize.  `in` is read
 *    and *out / *outsize are written on rank 0 only (same ownership rule as ZopfliCompress).
 * All return 0 on success. */
#define ZOPFLI_B200_UNIQUE_ID_BYTES 128
int ZopfliB200DistUniqueId(unsigned char* id128);
int ZopfliB200DistInit(int rank, int world, const unsigned char* id128);
#define ZOPFLI_B200_DIST_STAGED 1 /* flags: the shards of this very input are still on the GPUs from the
                                    previous ZopfliB200DistCompress call -- skip the H2D copy and the scatter
                                    (the bench's "input resident in HBM" leg); `in` is still read for the checksum */
int ZopfliB200DistCompress(const ZopfliOptions* options, ZopfliFormat output_type, const unsigned char* in,
                           size_t insize, int flags, unsigned char** out, size_t* outsize);
void ZopfliB200DistFinalize(void);
/* The placement arithmetic of the multi-GPU path as pure functions (csrc/dist_layout.hpp; no GPU needed):
 * the byte range [a, b) rank `rank` owns and the start `base` of its device copy (shard + dictionary);
 * and, from len8[world][8] (bits of every rank's blocks for each start phase), the absolute bit offsets
 * start[0..world] of the ranks in the one stream. */
void ZopfliB200DistShard(size_t insize, int world, int rank, size_t* a, size_t* b, size_t* base);
void ZopfliB200DistPlacement(const uint64_t* len8, int world, unsigned phase0, uint64_t* start);

/* CRC-32 of the gzip trailer (gzip_container.c:27-81) and its combination across shards. */
uint32_t ZopfliB200Crc32(const unsigned char* data, size_t size);
uint32_t ZopfliB200Crc32Combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
/* Adler-32 of the zlib trailer (zlib_container.c:29-48), threaded with an exact combination. */
uint32_t ZopfliB200Adler32(const unsigned char* data, size_t size);

/* Engine control / introspection. */
typedef struct ZopfliB200Stats {
  double ms_same, ms_keys, ms_scan, ms_scatter, ms_match, ms_greedy, ms_iterate, ms_pack, ms_h2d, ms_d2h;
  double ms_host_split, ms_host_emit, ms_host_other, ms_total;
  uint64_t launches, match_positions, iterate_positions, iterate_steps, h2d_bytes, d2h_bytes;
  uint64_t cyc_sum[6], cyc_max[6], max_block_positions; /* k_iterate phase cycles, see engine.hpp */
  double ms_split; uint64_t split_evals, split_rounds;  /* device split-cost service */
  uint64_t iterate_launches;                             /* k_iterate launches (ms_iterate is the sum of their durations) */
  uint64_t int_steps;                                    /* forward-DP steps that ran in the integer window (see iterate.cuh) */
  uint64_t dp_cyc_sum[5], dp_cnt_sum[6];                 /* DP cycles / groups of 32 steps by kind: integer window, fp64 magic, fp64 plain, */
  uint64_t dp_cyc_max[5], dp_cnt_max[6];                 /* fp64 ring-joining, general; cnt[5] = positions in the per-step loop; sum over blocks / critical block */
} ZopfliB200Stats;
void ZopfliB200GetStats(ZopfliB200Stats* out);
void ZopfliB200ResetStats(void);
void ZopfliB200SetStream(void* cuda_stream);
int ZopfliB200Device(void);
const char* ZopfliB200Version(void);

#ifdef __cplusplus
}
#endif
#endif

// Split-cost evaluation on the device: batched ZopfliCalculateBlockSizeAutoType
// (/root/reference/src/zopfli/deflate.c:610-621) over many symbol ranges of many LZ77 stores --
// the inner loop of ZopfliBlockSplitLZ77's FindMinimum (blocksplitter.c:43-128).
//
//   k_split_prep_sym   per symbol: alphabet symbols and byte length            lz77.c:98-149
//   k_split_prep_pos   per store: byte positions (exclusive scan of lengths)   lz77.h:49
//   k_split_prep_snap  per 256-symbol chunk: histogram; then prefix over chunks lz77.c:107-124 (role)
//   k_split_eval       one warp per (store, lstart, lend): range histogram from the snapshots,
//                      stored / fixed / dynamic bit costs, AutoType minimum
#pragma once
#include "iterate.cuh"

namespace zb {

constexpr uint32_t kSnap = 256;

struct SplitStoreDesc {
  uint64_t sym_off;    // offset of this store's symbols in the flat arrays
  uint64_t snap_off;   // offset (in snapshots of 320 u32) of this store's snapshot table
  uint64_t pos_off;    // offset of this store's n+1 byte positions in `pos`
  uint32_t n;          // symbols
  uint32_t nsnap;      // n / kSnap + 1
};

struct SplitEval { uint32_t store, lstart, lend, pad; };

struct SplitBatch {
  const uint16_t* ll;
  const uint16_t* d;
  uint16_t* llsym;
  uint8_t* dsym;
  uint32_t* pos;         // byte offset of each symbol relative to its store start; n+1 entries at pos_off
  uint32_t* snaps;       // prefix histograms at multiples of kSnap
  const SplitStoreDesc* stores;
};

__global__ void k_split_prep_sym(SplitBatch b, uint32_t store) {
  const SplitStoreDesc sd = b.stores[store];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sd.n) return;
  const uint32_t l = b.ll[sd.sym_off + i], dd = b.d[sd.sym_off + i];
  b.llsym[sd.sym_off + i] = (uint16_t)(dd == 0 ? l : (uint32_t)length_symbol((int)l));
  b.dsym[sd.sym_off + i] = (uint8_t)(dd == 0 ? 0 : dist_symbol((int)dd));
}

// one CTA per store: byte positions by a chunked scan (1024 threads x items)
__global__ void k_split_prep_pos(SplitBatch b) {
  const uint32_t store = blockIdx.x;
  const SplitStoreDesc sd = b.stores[store];
  __shared__ uint32_t part[1024];
  __shared__ uint32_t carry_s;
  uint32_t* pos = b.pos + sd.pos_off;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < sd.n; base += 1024 * 8) {
    uint32_t v[8], sum = 0;
    const uint32_t i0 = base + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t i = i0 + k;
      v[k] = i < sd.n ? (b.d[sd.sym_off + i] == 0 ? 1u : (uint32_t)b.ll[sd.sym_off + i]) : 0u;
      sum += v[k];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t t = threadIdx.x >= (uint32_t)off ? part[threadIdx.x - off] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t run = carry_s + part[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t i = i0 + k;
      if (i < sd.n) pos[i] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) pos[sd.n] = carry_s;
}

// chunk histograms: one warp per 256-symbol chunk (work list = (store, chunk))
struct SnapWork { uint32_t store, chunk; };
__global__ void k_split_prep_snap(SplitBatch b, const SnapWork* __restrict__ work, uint32_t nwork) {
  __shared__ uint32_t h[8][320];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wi = blockIdx.x * 8 + warp;
  if (wi >= nwork) return;
  const SnapWork w = work[wi];
  const SplitStoreDesc sd = b.stores[w.store];
  for (int i = lane; i < 320; i += 32) h[warp][i] = 0;
  __syncwarp();
  const uint32_t i0 = w.chunk * kSnap;
  for (uint32_t t = lane; t < kSnap; t += 32) {
    const uint32_t i = i0 + t;
    if (i < sd.n) {
      atomicAdd(&h[warp][b.llsym[sd.sym_off + i]], 1u);
      if (b.d[sd.sym_off + i]) atomicAdd(&h[warp][288 + b.dsym[sd.sym_off + i]], 1u);
    }
  }
  __syncwarp();
  // chunk c's own histogram goes to slot c+1; the prefix pass turns slots into histograms of [0, c*256)
  uint32_t* dst = b.snaps + (sd.snap_off + w.chunk + 1) * 320;
  if (w.chunk + 1 < sd.nsnap)
    for (int i = lane; i < 320; i += 32) dst[i] = h[warp][i];
}
// prefix over chunks: one CTA per store, 320 threads (one per bin)
__global__ void k_split_prep_prefix(SplitBatch b) {
  const SplitStoreDesc sd = b.stores[blockIdx.x];
  const uint32_t bin = threadIdx.x;
  if (bin >= 320) return;
  uint32_t* s0 = b.snaps + sd.snap_off * 320;
  uint32_t run = 0;
  s0[bin] = 0;
  for (uint32_t c = 1; c < sd.nsnap; c++) {
    run += s0[(uint64_t)c * 320 + bin];
    s0[(uint64_t)c * 320 + bin] = run;
  }
}

struct SplitSmem {
  uint32_t hist[320];
  CostStage cs;
};

constexpr int kEvalWarps = 4;

__global__ void __launch_bounds__(kEvalWarps * 32) k_split_eval(SplitBatch b, const SplitEval* __restrict__ evals,
                                                               uint32_t nevals, uint64_t* __restrict__ out) {
  __shared__ SplitSmem sm[kEvalWarps];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ei = blockIdx.x * kEvalWarps + warp;
  if (ei >= nevals) return;
  SplitSmem& s = sm[warp];
  const SplitEval e = evals[ei];
  const SplitStoreDesc sd = b.stores[e.store];
  const uint16_t* llsym = b.llsym + sd.sym_off;
  const uint8_t* dsym = b.dsym + sd.sym_off;
  const uint16_t* dd = b.d + sd.sym_off;
  const uint32_t* pos = b.pos + sd.pos_off;
  // ---- range histogram (ZopfliLZ77GetHistogram lz77.c:189-217) ----
  for (int i = lane; i < 320; i += 32) s.hist[i] = 0;
  __syncwarp();
  if (e.lend - e.lstart < 2 * kSnap) {
    for (uint32_t i = e.lstart + lane; i < e.lend; i += 32) {
      atomicAdd(&s.hist[llsym[i]], 1u);
      if (dd[i]) atomicAdd(&s.hist[288 + dsym[i]], 1u);
    }
  } else {
    const uint32_t ka = e.lstart / kSnap, kb = e.lend / kSnap;
    const uint32_t* sa = b.snaps + (sd.snap_off + ka) * 320;
    const uint32_t* sb = b.snaps + (sd.snap_off + kb) * 320;
    for (int i = lane; i < 320; i += 32) s.hist[i] = sb[i] - sa[i];
    __syncwarp();
    for (uint32_t i = kb * kSnap + lane; i < e.lend; i += 32) {   // + [kb*256, lend)
      atomicAdd(&s.hist[llsym[i]], 1u);
      if (dd[i]) atomicAdd(&s.hist[288 + dsym[i]], 1u);
    }
    for (uint32_t i = ka * kSnap + lane; i < e.lstart; i += 32) {  // - [ka*256, lstart)
      atomicSub(&s.hist[llsym[i]], 1u);
      if (dd[i]) atomicSub(&s.hist[288 + dsym[i]], 1u);
    }
  }
  __syncwarp();
  // ---- costs (deflate.c:584-621) ----
  uint64_t nbytes = 0;
  if (e.lend > e.lstart) nbytes = pos[e.lend] - pos[e.lstart];  // ZopfliLZ77GetByteRange lz77.c:160-166
  const uint64_t unc = stored_bits(nbytes);
  uint64_t fixed = unc;
  if (sd.n <= 1000) {  // deflate.c:615: lz77->size of the WHOLE store
    uint64_t f = 0;
    for (int i = lane; i < 320; i += 32) {
      const uint32_t c = s.hist[i];
      if (i < 288) {
        if (i < 256) f += (uint64_t)fixed_ll_length(i) * c;
        else if (i >= 257 && i < 286) f += (uint64_t)(fixed_ll_length(i) + length_symbol_extra_bits(i)) * c;
      } else if (i - 288 < 30) {
        f += (uint64_t)(5 + dist_symbol_extra_bits(i - 288)) * c;
      }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) f += __shfl_xor_sync(0xffffffffu, f, dlt);
    fixed = 3 + f + 7;
  }
  if (lane == 0) s.hist[256] = 1;  // deflate.c:575
  __syncwarp();
  const uint64_t dyn = warp_dynamic_bits(s.hist, s.cs, lane);
  if (lane == 0) out[ei] = (unc < fixed && unc < dyn) ? unc : (fixed < dyn ? fixed : dyn);
}

}  // namespace zb

// Multi-GPU entry points used by the C-ABI layer (api.cpp); implemented in dist.cpp over NCCL.
#pragma once
#include <stddef.h>

#include "../../include/zopfli.h"

namespace zb {

// one process, several GPUs: ZOPFLI_B200_GPUS (default 1)
int dist_local_gpus();
// ZopfliDeflate(btype 2) over `ngpus` GPUs of this process; false if NCCL / the devices are unavailable
bool dist_local_deflate(int ngpus, const ZopfliOptions* opt, int final, const unsigned char* in, size_t insize, unsigned char* bp,
                        unsigned char** out, size_t* outsize);

// one process per GPU (ZopfliB200DistInit): collective ZopfliDeflate(btype 2); in/bp/out/outsize on rank 0 only
bool dist_rank_ready();
int dist_rank();
// staged: the shards of the same input are still on the devices from the previous call (skip H2D + scatter)
void dist_rank_deflate(const ZopfliOptions* opt, int final, const unsigned char* in, size_t insize, unsigned char* bp,
                       unsigned char** out, size_t* outsize, bool staged);

}  // namespace zb

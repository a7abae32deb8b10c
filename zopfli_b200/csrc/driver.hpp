// Internal interface between the C-ABI layer (api.cpp) and the batched deflate driver.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <utility>
#include <vector>

#include "../../include/zopfli.h"
#include "emit.hpp"

namespace zb {

// One deflate block of the output: either an already encoded bit string that starts at bit 0,
// or a stored block that is written at splice time because its padding depends on the absolute
// bit offset (/root/reference/src/zopfli/deflate.c:643-649).
struct Piece {
  bool stored = false;
  BitString bits;
  size_t instart = 0, inend = 0;
  bool final = false;
  uint32_t unit = 0;   // index of the unit (master block) the piece belongs to
};

class Engine;

struct HostTimes { double split = 0, emit = 0, other = 0; };
extern HostTimes g_host_times;

// Compresses the given byte ranges ("units": master blocks, or the single range of a
// ZopfliDeflatePart call) of `in`; the engine must already hold in[in_base ...).
void deflate_units(Engine& eng, const ZopfliOptions* opt, int btype, bool final_last, const unsigned char* in,
                   const std::vector<std::pair<size_t, size_t>>& units, size_t in_base,
                   std::vector<Piece>& pieces);

// unit_bits (optional): receives, relative to the first bit written by this call, the bit offset at
// which every unit's first piece starts, plus the end offset as last entry
void splice_pieces(const std::vector<Piece>& pieces, const unsigned char* in, unsigned char* bp,
                   unsigned char** out, size_t* outsize, std::vector<uint64_t>* unit_bits = nullptr);

unsigned char* append_reserve(size_t n, unsigned char** out, size_t* outsize);
void append_bytes(const unsigned char* src, size_t n, unsigned char** out, size_t* outsize);

}  // namespace zb

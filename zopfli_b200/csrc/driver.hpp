// Internal interface between the C-ABI layer (api.cpp) and the batched deflate driver.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <utility>
#include <vector>

#include "../../include/zopfli.h"
#include "engine.hpp"

namespace zb {

// One deflate block of the output stream.  Compressed blocks live on the device as (store buffer,
// symbol range, plan); their exact bit size is known, so the host can place every block before a bit
// is written.  Stored blocks name input bytes; their size depends on the bit offset (deflate.c:643-649).
struct Piece {
  uint8_t type = 2;          // 0 stored, 1 fixed, 2 dynamic
  bool final = false;
  uint8_t buf = 0;           // Engine::StoreBuf
  uint32_t unit = 0;         // index of the unit (master block) the piece belongs to
  uint32_t n = 0;            // symbols
  uint64_t off = 0;          // first symbol in `buf`
  uint64_t plan = 0;         // Engine::plan_blocks handle (dynamic blocks)
  uint64_t nbits = 0;        // exact size (compressed blocks)
  uint32_t tree_bits = 0;    // dynamic blocks: size of the tree header inside nbits (verbose report only)
  size_t instart = 0, inend = 0;  // stored blocks: absolute input positions
};

// A record of a position-independent span (ZopfliB200DeflateSpan / AppendSpan): bits that start at bit
// offset 0, or stored bytes.
struct BitString {
  std::vector<uint8_t> bytes;
  uint64_t nbits = 0;
};
struct SpanPiece {
  bool stored = false;
  BitString bits;
  size_t instart = 0, inend = 0;
  bool final = false;
};

struct HostTimes { double split = 0, emit = 0, other = 0; };
extern HostTimes g_host_times;

// Compresses the given byte ranges ("units": master blocks, or the single range of a
// ZopfliDeflatePart call) of `in`; the engine must already hold in[in_base ...).
void deflate_units(Engine& eng, const ZopfliOptions* opt, int btype, bool final_last, const unsigned char* in,
                   const std::vector<std::pair<size_t, size_t>>& units, size_t in_base,
                   std::vector<Piece>& pieces);

// Places the pieces behind the caller's stream (bit phase *bp), emits them on the device and appends
// the bytes to *out.  unit_bits (optional) receives the bit offset of every unit's first piece relative
// to the first bit written, plus the end offset as last entry.
void assemble(Engine& eng, const std::vector<Piece>& pieces, size_t in_base, unsigned char* bp, unsigned char** out,
              size_t* outsize, std::vector<uint64_t>* unit_bits = nullptr, bool verbose = false);

// bit position of every piece when the first one starts at bit0; returns the end position
uint64_t layout_pieces(const std::vector<Piece>& pieces, size_t in_base, uint64_t bit0, std::vector<Engine::EmitPiece>& ep,
                       std::vector<uint64_t>* unit_bits);

void splice_pieces(const std::vector<SpanPiece>& pieces, const unsigned char* in, unsigned char* bp,
                   unsigned char** out, size_t* outsize);

unsigned char* append_reserve(size_t n, unsigned char** out, size_t* outsize);
void append_bytes(const unsigned char* src, size_t n, unsigned char** out, size_t* outsize);

}  // namespace zb

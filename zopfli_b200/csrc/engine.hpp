// Host-visible interface of the CUDA engine (no CUDA types leak out of engine.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace zb {

struct ParseRange {
  uint64_t instart, inend;
  int mode;           // 0 greedy parse only (ZopfliLZ77Greedy), 1 optimal parse (ZopfliLZ77Optimal),
                      // 2 optimal parse with fixed-tree costs (ZopfliLZ77OptimalFixed)
  int numiterations;  // mode 1 only
};

struct ParseResult {        // symbols of range i: [off[i], off[i]+size[i]) in ll / d
  std::vector<uint32_t> off, size;
  std::vector<uint16_t> ll, d;
  std::vector<uint64_t> cost;  // mode 1: exact dynamic-block bit cost of the returned parse
};

struct EngineStats {  // accumulated since the last reset; times from CUDA events on the engine stream
  double ms_same, ms_keys, ms_scan, ms_scatter, ms_match, ms_greedy, ms_iterate, ms_pack, ms_h2d, ms_d2h, ms_split;
  uint64_t split_evals, split_rounds, iterate_launches;
  uint64_t launches;
  uint64_t match_positions, iterate_positions, iterate_steps;  // steps = positions x iterations
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t cyc_sum[6];   // k_iterate phase cycles summed over blocks (model, DP, trace, follow, cost, stats)
  uint64_t cyc_max[6];   // same for the block with the largest total (the critical path)
  uint64_t max_block_positions;
  uint64_t int_steps;    // DP steps that ran in the integer window
  uint64_t dp_cyc_sum[5], dp_cnt_sum[6];   // DP cycles / groups by kind of group (JobState::dpc, dpn), summed over blocks
  uint64_t dp_cyc_max[5], dp_cnt_max[6];   // same for the critical-path block
};

class Engine {
 public:
  // Engine contexts are leased for the duration of one API call (Engine::Lease): a context owns its
  // input buffers, lanes and arenas, so concurrent calls never share mutable device state.
  // dev < 0: the process's default device (ZOPFLI_B200_DEVICE, else LOCAL_RANK, else 0)
  static Engine* acquire(int dev = -1);
  static void release(Engine* e);
  static int default_device();
  static int device_count();
  struct Lease {
    Engine* e;
    explicit Lease(int dev = -1) : e(Engine::acquire(dev)) {}
    ~Lease() { Engine::release(e); }
    Lease(const Lease&) = delete;
    Lease& operator=(const Lease&) = delete;
    Engine& operator*() { return *e; }
    Engine* operator->() { return e; }
  };
  static EngineStats stats_all();   // summed over every context of the process
  static void reset_stats_all();

  // Input residency. set_input_host copies to the device (inside the caller's timed region);
  // set_input_device adopts an existing device buffer that must stay valid and be readable
  // 16 bytes past insize (e.g. a padded torch tensor).
  void set_input_host(const uint8_t* in, size_t insize);
  void set_input_device(const uint8_t* dev_in, size_t insize);

  // `lane` selects one of kLanes independent stream + arena sets; calls on different lanes may run
  // concurrently from different host threads (chunk pipelines; giant blocks beside the rest).
  static constexpr int kLanes = 8;
  void parse(const std::vector<ParseRange>& ranges, ParseResult& out, int lane = 0);

  // test seam: raw match table of one range (length, dist, expanded sublen[259] per position)
  void match_table(uint64_t instart, uint64_t inend, std::vector<uint16_t>& len,
                   std::vector<uint16_t>& dist, std::vector<uint16_t>& sublen,
                   std::vector<uint16_t>& same, std::vector<uint16_t>& hv, std::vector<uint16_t>& hv2);
  // test seam: device-side exact dynamic block size of a 320-bin histogram
  uint64_t device_block_bits(const uint32_t* hist320);

  // Split-cost service (ZopfliCalculateBlockSizeAutoType over symbol ranges, deflate.c:610-621):
  // split_begin uploads n stores given as flat symbol arrays with per-store offset/size,
  // split_eval prices a batch of (store, lstart, lend) ranges on the device.
  void split_begin(const uint16_t* ll, const uint16_t* d, const std::vector<uint64_t>& off,
                   const std::vector<uint32_t>& size, int lane = 1);
  // Stage A + B without a symbol round trip: greedy parse of `ranges` whose stores stay on the device
  // and become the stores of the split service of the same lane.  sizes[i] = symbols of range i.
  void greedy_to_split(const std::vector<ParseRange>& ranges, std::vector<uint32_t>& sizes, int lane = 1);
  // byte offset (from the start of its range) of symbol idx of store `store`, for many (store, idx)
  struct SplitPos { uint32_t store, idx; };
  void split_positions(const std::vector<SplitPos>& q, std::vector<uint32_t>& bytepos, int lane = 1);
  struct SplitReq { uint32_t store, lstart, lend; };
  void split_eval(const SplitReq* reqs, size_t n, uint64_t* costs, int lane = 1);

  // ---- device-resident finish (finish.cuh): stores, plans and the output stay on the GPU ----
  // Three symbol buffers, all indexed like the input (a block never has more symbols than bytes):
  //   kPack  best parse of every first-split block at [block start ...)
  //   kFin   the master block's concatenated store at [master start ...)
  //   kFix   fixed-tree re-parse of a final block at [block start ...)
  enum StoreBuf { kPack = 0, kFin = 1, kFix = 2 };
  // parse() whose results stay on the device in buffer `dest`; sizes / costs as in ParseResult
  // iter_costs (optional, for the verbose report of squeeze.c:493-495): cost of iteration i of range r at
  // [r * numiterations_max + i], ~0 where an iteration did not run
  void parse_keep(const std::vector<ParseRange>& ranges, int dest, std::vector<uint32_t>& sizes,
                  std::vector<uint64_t>& costs, int lane = 0, std::vector<uint64_t>* iter_costs = nullptr);
  // kPack -> kFin copies (ZopfliAppendLZ77Store), then the split service of `lane` over the kFin stores
  // [store_off[i], store_off[i] + store_size[i]) (skipped when store_off is empty)
  struct SymCopy { uint64_t src_off, dst_off; uint32_t n, pad; };
  void concat_stores(const std::vector<SymCopy>& copies, const std::vector<uint64_t>& store_off,
                     const std::vector<uint32_t>& store_size, int lane = 1);
  // stored / fixed / dynamic sizes of symbol ranges + the emission plan of each (kept on the device);
  // handles[i] identifies range i's plan in emit()
  struct PlanReq { uint64_t off; uint32_t n; uint32_t buf; };
  struct PlanCost { uint64_t unc, fixed, dyn, tree; };  // tree: bits of the dynamic block's tree header (part of dyn)
  void plan_blocks(const std::vector<PlanReq>& reqs, std::vector<PlanCost>& costs, std::vector<uint64_t>& handles,
                   int lane = 1);
  // Writes every piece at its final bit position and copies bytes [0, ceil(total_bits / 8)) of the
  // stream to host_dst.  Positions are absolute within the stream of this call (the first piece starts
  // at the caller's bit phase 0..7); stored pieces name bytes of the engine's input.
  struct EmitPiece {
    uint64_t bit_start, nbits;   // nbits: predicted size of a compressed piece (checked by the kernel)
    uint64_t off;                // symbols [off, off + n) of buffer `buf`
    uint64_t plan;               // handle from plan_blocks (dynamic pieces)
    uint64_t in_start, in_len;   // stored pieces: engine-relative input bytes
    uint32_t n;
    uint8_t buf, type, final, pad;   // type: 0 stored, 1 fixed, 2 dynamic
  };
  void emit(const std::vector<EmitPiece>& pieces, uint64_t total_bits, uint8_t* host_dst);
  // The two halves of emit(), for callers that move the stream between GPUs first (dist.cpp): the
  // emission leaves the stream in this context's output buffer (capacity >= reserve_bytes, zeroed up to
  // the emitted size) and returns its device address; download() copies device bytes to the host.
  void* emit_device(const std::vector<EmitPiece>& pieces, uint64_t total_bits, size_t reserve_bytes);
  void download(const void* dev_src, uint8_t* host_dst, size_t nbytes);
  // host -> device through the context's threaded pinned staging (synchronous)
  void upload(void* dev_dst, const uint8_t* host_src, size_t nbytes);
  void* stream();   // cudaStream_t of lane 0, the stream emit / download / set_input run on
  uint64_t input_size() const;
  uint64_t device_memory_total() const;   // bytes of device memory (sizes the master-block batches of a pipeline)

  void set_stream(void* cuda_stream);  // optional: run on the caller's stream
  EngineStats stats();
  void reset_stats();
  int device() const;

 private:
  void parse_common(const std::vector<ParseRange>& ranges, ParseResult& out, int dest, int lane,
                    std::vector<uint64_t>* iter_costs = nullptr);
  explicit Engine(int dev);
  struct Impl;
  Impl* p_;
};

}  // namespace zb

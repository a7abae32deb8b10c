// TEST INFRASTRUCTURE ONLY -- never part of libzopfli.so.
//
// A stand-in for zb::Engine (zopfli_b200/csrc/engine.hpp) that answers parse requests with the
// plain-C oracle (oracle/zopfli_oracle.c).  tests/hostmock/Makefile links it with the PRODUCT's
// host sources (driver.cpp, api.cpp) into tests/_build/libzopfli_hostmock.so so that the host
// logic -- block splitting, block-type choice, bit emission, containers, splice -- can be checked
// against the reference on a box without a GPU.  The shipped library links engine.cu instead and
// has no such path.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include <vector>

#include "../../oracle/zopfli_oracle.h"
#include "../../zopfli_b200/csrc/engine.hpp"
#include <memory>

#include "../../zopfli_b200/csrc/host_emit.hpp"
#include "../../zopfli_b200/csrc/lz77_store.hpp"

namespace zb {

struct Engine::Impl {
  std::vector<uint8_t> in;
  size_t insize = 0;
  std::vector<zb::Lz77Store> split_stores[Engine::kLanes];  // one set per lane, as in the engine
  std::vector<uint16_t> sym_ll[3], sym_d[3];                // kPack / kFin / kFix, indexed like the input
  std::vector<std::unique_ptr<BlockPlan>> plans;
  std::mutex mu;
  void ensure_sym() {
    for (int b = 0; b < 3; b++)
      if (sym_ll[b].size() < insize + 64) { sym_ll[b].assign(insize + 64, 0); sym_d[b].assign(insize + 64, 0); }
  }
};

Engine::Engine(int) : p_(new Impl) {}
int Engine::default_device() { return -1; }
int Engine::device_count() { return 0; }
Engine* Engine::acquire(int dev) { return new Engine(dev); }  // one fresh mock context per call
void Engine::release(Engine* e) { delete e->p_; e->p_ = nullptr; }
EngineStats Engine::stats_all() {
  EngineStats s;
  memset(&s, 0, sizeof(s));
  return s;
}
void Engine::reset_stats_all() {}
int Engine::device() const { return -1; }
void Engine::set_stream(void*) {}
EngineStats Engine::stats() {
  EngineStats s;
  memset(&s, 0, sizeof(s));
  return s;
}
void Engine::reset_stats() {}
void Engine::set_input_host(const uint8_t* in, size_t n) {
  p_->in.assign(in, in + n);
  p_->in.resize(n + 64, 0);
  p_->insize = n;
  p_->ensure_sym();
}
uint64_t Engine::input_size() const { return p_->insize; }
uint64_t Engine::device_memory_total() const { return 16ull << 30; }
void Engine::set_input_device(const uint8_t*, size_t) {}
void Engine::parse(const std::vector<ParseRange>& r, ParseResult& out, int) {
  out.off.assign(r.size(), 0);
  out.size.assign(r.size(), 0);
  out.cost.assign(r.size(), 0);
  out.ll.clear();
  out.d.clear();
  for (size_t i = 0; i < r.size(); i++) {
    ZoStore st;
    zo_store_init(&st);
    if (r[i].mode == 0) zo_lz77_greedy(p_->in.data(), r[i].instart, r[i].inend, &st);
    else if (r[i].mode == 1) zo_lz77_optimal(p_->in.data(), r[i].instart, r[i].inend, r[i].numiterations, &st);
    else zo_lz77_optimal_fixed(p_->in.data(), r[i].instart, r[i].inend, &st);
    out.off[i] = (uint32_t)out.ll.size();
    out.size[i] = (uint32_t)st.size;
    out.ll.insert(out.ll.end(), st.litlens, st.litlens + st.size);
    out.d.insert(out.d.end(), st.dists, st.dists + st.size);
    zo_store_free(&st);
  }
}
void Engine::parse_keep(const std::vector<ParseRange>& r, int dest, std::vector<uint32_t>& sizes,
                        std::vector<uint64_t>& costs, int lane, std::vector<uint64_t>* iter_costs) {
  if (iter_costs) iter_costs->clear();  // the oracle does not expose per-iteration costs
  ParseResult res;
  parse(r, res, lane);
  sizes = res.size;
  costs = res.cost;
  std::lock_guard<std::mutex> g(p_->mu);
  for (size_t i = 0; i < r.size(); i++) {
    memcpy(p_->sym_ll[dest].data() + r[i].instart, res.ll.data() + res.off[i], res.size[i] * 2);
    memcpy(p_->sym_d[dest].data() + r[i].instart, res.d.data() + res.off[i], res.size[i] * 2);
  }
}
void Engine::concat_stores(const std::vector<SymCopy>& copies, const std::vector<uint64_t>& store_off,
                           const std::vector<uint32_t>& store_size, int lane) {
  for (const SymCopy& c : copies) {
    memcpy(p_->sym_ll[kFin].data() + c.dst_off, p_->sym_ll[kPack].data() + c.src_off, (size_t)c.n * 2);
    memcpy(p_->sym_d[kFin].data() + c.dst_off, p_->sym_d[kPack].data() + c.src_off, (size_t)c.n * 2);
  }
  std::vector<zb::Lz77Store>& st = p_->split_stores[(unsigned)lane % kLanes];
  st.clear();
  st.resize(store_off.size());
  for (size_t i = 0; i < store_off.size(); i++) {
    st[i].append(p_->sym_ll[kFin].data() + store_off[i], p_->sym_d[kFin].data() + store_off[i], store_size[i], 0);
    st[i].finalize();
  }
}
void Engine::plan_blocks(const std::vector<PlanReq>& reqs, std::vector<PlanCost>& costs, std::vector<uint64_t>& handles,
                         int) {
  costs.resize(reqs.size());
  handles.resize(reqs.size());
  for (size_t i = 0; i < reqs.size(); i++) {
    std::unique_ptr<BlockPlan> p(new BlockPlan);
    host_block_plan(p_->sym_ll[reqs[i].buf].data() + reqs[i].off, p_->sym_d[reqs[i].buf].data() + reqs[i].off, reqs[i].n, *p);
    costs[i] = PlanCost{p->unc_bits, p->fixed_bits, p->dyn_bits, p->tree_bits};
    handles[i] = (uint64_t)(uintptr_t)p.get();
    std::lock_guard<std::mutex> g(p_->mu);
    p_->plans.push_back(std::move(p));
  }
}
void Engine::emit(const std::vector<EmitPiece>& pieces, uint64_t total_bits, uint8_t* host_dst) {
  const size_t nbytes = (size_t)((total_bits + 7) / 8);
  std::vector<uint8_t> buf(nbytes + 16, 0);
  for (const EmitPiece& p : pieces) {
    HostBitSink sink{buf.data(), p.bit_start};
    if (p.type == 0) {
      host_emit_stored(p.final != 0, p_->in.data() + p.in_start, p.in_len, sink);
    } else {
      const uint64_t nb = host_emit_block(p.type, p.final != 0, p_->sym_ll[p.buf].data() + p.off, p_->sym_d[p.buf].data() + p.off,
                                          p.n, (const BlockPlan*)(uintptr_t)p.plan, sink);
      if (nb != p.nbits) { fprintf(stderr, "mock: emitted %llu bits, predicted %llu\n", (unsigned long long)nb, (unsigned long long)p.nbits); abort(); }
    }
  }
  memcpy(host_dst, buf.data(), nbytes);
}
void* Engine::emit_device(const std::vector<EmitPiece>&, uint64_t, size_t) { return nullptr; }
void Engine::download(const void*, uint8_t*, size_t) {}
void Engine::upload(void*, const uint8_t*, size_t) {}
void* Engine::stream() { return nullptr; }
void Engine::match_table(uint64_t, uint64_t, std::vector<uint16_t>&, std::vector<uint16_t>&,
                         std::vector<uint16_t>&, std::vector<uint16_t>&, std::vector<uint16_t>&,
                         std::vector<uint16_t>&) {}
uint64_t Engine::device_block_bits(const uint32_t*) { return 0; }

// split service of the mock: the product's own HOST estimators (lz77_store.hpp), so the batched
// scheduler (batched_split.hpp) is exercised on CPU exactly as the driver uses it
void Engine::split_begin(const uint16_t* ll, const uint16_t* d, const std::vector<uint64_t>& off,
                         const std::vector<uint32_t>& size, int lane) {
  std::vector<zb::Lz77Store>& g_split_stores = p_->split_stores[(unsigned)lane % kLanes];
  g_split_stores.clear();
  g_split_stores.resize(off.size());
  for (size_t i = 0; i < off.size(); i++) {
    g_split_stores[i].append(ll + off[i], d + off[i], size[i], 0);
    g_split_stores[i].finalize();
  }
}
void Engine::greedy_to_split(const std::vector<ParseRange>& r, std::vector<uint32_t>& sizes, int lane) {
  std::vector<zb::Lz77Store>& g_split_stores = p_->split_stores[(unsigned)lane % kLanes];
  g_split_stores.clear();
  g_split_stores.resize(r.size());
  sizes.assign(r.size(), 0);
  for (size_t i = 0; i < r.size(); i++) {
    ZoStore st;
    zo_store_init(&st);
    zo_lz77_greedy(p_->in.data(), r[i].instart, r[i].inend, &st);
    g_split_stores[i].append(st.litlens, st.dists, st.size, 0);
    g_split_stores[i].finalize();
    sizes[i] = (uint32_t)st.size;
    zo_store_free(&st);
  }
}
void Engine::split_positions(const std::vector<SplitPos>& q, std::vector<uint32_t>& bytepos, int lane) {
  std::vector<zb::Lz77Store>& g_split_stores = p_->split_stores[(unsigned)lane % kLanes];
  bytepos.assign(q.size(), 0);
  for (size_t i = 0; i < q.size(); i++) bytepos[i] = (uint32_t)g_split_stores[q[i].store].pos[q[i].idx];
}
void Engine::split_eval(const SplitReq* reqs, size_t n, uint64_t* costs, int lane) {
  std::vector<zb::Lz77Store>& g_split_stores = p_->split_stores[(unsigned)lane % kLanes];
  static thread_local DynScratch s;
  for (size_t i = 0; i < n; i++) costs[i] = auto_type_bits(g_split_stores[reqs[i].store], reqs[i].lstart, reqs[i].lend, s);
}

}  // namespace zb

// multi-GPU entry points (dist.cpp needs CUDA + NCCL): absent from the mock
#include "../../zopfli_b200/csrc/dist.hpp"
namespace zb {
int dist_local_gpus() { return 1; }
bool dist_local_deflate(int, const ZopfliOptions*, int, const unsigned char*, size_t, unsigned char*, unsigned char**, size_t*) { return false; }
bool dist_rank_ready() { return false; }
int dist_rank() { return 0; }
void dist_rank_deflate(const ZopfliOptions*, int, const unsigned char*, size_t, unsigned char*, unsigned char**, size_t*, bool) {}
}  // namespace zb
extern "C" {
int ZopfliB200DistUniqueId(unsigned char*) { return 1; }
int ZopfliB200DistInit(int, int, const unsigned char*) { return 1; }
void ZopfliB200DistFinalize(void) {}
}

// TEST INFRASTRUCTURE ONLY -- never part of libzopfli.so.
//
// A stand-in for zb::Engine (zopfli_b200/csrc/engine.hpp) that answers parse requests with the
// plain-C oracle (oracle/zopfli_oracle.c).  tests/hostmock/Makefile links it with the PRODUCT's
// host sources (driver.cpp, api.cpp) into tests/_build/libzopfli_hostmock.so so that the host
// logic -- block splitting, block-type choice, bit emission, containers, splice -- can be checked
// against the reference on a box without a GPU.  The shipped library links engine.cu instead and
// has no such path.
#include <string.h>

#include <vector>

#include "../../oracle/zopfli_oracle.h"
#include "../../zopfli_b200/csrc/engine.hpp"
#include "../../zopfli_b200/csrc/lz77_store.hpp"

namespace zb {

struct Engine::Impl {
  std::vector<uint8_t> in;
  std::vector<zb::Lz77Store> split_stores[Engine::kLanes];  // one set per lane, as in the engine
};

Engine::Engine() : p_(new Impl) {}
Engine* Engine::acquire() { return new Engine; }  // one fresh mock context per call
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
}
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

// Optional per-category device timing with CUDA events on the launching stream (bench.py's roofline legs).
// Off by default; when on, every kernel wrapper records an event pair around its launch(es).
#include <mutex>
#include <vector>

#include "../../include/agb.h"
#include "agb_internal.h"

namespace agb {

namespace {
struct Rec {
  int cat;
  double flops;
  double bytes;
  cudaEvent_t a, b;
};
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool profiling_on() { return g_on; }

LaunchTimer::LaunchTimer(int cat, double flops, double bytes, cudaStream_t s) : idx_(-1), s_(s) {
  if (!g_on) return;
  Rec r{cat, flops, bytes, get_event(), get_event()};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  idx_ = (int)g_recs.size() - 1;
}
LaunchTimer::~LaunchTimer() {
  if (idx_ >= 0) cudaEventRecord(g_recs[idx_].b, s_);
}

}  // namespace agb

extern "C" {

int agb_profile_begin(void) {
  for (auto& r : agb::g_recs) {
    agb::g_pool.push_back(r.a);
    agb::g_pool.push_back(r.b);
  }
  agb::g_recs.clear();
  agb::g_on = true;
  return 0;
}

// Stops profiling, synchronises the device and sums per category (arrays of AGB_PROFILE_CATEGORIES entries).
int agb_profile_end(double* ms, int64_t* count, double* flops, double* bytes) {
  agb::g_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) {
    agb::set_error("agb_profile_end: device synchronize failed");
    return 1;
  }
  for (int i = 0; i < AGB_PROFILE_CATEGORIES; ++i) {
    if (ms) ms[i] = 0;
    if (count) count[i] = 0;
    if (flops) flops[i] = 0;
    if (bytes) bytes[i] = 0;
  }
  for (auto& r : agb::g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    if (r.cat < 0 || r.cat >= AGB_PROFILE_CATEGORIES) continue;
    if (ms) ms[r.cat] += t;
    if (count) count[r.cat] += 1;
    if (flops) flops[r.cat] += r.flops;
    if (bytes) bytes[r.cat] += r.bytes;
  }
  return 0;
}

}  // extern "C"

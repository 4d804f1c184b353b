// Library-wide pieces of the C ABI: version, thread-local error string, launch counter.
#include <atomic>
#include <mutex>
#include <vector>
#include <stdarg.h>
#include <string.h>

#include "wmk_common.cuh"

namespace wmk {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------ profiling
static const char* kFamilyNames[FAM_COUNT] = {"gemm", "window_attention", "layernorm", "dwconv_gelu", "layout",
                                              "small", "stft", "istft", "attack", "stats", "gemm_hbm"};
struct ProfRec { cudaEvent_t a, b; int family; double work, work2; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int family, double work, cudaStream_t s, double work2) : slot(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{prof_event(), prof_event(), family, work, work2};
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace wmk

extern "C" int wmk_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(wmk::g_prof_mu);
  wmk::g_prof_on = on != 0;
  return 0;
}
extern "C" int wmk_profile_num_families(void) { return wmk::FAM_COUNT; }
extern "C" const char* wmk_profile_family_name(int f) { return f >= 0 && f < wmk::FAM_COUNT ? wmk::kFamilyNames[f] : ""; }
extern "C" int wmk_profile_collect(double* ms, double* work, double* work2, uint64_t* launches) {
  using namespace wmk;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int f = 0; f < FAM_COUNT; ++f) { ms[f] = 0; work[f] = 0; launches[f] = 0; if (work2) work2[f] = 0; }
  for (ProfRec& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) { set_error("profile: event sync failed"); return WMK_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.family] += t; work[r.family] += r.work; launches[r.family] += 1;
    if (work2) work2[r.family] += r.work2;
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof.clear();
  return 0;
}

extern "C" int wmk_version(void) { return 100; }
extern "C" const char* wmk_last_error(void) { return wmk::g_err; }
extern "C" uint64_t wmk_launch_count(void) { return wmk::g_launches.load(std::memory_order_relaxed); }

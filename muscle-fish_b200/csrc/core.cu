// Error string, launch accounting, device properties.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

namespace mfish {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  // per device: a process that drives several GPUs sizes each one's grids with its own SM count
  static std::mutex mu;
  static std::map<int, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  // the small per-call scratch (sort buffers, tap tables) comes from the stream-ordered pool;
  // keep freed blocks cached instead of returning them to the driver at every synchronise
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cache[dev] = n;
  return n;
}

void count_launch(int k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_prof;

ProfScope::ProfScope(const char* n, cudaStream_t s) : name(n), st(s), rec(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec* r = new ProfRec{n, nullptr, nullptr};
  if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) {
    delete r;
    return;
  }
  cudaEventRecord(r->a, s);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  ProfRec* r = reinterpret_cast<ProfRec*>(rec);
  cudaEventRecord(r->b, st);
  std::lock_guard<std::mutex> g(g_prof_mu);
  g_prof.push_back(r);
}

}  // namespace mfish

extern "C" {

int mfish_abi_version(void) { return MFISH_ABI_VERSION; }

const char* mfish_last_error(void) { return mfish::g_err; }

int64_t mfish_launch_count(void) { return mfish::g_launches.load(); }

void mfish_profile_enable(int on) { mfish::g_prof_on.store(on ? 1 : 0); }

// Waits for the recorded events and writes one line per kernel name: "name calls total_ms\n".
// Returns the number of bytes needed (excluding the terminating 0); clears the records.
int64_t mfish_profile_collect(char* buf, int64_t cap) {
  using namespace mfish;
  std::vector<ProfRec*> recs;
  {
    std::lock_guard<std::mutex> g(g_prof_mu);
    recs.swap(g_prof);
  }
  std::map<std::string, std::pair<int64_t, double>> agg;
  std::vector<std::string> order;
  for (ProfRec* r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r->b) == cudaSuccess && cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) {
      auto it = agg.find(r->name);
      if (it == agg.end()) {
        order.push_back(r->name);
        agg[r->name] = {1, (double)ms};
      } else {
        it->second.first += 1;
        it->second.second += ms;
      }
    }
    cudaEventDestroy(r->a);
    cudaEventDestroy(r->b);
    delete r;
  }
  std::string out;
  char line[256];
  for (const std::string& n : order) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), (long long)agg[n].first, agg[n].second);
    out += line;
  }
  if (buf && cap > 0) {
    int64_t k = std::min<int64_t>((int64_t)out.size(), cap - 1);
    memcpy(buf, out.data(), (size_t)k);
    buf[k] = 0;
  }
  return (int64_t)out.size();
}

}  // extern "C"

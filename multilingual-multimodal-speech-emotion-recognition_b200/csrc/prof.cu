#include "prof.cuh"
#include "common.cuh"
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace ser {

namespace {
struct Rec { std::string family; double flops, bytes; cudaEvent_t e0, e1; };
struct State {
  bool enabled = false;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  std::mutex mu;
};
State& st() { static State s; return s; }

cudaEvent_t get_event() {
  State& s = st();
  if (!s.pool.empty()) { cudaEvent_t e = s.pool.back(); s.pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_enabled() { return st().enabled; }

void prof_begin(const char* family, double flops, double bytes, cudaStream_t s) {
  State& S = st();
  std::lock_guard<std::mutex> lk(S.mu);
  Rec r{family, flops, bytes, get_event(), get_event()};
  cudaEventRecord(r.e0, s);
  S.recs.push_back(r);
}

void prof_end(cudaStream_t s) {
  State& S = st();
  std::lock_guard<std::mutex> lk(S.mu);
  if (!S.recs.empty()) cudaEventRecord(S.recs.back().e1, s);
}

namespace {
__global__ void prof_stall_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}
__global__ void prof_null_kernel() { pdl_sync(); }
}  // namespace

}  // namespace ser

extern "C" {

// n profiled launches of an empty kernel (family "prof_null"): the floor of one event interval -- what the events
// themselves and an isolated launch add to every record of the profile
int ser_prof_null(int n, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    ser::ProfScope prof("prof_null", 0.0, 0.0, s);
    SER_CUDA_CHECK(ser::launch_pdl(ser::prof_null_kernel, dim3(1), dim3(32), 0, s));
  }
  return SER_OK;
}

int ser_prof_stall(double microseconds, void* stream) {
  if (microseconds <= 0.0) return SER_OK;
  if (microseconds > 2.0e5) microseconds = 2.0e5;          // never more than 0.2 s
  ser::prof_stall_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long>(microseconds * 1e3));
  SER_CUDA_CHECK(cudaGetLastError());
  return SER_OK;
}

int ser_prof_enable(int on) {
  ser::st().enabled = (on != 0);
  return SER_OK;
}

// Synchronises the device, aggregates by family and writes one line per family:
//   "<family> <launches> <total_ms> <flops> <bytes>\n".  Returns the number of bytes written (or needed).
int ser_prof_report(char* buf, int cap) {
  ser::State& S = ser::st();
  std::lock_guard<std::mutex> lk(S.mu);
  cudaDeviceSynchronize();
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (ser::Rec& r : S.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      Agg& a = agg[r.family];
      a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    }
    S.pool.push_back(r.e0);
    S.pool.push_back(r.e1);
  }
  S.recs.clear();
  cudaGetLastError();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops,
             kv.second.bytes);
    out += line;
  }
  if (buf != nullptr && cap > 0) {
    const int n = static_cast<int>(out.size()) < cap - 1 ? static_cast<int>(out.size()) : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return static_cast<int>(out.size());
}

}  // extern "C"

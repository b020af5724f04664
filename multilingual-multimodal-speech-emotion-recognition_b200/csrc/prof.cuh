// Opt-in per-kernel-family profiler: CUDA events recorded on the launching stream around each launch
// (B200_PROFILING.md: "time kernels with CUDA events on the launching stream").  Disabled by default --
// a single predictable branch per launch.  bench.py enables it for a short pass after the timed region
// to obtain per-family launch counts, device time, algorithmic FLOPs and algorithmic bytes.
#pragma once
#include <cuda_runtime.h>

namespace ser {

bool prof_enabled();
void prof_begin(const char* family, double flops, double bytes, cudaStream_t s);
void prof_end(cudaStream_t s);

struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(const char* family, double flops, double bytes, cudaStream_t stream) : s(stream), on(prof_enabled()) {
    if (on) prof_begin(family, flops, bytes, s);
  }
  ~ProfScope() { if (on) prof_end(s); }
};

}  // namespace ser

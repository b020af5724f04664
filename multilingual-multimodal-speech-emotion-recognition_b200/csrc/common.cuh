// Shared device/host helpers for the fusion-head kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ser_head.h"

namespace ser {

// dtype tags used across the C-ABI (include/ser_head.h: SER_F32 / SER_BF16)
enum : int { DT_F32 = 0, DT_BF16 = 1 };

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3 };
// gate modes: multiply the GEMM result by a derivative read from a saved activation
enum : int { GATE_NONE = 0, GATE_RELU = 1 /* g > 0 */, GATE_TANH = 2 /* 1 - g^2 */ };

// error codes returned through the C-ABI are the SER_OK / SER_ERR_* macros of include/ser_head.h

#define SER_CUDA_CHECK(...)                                                          \
  do {                                                                               \
    cudaError_t _e = (__VA_ARGS__);                                                       \
    if (_e != cudaSuccess) {                                                         \
      ser::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));               \
      return SER_ERR_CUDA;                                                      \
    }                                                                                \
  } while (0)

// every kernel launch in the library is followed by this macro, so the counter is the number of launches
#define SER_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    ser::count_launch();                                                             \
    SER_CUDA_CHECK(cudaGetLastError());                                              \
  } while (0)

#define SER_REQUIRE(cond, msg)                                                       \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ser::set_last_error(__FILE__, __LINE__, msg);                                  \
      return SER_ERR_ARG;                                                       \
    }                                                                                \
  } while (0)

#define SER_TRY(expr)                                                                \
  do {                                                                               \
    int _rc = (expr);                                                                \
    if (_rc != SER_OK) return _rc;                                              \
  } while (0)

void set_last_error(const char* file, int line, const char* msg);
void count_launch();
const char* last_error();

// ---------------------------------------------------------------------------------
// element load/store with on-the-fly conversion to fp32
// ---------------------------------------------------------------------------------
template <typename T> struct DTypeOf;
template <> struct DTypeOf<float> { static constexpr int value = DT_F32; };
template <> struct DTypeOf<__nv_bfloat16> { static constexpr int value = DT_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// load element i of a buffer whose dtype is only known at run time
__device__ __forceinline__ float ld_dyn(const void* p, size_t i, int is_f32) {
  return is_f32 ? reinterpret_cast<const float*>(p)[i]
                : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_dyn(void* p, size_t i, int is_f32, float v) {
  if (is_f32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// 8 consecutive elements -> fp32 registers (16-byte aligned for bf16, 32-byte for fp32)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// ---------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ float apply_gate(float v, float g, int mode) {
  if (mode == GATE_RELU) return g > 0.f ? v : 0.f;
  if (mode == GATE_TANH) return v * (1.f - g * g);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A training step is ~100 dependent launches replayed as one CUDA graph; with a
// plain kernel -> kernel edge the next grid is only scheduled after the previous one has drained and flushed, which
// costs 2-4 us per boundary -- more than most of the small kernels of the classifier tail.  Every kernel of the library
// therefore (1) runs its data-independent prologue (barrier init, tensor-memory allocation, descriptor prefetch, index
// arithmetic), (2) executes pdl_wait() BEFORE its first global-memory access that can alias another kernel's output --
// it returns once every prerequisite grid has completed and its writes are visible --, and (3) executes pdl_trigger()
// right after, which lets the NEXT grid of the stream be scheduled as soon as SM resources allow and run its own
// prologue up to its own pdl_wait().  Since a grid cannot complete before its wait has returned, completion stays
// transitive: kernel N+2 never runs ahead of kernel N.  launch_pdl() attaches the launch attribute; kernels launched
// without it (pdl_enabled() false: SER_PDL=0) see both instructions as no-ops.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---------------------------------------------------------------------------------
// GEMM front end shared by both tiers.  C[M,N] = epilogue(alpha * op(A) op(B)^T)
//   a_trans = 0 : A stored [M,K] row-major (lda = row stride)     "K-major"
//   a_trans = 1 : A stored [K,M] row-major                         "MN-major"
//   b_trans = 0 : B stored [N,K] row-major (an nn.Linear weight)   "K-major"
//   b_trans = 1 : B stored [K,N] row-major                         "MN-major"
// epilogue: v = alpha*acc + bias[n]; v = act(v); v = gate'(v, G[m,n]); v += R[m,n];
//           C = v  (or C += v when accumulate / split-K; fp32 C only)
// ---------------------------------------------------------------------------------
struct GemmArgs {
  int dtype = DT_F32;            // dtype of A, B (and default for C, R, G)
  int M = 0, N = 0, K = 0;
  const void* A = nullptr; long long lda = 0; int a_trans = 0;
  const void* B = nullptr; long long ldb = 0; int b_trans = 0;
  void* C = nullptr; long long ldc = 0; int c_f32 = 1;
  const float* bias = nullptr;
  const void* R = nullptr; long long ldr = 0; int r_f32 = 1;
  const void* G = nullptr; long long ldg = 0; int g_f32 = 1; int gate_mode = GATE_NONE;
  int act = ACT_NONE;
  int accumulate = 0;            // C += (fp32 C only)
  float alpha = 1.f;
  int splits = 0;                // 0 = choose automatically
  // batched GEMM: `batch` independent problems, operand / output b at base + b * stride (elements)
  int batch = 1;
  long long strideA = 0, strideB = 0, strideC = 0, strideR = 0, strideG = 0;
  long long strideBias = 0;      // batch stride of the bias vector (0: all problems share one bias)
  // optional by-product of dW-type GEMMs (a_trans): rowsum[m] = sum_k op(A)[m, k], i.e. the bias gradient when
  // A = dY stored [tokens, features].  Overwritten.  Batched: rowsum + b * strideRS.
  float* rowsum = nullptr; long long strideRS = 0;
  // the caller guarantees that C (when split-K accumulates into it) and rowsum already hold zeros: skip the memsets
  int out_zeroed = 0;
};

int gemm(const GemmArgs& a, cudaStream_t stream);
int gemm_simt_f32(const GemmArgs& a, cudaStream_t stream);
int gemm_tc_bf16(const GemmArgs& a, cudaStream_t stream);
bool gemm_tc_rowsum_ok(const GemmArgs& a);
// ---------------------------------------------------------------------------------
// Side branch: weight-gradient kernels are leaves of the backward dependency graph (nothing but the optimizer / the
// gradient all-reduce consumes them), yet in stream order they sit ON the chain of small dX kernels of the classifier
// tail and the fusion MLP.  A module's backward forks them onto a library-owned second stream (event record on the
// caller's stream -> wait on the side stream), keeps the dX chain on the caller's stream, and joins before it
// returns, so the caller sees ordinary stream semantics; under CUDA-graph capture the fork / join events become
// parallel branches of the step graph.  SER_SIDE_STREAM=0 keeps everything on the caller's stream.
// ---------------------------------------------------------------------------------
struct SideBranch {
  cudaStream_t main_stream = nullptr, side_stream = nullptr;
  bool on = false, used = false;
  explicit SideBranch(cudaStream_t main_s);
  int fork();                         // work enqueued on side() from now on sees everything enqueued on main so far
  int join();                         // main waits for everything enqueued on the side stream (no-op if never forked)
  cudaStream_t side() const { return on ? side_stream : main_stream; }
};

int device_sm_count();
int gemm_grid_sms();               // device_sm_count() minus the SMs reserved for concurrent collectives
void set_reserved_sms(int n);

}  // namespace ser

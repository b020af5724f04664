// fp32 tier GEMM: CUDA-core FFMA, 64x64x32 tiles, 4x4 register micro-tiles, register-prefetched operand tiles
// (the next K tile is in flight while the current one is multiplied), optional split-K.
// This is the "within 1e-4 of the fp32 reference" path (TF32 tensor cores are not accurate enough,
// see SURVEY.md section 7 "Hard parts"); it shares the epilogue contract of the tcgen05 kernel.
#include "kernels.cuh"
#include "prof.cuh"
#include <mutex>

namespace ser {

namespace {

constexpr int TM = 64, TN = 64, TK = 32;
constexpr int EPT = TM * TK / 256;     // operand elements per thread and tile

struct SimtEpilogue {
  float* C; long long ldc;
  const float* bias;
  const float* R; long long ldr;
  const float* G; long long ldg; int gate_mode;
  int act;
  int atomic;
  float alpha;
};

// A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
template <bool A_KCONT, bool B_KCONT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                 long long sbn, long long sbk, SimtEpilogue ep, int M, int N, int K, int splits) {
  pdl_sync();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int sp = blockIdx.z;
  const int ktiles = (K + TK - 1) / TK;
  const int kt0 = static_cast<int>((static_cast<long long>(sp) * ktiles) / splits);
  const int kt1 = static_cast<int>((static_cast<long long>(sp + 1) * ktiles) / splits);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // element e of a tile -> (row, k): thread order follows the contiguous axis of the operand
  auto fetch = [&](int kt, float (&ra)[EPT], float (&rb)[EPT]) {
    const int k0 = kt * TK;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
      const int e = tid + r * 256;
      int mm, kk;
      if (A_KCONT) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
      int nn, kb;
      if (B_KCONT) { kb = e % TK; nn = e / TK; } else { nn = e % TN; kb = e / TN; }
      const int gn = n0 + nn, gkb = k0 + kb;
      rb[r] = (gn < N && gkb < K) ? B[gn * sbn + gkb * sbk] : 0.f;
    }
  };
  float ra[EPT], rb[EPT];
  if (kt0 < kt1) fetch(kt0, ra, rb);
  for (int kt = kt0; kt < kt1; ++kt) {
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
      const int e = tid + r * 256;
      int mm, kk;
      if (A_KCONT) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      As[kk][mm] = ra[r];
      int nn, kb;
      if (B_KCONT) { kb = e % TK; nn = e / TK; } else { nn = e % TN; kb = e / TN; }
      Bs[kb][nn] = rb[r];
    }
    __syncthreads();
    if (kt + 1 < kt1) fetch(kt + 1, ra, rb);          // global loads overlap the FFMA block below
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool lead = (sp == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] * ep.alpha;
      if (ep.bias != nullptr && lead) v += ep.bias[n];
      v = apply_act(v, ep.act);
      if (ep.gate_mode != GATE_NONE) v = apply_gate(v, ep.G[m * ep.ldg + n], ep.gate_mode);
      if (ep.R != nullptr && lead) v += ep.R[m * ep.ldr + n];
      float* c = ep.C + m * ep.ldc + n;
      if (ep.atomic) atomicAdd(c, v); else *c = v;
    }
  }
}

}  // namespace

int gemm_simt_f32(const GemmArgs& a0, cudaStream_t stream) {
  if (a0.batch > 1) {          // fp32 tier: batched problems are simply issued one by one
    for (int b = 0; b < a0.batch; ++b) {
      GemmArgs a = a0;
      a.batch = 1;
      a.A = reinterpret_cast<const float*>(a0.A) + b * a0.strideA;
      a.B = reinterpret_cast<const float*>(a0.B) + b * a0.strideB;
      a.C = reinterpret_cast<float*>(a0.C) + b * a0.strideC;
      if (a0.bias) a.bias = a0.bias + b * a0.strideBias;
      if (a0.rowsum) a.rowsum = a0.rowsum + b * a0.strideRS;
      if (a0.R) a.R = reinterpret_cast<const float*>(a0.R) + b * a0.strideR;
      if (a0.G) a.G = reinterpret_cast<const float*>(a0.G) + b * a0.strideG;
      SER_TRY(gemm_simt_f32(a, stream));
    }
    return SER_OK;
  }
  const GemmArgs& a = a0;
  SER_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm_simt: empty problem");
  SER_REQUIRE(a.c_f32 && (a.R == nullptr || a.r_f32) && (a.G == nullptr || a.g_f32),
              "gemm_simt: fp32 tier expects fp32 C / residual / gate");
  const int mt = ceil_div(a.M, TM), nt = ceil_div(a.N, TN), ktiles = ceil_div(a.K, TK);
  const bool linear = (a.act == ACT_NONE && a.gate_mode == GATE_NONE);
  int splits = a.splits;
  if (splits <= 0) {
    splits = 1;
    const int tiles = mt * nt;
    const int target = 2 * device_sm_count();
    if (linear && tiles < target && ktiles >= 64) {
      splits = target / tiles;
      if (splits > ktiles / 4) splits = ktiles / 4;
      if (splits < 1) splits = 1;
    }
  }
  if (!linear) splits = 1;
  if (splits > ktiles) splits = ktiles;

  SimtEpilogue ep;
  ep.C = reinterpret_cast<float*>(a.C); ep.ldc = a.ldc;
  ep.bias = a.bias;
  ep.R = reinterpret_cast<const float*>(a.R); ep.ldr = a.ldr;
  ep.G = reinterpret_cast<const float*>(a.G); ep.ldg = a.ldg; ep.gate_mode = a.gate_mode;
  ep.act = a.act; ep.alpha = a.alpha;
  int accumulate = a.accumulate;
  if (splits > 1 && a.R != nullptr && a.R == a.C) {
    // in-place residual with split-K: C already holds R, so every split simply accumulates into it
    ep.R = nullptr;
    accumulate = 1;
  }
  ep.atomic = (splits > 1 || accumulate) ? 1 : 0;
  if (ep.atomic && !accumulate && !a.out_zeroed) {
    SER_CUDA_CHECK(cudaMemset2DAsync(a.C, a.ldc * sizeof(float), 0, a.N * sizeof(float), a.M, stream));
  }
  const float* A = reinterpret_cast<const float*>(a.A);
  const float* B = reinterpret_cast<const float*>(a.B);
  const long long sam = a.a_trans ? 1 : a.lda, sak = a.a_trans ? a.lda : 1;
  const long long sbn = a.b_trans ? 1 : a.ldb, sbk = a.b_trans ? a.ldb : 1;
  const double gflops = 2.0 * a.M * a.N * a.K;
  const double gesz = (a.dtype == DT_F32) ? 4.0 : 2.0;
  const double gbytes = (static_cast<double>(a.M) * a.K + static_cast<double>(a.N) * a.K) * gesz +
                        static_cast<double>(a.M) * a.N * ((a.c_f32 ? 4.0 : 2.0) + (a.R ? (a.r_f32 ? 4.0 : 2.0) : 0.0) +
                                                         (a.G ? (a.g_f32 ? 4.0 : 2.0) : 0.0));
  ProfScope prof(a.a_trans ? "gemm_simt_wgrad" : (a.b_trans ? "gemm_simt_dgrad" : "gemm_simt_fwd"), gflops, gbytes, stream);
  dim3 grid(nt, mt, splits);
  if (!a.a_trans && !a.b_trans)
    SER_CUDA_CHECK(launch_pdl(gemm_simt_kernel<true, true>, dim3(grid), dim3(256), 0, stream, A, sam, sak, B, sbn, sbk, ep, a.M, a.N, a.K, splits));
  else if (!a.a_trans && a.b_trans)
    SER_CUDA_CHECK(launch_pdl(gemm_simt_kernel<true, false>, dim3(grid), dim3(256), 0, stream, A, sam, sak, B, sbn, sbk, ep, a.M, a.N, a.K, splits));
  else if (a.a_trans && a.b_trans)
    SER_CUDA_CHECK(launch_pdl(gemm_simt_kernel<false, false>, dim3(grid), dim3(256), 0, stream, A, sam, sak, B, sbn, sbk, ep, a.M, a.N, a.K, splits));
  else
    SER_CUDA_CHECK(launch_pdl(gemm_simt_kernel<false, true>, dim3(grid), dim3(256), 0, stream, A, sam, sak, B, sbn, sbk, ep, a.M, a.N, a.K, splits));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.rowsum != nullptr && !gemm_tc_rowsum_ok(a)) {
    // no fused row sum on this path: rowsum of op(A) = column sums of the stored [K, M] operand, as its own launch
    SER_REQUIRE(a.a_trans, "gemm: rowsum needs an MN-major (a_trans) A operand");
    const int f32 = a.dtype == DT_F32 ? 1 : 0;
    if (a.batch > 1) SER_TRY(colsum_batched(a.A, f32, a.lda, a.K, a.M, a.rowsum, a.batch, a.strideA, a.strideRS, stream));
    else SER_TRY(colsum(a.A, f32, a.lda, a.K, a.M, a.rowsum, stream));
    GemmArgs b = a;
    b.rowsum = nullptr;
    return gemm(b, stream);
  }
  if (a.dtype == DT_F32) return gemm_simt_f32(a, stream);
  if (a.dtype == DT_BF16) return gemm_tc_bf16(a, stream);
  set_last_error(__FILE__, __LINE__, "gemm: unknown dtype");
  return SER_ERR_ARG;
}

int device_sm_count() {
  static int sms = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { sms = 148; return; }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  });
  return sms;
}

// SMs the persistent tcgen05 GEMM grids leave free.  In data-parallel runs an NCCL all-reduce kernel is resident for
// most of the backward pass (one CTA per channel); a persistent GEMM CTA assigned to an SM NCCL occupies cannot start
// until another CTA of its own grid has finished ALL its tiles, which stretches that GEMM by up to one full tile
// sequence.  parallel.DataParallelHead reserves as many SMs as NCCL has channels (ser_set_reserved_sms).
static int g_reserved_sms = 0;
void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int gemm_grid_sms() {
  const int n = device_sm_count() - g_reserved_sms;
  return n < 16 ? 16 : n;
}

}  // namespace ser

// AttentiveStatsPooling (src/models/pooling.py:15-28) after the 768->128 tanh GEMM, and the gated
// fusion mix (src/models/fusion.py:21-25).  These are the bandwidth-bound kernels of the head:
// 128-bit loads, warp-shuffle / shared-memory reductions over time and feature axes, masks in-kernel.
//
// Pooling layout: x is [B, T, D] row-major.  The statistics kernels run one CTA per (sample, 256-column
// slab): a warp covers the slab width (512 contiguous bytes of bf16 per row), the 8 warps walk T with 4 rows in flight.
// The weighted variance sum_t a_t (x_t - mu)^2 (pooling.py:26) is computed in one pass over x with the shifted-data
// identity (pivot = first frame of the column), see asp_stats_kernel.
// A sample whose frames are all padded yields NaN (softmax over all -inf), as in the reference.
#include "kernels.cuh"
#include "prof.cuh"
#include <stdlib.h>

namespace ser {

namespace {

constexpr int NTH = 256;      // threads per CTA
// A CTA owns one (sample, SLAB-column slab).  SLAB = 256 (the fast path, D % 256 == 0): a warp reads 512 contiguous
// bytes of a row (32 lanes x 8 elements), the 8 warps walk the time axis with 4 rows in flight each.  SLAB = 64 keeps
// any D % 64 == 0 working (8 lanes per row, 32 row groups).
template <int SLAB> struct SlabCfg {
  static constexpr int kLanes = SLAB / 8;        // lanes that share a row
  static constexpr int kRG = NTH / kLanes;       // row groups walking T
};

// e[b,t] = u[b,t,:] . w2 + b2 ; Hd / 8 lanes per row (16 for Hd = 128 -> two rows per warp), 128-bit loads
template <typename T>
__global__ void __launch_bounds__(256)
asp_score_kernel(const T* __restrict__ u, const float* __restrict__ w2, const float* __restrict__ b2,
                 float* __restrict__ e, int M, int Hd) {
  pdl_sync();
  const int lpr = Hd >> 3;                         // lanes per row (power of two <= 32)
  const int rpw = 32 / lpr;                        // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / lpr, cl = (lane % lpr) * 8;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + sub;
  float acc = 0.f;
  if (row < M) {
    float v[8], w[8];
    load8(u + static_cast<size_t>(row) * Hd + cl, v);
    load8(w2 + cl, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(v[i], w[i], acc);
  }
  for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < M && (lane % lpr) == 0) e[row] = acc + b2[0];
}

// softmax over T of the masked scores of sample b into smem `sa` (all threads participate)
__device__ __forceinline__ void softmax_row(const float* __restrict__ e, const float* __restrict__ mask, int T,
                                            float* sa, float* red) {
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float v = e[t];
    if (mask != nullptr && mask[t] == 0.f) v = -INFINITY;
    sa[t] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float p = expf(sa[t] - mx);      // all-masked: exp(-inf - -inf) = NaN, as torch.softmax
    sa[t] = p;
    sum += p;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int t = threadIdx.x; t < T; t += blockDim.x) sa[t] *= inv;
  __syncthreads();
}

template <typename T, int SLAB>
__global__ void __launch_bounds__(NTH)
asp_stats_kernel(const T* __restrict__ x, const float* __restrict__ e, const float* __restrict__ mask,
                 float* __restrict__ alpha, void* __restrict__ out, int out_f32, int Tlen, int D) {
  pdl_sync();
  constexpr int RG = SlabCfg<SLAB>::kRG, KL = SlabCfg<SLAB>::kLanes;
  extern __shared__ float smem[];
  float* sa = smem;                      // [T]
  float* red = sa + Tlen;                // [32]
  float* part = red + 32;                // [RG][SLAB]
  float* smean = part + RG * SLAB;       // [RG][SLAB] (second partial array)
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * SLAB;
  const int cl = (threadIdx.x % KL) * 8; // column offset inside the slab
  const int rg = threadIdx.x / KL;
  softmax_row(e + static_cast<size_t>(b) * Tlen, mask ? mask + static_cast<size_t>(b) * Tlen : nullptr, Tlen, sa, red);
  if (blockIdx.x == 0 && alpha != nullptr)
    for (int t = threadIdx.x; t < Tlen; t += blockDim.x) alpha[static_cast<size_t>(b) * Tlen + t] = sa[t];

  const T* xb = x + static_cast<size_t>(b) * Tlen * D + c0 + cl;
  // ONE pass over the slab: with weights that sum to 1, sum_t a_t (x_t - mu)^2 = sum_t a_t (x_t - p)^2 - (mu - p)^2 for
  // any pivot p; p = the column's first frame keeps both terms of the order of the variance itself (the textbook
  // shifted-data form: no cancellation beyond a factor of a few), and x is read exactly once -- the reference's
  // two-pass form (pooling.py:24-26) read it twice, the second time from L2.
  float piv[8], acc[8], acq[8];
  load8(xb, piv);
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = 0.f; acq[i] = 0.f; }
  constexpr int U = 4;                    // independent row loads in flight per thread
  int t = rg;
  for (; t + (U - 1) * RG < Tlen; t += U * RG) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) load8(xb + static_cast<size_t>(t + u * RG) * D, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float a = sa[t + u * RG];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[u][i] - piv[i];
        const float ad = a * d;
        acc[i] += ad;
        acq[i] = fmaf(ad, d, acq[i]);
      }
    }
  }
  for (; t < Tlen; t += RG) {
    float v[8];
    load8(xb + static_cast<size_t>(t) * D, v);
    const float a = sa[t];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - piv[i];
      const float ad = a * d;
      acc[i] += ad;
      acq[i] = fmaf(ad, d, acq[i]);
    }
  }
  float* part2 = smean;                  // second partial array follows the first: [RG][SLAB] each (see asp_fwd's smem size)
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[rg * SLAB + cl + i] = acc[i]; part2[rg * SLAB + cl + i] = acq[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < SLAB; c += NTH) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < RG; ++r) { s1 += part[r * SLAB + c]; s2 += part2[r * SLAB + c]; }
    const float p0 = to_f32(x[static_cast<size_t>(b) * Tlen * D + c0 + c]);
    const size_t o = static_cast<size_t>(b) * 2 * D + c0 + c;
    st_dyn(out, o, out_f32, p0 + s1);
    float var = s2 - s1 * s1;
    var = var < 0.f ? 0.f : var;          // (a NaN -- all frames padded -- stays NaN, as in the reference)
    st_dyn(out, o + D, out_f32, sqrtf(var + 1e-6f));
  }
}

// backward, part A: per (sample, slab): dx_stats and partial dalpha
template <typename T, int SLAB>
__global__ void __launch_bounds__(NTH)
asp_bwd_stats_kernel(const T* __restrict__ x, const float* __restrict__ alpha, const void* __restrict__ out,
                     int out_f32, const void* __restrict__ dout, int dout_f32, T* __restrict__ dx,
                     float* __restrict__ dalpha, int Tlen, int D) {
  pdl_sync();
  constexpr int RG = SlabCfg<SLAB>::kRG, KL = SlabCfg<SLAB>::kLanes;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * SLAB;
  const int cl = (threadIdx.x % KL) * 8;
  const int rg = threadIdx.x / KL;
  float mu[8], dmu[8], dvar[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t o = static_cast<size_t>(b) * 2 * D + c0 + cl + i;
    mu[i] = ld_dyn(out, o, out_f32);
    const float sd = ld_dyn(out, o + D, out_f32);
    dmu[i] = ld_dyn(dout, o, dout_f32);
    dvar[i] = ld_dyn(dout, o + D, dout_f32) / (2.f * sd);
  }
  const size_t base = static_cast<size_t>(b) * Tlen * D + c0 + cl;
  auto one_row = [&](int t, const float (&v)[8]) {
    float g[8];
    const float a = alpha[static_cast<size_t>(b) * Tlen + t];
    float da = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - mu[i];
      g[i] = a * (dmu[i] + 2.f * d * dvar[i]);
      da = fmaf(v[i], dmu[i], da);
      da = fmaf(d * d, dvar[i], da);
    }
    store8(dx + base + static_cast<size_t>(t) * D, g);
    // reduce over the KL column-lanes that share this row (consecutive lanes of one warp)
#pragma unroll
    for (int o = KL >> 1; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    if ((threadIdx.x % KL) == 0) atomicAdd(dalpha + static_cast<size_t>(b) * Tlen + t, da);
  };
  constexpr int U = 4;
  int t = rg;
  for (; t + (U - 1) * RG < Tlen; t += U * RG) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) load8(x + base + static_cast<size_t>(t + u * RG) * D, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) one_row(t + u * RG, v[u]);
  }
  for (; t < Tlen; t += RG) {
    float v[8];
    load8(x + base + static_cast<size_t>(t) * D, v);
    one_row(t, v);
  }
}

// backward, part B: softmax backward + scorer pre-activation gradient; one CTA per sample
template <typename T>
__global__ void __launch_bounds__(256)
asp_bwd_score_kernel(const T* __restrict__ u, const float* __restrict__ w2, const float* __restrict__ alpha,
                     const float* __restrict__ dalpha, T* __restrict__ dpre, float* __restrict__ dw2,
                     float* __restrict__ db2, int Tlen, int Hd) {
  pdl_sync();
  extern __shared__ float smem[];
  float* sde = smem;             // [T]
  float* red = sde + Tlen;       // [32]
  float* sw = red + 32;          // [Hd] local dw2
  const int b = blockIdx.x;
  const float* al = alpha + static_cast<size_t>(b) * Tlen;
  const float* da = dalpha + static_cast<size_t>(b) * Tlen;
  float dot = 0.f;
  for (int t = threadIdx.x; t < Tlen; t += blockDim.x) {
    const float a = al[t];
    if (a != 0.f) dot = fmaf(a, da[t], dot);       // padded frames have alpha = 0 exactly
  }
  dot = block_sum(dot, red);
  float dbl = 0.f;
  for (int t = threadIdx.x; t < Tlen; t += blockDim.x) {
    const float a = al[t];
    const float de = (a != 0.f) ? a * (da[t] - dot) : 0.f;
    sde[t] = de;
    dbl += de;
  }
  for (int c = threadIdx.x; c < Hd; c += blockDim.x) sw[c] = 0.f;
  dbl = block_sum(dbl, red);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(db2, dbl);
  // Hd / 8 lanes per row, 128-bit loads / stores; every lane owns 8 scorer columns and accumulates their dw2
  const int lpr = Hd >> 3;
  const int rows_per_pass = blockDim.x / lpr;
  const int cl = (threadIdx.x % lpr) * 8;
  float wv[8], accw[8];
  load8(w2 + cl, wv);
#pragma unroll
  for (int i = 0; i < 8; ++i) accw[i] = 0.f;
  for (int t = threadIdx.x / lpr; t < Tlen; t += rows_per_pass) {
    const size_t o = (static_cast<size_t>(b) * Tlen + t) * Hd + cl;
    float uv[8], g[8];
    load8(u + o, uv);
    const float de = sde[t];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] = de * wv[i] * (1.f - uv[i] * uv[i]);
      accw[i] = fmaf(de, uv[i], accw[i]);
    }
    store8(dpre + o, g);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sw[cl + i], accw[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < Hd; i += blockDim.x) atomicAdd(dw2 + i, sw[i]);
}

// ---------------------------------------------------------------------------------------------
// gated fusion mix: one warp per sample
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
mix_fwd_kernel(const T* __restrict__ pa, const T* __restrict__ pt, const T* __restrict__ ga, const T* __restrict__ gt,
               const float* __restrict__ wga, const float* __restrict__ bga, const float* __restrict__ wgt,
               const float* __restrict__ bgt, float* __restrict__ gates, T* __restrict__ fused, int B, int P, int G) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float sa = 0.f, st = 0.f;
  for (int c = lane; c < G; c += 32) {
    sa = fmaf(to_f32(ga[static_cast<size_t>(row) * G + c]), wga[c], sa);
    st = fmaf(to_f32(gt[static_cast<size_t>(row) * G + c]), wgt[c], st);
  }
  sa = warp_sum(sa) + bga[0];
  st = warp_sum(st) + bgt[0];
  const float wa = 1.f / (1.f + expf(-sa)), wt = 1.f / (1.f + expf(-st));
  const float ws = wa + wt + 1e-8f;
  const float na = wa / ws, nt = wt / ws;
  if (lane == 0) { gates[2 * row] = wa; gates[2 * row + 1] = wt; }
  for (int c = lane; c < P; c += 32) {
    const size_t o = static_cast<size_t>(row) * P + c;
    fused[o] = from_f32<T>(na * to_f32(pa[o]) + nt * to_f32(pt[o]));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mix_bwd_kernel(const T* __restrict__ pa, const T* __restrict__ pt, const T* __restrict__ ga, const T* __restrict__ gt,
               const float* __restrict__ wga, const float* __restrict__ wgt, const float* __restrict__ gates,
               const T* __restrict__ dfused, T* __restrict__ dpa, T* __restrict__ dpt, T* __restrict__ dga,
               T* __restrict__ dgt, float* __restrict__ dwga, float* __restrict__ dbga, float* __restrict__ dwgt,
               float* __restrict__ dbgt, int B, int P, int G) {
  pdl_sync();
  extern __shared__ float smem[];       // [2*G + 2] block-local accumulation of gate-weight grads
  for (int i = threadIdx.x; i < 2 * G + 2; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < B) {
    const float wa = gates[2 * row], wt = gates[2 * row + 1];
    const float ws = wa + wt + 1e-8f;
    const float na = wa / ws, nt = wt / ws;
    float dna = 0.f, dnt = 0.f;
    for (int c = lane; c < P; c += 32) {
      const size_t o = static_cast<size_t>(row) * P + c;
      const float g = to_f32(dfused[o]);
      dna = fmaf(g, to_f32(pa[o]), dna);
      dnt = fmaf(g, to_f32(pt[o]), dnt);
      dpa[o] = from_f32<T>(na * g);
      dpt[o] = from_f32<T>(nt * g);
    }
    dna = warp_sum(dna);
    dnt = warp_sum(dnt);
    const float inv2 = 1.f / (ws * ws);
    const float dwa = dna * (ws - wa) * inv2 - dnt * wt * inv2;
    const float dwt = dnt * (ws - wt) * inv2 - dna * wa * inv2;
    const float dsa = dwa * wa * (1.f - wa), dst = dwt * wt * (1.f - wt);
    for (int c = lane; c < G; c += 32) {
      const size_t o = static_cast<size_t>(row) * G + c;
      const float hav = to_f32(ga[o]), htv = to_f32(gt[o]);
      dga[o] = from_f32<T>(hav > 0.f ? dsa * wga[c] : 0.f);
      dgt[o] = from_f32<T>(htv > 0.f ? dst * wgt[c] : 0.f);
      atomicAdd(&smem[c], dsa * hav);
      atomicAdd(&smem[G + c], dst * htv);
    }
    if (lane == 0) { atomicAdd(&smem[2 * G], dsa); atomicAdd(&smem[2 * G + 1], dst); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) { atomicAdd(dwga + i, smem[i]); atomicAdd(dwgt + i, smem[G + i]); }
  if (threadIdx.x == 0) { atomicAdd(dbga, smem[2 * G]); atomicAdd(dbgt, smem[2 * G + 1]); }
}

// column slab of the statistics kernels (SER_ASP_SLAB: A/B switch)
int asp_slab(int D) {
  static const int env = getenv("SER_ASP_SLAB") ? atoi(getenv("SER_ASP_SLAB")) : 0;
  if ((env == 256 || env == 128 || env == 64) && D % env == 0) return env;
  return (D % 256 == 0) ? 256 : (D % 128 == 0) ? 128 : 64;
}

template <typename T, int SLAB>
int asp_bwd_stats_launch(const AspArgs& a, cudaStream_t s) {
  SER_CUDA_CHECK(launch_pdl(asp_bwd_stats_kernel<T, SLAB>, dim3(a.D / SLAB, a.B), dim3(NTH), 0, s, reinterpret_cast<const T*>(a.x),
                            a.alpha, a.out, a.out_f32, a.dout, a.dout_f32, reinterpret_cast<T*>(a.dx), a.dalpha, a.T, a.D));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

template <typename T, int SLAB>
int asp_stats_launch(const AspArgs& a, cudaStream_t s) {
  constexpr int RG = SlabCfg<SLAB>::kRG;
  const size_t smem = sizeof(float) * (a.T + 32 + 2 * RG * SLAB);
  auto kern = asp_stats_kernel<T, SLAB>;
  if (smem > 48 * 1024) SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SER_CUDA_CHECK(launch_pdl(kern, dim3(dim3(a.D / SLAB, a.B)), dim3(NTH), smem, s, reinterpret_cast<const T*>(a.x), a.e, a.mask, a.alpha, a.out,
                                                a.out_f32, a.T, a.D));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

template <typename T>
int asp_fwd_impl(const AspArgs& a, cudaStream_t s) {
  const int M = a.B * a.T;
  // algorithmic bytes: x read once + u read once (single-pass minimum)
  ProfScope prof("asp_fwd", 0.0, sizeof(T) * static_cast<double>(M) * (a.D + a.Hd), s);
  const int rows_per_cta = 8 * (32 / (a.Hd >> 3));
  SER_CUDA_CHECK(launch_pdl(asp_score_kernel<T>, dim3(ceil_div(M, rows_per_cta)), dim3(256), 0, s, reinterpret_cast<const T*>(a.u), a.w2, a.b2, a.e, M, a.Hd));
  SER_LAUNCH_CHECK();
  const int slab = asp_slab(a.D);
  return slab == 256 ? asp_stats_launch<T, 256>(a, s) : slab == 128 ? asp_stats_launch<T, 128>(a, s) : asp_stats_launch<T, 64>(a, s);
}

template <typename T>
int asp_bwd_impl(const AspArgs& a, cudaStream_t s) {
  ProfScope prof("asp_bwd", 0.0, sizeof(T) * static_cast<double>(a.B) * a.T * (2.0 * a.D + 2.0 * a.Hd), s);
  SER_TRY(zero_async(a.dalpha, sizeof(float) * a.B * a.T, s));
  const int slab = asp_slab(a.D);
  SER_TRY(slab == 256 ? (asp_bwd_stats_launch<T, 256>(a, s)) : slab == 128 ? (asp_bwd_stats_launch<T, 128>(a, s))
                                                                         : (asp_bwd_stats_launch<T, 64>(a, s)));
  const size_t smem = sizeof(float) * (a.T + 32 + a.Hd);
  auto kern = asp_bwd_score_kernel<T>;
  if (smem > 48 * 1024) SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SER_CUDA_CHECK(launch_pdl(kern, dim3(a.B), dim3(256), smem, s, reinterpret_cast<const T*>(a.u), a.w2, a.alpha, a.dalpha, reinterpret_cast<T*>(a.dpre),
                              a.dw2, a.db2, a.T, a.Hd));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace

int asp_fwd(const AspArgs& a, cudaStream_t s) {
  SER_REQUIRE(a.D % 64 == 0, "asp: feature dim must be a multiple of 64");
  SER_REQUIRE(a.Hd >= 8 && a.Hd <= 256 && 256 % a.Hd == 0, "asp: scorer hidden dim must be a power of two in [8, 256]");
  SER_REQUIRE(a.T > 0 && a.B > 0, "asp: empty input");
  return a.dtype == DT_F32 ? asp_fwd_impl<float>(a, s) : asp_fwd_impl<__nv_bfloat16>(a, s);
}
int asp_bwd(const AspArgs& a, cudaStream_t s) {
  SER_REQUIRE(a.D % 64 == 0, "asp: feature dim must be a multiple of 64");
  SER_REQUIRE(a.Hd >= 8 && a.Hd <= 256 && 256 % a.Hd == 0, "asp: scorer hidden dim must be a power of two in [8, 256]");
  return a.dtype == DT_F32 ? asp_bwd_impl<float>(a, s) : asp_bwd_impl<__nv_bfloat16>(a, s);
}

int fusion_mix_fwd(const MixArgs& a, cudaStream_t s) {
  ProfScope prof("fusion_mix_fwd", 0.0, (a.dtype == DT_F32 ? 4.0 : 2.0) * a.B * (3.0 * a.P + 2.0 * a.G), s);
#define SER_MIX_FWD(T)                                                                                             \
  SER_CUDA_CHECK(launch_pdl(mix_fwd_kernel<T>, dim3(ceil_div(a.B, 8)), dim3(256), 0, s, \
      reinterpret_cast<const T*>(a.pa), reinterpret_cast<const T*>(a.pt), reinterpret_cast<const T*>(a.ga),       \
      reinterpret_cast<const T*>(a.gt), a.wga, a.bga, a.wgt, a.bgt, a.gates, reinterpret_cast<T*>(a.fused), a.B,  \
      a.P, a.G))
  if (a.dtype == DT_F32) SER_MIX_FWD(float); else SER_MIX_FWD(__nv_bfloat16);
#undef SER_MIX_FWD
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int fusion_mix_bwd(const MixArgs& a, cudaStream_t s) {
  ProfScope prof("fusion_mix_bwd", 0.0, (a.dtype == DT_F32 ? 4.0 : 2.0) * a.B * (5.0 * a.P + 4.0 * a.G), s);
  const size_t smem = sizeof(float) * (2 * a.G + 2);
#define SER_MIX_BWD(T)                                                                                             \
  SER_CUDA_CHECK(launch_pdl(mix_bwd_kernel<T>, dim3(ceil_div(a.B, 8)), dim3(256), smem, s, \
      reinterpret_cast<const T*>(a.pa), reinterpret_cast<const T*>(a.pt), reinterpret_cast<const T*>(a.ga),       \
      reinterpret_cast<const T*>(a.gt), a.wga, a.wgt, a.gates, reinterpret_cast<const T*>(a.dfused),              \
      reinterpret_cast<T*>(a.dpa), reinterpret_cast<T*>(a.dpt), reinterpret_cast<T*>(a.dga),                      \
      reinterpret_cast<T*>(a.dgt), a.dwga, a.dbga, a.dwgt, a.dbgt, a.B, a.P, a.G))
  if (a.dtype == DT_F32) SER_MIX_BWD(float); else SER_MIX_BWD(__nv_bfloat16);
#undef SER_MIX_BWD
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

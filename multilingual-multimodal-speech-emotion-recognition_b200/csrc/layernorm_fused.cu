// Single-pass LayerNorm backward for the token-level LayerNorms of CrossModalAttention (norm_a / norm_t,
// cross_attention.py:43,51): bf16 rows of N = 256 / 512 / 768 columns.
//
// The generic path (elementwise.cu) reads dy and x twice: once per row for dx, once per column block for
// dgamma / dbeta.  Here one warp owns a row at a time and every lane owns the same 8 * N/256 columns of every row it
// visits, so the parameter gradients accumulate in registers while dx streams out: dy and x are read once, dx is
// written once -- 3 x M x N x 2 bytes, the algorithmic minimum.  dy / x stay packed (bf16) in registers and the next
// row is requested before the current one is reduced, so a warp keeps two rows (6 KB at N = 768) in flight; with 12
// resident warps per SM that is ~72 KB per SM outstanding, enough to cover HBM latency at full bandwidth.
// Optionally the kernel also writes mask * dx (the gradient entering a dropout-ed branch, dropout.cuh) so the residual
// dropout of the backward pass costs one extra store instead of one extra pass.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

constexpr float kEps = 1e-5f;

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
  v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

constexpr int kWarps = 4;      // 128 threads: ~150 registers per thread still leaves 12 warps per SM

template <int NCH, bool MASKED>
__global__ void __launch_bounds__(kWarps * 32, 3)
ln_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ stats, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                    __nv_bfloat16* __restrict__ dxm, DropSpec drop, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int M) {
  pdl_sync();
  constexpr int N = NCH * 256;
  __shared__ __align__(16) float sg[N];
  __shared__ __align__(16) float red[kWarps][N];
  for (int i = threadIdx.x; i < N; i += blockDim.x) sg[i] = gamma[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float invN = 1.f / static_cast<float>(N);
  DropKey dkey{0u, 1u};
  if (MASKED) dkey = drop_key(drop);

  float ag[NCH][8], ab[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[c][i] = 0.f; ab[c][i] = 0.f; }

  // contiguous block of rows per CTA, warps interleaved inside it
  const int rows_per_cta = (M + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  int row = r_begin + w;
  uint4 xr[NCH], gr[NCH];
  float2 st = make_float2(0.f, 0.f);
  if (row < r_end) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const size_t off = static_cast<size_t>(row) * N + c * 256 + lane * 8;
      xr[c] = __ldcs(reinterpret_cast<const uint4*>(x + off));
      gr[c] = __ldcs(reinterpret_cast<const uint4*>(dy + off));
    }
    st = *reinterpret_cast<const float2*>(stats + 2 * row);
  }
  for (; row < r_end; row += kWarps) {
    // request the next row before touching this one
    const int nrow = row + kWarps;
    uint4 xn[NCH], gn[NCH];
    float2 stn = make_float2(0.f, 0.f);
    if (nrow < r_end) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const size_t off = static_cast<size_t>(nrow) * N + c * 256 + lane * 8;
        xn[c] = __ldcs(reinterpret_cast<const uint4*>(x + off));
        gn[c] = __ldcs(reinterpret_cast<const uint4*>(dy + off));
      }
      stn = *reinterpret_cast<const float2*>(stats + 2 * nrow);
    }
    const float mu = st.x, rs = st.y;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float xv[8], gv[8], gm[8];
      unpack8(xr[c], xv);
      unpack8(gr[c], gv);
      load8(sg + c * 256 + lane * 8, gm);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = (xv[i] - mu) * rs;
        const float a = gv[i] * gm[i];
        s1 += a;
        s2 = fmaf(a, h, s2);
        ag[c][i] = fmaf(gv[i], h, ag[c][i]);
        ab[c][i] += gv[i];
      }
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float xv[8], gv[8], gm[8], o[8];
      unpack8(xr[c], xv);
      unpack8(gr[c], gv);
      load8(sg + c * 256 + lane * 8, gm);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = (xv[i] - mu) * rs;
        o[i] = rs * (gv[i] * gm[i] - s1 - h * s2);
      }
      const size_t off = static_cast<size_t>(row) * N + c * 256 + lane * 8;
      __stcs(reinterpret_cast<uint4*>(dx + off), pack8(o));
      if (MASKED) {
        const unsigned pair0 = static_cast<unsigned>(row) * (N / 2) + static_cast<unsigned>(c * 128 + lane * 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 m = drop_pair(dkey, pair0 + k, drop.thr, drop.scale);
          o[2 * k] *= m.x; o[2 * k + 1] *= m.y;
        }
        __stcs(reinterpret_cast<uint4*>(dxm + off), pack8(o));
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) { xr[c] = xn[c]; gr[c] = gn[c]; }
    st = stn;
  }
  // CTA-level reduction of the column accumulators, then one atomic per column
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NCH; ++c) store8(&red[w][c * 256 + lane * 8], pass == 0 ? ag[c] : ab[c]);
    __syncthreads();
    float* dst = pass == 0 ? dgamma : dbeta;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kWarps; ++j) s += red[j][n];
      atomicAdd(dst + n, s);
    }
  }
}

}  // namespace

bool layernorm_bwd_fused_ok(int dy_f32, int x_f32, int dx_f32, const void* add, const void* dx2, const float* dgamma,
                            int M, int N, int relu) {
  static const bool disabled = (getenv("SER_NO_FUSED_LN") != nullptr);     // A/B switch
  return !disabled && !dy_f32 && !x_f32 && !dx_f32 && add == nullptr && dx2 == nullptr && dgamma != nullptr &&
         relu == 0 && (N == 256 || N == 512 || N == 768) && M >= 1024;
}

// dgamma / dbeta accumulate (+=): the caller zeroes them.  dxm / drop: optional masked copy of dx (dropout.cuh).
int layernorm_bwd_fused(const void* dy, const void* x, const float* stats, const float* gamma, void* dx, void* dxm,
                        const DropSpec& drop, float* dgamma, float* dbeta, int M, int N, cudaStream_t s) {
  const bool masked = dxm != nullptr && drop.on();
  SER_REQUIRE(!masked || static_cast<long long>(M) * (N / 2) < (1LL << 32), "layernorm_bwd: dropout site too large");
  char pname[64];
  if (prof_enabled()) snprintf(pname, sizeof(pname), "layernorm_bwd:%dx%d", M, N);
  ProfScope prof(pname, 0.0, static_cast<double>(M) * N * 2.0 * (masked ? 4.0 : 3.0), s);
  int blocks = 3 * device_sm_count();            // 3 resident CTAs per SM, one wave
  const int max_blocks = ceil_div(M, 4 * kWarps);
  if (blocks > max_blocks) blocks = max_blocks;
  const __nv_bfloat16* pdy = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* px = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* pdx = reinterpret_cast<__nv_bfloat16*>(dx);
  __nv_bfloat16* pdm = reinterpret_cast<__nv_bfloat16*>(dxm);
#define SER_LN_FUSED(NCH)                                                                                              \
  do {                                                                                                                 \
    if (masked) SER_CUDA_CHECK(launch_pdl(ln_bwd_fused_kernel<NCH, true>, dim3(blocks), dim3(kWarps * 32), 0, s, pdy, px, stats, gamma, pdx, pdm, drop,   \
                                                                              dgamma, dbeta, M));                       \
    else SER_CUDA_CHECK(launch_pdl(ln_bwd_fused_kernel<NCH, false>, dim3(blocks), dim3(kWarps * 32), 0, s, pdy, px, stats, gamma, pdx, pdm, drop, dgamma, \
                                                                        dbeta, M));                                     \
  } while (0)
  if (N == 256) SER_LN_FUSED(1); else if (N == 512) SER_LN_FUSED(2); else SER_LN_FUSED(3);
#undef SER_LN_FUSED
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

// f1: per-utterance feature fusion (ser_featfuse_*, include/ser_head.h): the Linear(768 + F -> 768) . ReLU . Dropout
// the encoders apply per frame to [hidden ; utterance features] between the adapter and cross attention
// (src/models/audio_encoder.py:29-52,114-138; src/models/text_encoder.py:26-30,60-73).
//
// The F features are the same for every frame of an utterance, so the reference's expand + cat + 788-wide Linear is
//   c[u] = W[:, D:] f[u] + b        one B x D x F GEMM (fp32, CUDA cores: K = F <= 20)
//   z    = x W[:, :D]^T             the token-level GEMM of the tier (tcgen05 for bf16), 768-wide operands only
//   y    = dropout(relu(z + c[u]))  one streaming pass, in place on the GEMM's output
// and the backward never touches a [tokens, D+F] tensor either: dz = gate(dy, y); dx = dz W[:, :D];
// dW[:, :D] = dz^T x (+ db as the GEMM's row sums); dc[u] = sum_t dz[u,t]; dW[:, D:] = dc^T f.
#include "kernels.cuh"
#include "modules.cuh"
#include "dropout.cuh"
#include "prof.cuh"

namespace ser {

namespace {

// dst[r, c] = src[r, c] for a [rows, cols] fp32 matrix with row pitches lds / ldd (cols % 4 == 0): packs W[:, :D]
// into the tier's dtype and scatters the packed dW[:, :D] back into the nn.Linear layout
template <typename TD>
__global__ void __launch_bounds__(256)
copy2d_kernel(const float* __restrict__ src, long long lds, TD* __restrict__ dst, long long ldd, int rows, int cols4) {
  pdl_sync();
  const long long n = static_cast<long long>(rows) * cols4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols4), c = static_cast<int>(i % cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    TD* o = dst + r * ldd + c;
    o[0] = from_f32<TD>(v.x); o[1] = from_f32<TD>(v.y); o[2] = from_f32<TD>(v.z); o[3] = from_f32<TD>(v.w);
  }
}

template <typename TD>
int copy2d(const float* src, long long lds, TD* dst, long long ldd, int rows, int cols, cudaStream_t s) {
  const long long n = static_cast<long long>(rows) * (cols / 4);
  const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 4, (n + 255) / 256));
  SER_CUDA_CHECK(launch_pdl(copy2d_kernel<TD>, dim3(blocks), dim3(256), 0, s, src, lds, dst, ldd, rows, cols / 4));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

// y[r, :] = dropout(relu(y[r, :] + c[r / T, :])), 8 columns per thread and iteration (32-bit index arithmetic: the
// row / utterance divisions are the only non-streaming work of the pass)
template <typename T, bool DROP>
__global__ void __launch_bounds__(256)
bias_relu_drop_kernel(T* __restrict__ y, const float* __restrict__ c, unsigned n8, unsigned cols8, unsigned frames,
                      DropSpec d) {
  pdl_sync();
  DropKey key{0u, 1u};
  if (DROP) key = drop_key(d);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const unsigned row = i / cols8;
    const unsigned col = (i - row * cols8) * 8;
    float v[8], b[8];
    load8(y + static_cast<size_t>(i) * 8, v);
    load8(c + static_cast<size_t>(row / frames) * (cols8 * 8) + col, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k] + b[k], 0.f);
    if (DROP) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 m = drop_pair(key, i * 4 + k, d.thr, d.scale);
        v[2 * k] *= m.x; v[2 * k + 1] *= m.y;
      }
    }
    store8(y + static_cast<size_t>(i) * 8, v);
  }
}

// dz = y > 0 ? dy * scale : 0    (y is the post-dropout output: positive <=> ReLU open and element kept)
template <typename T>
__global__ void __launch_bounds__(256)
relu_keep_gate_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz, long long n8, float scale) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[8], v[8];
    load8(dy + i * 8, g);
    load8(y + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] * scale : 0.f;
    store8(dz + i * 8, g);
  }
}

int stream_grid(long long n8) {
  return static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (n8 + 255) / 256));
}

int check_dims(const ser_featfuse_desc& d, const char* who) {
  SER_REQUIRE(d.dtype == DT_F32 || d.dtype == DT_BF16, who);
  SER_REQUIRE(d.B > 0 && d.T > 0 && d.D > 0 && d.F > 0, "featfuse: empty problem");
  SER_REQUIRE(d.D % 128 == 0 && d.F % 4 == 0, "featfuse: D must be a multiple of 128 and F a multiple of 4");
  SER_REQUIRE(static_cast<long long>(d.B) * d.T * (d.D / 8) < (1LL << 30), "featfuse: too many tokens");
  return SER_OK;
}

}  // namespace

int featfuse_fwd(const ser_featfuse_desc& d, cudaStream_t s) {
  SER_TRY(check_dims(d, "featfuse_fwd: dtype"));
  SER_REQUIRE(d.x && d.feats && d.w && d.b && d.wx && d.c && d.y, "featfuse_fwd: null tensor");
  const int f = d.dtype == DT_F32 ? 1 : 0;
  const int M = d.B * d.T, D = d.D, F = d.F;
  const long long ldw = D + F;
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, DS_FEAT);
  SER_REQUIRE(!drop.on() || static_cast<long long>(M) * (D / 2) < (1LL << 32), "featfuse_fwd: dropout site too large");

  // c = feats W[:, D:]^T + b     (fp32 tier arithmetic in both tiers: B x D x F is tiny)
  GemmArgs gc;
  gc.dtype = DT_F32; gc.M = d.B; gc.N = D; gc.K = F;
  gc.A = d.feats; gc.lda = F; gc.B = d.w + D; gc.ldb = ldw;
  gc.C = d.c; gc.ldc = D; gc.c_f32 = 1; gc.bias = d.b;
  SER_TRY(gemm(gc, s));

  // wx = W[:, :D] in the tier's dtype, then z = x wx^T into y
  if (f) SER_TRY(copy2d<float>(d.w, ldw, reinterpret_cast<float*>(d.wx), D, D, D, s));
  else   SER_TRY(copy2d<__nv_bfloat16>(d.w, ldw, reinterpret_cast<__nv_bfloat16*>(d.wx), D, D, D, s));
  GemmArgs gz;
  gz.dtype = d.dtype; gz.M = M; gz.N = D; gz.K = D;
  gz.A = d.x; gz.lda = D; gz.B = d.wx; gz.ldb = D;
  gz.C = d.y; gz.ldc = D; gz.c_f32 = f;
  SER_TRY(gemm(gz, s));

  const unsigned n8 = static_cast<unsigned>(static_cast<long long>(M) * D / 8);
  ProfScope prof("featfuse_bias_relu_drop", 0.0, static_cast<double>(M) * D * (f ? 8.0 : 4.0), s);
  const unsigned frames = static_cast<unsigned>(d.T), cols8 = static_cast<unsigned>(D / 8);
#define SER_FF_EPI(TY_)                                                                                             \
  do {                                                                                                              \
    if (drop.on()) SER_CUDA_CHECK(launch_pdl(bias_relu_drop_kernel<TY_, true>, dim3(stream_grid(n8)), dim3(256), 0, s, reinterpret_cast<TY_*>(d.y),    \
                                                                                     d.c, n8, cols8, frames, drop));  \
    else SER_CUDA_CHECK(launch_pdl(bias_relu_drop_kernel<TY_, false>, dim3(stream_grid(n8)), dim3(256), 0, s, reinterpret_cast<TY_*>(d.y), d.c, n8,    \
                                                                           cols8, frames, drop));                    \
  } while (0)
  if (f) SER_FF_EPI(float); else SER_FF_EPI(__nv_bfloat16);
#undef SER_FF_EPI
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int featfuse_bwd(const ser_featfuse_desc& d, cudaStream_t s) {
  SER_TRY(check_dims(d, "featfuse_bwd: dtype"));
  SER_REQUIRE(d.x && d.feats && d.wx && d.y && d.dy && d.dz && d.dc && d.dwx && d.dw && d.db, "featfuse_bwd: null tensor");
  const int f = d.dtype == DT_F32 ? 1 : 0;
  const int M = d.B * d.T, D = d.D, F = d.F;
  const long long ldw = D + F;
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, DS_FEAT);

  // dz = gate(dy, y)
  {
    const long long n8 = static_cast<long long>(M) * D / 8;
    ProfScope prof("featfuse_gate", 0.0, static_cast<double>(M) * D * (f ? 12.0 : 6.0), s);
    if (f) SER_CUDA_CHECK(launch_pdl(relu_keep_gate_kernel<float>, dim3(stream_grid(n8)), dim3(256), 0, s, reinterpret_cast<const float*>(d.dy), reinterpret_cast<const float*>(d.y), reinterpret_cast<float*>(d.dz), n8, drop.scale));
    else SER_CUDA_CHECK(launch_pdl(relu_keep_gate_kernel<__nv_bfloat16>, dim3(stream_grid(n8)), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(d.dy), reinterpret_cast<const __nv_bfloat16*>(d.y),
        reinterpret_cast<__nv_bfloat16*>(d.dz), n8, drop.scale));
    SER_LAUNCH_CHECK();
  }

  // dW[:, :D] = dz^T x (packed, then scattered into the nn.Linear layout);  db = column sums of dz (the GEMM's row sums)
  GemmArgs gw;
  gw.dtype = d.dtype; gw.M = D; gw.N = D; gw.K = M;
  gw.A = d.dz; gw.lda = D; gw.a_trans = 1;
  gw.B = d.x; gw.ldb = D; gw.b_trans = 1;
  gw.C = d.dwx; gw.ldc = D; gw.c_f32 = 1;
  gw.rowsum = d.db; gw.out_zeroed = d.grads_zeroed;
  SER_TRY(gemm(gw, s));
  SER_TRY(copy2d<float>(d.dwx, D, d.dw, ldw, D, D, s));

  // dc[u] = sum_t dz[u, t];  dW[:, D:] = dc^T feats
  SER_TRY(colsum_batched(d.dz, f, D, d.T, D, d.dc, d.B, static_cast<long long>(d.T) * D, D, s));
  GemmArgs gf;
  gf.dtype = DT_F32; gf.M = D; gf.N = F; gf.K = d.B;
  gf.A = d.dc; gf.lda = D; gf.a_trans = 1;
  gf.B = d.feats; gf.ldb = F; gf.b_trans = 1;
  gf.C = d.dw + D; gf.ldc = ldw; gf.c_f32 = 1;
  gf.splits = 1;
  SER_TRY(gemm(gf, s));

  // dx = dz W[:, :D]
  if (d.dx != nullptr) {
    GemmArgs gx;
    gx.dtype = d.dtype; gx.M = M; gx.N = D; gx.K = D;
    gx.A = d.dz; gx.lda = D;
    gx.B = d.wx; gx.ldb = D; gx.b_trans = 1;
    gx.C = d.dx; gx.ldc = D; gx.c_f32 = f;
    SER_TRY(gemm(gx, s));
  }
  return SER_OK;
}

}  // namespace ser

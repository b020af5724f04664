// Memory-bound row/column kernels: casts, column sums (bias gradients), LayerNorm forward/backward.
// LayerNorm is used by cross_attention.py:43,51 (norm_a / norm_t, N = 768) and by every block of the
// classifier (classifier.py:80,107,119,125; N = 512 / 256).  One warp owns one row; each lane moves
// 8 consecutive elements per 128-bit (bf16) / 2x128-bit (fp32) access; statistics are fp32 and the
// variance is the two-pass form on register-resident data.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

constexpr float kLnEps = 1e-5f;
constexpr int kMaxChunks = 4;     // N <= 1024
constexpr int kCastSegs = 16;

__device__ __forceinline__ void load8_dyn(const void* p, size_t idx, int f32, float (&v)[8]) {
  if (f32) load8(reinterpret_cast<const float*>(p) + idx, v);
  else load8(reinterpret_cast<const __nv_bfloat16*>(p) + idx, v);
}
__device__ __forceinline__ void store8_dyn(void* p, size_t idx, int f32, const float (&v)[8]) {
  if (f32) store8(reinterpret_cast<float*>(p) + idx, v);
  else store8(reinterpret_cast<__nv_bfloat16*>(p) + idx, v);
}

__global__ void cast_kernel(const void* __restrict__ src, int src_f32, void* __restrict__ dst, int dst_f32,
                            long long n) {
  pdl_sync();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      load8_dyn(src, i, src_f32, v);
      store8_dyn(dst, i, dst_f32, v);
    } else {
      for (long long j = i; j < n; ++j) st_dyn(dst, j, dst_f32, ld_dyn(src, j, src_f32));
    }
  }
}

// fp32 -> bf16 copies of up to kCastSegs parameter buffers in ONE launch (every module's flat buffer is a multiple of
// 64 elements, so an 8-element vector never straddles two segments)
struct CastSegs {
  const float* src[kCastSegs];
  __nv_bfloat16* dst[kCastSegs];
  long long end[kCastSegs];        // exclusive prefix sums of the segment lengths, in 8-element vectors
  int n;
};
__global__ void __launch_bounds__(256)
cast_multi_kernel(const CastSegs segs) {
  pdl_sync();
  const long long total = segs.end[segs.n - 1];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int k = 0;
    while (i >= segs.end[k]) ++k;
    const long long local = i - (k == 0 ? 0 : segs.end[k - 1]);
    float v[8];
    load8(segs.src[k] + local * 8, v);
    store8(segs.dst[k] + local * 8, v);
  }
}

// zero a 16-byte aligned region: a kernel node instead of a memset node (inside a captured graph a memset between two
// kernels costs ~4 us of serialisation, a small kernel well under 2)
__global__ void __launch_bounds__(256)
zero_kernel(uint4* __restrict__ p, long long n16) {
  pdl_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// blockDim = (32, 8): x walks columns, y walks rows; grid = (ceil(N/32), row_splits)
__global__ void colsum_kernel(const void* __restrict__ X, int x_f32, long long ld, int M, int N,
                              float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N) {
    for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += gridDim.y * 8)
      acc += ld_dyn(X, static_cast<size_t>(m) * ld + n, x_f32);
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x];
    if (gridDim.y == 1) out[n] = s; else atomicAdd(out + n, s);
  }
}

// vectorised variant: each lane owns 8 consecutive columns (one 128-bit load of bf16 / two of fp32), a warp spans
// 256 columns.  Every CTA streams ONE CONTIGUOUS block of rows (DRAM-page friendly, like a copy kernel): its 8 warps
// interleave over the block with 8 independent row loads in flight per lane.
// grid = (ceil(N/256), row_blocks, batch)
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const void* __restrict__ X0, int x_f32, long long ld, int M, int N, float* __restrict__ out0,
                  long long strideX, long long strideOut) {
  pdl_sync();
  __shared__ float red[8][256 + 8];
  const void* X = x_f32 ? static_cast<const void*>(reinterpret_cast<const float*>(X0) + blockIdx.z * strideX)
                        : static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(X0) + blockIdx.z * strideX);
  float* out = out0 + blockIdx.z * strideOut;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    const int rows_per_cta = (M + gridDim.y - 1) / gridDim.y;
    const int r_begin = blockIdx.y * rows_per_cta;
    const int r_end = min(M, r_begin + rows_per_cta);
    int m = r_begin + w;
    constexpr int U = 8;
    for (; m + (U - 1) * 8 < r_end; m += U * 8) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) load8_dyn(X, static_cast<size_t>(m + u * 8) * ld + col, x_f32, v[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        acc[i] += ((v[0][i] + v[1][i]) + (v[2][i] + v[3][i])) + ((v[4][i] + v[5][i]) + (v[6][i] + v[7][i]));
    }
    for (; m < r_end; m += 8) {
      float v0[8];
      load8_dyn(X, static_cast<size_t>(m) * ld + col, x_f32, v0);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v0[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  const int n = blockIdx.x * 256 + c;
  if (n < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][c];
    if (gridDim.y == 1) out[n] = s; else atomicAdd(out + n, s);
  }
}

template <bool RES>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const void* __restrict__ x, int x_f32, void* __restrict__ y, int y_f32, void* __restrict__ y2,
              int y2_f32, const float* __restrict__ gamma, const float* __restrict__ beta,
              float* __restrict__ stats, int M, int N, int relu, const void* __restrict__ res, void* xout,
              DropSpec drop) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nch = (N + 255) >> 8;
  const float invN = 1.f / static_cast<float>(N);
  // optional prologue (res != NULL): the row that is normalised is res + dropout(x), also written to xout
  // (norm_a(audio_seq + self.dropout(a_out)), cross_attention.py:43) -- x, res and xout share x's dtype
  DropKey dkey{0u, 1u};
  if (RES && drop.on()) dkey = drop_key(drop);
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    float v[kMaxChunks][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        load8_dyn(x, static_cast<size_t>(row) * N + col, x_f32, v[c]);
        if (RES) {
          float r[8];
          load8_dyn(res, static_cast<size_t>(row) * N + col, x_f32, r);
          if (drop.on()) {
            const unsigned pair0 = static_cast<unsigned>(row) * static_cast<unsigned>(N >> 1) + static_cast<unsigned>(col >> 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 m = drop_pair(dkey, pair0 + k, drop.thr, drop.scale);
              v[c][2 * k] *= m.x; v[c][2 * k + 1] *= m.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[c][i] += r[i];
          store8_dyn(xout, static_cast<size_t>(row) * N + col, x_f32, v[c]);
          if (!x_f32) {                    // the saved row is bf16: normalise exactly what the backward will read
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c][i] = __bfloat162float(__float2bfloat16_rn(v[c][i]));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[c][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
      }
    }
    const float mean = warp_sum(s) * invN;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[c][i] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invN + kLnEps);
    if (lane == 0 && stats != nullptr) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float g[8], b[8], o[8];
        load8(gamma + col, g);
        load8(beta + col, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = (v[c][i] - mean) * rstd * g[i] + b[i];
          o[i] = (relu && t < 0.f) ? 0.f : t;
        }
        store8_dyn(y, static_cast<size_t>(row) * N + col, y_f32, o);
        if (y2 != nullptr) store8_dyn(y2, static_cast<size_t>(row) * N + col, y2_f32, o);
      }
    }
  }
}

// LayerNorm backward is split in two streaming kernels so that neither carries the other's register state:
//   ln_bwd_dx_kernel    : one warp per row, dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)) [+ add];
//                         the next row's operands are requested before the current row's reductions (2 rows in flight)
//   ln_bwd_param_kernel : column owners (8 columns per lane), contiguous row blocks per CTA, dgamma += g*xhat,
//                         dbeta += g; re-reads dy and x once (they sit in L2 / HBM) instead of holding 2 x N/32
//                         accumulators per lane in the row kernel.
template <int NCH>
__global__ void __launch_bounds__(256, 3)
ln_bwd_dx_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ x, int x_f32,
                 const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const void* __restrict__ add, int add_f32, void* __restrict__ dx, int dx_f32, void* __restrict__ dx2,
                 int dx2_f32, int M, int N, int relu) {
  pdl_sync();
  // gamma / beta live in shared memory (broadcast-free 128-bit reads) so a row costs only its own x and dy in
  // registers: <= 85 registers per thread keeps 24 warps per SM resident, which is what hides the HBM latency here
  __shared__ __align__(16) float sg[NCH * 256];
  __shared__ __align__(16) float sb[NCH * 256];
  for (int i = threadIdx.x; i < NCH * 256; i += blockDim.x) {
    sg[i] = (i < N) ? gamma[i] : 0.f;
    sb[i] = (i < N && relu) ? beta[i] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float invN = 1.f / static_cast<float>(N);
  // contiguous block of rows per CTA, warps interleaved inside it
  const int rows_per_cta = (M + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  for (int row = r_begin + (threadIdx.x >> 5); row < r_end; row += wpb) {
    float xv[NCH][8], gv[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < N) {
        load8_dyn(x, static_cast<size_t>(row) * N + col, x_f32, xv[c]);
        load8_dyn(dy, static_cast<size_t>(row) * N + col, dy_f32, gv[c]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { xv[c][i] = 0.f; gv[c][i] = 0.f; }
      }
    }
    const float mu = stats[2 * row], rs = stats[2 * row + 1];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      float gm[8];
      load8(sg + col, gm);
      float bt[8];
      if (relu) load8(sb + col, bt);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = (xv[c][i] - mu) * rs;
        if (relu && !(h * gm[i] + bt[i] > 0.f)) gv[c][i] = 0.f;
        const float a = gv[c][i] * gm[i];
        xv[c][i] = h;                      // x-hat replaces x
        gv[c][i] = a;                      // g * gamma replaces g
        s1 += a;
        s2 = fmaf(a, h, s2);
      }
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < N) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (gv[c][i] - s1 - xv[c][i] * s2);
        if (add != nullptr) {
          float r[8];
          load8_dyn(add, static_cast<size_t>(row) * N + col, add_f32, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r[i];
        }
        store8_dyn(dx, static_cast<size_t>(row) * N + col, dx_f32, o);
        if (dx2 != nullptr) store8_dyn(dx2, static_cast<size_t>(row) * N + col, dx2_f32, o);
      }
    }
  }
}

// grid = (ceil(N/256), row_blocks); dgamma / dbeta accumulate with one atomic per column and CTA
__global__ void __launch_bounds__(256)
ln_bwd_param_kernel(const void* __restrict__ dy, int dy_f32, const void* __restrict__ x, int x_f32,
                    const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int N, int relu) {
  pdl_sync();
  __shared__ float red[2][8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  if (col < N) {
    float gm[8], bt[8];
    load8(gamma + col, gm);
    load8(beta + col, bt);
    const int rows_per_cta = (M + gridDim.y - 1) / gridDim.y;
    const int r_begin = blockIdx.y * rows_per_cta;
    const int r_end = min(M, r_begin + rows_per_cta);
    constexpr int U = 4;
    int m = r_begin + w;
    for (; m + (U - 1) * 8 < r_end; m += U * 8) {
      float xv[U][8], gv[U][8];
      float2 st[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        load8_dyn(x, static_cast<size_t>(m + u * 8) * N + col, x_f32, xv[u]);
        load8_dyn(dy, static_cast<size_t>(m + u * 8) * N + col, dy_f32, gv[u]);
        st[u] = *reinterpret_cast<const float2*>(stats + 2 * (m + u * 8));
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h = (xv[u][i] - st[u].x) * st[u].y;
          float g = gv[u][i];
          if (relu && !(h * gm[i] + bt[i] > 0.f)) g = 0.f;
          ag[i] = fmaf(g, h, ag[i]);
          ab[i] += g;
        }
    }
    for (; m < r_end; m += 8) {
      float xv[8], gv[8];
      load8_dyn(x, static_cast<size_t>(m) * N + col, x_f32, xv);
      load8_dyn(dy, static_cast<size_t>(m) * N + col, dy_f32, gv);
      const float2 st = *reinterpret_cast<const float2*>(stats + 2 * m);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = (xv[i] - st.x) * st.y;
        float g = gv[i];
        if (relu && !(h * gm[i] + bt[i] > 0.f)) g = 0.f;
        ag[i] = fmaf(g, h, ag[i]);
        ab[i] += g;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][w][lane * 8 + i] = ag[i]; red[1][w][lane * 8 + i] = ab[i]; }
  __syncthreads();
  const int c = threadIdx.x;
  const int n = blockIdx.x * 256 + c;
  if (n < N) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sg += red[0][j][c]; sb += red[1][j][c]; }
    atomicAdd(dgamma + n, sg);
    atomicAdd(dbeta + n, sb);
  }
}

// ---------------------------------------------------------------------------------------------------
// Classifier block prologue: y = LN_outer(h); n = LN_inner(y)   (classifier.py:207-212 + :80), one pass over h.
// fp32 residual stream in / out, n in the tier's dtype (GEMM operand).  One warp per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln2_fwd_kernel(const float* __restrict__ h, float* __restrict__ y, void* __restrict__ n, int n_f32,
               const float* __restrict__ go, const float* __restrict__ bo, const float* __restrict__ gi,
               const float* __restrict__ bi, float* __restrict__ stats_o, float* __restrict__ stats_i, int M, int N) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nch = (N + 255) >> 8;
  const float invN = 1.f / static_cast<float>(N);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float v[kMaxChunks][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        load8(h + static_cast<size_t>(row) * N + col, v[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[c][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
      }
    }
    float mean = warp_sum(s) * invN;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[c][i] - mean; q += d * d; }
      }
    }
    float rstd = rsqrtf(warp_sum(q) * invN + kLnEps);
    if (lane == 0) { stats_o[2 * row] = mean; stats_o[2 * row + 1] = rstd; }
    s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float g[8], b[8];
        load8(go + col, g); load8(bo + col, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[c][i] = (v[c][i] - mean) * rstd * g[i] + b[i]; s += v[c][i]; }
        store8(y + static_cast<size_t>(row) * N + col, v[c]);
      }
    }
    mean = warp_sum(s) * invN;
    q = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[c][i] - mean; q += d * d; }
      }
    }
    rstd = rsqrtf(warp_sum(q) * invN + kLnEps);
    if (lane == 0) { stats_i[2 * row] = mean; stats_i[2 * row + 1] = rstd; }
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float g[8], b[8], o[8];
        load8(gi + col, g); load8(bi + col, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd * g[i] + b[i];
        store8_dyn(n, static_cast<size_t>(row) * N + col, n_f32, o);
      }
    }
  }
}

// Backward of the pair: dy = dskip + LN_inner'(dn) ; dh = LN_outer'(dy).  y is recomputed from h.
// dgi/dbi/dgo/dbo accumulate (+=) into fp32 [N] buffers.
__global__ void __launch_bounds__(256)
ln2_bwd_kernel(const float* __restrict__ dn, const float* __restrict__ dskip, const float* __restrict__ h,
               const float* __restrict__ stats_o, const float* __restrict__ stats_i, const float* __restrict__ go,
               const float* __restrict__ bo, const float* __restrict__ gi, float* __restrict__ dh,
               void* __restrict__ dh2, int dh2_f32, float* __restrict__ dgi, float* __restrict__ dbi,
               float* __restrict__ dgo, float* __restrict__ dbo, int M, int N) {
  pdl_sync();
  __shared__ float sacc[4][kMaxChunks * 256];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nch = (N + 255) >> 8;
  const float invN = 1.f / static_cast<float>(N);
  for (int i = threadIdx.x; i < 4 * kMaxChunks * 256; i += blockDim.x) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  float agi[kMaxChunks][8], abi[kMaxChunks][8], ago[kMaxChunks][8], abo[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) { agi[c][i] = 0.f; abi[c][i] = 0.f; ago[c][i] = 0.f; abo[c][i] = 0.f; }

  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const float mo = stats_o[2 * row], ro = stats_o[2 * row + 1];
    const float mi = stats_i[2 * row], ri = stats_i[2 * row + 1];
    float xo[kMaxChunks][8], xi[kMaxChunks][8], a[kMaxChunks][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float hv[8], gv[8], g_o[8], b_o[8], g_i[8];
        load8(h + static_cast<size_t>(row) * N + col, hv);
        load8(dn + static_cast<size_t>(row) * N + col, gv);
        load8(go + col, g_o); load8(bo + col, b_o); load8(gi + col, g_i);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (hv[i] - mo) * ro;
          const float yv = xh * g_o[i] + b_o[i];
          const float xih = (yv - mi) * ri;
          xo[c][i] = xh; xi[c][i] = xih;
          a[c][i] = gv[i] * g_i[i];
          s1 += a[c][i]; s2 += a[c][i] * xih;
          agi[c][i] += gv[i] * xih; abi[c][i] += gv[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { xo[c][i] = 0.f; xi[c][i] = 0.f; a[c][i] = 0.f; }
      }
    }
    s1 = warp_sum(s1) * invN; s2 = warp_sum(s2) * invN;
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float sk[8], g_o[8];
        load8(dskip + static_cast<size_t>(row) * N + col, sk);
        load8(go + col, g_o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dy = ri * (a[c][i] - s1 - xi[c][i] * s2) + sk[i];
          ago[c][i] += dy * xo[c][i]; abo[c][i] += dy;
          a[c][i] = dy * g_o[i];
          t1 += a[c][i]; t2 += a[c][i] * xo[c][i];
        }
      }
    }
    t1 = warp_sum(t1) * invN; t2 = warp_sum(t2) * invN;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = c * 256 + lane * 8;
      if (c < nch && col < N) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ro * (a[c][i] - t1 - xo[c][i] * t2);
        store8(dh + static_cast<size_t>(row) * N + col, o);
        if (dh2 != nullptr) store8_dyn(dh2, static_cast<size_t>(row) * N + col, dh2_f32, o);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int col = c * 256 + lane * 8;
    if (c < nch && col < N) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&sacc[0][col + i], agi[c][i]); atomicAdd(&sacc[1][col + i], abi[c][i]);
        atomicAdd(&sacc[2][col + i], ago[c][i]); atomicAdd(&sacc[3][col + i], abo[c][i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    atomicAdd(dgi + i, sacc[0][i]); atomicAdd(dbi + i, sacc[1][i]);
    atomicAdd(dgo + i, sacc[2][i]); atomicAdd(dbo + i, sacc[3][i]);
  }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ ds,
                                   long long n) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) ds[i] = dy[i] * y[i] * (1.f - y[i]);
}

}  // namespace

static int ln_grid(int M);

int cast_any(const void* src, int src_f32, void* dst, int dst_f32, long long n, cudaStream_t s) {
  if (n <= 0) return SER_OK;
  SER_REQUIRE((reinterpret_cast<uintptr_t>(src) & 31) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0,
              "cast: buffers must be 32-byte aligned");
  long long blocks = (n / 8 + 255) / 256;
  const long long cap = 8LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ProfScope prof("cast", 0.0, static_cast<double>(n) * ((src_f32 ? 4 : 2) + (dst_f32 ? 4 : 2)), s);
  SER_CUDA_CHECK(launch_pdl(cast_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, s, src, src_f32, dst, dst_f32, n));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Packed valid frames -> zero-padded batch.  The encoders pad every utterance of a batch to the longest one
// (src/models/audio_encoder.py:140-163, text_encoder.py:75-78: zero frames + a 1/0 mask); a host that feeds hidden
// states to the head therefore ships 25-40 % zeros.  Here the host sends only the valid frames, concatenated, plus the
// B + 1 row offsets, and one pass rebuilds out[b, t, :] = t < len_b ? packed[off_b + t, :] : 0 and mask[b, t].
// Warp per output row, 16-byte accesses.
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
unpack_frames_kernel(const uint4* __restrict__ packed, const long long* __restrict__ offsets, uint4* __restrict__ out,
                     float* __restrict__ mask, int B, int T, int row16) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long nrows = static_cast<long long>(B) * T;
  for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows;
       r += static_cast<long long>(gridDim.x) * (blockDim.x >> 5)) {
    const int b = static_cast<int>(r / T), t = static_cast<int>(r % T);
    const long long o0 = __ldg(offsets + b), o1 = __ldg(offsets + b + 1);
    const bool ok = t < (o1 - o0);
    uint4* dst = out + r * row16;
    if (ok) {
      const uint4* src = packed + (o0 + t) * row16;
      for (int i = lane; i < row16; i += 32) dst[i] = __ldg(src + i);
    } else {
      for (int i = lane; i < row16; i += 32) dst[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (mask != nullptr && lane == 0) mask[r] = ok ? 1.f : 0.f;
  }
}
}  // namespace

int unpack_frames(const void* packed, const long long* offsets, void* out, float* mask, int B, int T, int D, int elem_bytes,
                  cudaStream_t s) {
  SER_REQUIRE(B > 0 && T > 0 && D > 0, "unpack_frames: empty problem");
  SER_REQUIRE((static_cast<long long>(D) * elem_bytes) % 16 == 0, "unpack_frames: rows must be multiples of 16 bytes");
  SER_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "unpack_frames: buffers must be 16-byte aligned");
  const int row16 = static_cast<int>(static_cast<long long>(D) * elem_bytes / 16);
  const long long nrows = static_cast<long long>(B) * T;
  long long blocks = (nrows + 7) / 8;
  const long long cap = 8LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  // algorithmic bytes: every output row written once, valid rows read once (upper bound: all rows valid)
  ProfScope prof("unpack_frames", 0.0, 2.0 * nrows * D * elem_bytes, s);
  SER_CUDA_CHECK(launch_pdl(unpack_frames_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, s,
                            reinterpret_cast<const uint4*>(packed), offsets, reinterpret_cast<uint4*>(out), mask, B, T, row16));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int zero_async(void* p, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return SER_OK;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (bytes & 15) != 0) {
    SER_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, s));
    return SER_OK;
  }
  const long long n16 = static_cast<long long>(bytes / 16);
  const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 4, (n16 + 255) / 256));
  SER_CUDA_CHECK(launch_pdl(zero_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<uint4*>(p), n16));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int cast_multi(int n, const void* const* src, void* const* dst, const long long* counts, cudaStream_t s) {
  SER_REQUIRE(n >= 1 && n <= kCastSegs, "cast_multi: 1..16 segments");
  CastSegs segs{};
  long long acc = 0, bytes = 0;
  for (int i = 0; i < n; ++i) {
    SER_REQUIRE(counts[i] > 0 && counts[i] % 8 == 0, "cast_multi: segment lengths must be positive multiples of 8");
    SER_REQUIRE((reinterpret_cast<uintptr_t>(src[i]) & 31) == 0 && (reinterpret_cast<uintptr_t>(dst[i]) & 15) == 0,
                "cast_multi: segments must be 32-byte (fp32) / 16-byte (bf16) aligned");
    segs.src[i] = reinterpret_cast<const float*>(src[i]);
    segs.dst[i] = reinterpret_cast<__nv_bfloat16*>(dst[i]);
    acc += counts[i] / 8;
    segs.end[i] = acc;
    bytes += counts[i] * 6;
  }
  segs.n = n;
  ProfScope prof("cast", 0.0, static_cast<double>(bytes), s);
  const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (acc + 255) / 256));
  SER_CUDA_CHECK(launch_pdl(cast_multi_kernel, dim3(blocks), dim3(256), 0, s, segs));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t s) {
  return cast_any(src, 1, dst, 0, n, s);
}

int colsum(const void* X, int x_f32, long long ld, int M, int N, float* out, cudaStream_t s) {
  SER_REQUIRE(M > 0 && N > 0, "colsum: empty");
  char pname[64];
  if (prof_enabled()) snprintf(pname, sizeof(pname), "colsum:%dx%d", M, N);
  ProfScope prof(pname, 0.0, static_cast<double>(M) * N * (x_f32 ? 4 : 2), s);
  const bool vec_ok = (N % 8 == 0) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(X) & (x_f32 ? 31 : 15)) == 0);
  if (vec_ok) {
    // 128-bit loads, 4 rows in flight per lane; enough row splits to put ~4 CTAs on every SM
    const int gx = ceil_div(N, 256);
    int gy = ceil_div(4 * device_sm_count(), gx);
    const int max_gy = ceil_div(M, 128);
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    if (gy > 1) SER_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * N, s));
    SER_CUDA_CHECK(launch_pdl(colsum_vec_kernel, dim3(dim3(gx, gy, 1)), dim3(256), 0, s, X, x_f32, ld, M, N, out, 0, 0));
    SER_LAUNCH_CHECK();
    return SER_OK;
  }
  const int gx = ceil_div(N, 32);
  int gy = ceil_div(4 * device_sm_count(), gx);
  const int max_gy = ceil_div(M, 64);
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  if (gy > 1) SER_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * N, s));
  SER_CUDA_CHECK(launch_pdl(colsum_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, s, X, x_f32, ld, M, N, out));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

// `batch` column sums in one launch: X_b = X + b*strideX ([M,N], ld), out_b = out + b*strideOut
int colsum_batched(const void* X, int x_f32, long long ld, int M, int N, float* out, int batch, long long strideX,
                   long long strideOut, cudaStream_t s) {
  SER_REQUIRE(M > 0 && N > 0 && batch > 0, "colsum_batched: empty");
  SER_REQUIRE((N % 8 == 0) && (ld % 8 == 0) && (strideX % 8 == 0) && ((reinterpret_cast<uintptr_t>(X) & 31) == 0),
              "colsum_batched: needs 8-element aligned rows");
  const int gx = ceil_div(N, 256);
  int gy = ceil_div(M, 64);
  if (gy > 8) gy = 8;
  if (gy > 1) {
    if (strideOut == N) SER_TRY(zero_async(out, sizeof(float) * N * static_cast<size_t>(batch), s));   // contiguous: one launch
    else for (int b = 0; b < batch; ++b) SER_CUDA_CHECK(cudaMemsetAsync(out + b * strideOut, 0, sizeof(float) * N, s));
  }
  ProfScope prof("colsum_batched", 0.0, static_cast<double>(batch) * M * N * (x_f32 ? 4 : 2), s);
  SER_CUDA_CHECK(launch_pdl(colsum_vec_kernel, dim3(dim3(gx, gy, batch)), dim3(256), 0, s, X, x_f32, ld, M, N, out, strideX, strideOut));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

static int ln_grid(int M) {
  int blocks = ceil_div(M, 8);
  const int cap = 8 * device_sm_count();
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

int layernorm_fwd(const void* x, int x_f32, void* y, int y_f32, void* y2, int y2_f32, const float* gamma,
                  const float* beta, float* stats, int M, int N, int relu, cudaStream_t s, const void* res, void* xout,
                  const DropSpec* drop) {
  SER_REQUIRE(N % 8 == 0 && N <= kMaxChunks * 256, "layernorm: N must be a multiple of 8 and <= 1024");
  SER_REQUIRE((res == nullptr) == (xout == nullptr), "layernorm_fwd: res and xout go together");
  SER_REQUIRE(res == nullptr || static_cast<long long>(M) * (N / 2) < (1LL << 32), "layernorm_fwd: dropout site too large");
  char pname[64];
  if (prof_enabled()) snprintf(pname, sizeof(pname), "layernorm_fwd:%dx%d", M, N);
  // algorithmic bytes: x read, y (and its copy) written; with the residual-dropout prologue also res read and xout written
  ProfScope prof(pname, 0.0, static_cast<double>(M) * N * ((x_f32 ? 4 : 2) * (res ? 3 : 1) + (y_f32 ? 4 : 2) + (y2 ? (y2_f32 ? 4 : 2) : 0)), s);
  if (res != nullptr)
    SER_CUDA_CHECK(launch_pdl(ln_fwd_kernel<true>, dim3(ln_grid(M)), dim3(256), 0, s, x, x_f32, y, y_f32, y2, y2_f32, gamma, beta, stats, M, N, relu, res, xout,
                                                   drop != nullptr ? *drop : DropSpec{}));
  else
    SER_CUDA_CHECK(launch_pdl(ln_fwd_kernel<false>, dim3(ln_grid(M)), dim3(256), 0, s, x, x_f32, y, y_f32, y2, y2_f32, gamma, beta, stats, M, N, relu, nullptr,
                                                    nullptr, DropSpec{}));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int layernorm_bwd(const void* dy, int dy_f32, const void* x, int x_f32, const float* stats, const float* gamma,
                  const float* beta, const void* add, int add_f32, void* dx, int dx_f32, void* dx2, int dx2_f32,
                  float* dgamma, float* dbeta, int M, int N, int relu, cudaStream_t s, void* dxm, const DropSpec* drop) {
  SER_REQUIRE(N % 8 == 0 && N <= kMaxChunks * 256, "layernorm: N must be a multiple of 8 and <= 1024");
  SER_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  const bool masked = dxm != nullptr && drop != nullptr && drop->on();
  // token-level bf16 LayerNorms: single-pass kernel (layernorm_fused.cu)
  if (layernorm_bwd_fused_ok(dy_f32, x_f32, dx_f32, add, dx2, dgamma, M, N, relu))
    return layernorm_bwd_fused(dy, x, stats, gamma, dx, masked ? dxm : nullptr, masked ? *drop : DropSpec{}, dgamma, dbeta,
                               M, N, s);
  if (masked) {       // generic path: dx first, the masked copy as a separate pass
    SER_TRY(layernorm_bwd(dy, dy_f32, x, x_f32, stats, gamma, beta, add, add_f32, dx, dx_f32, dx2, dx2_f32, dgamma, dbeta,
                          M, N, relu, s, nullptr, nullptr));
    return dropout_apply(dx, dxm, nullptr, dx_f32, M, N, *drop, s);
  }
  char pname[64];
  if (prof_enabled()) snprintf(pname, sizeof(pname), "layernorm_bwd:%dx%d", M, N);
  // algorithmic bytes: dy, x read once, dx (and its copy / the added tensor) moved once
  ProfScope prof(pname, 0.0, static_cast<double>(M) * N * ((dy_f32 ? 4 : 2) + (x_f32 ? 4 : 2) + (dx_f32 ? 4 : 2) +
                                          (add ? (add_f32 ? 4 : 2) : 0) + (dx2 ? (dx2_f32 ? 4 : 2) : 0)), s);
  int blocks = ceil_div(M, 16);
  const int cap = 12 * device_sm_count();       // 3 resident CTAs per SM, 4 waves
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int nch = ceil_div(N, 256);
#define SER_LN_DX(NCH)                                                                                              \
  SER_CUDA_CHECK(launch_pdl(ln_bwd_dx_kernel<NCH>, dim3(blocks), dim3(256), 0, s, dy, dy_f32, x, x_f32, stats, gamma, beta, add, add_f32, dx, dx_f32,  \
                                               dx2, dx2_f32, M, N, relu))
  if (nch == 1) SER_LN_DX(1); else if (nch == 2) SER_LN_DX(2); else if (nch == 3) SER_LN_DX(3); else SER_LN_DX(4);
#undef SER_LN_DX
  SER_LAUNCH_CHECK();
  if (dgamma != nullptr) {
    const int gx = ceil_div(N, 256);
    int gy = ceil_div(4 * device_sm_count(), gx);
    const int max_gy = ceil_div(M, 64);
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    SER_CUDA_CHECK(launch_pdl(ln_bwd_param_kernel, dim3(dim3(gx, gy)), dim3(256), 0, s, dy, dy_f32, x, x_f32, stats, gamma, beta, dgamma, dbeta, M, N, relu));
    SER_LAUNCH_CHECK();
  }
  return SER_OK;
}

int layernorm2_fwd(const float* h, float* y, void* n, int n_f32, const float* go, const float* bo, const float* gi,
                   const float* bi, float* stats_o, float* stats_i, int M, int N, cudaStream_t s) {
  SER_REQUIRE(N % 8 == 0 && N <= kMaxChunks * 256 && M > 0, "layernorm2: N must be a multiple of 8 and <= 1024");
  ProfScope prof("layernorm2_fwd", 0.0, static_cast<double>(M) * N * (8 + (n_f32 ? 4 : 2)), s);
  SER_CUDA_CHECK(launch_pdl(ln2_fwd_kernel, dim3(ln_grid(M)), dim3(256), 0, s, h, y, n, n_f32, go, bo, gi, bi, stats_o, stats_i, M, N));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int layernorm2_bwd(const float* dn, const float* dskip, const float* h, const float* stats_o, const float* stats_i,
                   const float* go, const float* bo, const float* gi, float* dh, void* dh2, int dh2_f32, float* dgi,
                   float* dbi, float* dgo, float* dbo, int M, int N, cudaStream_t s) {
  SER_REQUIRE(N % 8 == 0 && N <= kMaxChunks * 256 && M > 0, "layernorm2: N must be a multiple of 8 and <= 1024");
  int blocks = ceil_div(M, 8);
  const int cap = 2 * device_sm_count();
  if (blocks > cap) blocks = cap;
  ProfScope prof("layernorm2_bwd", 0.0, static_cast<double>(M) * N * (16 + (dh2 ? (dh2_f32 ? 4 : 2) : 0)), s);
  SER_CUDA_CHECK(launch_pdl(ln2_bwd_kernel, dim3(blocks), dim3(256), 0, s, dn, dskip, h, stats_o, stats_i, go, bo, gi, dh, dh2, dh2_f32, dgi, dbi, dgo, dbo, M, N));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int sigmoid_bwd(const float* dy, const float* y, float* ds, long long n, cudaStream_t s) {
  if (n <= 0) return SER_OK;
  SER_CUDA_CHECK(launch_pdl(sigmoid_bwd_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, s, dy, y, ds, n));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

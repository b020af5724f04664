// SupConLoss (src/models/losses.py:67-88): supervised contrastive loss over one batch of embeddings.
//   fn = normalize(f);  S = fn fn^T / T;  S -= rowmax(S);  positives = same label, off-diagonal;
//   log_prob = S - log(sum_{j != i} exp(S_ij) + 1e-12);  loss = -mean_i( sum_pos log_prob / (n_pos + 1e-12) )
// The reference's training scripts import and construct it (train.py:8,86) but never add it to the loss; it is part of
// the drop-in surface of models/losses.py all the same.  Everything is fp32 (B x B similarities through the fp32
// CUDA-core GEMM); row kernels are one CTA per sample with the similarity row staged in shared memory.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

constexpr float kNormEps = 1e-12f;      // F.normalize eps
constexpr float kLogEps = 1e-12f;

// fn[i,:] = f[i,:] / max(||f_i||, eps); inv[i] = 1 / max(||f_i||, eps).  One warp per row.
__global__ void __launch_bounds__(256)
supcon_normalize_kernel(const void* __restrict__ f, int f_f32, float* __restrict__ fn, float* __restrict__ inv, int B, int D) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float ss = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = ld_dyn(f, static_cast<size_t>(row) * D + k, f_f32); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float r = 1.f / fmaxf(sqrtf(ss), kNormEps);
  for (int k = lane; k < D; k += 32) fn[static_cast<size_t>(row) * D + k] = ld_dyn(f, static_cast<size_t>(row) * D + k, f_f32) * r;
  if (lane == 0) inv[row] = r;
}

// one CTA per row i of S: row statistics and the row's loss term; backward: S row -> dL/dS row (in place)
template <bool BWD>
__global__ void __launch_bounds__(256)
supcon_rows_kernel(float* __restrict__ S, const long long* __restrict__ labels, float* __restrict__ stats,
                   float* __restrict__ loss, const float* __restrict__ gscale, int B) {
  pdl_sync();
  extern __shared__ float srow[];            // [B]
  __shared__ float red[32];
  const int i = blockIdx.x;
  float* row = S + static_cast<size_t>(i) * B;
  const long long yi = labels[i];
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) { const float v = row[j]; srow[j] = v; mx = fmaxf(mx, v); }
  mx = block_max(mx, red);                   // the max includes the diagonal (losses.py:79)
  float den = 0.f, pos = 0.f, cnt = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    if (j == i) continue;
    const float z = srow[j] - mx;
    den += expf(z);
    if (labels[j] == yi) { pos += z; cnt += 1.f; }
  }
  den = block_sum(den, red);
  pos = block_sum(pos, red);
  cnt = block_sum(cnt, red);
  const float logden = logf(den + kLogEps);
  if (!BWD) {
    if (threadIdx.x == 0) {
      const float mlpp = (pos - cnt * logden) / (cnt + kLogEps);
      atomicAdd(loss, -mlpp / static_cast<float>(B));
      stats[3 * i] = mx; stats[3 * i + 1] = den; stats[3 * i + 2] = cnt;
    }
  } else {
    // d loss / d S_ij = -(g/B) * ( [pos_ij] / (cnt+eps) - (cnt/(cnt+eps)) * exp(z_ij) / (den+eps) ),  j != i
    const float g = (gscale != nullptr ? gscale[0] : 1.f) / static_cast<float>(B);
    const float wpos = 1.f / (cnt + kLogEps), wden = cnt / (cnt + kLogEps) / (den + kLogEps);
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
      float d = 0.f;
      if (j != i) {
        d = -wden * expf(srow[j] - mx);
        if (labels[j] == yi) d += wpos;
      }
      row[j] = -g * d;
    }
  }
}

// df_i = inv_i * (dfn_i - fn_i (fn_i . dfn_i))   (backward of F.normalize for ||f|| >= eps); one warp per row
__global__ void __launch_bounds__(256)
supcon_normalize_bwd_kernel(const float* __restrict__ fn, const float* __restrict__ inv, const float* __restrict__ dfn,
                            void* __restrict__ df, int df_f32, int B, int D) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float dot = 0.f;
  for (int k = lane; k < D; k += 32) dot = fmaf(fn[static_cast<size_t>(row) * D + k], dfn[static_cast<size_t>(row) * D + k], dot);
  dot = warp_sum(dot);
  const float r = inv[row];
  for (int k = lane; k < D; k += 32) {
    const size_t o = static_cast<size_t>(row) * D + k;
    st_dyn(df, o, df_f32, r * (dfn[o] - fn[o] * dot));
  }
}

}  // namespace

// ws layout (fp32): fn [B,D] | inv [B] | stats [B,3] | S [B,B] | dfn [B,D]
size_t supcon_ws_bytes(int B, int D) {
  return sizeof(float) * (2 * static_cast<size_t>(B) * D + 4 * static_cast<size_t>(B) + static_cast<size_t>(B) * B) + 1024;
}

struct SupconBufs { float* fn; float* inv; float* stats; float* S; float* dfn; };
static SupconBufs supcon_bufs(void* ws, int B, int D) {
  SupconBufs b;
  float* p = reinterpret_cast<float*>(ws);
  b.fn = p; p += static_cast<size_t>(B) * D;
  b.inv = p; p += B;
  b.stats = p; p += 3 * static_cast<size_t>(B);
  p = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
  b.S = p; p += static_cast<size_t>(B) * B;
  p = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
  b.dfn = p;
  return b;
}

static int supcon_similarities(const SupconBufs& b, const void* f, int f_f32, int B, int D, float temperature, cudaStream_t s) {
  SER_CUDA_CHECK(launch_pdl(supcon_normalize_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, s, f, f_f32, b.fn, b.inv, B, D));
  SER_LAUNCH_CHECK();
  GemmArgs g;
  g.dtype = DT_F32; g.M = B; g.N = B; g.K = D;
  g.A = b.fn; g.lda = D; g.B = b.fn; g.ldb = D; g.C = b.S; g.ldc = B; g.c_f32 = 1; g.alpha = 1.f / temperature; g.splits = 1;
  return gemm(g, s);
}

int supcon_fwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, float* loss, void* ws,
               size_t ws_bytes, cudaStream_t s) {
  SER_REQUIRE(B > 0 && D > 0 && f != nullptr && labels != nullptr && loss != nullptr, "supcon_fwd: null tensor");
  SER_REQUIRE(B <= 8192, "supcon: at most 8192 samples per batch (similarity row staged in shared memory)");
  SER_REQUIRE(ws != nullptr && ws_bytes >= supcon_ws_bytes(B, D), "supcon_fwd: workspace too small");
  ProfScope prof("supcon_fwd", 2.0 * B * B * D, 4.0 * (static_cast<double>(B) * D + static_cast<double>(B) * B), s);
  const SupconBufs b = supcon_bufs(ws, B, D);
  SER_TRY(supcon_similarities(b, f, f_f32, B, D, temperature, s));
  SER_CUDA_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), s));
  SER_CUDA_CHECK(launch_pdl(supcon_rows_kernel<false>, dim3(B), dim3(256), sizeof(float) * B, s, b.S, labels, b.stats, loss, nullptr, B));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

// recomputes the similarities (B x B x D flops: cheaper than keeping S alive between forward and backward)
int supcon_bwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, const float* gscale,
               void* df, int df_f32, void* ws, size_t ws_bytes, cudaStream_t s) {
  SER_REQUIRE(B > 0 && D > 0 && f != nullptr && labels != nullptr && df != nullptr, "supcon_bwd: null tensor");
  SER_REQUIRE(B <= 8192, "supcon: at most 8192 samples per batch");
  SER_REQUIRE(ws != nullptr && ws_bytes >= supcon_ws_bytes(B, D), "supcon_bwd: workspace too small");
  ProfScope prof("supcon_bwd", 6.0 * B * B * D, 4.0 * (3.0 * B * D + 2.0 * B * B), s);
  const SupconBufs b = supcon_bufs(ws, B, D);
  SER_TRY(supcon_similarities(b, f, f_f32, B, D, temperature, s));
  SER_CUDA_CHECK(launch_pdl(supcon_rows_kernel<true>, dim3(B), dim3(256), sizeof(float) * B, s, b.S, labels, b.stats, nullptr, gscale, B));
  SER_LAUNCH_CHECK();
  // S = fn fn^T / T  ->  dfn = (G + G^T) fn / T
  GemmArgs g;
  g.dtype = DT_F32; g.M = B; g.N = D; g.K = B; g.alpha = 1.f / temperature; g.splits = 1;
  g.A = b.S; g.lda = B; g.a_trans = 0;
  g.B = b.fn; g.ldb = D; g.b_trans = 1;
  g.C = b.dfn; g.ldc = D; g.c_f32 = 1;
  SER_TRY(gemm(g, s));
  g.a_trans = 1; g.accumulate = 1;
  SER_TRY(gemm(g, s));
  SER_CUDA_CHECK(launch_pdl(supcon_normalize_bwd_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, s, b.fn, b.inv, b.dfn, df, df_f32, B, D));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

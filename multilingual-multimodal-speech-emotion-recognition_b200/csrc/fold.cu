// Folding of consecutive Linear layers of CrossModalAttention (bf16 tier only):
//   (a) outer q/k/v projection (cross_attention.py:38-40,46-48) followed by the MHA in-projection
//       (torch/nn/functional.py:5798):  (x Wq^T + bq) Win^T + bin  =  x (Win Wq)^T + (Win bq + bin)
//   (b) MHA out_proj followed by out_a / out_t (cross_attention.py:42,50): likewise.
// The products are tiny weight-only GEMMs (tcgen05 kernel, a few tiles); this file holds the element-wise glue:
// block-diagonal assembly of the three in-projections a modality's tokens go through, the folded biases, and the
// backward of the folds' bias / rank-1 terms.  Exact algebra -- the only numerical difference to the unfolded path
// is one bf16 rounding of the folded weight instead of one rounding of the intermediate activation.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

typedef __nv_bfloat16 bf16;

// Wbd[mod][r][c] (3S x 3S): block b = r / S is win_{sel(mod,b)}[r, c - bS] inside its diagonal block, 0 elsewhere.
// audio tokens pass through attn_a's q rows and attn_t's k, v rows; text tokens the other way round.
__global__ void fold_assemble_kernel(const bf16* __restrict__ win_a, const bf16* __restrict__ win_t,
                                     bf16* __restrict__ wbd_a, bf16* __restrict__ wbd_t, int S) {
  pdl_sync();
  const int S3 = 3 * S;
  const int mod = blockIdx.y;                    // 0 audio, 1 text
  const int r = blockIdx.x;
  const int b = r / S;
  const bf16* src = ((b == 0) == (mod == 0)) ? win_a : win_t;
  bf16* dst = (mod == 0 ? wbd_a : wbd_t) + static_cast<size_t>(r) * S3;
  for (int c = threadIdx.x; c < S3; c += blockDim.x) {
    const int cl = c - b * S;
    dst[c] = (cl >= 0 && cl < S) ? src[static_cast<size_t>(r) * S + cl] : __float2bfloat16(0.f);
  }
}

// warp per output: bc[mod][r] = sum_s Win_sel[r, s] * bqkv_mod[b*S + s] + bin_sel[r]   (r < 3S)
//                  bz[mod][r] = sum_s wout_mod[r, s] * bo_mod[s] + bout_mod[r]          (r < D)
__global__ void fold_bias_fwd_kernel(const bf16* __restrict__ win_a, const bf16* __restrict__ win_t,
                                     const float* __restrict__ bin_a, const float* __restrict__ bin_t,
                                     const float* __restrict__ bqkv_a, const float* __restrict__ bqkv_t,
                                     const bf16* __restrict__ wout_a, const bf16* __restrict__ wout_t,
                                     const float* __restrict__ bo_a, const float* __restrict__ bo_t,
                                     const float* __restrict__ bout_a, const float* __restrict__ bout_t,
                                     float* __restrict__ bc_a, float* __restrict__ bc_t, float* __restrict__ bz_a,
                                     float* __restrict__ bz_t, int S, int D) {
  pdl_sync();
  const int S3 = 3 * S;
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // [0, 2*3S + 2*D)
  if (o >= 2 * S3 + 2 * D) return;
  float acc = 0.f;
  if (o < 2 * S3) {
    const int mod = o / S3, r = o % S3, b = r / S;
    const bool use_a = ((b == 0) == (mod == 0));
    const bf16* w = (use_a ? win_a : win_t) + static_cast<size_t>(r) * S;
    const float* x = (mod == 0 ? bqkv_a : bqkv_t) + b * S;
    for (int s = lane; s < S; s += 32) acc = fmaf(__bfloat162float(w[s]), x[s], acc);
    acc = warp_sum(acc);
    if (lane == 0) (mod == 0 ? bc_a : bc_t)[r] = acc + (use_a ? bin_a : bin_t)[r];
  } else {
    const int q = o - 2 * S3, mod = q / D, r = q % D;
    const bf16* w = (mod == 0 ? wout_a : wout_t) + static_cast<size_t>(r) * S;
    const float* x = (mod == 0 ? bo_a : bo_t);
    for (int s = lane; s < S; s += 32) acc = fmaf(__bfloat162float(w[s]), x[s], acc);
    acc = warp_sum(acc);
    if (lane == 0) (mod == 0 ? bz_a : bz_t)[r] = acc + (mod == 0 ? bout_a : bout_t)[r];
  }
}

// backward glue of fold (a): rows of the in-projection gradients out of the block-diagonal product, rank-1 bias term,
// in-projection bias gradient.  grid = (3S, 2): one CTA per (row r, modality).
//   dwin_sel[r, s] = dWbd[mod][r, bS + s] + dbc[mod][r] * bqkv_mod[bS + s];   dbin_sel[r] = dbc[mod][r]
__global__ void fold_in_bwd_rows_kernel(const float* __restrict__ dwbd_a, const float* __restrict__ dwbd_t,
                                        const float* __restrict__ dbc_a, const float* __restrict__ dbc_t,
                                        const float* __restrict__ bqkv_a, const float* __restrict__ bqkv_t,
                                        float* __restrict__ dwin_a, float* __restrict__ dwin_t,
                                        float* __restrict__ dbin_a, float* __restrict__ dbin_t, int S) {
  pdl_sync();
  const int S3 = 3 * S;
  const int mod = blockIdx.y, r = blockIdx.x, b = r / S;
  const bool use_a = ((b == 0) == (mod == 0));
  const float* dwbd = (mod == 0 ? dwbd_a : dwbd_t) + static_cast<size_t>(r) * S3 + b * S;
  const float g = (mod == 0 ? dbc_a : dbc_t)[r];
  const float* bq = (mod == 0 ? bqkv_a : bqkv_t) + b * S;
  float* dst = (use_a ? dwin_a : dwin_t) + static_cast<size_t>(r) * S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) dst[s] = dwbd[s] + g * bq[s];
  if (threadIdx.x == 0) (use_a ? dbin_a : dbin_t)[r] = g;
}

// column-wise mat-vecs of the folds' bias paths.  block = (32 columns, 8 row groups), grid = (ceil(4S / 32), 2):
//   dbqkv_mod[bS + s] = sum_{r in block b} Win_sel[r, s] * dbc[mod][r]          (3S outputs)
//   dbo_mod[s]        = sum_r wout_mod[r, s] * dbz[mod][r]                       (S outputs)
__global__ void __launch_bounds__(256)
fold_bias_bwd_cols_kernel(const bf16* __restrict__ win_a, const bf16* __restrict__ win_t,
                          const bf16* __restrict__ wout_a, const bf16* __restrict__ wout_t,
                          const float* __restrict__ dbc_a, const float* __restrict__ dbc_t,
                          const float* __restrict__ dbz_a, const float* __restrict__ dbz_t,
                          float* __restrict__ dbqkv_a, float* __restrict__ dbqkv_t,
                          float* __restrict__ dbo_a, float* __restrict__ dbo_t, int S, int D) {
  pdl_sync();
  __shared__ float red[8][33];
  const int S3 = 3 * S;
  const int mod = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;
  float acc = 0.f;
  float* dst = nullptr;
  if (o < S3) {
    const int b = o / S, s = o % S;
    const bool use_a = ((b == 0) == (mod == 0));
    const bf16* w = (use_a ? win_a : win_t) + static_cast<size_t>(b) * S * S + s;      // rows bS .. bS+S-1, column s
    const float* g = (mod == 0 ? dbc_a : dbc_t) + b * S;
    for (int r = ty; r < S; r += 8) acc = fmaf(__bfloat162float(w[static_cast<size_t>(r) * S]), g[r], acc);
    dst = (mod == 0 ? dbqkv_a : dbqkv_t) + o;
  } else if (o < S3 + S) {
    const int s = o - S3;
    const bf16* w = (mod == 0 ? wout_a : wout_t) + s;
    const float* g = (mod == 0 ? dbz_a : dbz_t);
    for (int r = ty; r < D; r += 8) acc = fmaf(__bfloat162float(w[static_cast<size_t>(r) * S]), g[r], acc);
    dst = (mod == 0 ? dbo_a : dbo_t) + s;
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && dst != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    *dst = t;
  }
}

// fold (b) rank-1 term and bias: dwout[r, s] += dbz[r] * bo[s];  dbout[r] = dbz[r].   grid = (D, 2)
__global__ void fold_out_bwd_rows_kernel(const float* __restrict__ dbz_a, const float* __restrict__ dbz_t,
                                         const float* __restrict__ bo_a, const float* __restrict__ bo_t,
                                         float* __restrict__ dwout_a, float* __restrict__ dwout_t,
                                         float* __restrict__ dbout_a, float* __restrict__ dbout_t, int S) {
  pdl_sync();
  const int mod = blockIdx.y, r = blockIdx.x;
  const float g = (mod == 0 ? dbz_a : dbz_t)[r];
  const float* bo = (mod == 0 ? bo_a : bo_t);
  float* dst = (mod == 0 ? dwout_a : dwout_t) + static_cast<size_t>(r) * S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) dst[s] += g * bo[s];
  if (threadIdx.x == 0) (mod == 0 ? dbout_a : dbout_t)[r] = g;
}

}  // namespace

int fold_assemble(const void* win_a, const void* win_t, void* wbd_a, void* wbd_t, int S, cudaStream_t s) {
  ProfScope prof("fold_glue", 0.0, 2.0 * 9.0 * S * S * 2.0, s);
  SER_CUDA_CHECK(launch_pdl(fold_assemble_kernel, dim3(dim3(3 * S, 2)), dim3(256), 0, s, reinterpret_cast<const bf16*>(win_a), reinterpret_cast<const bf16*>(win_t),
                                                      reinterpret_cast<bf16*>(wbd_a), reinterpret_cast<bf16*>(wbd_t), S));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int fold_bias_fwd(const FoldBiasArgs& a, cudaStream_t s) {
  ProfScope prof("fold_glue", 0.0, 0.0, s);
  const int outs = 2 * 3 * a.S + 2 * a.D;
  SER_CUDA_CHECK(launch_pdl(fold_bias_fwd_kernel, dim3(ceil_div(outs, 8)), dim3(256), 0, s, reinterpret_cast<const bf16*>(a.win_a), reinterpret_cast<const bf16*>(a.win_t), a.bin_a, a.bin_t, a.bqkv_a, a.bqkv_t,
      reinterpret_cast<const bf16*>(a.wout_a), reinterpret_cast<const bf16*>(a.wout_t), a.bo_a, a.bo_t, a.bout_a, a.bout_t,
      a.bc_a, a.bc_t, a.bz_a, a.bz_t, a.S, a.D));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int fold_bwd_glue(const FoldBwdArgs& a, cudaStream_t s) {
  ProfScope prof("fold_glue", 0.0, 0.0, s);
  SER_CUDA_CHECK(launch_pdl(fold_in_bwd_rows_kernel, dim3(dim3(3 * a.S, 2)), dim3(256), 0, s, a.dwbd_a, a.dwbd_t, a.dbc_a, a.dbc_t, a.bqkv_a, a.bqkv_t, a.dwin_a,
                                                           a.dwin_t, a.dbin_a, a.dbin_t, a.S));
  SER_LAUNCH_CHECK();
  SER_CUDA_CHECK(launch_pdl(fold_bias_bwd_cols_kernel, dim3(dim3(ceil_div(4 * a.S, 32), 2)), dim3(256), 0, s, reinterpret_cast<const bf16*>(a.win_a), reinterpret_cast<const bf16*>(a.win_t), reinterpret_cast<const bf16*>(a.wout_a),
      reinterpret_cast<const bf16*>(a.wout_t), a.dbc_a, a.dbc_t, a.dbz_a, a.dbz_t, a.dbqkv_a, a.dbqkv_t, a.dbo_a, a.dbo_t,
      a.S, a.D));
  SER_LAUNCH_CHECK();
  SER_CUDA_CHECK(launch_pdl(fold_out_bwd_rows_kernel, dim3(dim3(a.D, 2)), dim3(256), 0, s, a.dbz_a, a.dbz_t, a.bo_a, a.bo_t, a.dwout_a, a.dwout_t, a.dbout_a,
                                                        a.dbout_t, a.S));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

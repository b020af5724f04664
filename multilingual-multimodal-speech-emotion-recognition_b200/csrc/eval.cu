// Inference-side kernels: OpenMax re-scaling (src/models/classifier.py:240-275), TTA view mean,
// temperature scaling, softmax / argmax / energy (src/eval.py:186-206, src/utils.py:12-14) and the
// 100-point temperature sweep (src/eval.py:48-67).  The reference runs these as Python loops with one
// device->host sync per sample; here each is a single launch with one warp per sample.
#include "kernels.cuh"

namespace ser {

namespace {

constexpr int kMaxC = 32;

__global__ void __launch_bounds__(256)
openmax_kernel(const float* __restrict__ feats, const float* __restrict__ logits, const float* __restrict__ av,
               const float* __restrict__ w_alpha, const float* __restrict__ w_beta, const float* __restrict__ w_tau,
               float* __restrict__ out, int B, int C, int F) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float unknown = 0.f;                       // torch.zeros(...) then running maximum (classifier.py:252,264)
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int d = lane; d < F; d += 32) {
      const float df = feats[static_cast<size_t>(row) * F + d] - av[static_cast<size_t>(c) * F + d];
      s = fmaf(df, df, s);
    }
    const float dist = sqrtf(warp_sum(s));
    const float beta = fmaxf(w_beta[c], 1e-6f);
    const float sx = fmaxf(dist - w_tau[c], 0.f);
    const float cdf = 1.f - expf(-powf(sx / beta, w_alpha[c]));
    unknown = fmaxf(unknown, cdf);
  }
  const float scale = (unknown > 0.3f) ? 1.f - unknown * 0.8f : 1.f;
  if (lane < C) out[static_cast<size_t>(row) * C + lane] = logits[static_cast<size_t>(row) * C + lane] * scale;
}

__global__ void __launch_bounds__(256)
eval_post_kernel(const float* __restrict__ lv, int V, int B, int C, float temperature, float* __restrict__ mean_logits,
                 float* __restrict__ probs, long long* __restrict__ preds, float* __restrict__ energy) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float m = 0.f;
  if (lane < C) {
    for (int v = 0; v < V; ++v) m += lv[(static_cast<size_t>(v) * B + row) * C + lane];
    m /= static_cast<float>(V);                 // torch.stack(...).mean(0)
    if (mean_logits != nullptr) mean_logits[static_cast<size_t>(row) * C + lane] = m;
  }
  const float z = (lane < C) ? m / temperature : -INFINITY;
  const float mx = warp_max(z);
  const float ex = (lane < C) ? expf(z - mx) : 0.f;
  const float sum = warp_sum(ex);
  const float p = ex / sum;
  if (lane < C && probs != nullptr) probs[static_cast<size_t>(row) * C + lane] = p;
  // argmax of probs == first maximum
  float bv = (lane < C) ? p : -INFINITY;
  int bi = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    if (preds != nullptr) preds[row] = bi;
    if (energy != nullptr) energy[row] = -(mx + logf(sum));     // -logsumexp(logits / T)
  }
}

// grid = (ceil(B/8), nT); err[t] += sum_b |conf - correct| / B
__global__ void __launch_bounds__(256)
temp_sweep_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                  const float* __restrict__ temps, float* __restrict__ err) {
  pdl_sync();
  __shared__ float acc;
  if (threadIdx.x == 0) acc = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float T = temps[blockIdx.y];
  if (row < B) {
    const float z = (lane < C) ? logits[static_cast<size_t>(row) * C + lane] / T : -INFINITY;
    const float mx = warp_max(z);
    const float ex = (lane < C) ? expf(z - mx) : 0.f;
    const float sum = warp_sum(ex);
    const float p = ex / sum;
    float bv = (lane < C) ? p : -INFINITY;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) atomicAdd(&acc, fabsf(bv - ((static_cast<long long>(bi) == labels[row]) ? 1.f : 0.f)));
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(err + blockIdx.y, acc / static_cast<float>(B));
}

// late-stage OOD scores (src/models/dual_gate_ood.py:203-220, :280-312, :360-383): one warp per sample
__global__ void __launch_bounds__(256)
late_ood_kernel(const float* __restrict__ logits, const void* __restrict__ feats, int feats_f32,
                const float* __restrict__ protos, const float* __restrict__ cov, const float* __restrict__ temperature,
                const float* __restrict__ mix, float* __restrict__ distances, float* __restrict__ scores, int B, int C,
                int D) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  // energy of the temperature-scaled logits: -logsumexp(logits / T)
  const float z = (lane < C) ? logits[static_cast<size_t>(row) * C + lane] / temperature[0] : -INFINITY;
  const float mx = warp_max(z);
  const float se = warp_sum((lane < C) ? expf(z - mx) : 0.f);
  const float energy = -(mx + logf(se));
  // diagonal Mahalanobis distance to every class prototype
  float min_d = INFINITY;
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float df = ld_dyn(feats, static_cast<size_t>(row) * D + d, feats_f32) - protos[static_cast<size_t>(c) * D + d];
      s += df * df * (1.f / (cov[static_cast<size_t>(c) * D + d] + 1e-8f));
    }
    const float dist = sqrtf(warp_sum(s));
    if (lane == 0) distances[static_cast<size_t>(row) * C + c] = dist;
    min_d = fminf(min_d, dist);
  }
  if (lane == 0) {
    const float e_norm = 1.f / (1.f + expf(energy));          // sigmoid(-energy)
    const float d_norm = expf(-min_d);
    const float m = fmaxf(mix[0], mix[1]);
    const float w0 = expf(mix[0] - m), w1 = expf(mix[1] - m);
    float* o = scores + static_cast<size_t>(row) * 5;
    o[0] = energy; o[1] = min_d; o[2] = e_norm; o[3] = d_norm;
    o[4] = (w0 * e_norm + w1 * d_norm) / (w0 + w1);
  }
}

}  // namespace

int late_ood(const float* logits, const void* feats, int feats_f32, const float* prototypes, const float* covariances,
             const float* temperature, const float* mix, float* distances, float* scores, int B, int C, int D,
             cudaStream_t s) {
  SER_REQUIRE(B > 0 && C > 0 && C <= kMaxC && D > 0, "late_ood: bad shape (num_classes <= 32)");
  SER_REQUIRE(logits && feats && prototypes && covariances && temperature && mix && distances && scores,
              "late_ood: null tensor");
  SER_CUDA_CHECK(launch_pdl(late_ood_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, s, logits, feats, feats_f32, prototypes, covariances, temperature, mix,
                                                  distances, scores, B, C, D));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int openmax_fwd(const float* feats, const float* logits, const float* act_vecs, const float* w_alpha,
                const float* w_beta, const float* w_tau, float* out, int B, int C, int F, cudaStream_t s) {
  SER_REQUIRE(C <= kMaxC && B > 0, "openmax: num_classes <= 32");
  SER_CUDA_CHECK(launch_pdl(openmax_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, s, feats, logits, act_vecs, w_alpha, w_beta, w_tau, out, B, C, F));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int eval_post(const float* logits_views, int V, int B, int C, float temperature, float* mean_logits, float* probs,
              long long* preds, float* energy, cudaStream_t s) {
  SER_REQUIRE(C <= kMaxC && B > 0 && V > 0, "eval_post: bad shape");
  SER_CUDA_CHECK(launch_pdl(eval_post_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, s, logits_views, V, B, C, temperature, mean_logits, probs, preds, energy));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int temperature_sweep(const float* logits, const long long* labels, int B, int C, const float* temps, int nT,
                      float* err, cudaStream_t s) {
  SER_REQUIRE(C <= kMaxC && B > 0 && nT > 0, "temperature_sweep: bad shape");
  SER_CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(float) * nT, s));
  SER_CUDA_CHECK(launch_pdl(temp_sweep_kernel, dim3(dim3(ceil_div(B, 8), nT)), dim3(256), 0, s, logits, labels, B, C, temps, err));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

// Multi-tensor AdamW step and gradient-norm clipping: the consumers of the head's gradients in the reference's training
// loops (torch.optim.AdamW with per-group lr / weight decay, src/train.py:72-83,169-177; clip_grad_norm_ in
// train_crema.py).  Memory-bound: 28 bytes per parameter per step (read p, g, m, v; write p, m, v).  Up to kSegs
// tensors per launch (pointer table in the kernel arguments), 128-bit accesses on the aligned body of every tensor.
//
// Update rule = torch.optim.AdamW (decoupled weight decay, bias-corrected):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

constexpr int kSegs = 24;

struct AdamSegs {
  float* p[kSegs]; const float* g[kSegs]; float* m[kSegs]; float* v[kSegs];
  long long end[kSegs];          // exclusive prefix sums of the tensor lengths in 4-element groups (rounded up per tensor)
  long long len[kSegs];          // element counts
  int n;
};

// AMP hand-off (torch.amp.GradScaler.step(optimizer), src/train.py:88,169-177): `found_inf` != 0 skips the whole update
// (every thread returns before touching memory), `grad_scale` divides the gradients (the scaler's loss scale, when the
// caller did not unscale_ first), and the bias-correction step count is then read from DEVICE memory (`step_dev`, kept
// by the caller: incremented only on steps that were not skipped -- the host cannot know without a sync).
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamSegs s, float lr, float b1, float b2, float eps, float wd, float bc1, float rsqrt_bc2,
                   const float* __restrict__ gscale, const float* __restrict__ grad_scale,
                   const float* __restrict__ found_inf, const float* __restrict__ step_dev) {
  pdl_sync();
  if (found_inf != nullptr && found_inf[0] != 0.f) return;
  const long long total = s.end[s.n - 1];
  float gs = gscale != nullptr ? gscale[0] : 1.f;      // e.g. the clip coefficient of clip_grad_norm
  if (grad_scale != nullptr) gs /= grad_scale[0];
  if (step_dev != nullptr) {                           // block-uniform: bias corrections from the device step count
    __shared__ float sh_bc[2];
    if (threadIdx.x == 0) {
      const double t = static_cast<double>(step_dev[0]);
      sh_bc[0] = static_cast<float>(1.0 - pow(static_cast<double>(b1), t));
      sh_bc[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), t)));
    }
    __syncthreads();
    bc1 = sh_bc[0]; rsqrt_bc2 = sh_bc[1];
  }
  const float decay = 1.f - lr * wd, step = lr / bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int k = 0;
    while (i >= s.end[k]) ++k;
    const long long e0 = (i - (k == 0 ? 0 : s.end[k - 1])) * 4;
    float* p = s.p[k] + e0; const float* g = s.g[k] + e0; float* m = s.m[k] + e0; float* v = s.v[k] + e0;
    const int nv = static_cast<int>(min(4LL, s.len[k] - e0));
    float pv[4], gv[4], mv[4], vv[4];
    const bool vec = nv == 4 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                  reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(g);
      const float4 c = *reinterpret_cast<const float4*>(m), d = *reinterpret_cast<const float4*>(v);
      pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
      mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
      for (int j = 0; j < nv; ++j) { pv[j] = p[j]; gv[j] = g[j]; mv[j] = m[j]; vv[j] = v[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nv) {
        const float gj = gv[j] * gs;
        const float pj = pv[j] * decay;
        const float mj = fmaf(b1, mv[j], (1.f - b1) * gj);
        const float vj = fmaf(b2, vv[j], (1.f - b2) * gj * gj);
        const float denom = sqrtf(vj) * rsqrt_bc2 + eps;
        pv[j] = pj - step * (mj / denom);
        mv[j] = mj; vv[j] = vj;
      }
    }
    if (vec) {
      *reinterpret_cast<float4*>(p) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(m) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
      for (int j = 0; j < nv; ++j) { p[j] = pv[j]; m[j] = mv[j]; v[j] = vv[j]; }
    }
  }
}

struct NormSegs { const float* g[kSegs]; long long end[kSegs]; long long len[kSegs]; int n; };

// out[0] += sum of squares of all gradient elements (caller zeroes out)
__global__ void __launch_bounds__(256)
sumsq_multi_kernel(const NormSegs s, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32];
  const long long total = s.end[s.n - 1];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int k = 0;
    while (i >= s.end[k]) ++k;
    const long long e0 = (i - (k == 0 ? 0 : s.end[k - 1])) * 4;
    const float* g = s.g[k] + e0;
    const int nv = static_cast<int>(min(4LL, s.len[k] - e0));
    if (nv == 4 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(g);
      acc = fmaf(a.x, a.x, acc); acc = fmaf(a.y, a.y, acc); acc = fmaf(a.z, a.z, acc); acc = fmaf(a.w, a.w, acc);
    } else {
      for (int j = 0; j < nv; ++j) acc = fmaf(g[j], g[j], acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// coef[0] = min(1, max_norm / (sqrt(sumsq) + 1e-6)) -- torch.nn.utils.clip_grad_norm_; norm_out[0] = sqrt(sumsq)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef, float* __restrict__ norm_out) {
  pdl_sync();
  const float nrm = sqrtf(sumsq[0]);
  const float c = max_norm / (nrm + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
  if (norm_out != nullptr) norm_out[0] = nrm;
}

}  // namespace

int adamw_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* counts,
                float lr, float beta1, float beta2, float eps, float weight_decay, int step, const float* gscale,
                cudaStream_t s, const float* grad_scale, const float* found_inf, const float* step_dev) {
  if (step_dev != nullptr && step < 1) step = 1;       // the host value is a placeholder when the device count is used
  SER_REQUIRE(n >= 1 && step >= 1, "adamw: need at least one tensor and step >= 1");
  const float bc1 = 1.f - static_cast<float>(pow(static_cast<double>(beta1), step));
  const float bc2 = 1.f - static_cast<float>(pow(static_cast<double>(beta2), step));
  const float rsqrt_bc2 = 1.f / sqrtf(bc2);
  for (int i0 = 0; i0 < n; i0 += kSegs) {
    AdamSegs segs{};
    const int cnt = (n - i0 < kSegs) ? n - i0 : kSegs;
    long long acc = 0, elems = 0;
    for (int i = 0; i < cnt; ++i) {
      SER_REQUIRE(counts[i0 + i] > 0 && p[i0 + i] && g[i0 + i] && m[i0 + i] && v[i0 + i], "adamw: null / empty tensor");
      segs.p[i] = p[i0 + i]; segs.g[i] = g[i0 + i]; segs.m[i] = m[i0 + i]; segs.v[i] = v[i0 + i];
      segs.len[i] = counts[i0 + i];
      acc += (counts[i0 + i] + 3) / 4;
      segs.end[i] = acc;
      elems += counts[i0 + i];
    }
    segs.n = cnt;
    ProfScope prof("adamw", 12.0 * elems, 28.0 * elems, s);
    const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (acc + 255) / 256));
    SER_CUDA_CHECK(launch_pdl(adamw_multi_kernel, dim3(blocks), dim3(256), 0, s, segs, lr, beta1, beta2, eps, weight_decay, bc1, rsqrt_bc2, gscale, grad_scale,
                                              found_inf, step_dev));
    SER_LAUNCH_CHECK();
  }
  return SER_OK;
}

int grad_clip_coef(int n, const float* const* g, const long long* counts, float max_norm, float* scratch, float* coef,
                   float* norm_out, cudaStream_t s) {
  SER_REQUIRE(n >= 1 && scratch != nullptr && coef != nullptr, "grad_clip: null buffer");
  SER_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(float), s));
  for (int i0 = 0; i0 < n; i0 += kSegs) {
    NormSegs segs{};
    const int cnt = (n - i0 < kSegs) ? n - i0 : kSegs;
    long long acc = 0, elems = 0;
    for (int i = 0; i < cnt; ++i) {
      SER_REQUIRE(counts[i0 + i] > 0 && g[i0 + i], "grad_clip: null / empty tensor");
      segs.g[i] = g[i0 + i]; segs.len[i] = counts[i0 + i];
      acc += (counts[i0 + i] + 3) / 4;
      segs.end[i] = acc;
      elems += counts[i0 + i];
    }
    segs.n = cnt;
    ProfScope prof("grad_norm", 2.0 * elems, 4.0 * elems, s);
    const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (acc + 255) / 256));
    SER_CUDA_CHECK(launch_pdl(sumsq_multi_kernel, dim3(blocks), dim3(256), 0, s, segs, scratch));
    SER_LAUNCH_CHECK();
  }
  SER_CUDA_CHECK(launch_pdl(clip_coef_kernel, dim3(1), dim3(1), 0, s, scratch, max_norm, coef, norm_out));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

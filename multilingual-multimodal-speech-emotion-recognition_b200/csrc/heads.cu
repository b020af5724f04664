// Classifier heads on the 256-d penultimate features (reference: src/models/classifier.py:192-198 uncertainty_head,
// :224 output_projection[4], :229): logits = W_c f + b_c;  unc = sigmoid(w_u2 . relu(W_u1 f + b_u1) + b_u2).
// C + 64 + 1 outputs per sample: as GEMMs these are three launches of a few CTAs each in either tier; here they are
// one fp32 kernel forward and two backward (row-wise gradients, then the weight gradients), always on the fp32
// master weights -- the logits decide argmax, so they are never computed in reduced precision.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

// one CTA per sample; warp per output (C + U dot products of length F against the staged feature row).
// The weights do not depend on the preceding kernels of the step, so the CTA copies them into shared memory BEFORE
// griddepcontrol.wait (programmatic dependent launch: this prologue runs while the predecessor still executes) with all
// loads in flight at once -- the fp32 master weights are cold by this point of a step (one DRAM round trip instead of a
// chain of eight per output row); only the feature row is read after the wait.
template <bool STAGE_W>
__global__ void __launch_bounds__(128)
heads_fwd_kernel(const float* __restrict__ f, const float* __restrict__ w_c, const float* __restrict__ b_c,
                 const float* __restrict__ w_u1, const float* __restrict__ b_u1, const float* __restrict__ w_u2,
                 const float* __restrict__ b_u2, float* __restrict__ logits, float* __restrict__ u1,
                 float* __restrict__ unc, int F, int C, int U, DropSpec drop) {
  extern __shared__ __align__(16) float sm[];
  float* sf = sm;            // [F]
  float* su = sm + F;        // [U]
  float* sw = su + ((U + 3) & ~3);     // [C + U][F] (STAGE_W)
  const int nout = C + (unc != nullptr ? U : 0);
  if (STAGE_W) {
    const int nc4 = C * F / 4, nu4 = (nout - C) * F / 4;
    for (int i = threadIdx.x; i < nc4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sw + 4 * i))),
                   "l"(w_c + 4 * i) : "memory");
    for (int i = threadIdx.x; i < nu4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sw + C * F + 4 * i))),
                   "l"(w_u1 + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  pdl_sync();
  const int row = blockIdx.x;
  for (int k = threadIdx.x * 4; k < F; k += blockDim.x * 4)
    *reinterpret_cast<float4*>(sf + k) = __ldg(reinterpret_cast<const float4*>(f + static_cast<size_t>(row) * F + k));
  if (STAGE_W) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll 4
  for (int o = warp; o < nout; o += nwarps) {
    const float* w = STAGE_W ? sw + static_cast<size_t>(o) * F
                             : ((o < C) ? w_c + static_cast<size_t>(o) * F : w_u1 + static_cast<size_t>(o - C) * F);
    float acc = 0.f;
    for (int k = lane * 4; k < F; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(sf + k);
      const float4 b = STAGE_W ? *reinterpret_cast<const float4*>(w + k) : __ldg(reinterpret_cast<const float4*>(w + k));
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    acc = warp_sum(acc);
    if (lane != 0) continue;
    if (o < C) {
      logits[static_cast<size_t>(row) * C + o] = acc + b_c[o];
    } else {
      float v = fmaxf(acc + b_u1[o - C], 0.f);
      if (drop.on()) {                                   // uncertainty_head[2] (classifier.py:195); u1 is saved post-dropout
        const unsigned j = static_cast<unsigned>(o - C);
        v *= drop_one(drop_key(drop), static_cast<unsigned>(row) * ((U + 1) / 2) + (j >> 1), j & 1u, drop.thr, drop.scale);
      }
      su[o - C] = v;
      u1[static_cast<size_t>(row) * U + (o - C)] = v;
    }
  }
  if (unc == nullptr) return;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int j = threadIdx.x; j < U; j += 32) s = fmaf(su[j], w_u2[j], s);
    s = warp_sum(s);
    if (threadIdx.x == 0) unc[row] = 1.f / (1.f + expf(-(s + b_u2[0])));
  }
}

// row-wise backward: g[o] = gradient at the C + U pre-activations of this sample; df = g W; also saves du1, dsg
__global__ void __launch_bounds__(256)
heads_bwd_rows_kernel(const float* __restrict__ dlogits, const float* __restrict__ dunc, const float* __restrict__ unc,
                      const float* __restrict__ u1, const float* __restrict__ w_c, const float* __restrict__ w_u1,
                      const float* __restrict__ w_u2, float* __restrict__ df, float* __restrict__ du1,
                      float* __restrict__ dsg, int F, int C, int U, float uscale) {
  pdl_sync();
  extern __shared__ float sg[];       // [C + U]
  const int row = blockIdx.x;
  const bool have_u = (dunc != nullptr) && (unc != nullptr);
  float ds = 0.f;
  if (have_u) {
    const float y = unc[row];
    ds = dunc[row] * y * (1.f - y);
  }
  for (int o = threadIdx.x; o < C + U; o += blockDim.x) {
    float g;
    if (o < C) g = (dlogits != nullptr) ? dlogits[static_cast<size_t>(row) * C + o] : 0.f;
    else {
      const int j = o - C;
      // u1 is the post-dropout activation: > 0 exactly where the unit was kept AND the ReLU was open
      g = (have_u && u1[static_cast<size_t>(row) * U + j] > 0.f) ? ds * w_u2[j] * uscale : 0.f;
      du1[static_cast<size_t>(row) * U + j] = g;
    }
    sg[o] = g;
  }
  if (threadIdx.x == 0) dsg[row] = ds;
  __syncthreads();
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < C; ++o) acc = fmaf(sg[o], __ldg(w_c + static_cast<size_t>(o) * F + k), acc);
    if (have_u)
      for (int j = 0; j < U; ++j) acc = fmaf(sg[C + j], __ldg(w_u1 + static_cast<size_t>(j) * F + k), acc);
    df[static_cast<size_t>(row) * F + k] = acc;
  }
}

// weight gradients: CTA o < C + U owns row o of [dW_c ; dW_u1] (+ its bias gradient); CTA C + U owns dw_u2 / db_u2.
// Rows of the batch are walked in chunks staged through shared memory (gradient scalars) -- f is read coalesced.
__global__ void __launch_bounds__(256)
heads_bwd_w_kernel(const float* __restrict__ dlogits, const float* __restrict__ du1, const float* __restrict__ dsg,
                   const float* __restrict__ f, const float* __restrict__ u1, float* __restrict__ dw_c,
                   float* __restrict__ db_c, float* __restrict__ dw_u1, float* __restrict__ db_u1,
                   float* __restrict__ dw_u2, float* __restrict__ db_u2, int B, int F, int C, int U) {
  pdl_sync();
  __shared__ float sgr[256];
  __shared__ float red[32];
  const int o = blockIdx.x;
  const bool tail = (o == C + U);                     // dw_u2[j] = sum_rows dsg * u1[:, j]
  const int width = tail ? U : F;
  const float* x = tail ? u1 : f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                // up to 4 columns per thread (width <= 1024)
  float bsum = 0.f;
  for (int r0 = 0; r0 < B; r0 += 256) {
    const int r = r0 + threadIdx.x;
    float g = 0.f;
    if (r < B) {
      if (tail) g = dsg[r];
      else if (o < C) g = (dlogits != nullptr) ? dlogits[static_cast<size_t>(r) * C + o] : 0.f;
      else g = du1[static_cast<size_t>(r) * U + (o - C)];
    }
    __syncthreads();
    sgr[threadIdx.x] = g;
    bsum += g;
    __syncthreads();
    const int nr = min(256, B - r0);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k = threadIdx.x + c * 256;
      if (k < width) {
        // the loads are independent of the FMA chain: a deep unroll keeps 16 of them in flight per thread (the kernel
        // runs ~70 CTAs, so memory-level parallelism per thread is what hides the L2 latency)
        float a = acc[c];
        const float* xp = x + static_cast<size_t>(r0) * width + k;
        if (nr == 256) {
#pragma unroll 16
          for (int i = 0; i < 256; ++i) a = fmaf(sgr[i], xp[static_cast<size_t>(i) * width], a);
        } else {
          for (int i = 0; i < nr; ++i) a = fmaf(sgr[i], xp[static_cast<size_t>(i) * width], a);
        }
        acc[c] = a;
      }
    }
  }
  float* dst = tail ? dw_u2 : (o < C ? dw_c + static_cast<size_t>(o) * F : dw_u1 + static_cast<size_t>(o - C) * F);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int k = threadIdx.x + c * 256;
    if (k < width) dst[k] = acc[c];
  }
  bsum = block_sum(bsum, red);
  if (threadIdx.x == 0) {
    if (tail) db_u2[0] = bsum;
    else if (o < C) db_c[o] = bsum;
    else db_u1[o - C] = bsum;
  }
}

}  // namespace

int heads_fwd(const float* f, const float* w_c, const float* b_c, const float* w_u1, const float* b_u1,
              const float* w_u2, const float* b_u2, float* logits, float* u1, float* unc, int B, int F, int C, int U,
              const DropSpec& drop, cudaStream_t s) {
  SER_REQUIRE(B > 0 && F % 4 == 0 && F <= 4096 && C > 0 && U > 0 && U <= 1024, "heads_fwd: unsupported shape");
  ProfScope prof("heads_fwd", 2.0 * B * F * (C + U), 4.0 * (static_cast<double>(B) * F + (C + U) * F), s);
  // weights staged in shared memory when they fit beside the feature row (reference sizes: 68 x 256 fp32 = 68 KB)
  const size_t base = sizeof(float) * (F + ((U + 3) & ~3));
  const size_t staged = base + sizeof(float) * static_cast<size_t>(C + U) * F;
  const bool ok16 = (reinterpret_cast<uintptr_t>(w_c) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_u1) & 15) == 0;
  if (staged <= 96 * 1024 && ok16) {
    static bool configured = false;
    if (!configured) {
      SER_CUDA_CHECK(cudaFuncSetAttribute(heads_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      configured = true;
    }
    SER_CUDA_CHECK(launch_pdl(heads_fwd_kernel<true>, dim3(B), dim3(128), staged, s, f, w_c, b_c, w_u1, b_u1, w_u2, b_u2, logits, u1, unc, F, C, U, drop));
  } else {
    SER_CUDA_CHECK(launch_pdl(heads_fwd_kernel<false>, dim3(B), dim3(128), base, s, f, w_c, b_c, w_u1, b_u1, w_u2, b_u2, logits, u1, unc, F, C, U, drop));
  }
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int heads_bwd(const float* dlogits, const float* dunc, const float* unc, const float* u1, const float* f,
              const float* w_c, const float* w_u1, const float* w_u2, float* df, float* du1, float* dsg, float* dw_c,
              float* db_c, float* dw_u1, float* db_u1, float* dw_u2, float* db_u2, int B, int F, int C, int U,
              const DropSpec& drop, cudaStream_t s, SideBranch* sb) {
  SER_REQUIRE(B > 0 && F <= 1024 && U <= 1024 && C > 0, "heads_bwd: unsupported shape");
  ProfScope prof("heads_bwd", 4.0 * B * F * (C + U), 4.0 * (2.0 * B * F + 2.0 * (C + U) * F), s);
  SER_CUDA_CHECK(launch_pdl(heads_bwd_rows_kernel, dim3(B), dim3(256), sizeof(float) * (C + U), s, dlogits, dunc, unc, u1, w_c, w_u1, w_u2, df, du1, dsg,
                                                              F, C, U, drop.on() ? drop.scale : 1.f));
  SER_LAUNCH_CHECK();
  // the weight gradients are leaves: on the side branch they run beside the dX chain that continues on `s`
  cudaStream_t sw = s;
  if (sb != nullptr) { SER_TRY(sb->fork()); sw = sb->side(); }
  SER_CUDA_CHECK(launch_pdl(heads_bwd_w_kernel, dim3(C + U + 1), dim3(256), 0, sw, dlogits, du1, dsg, f, u1, dw_c, db_c, dw_u1, db_u1, dw_u2, db_u2, B, F, C, U));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

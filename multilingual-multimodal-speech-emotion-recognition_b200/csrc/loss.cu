// Training-loss kernels: label-smoothed CE (src/models/losses.py:12-30), class-balanced focal loss
// (losses.py:41-64), the uncertainty regulariser of src/train.py:161-163 and the PrototypeMemory pull/push
// loss (src/models/prototypes.py:13-53), with the composition of train.py:154-168.
//
// One warp per sample; logits rows are tiny (C <= 32) so lane c owns class c.  Batch reductions go
// through block-level shared memory and one atomic per CTA.  The per-term "non-finite => return a fresh
// zero" guards of the reference are evaluated on the (optionally all-reduced) batch sums, so that in
// data-parallel runs every rank takes the same branch.
#include "kernels.cuh"
#include "prof.cuh"

namespace ser {

namespace {

constexpr int kMaxC = 32;

__global__ void loss_prep_kernel(const long long* __restrict__ labels, const float* __restrict__ counts_in, int B,
                                 int C, float beta, int use_weights, float* __restrict__ class_w,
                                 float* __restrict__ sums) {
  pdl_sync();
  __shared__ float cnt[kMaxC];
  __shared__ float w[kMaxC];
  if (threadIdx.x < kMaxC) cnt[threadIdx.x] = 0.f;
  if (threadIdx.x < LS_COUNT) sums[threadIdx.x] = 0.f;
  __syncthreads();
  if (counts_in != nullptr) {
    if (threadIdx.x < C) cnt[threadIdx.x] = counts_in[threadIdx.x];
  } else {
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const long long y = labels[i];
      if (y >= 0 && y < C) atomicAdd(&cnt[static_cast<int>(y)], 1.f);
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    if (use_weights) {
      // effective-number weights (losses.py:44-49); pow evaluated in double and rounded once
      const float c = fmaxf(cnt[threadIdx.x], 1.f);
      const float pw = static_cast<float>(pow(static_cast<double>(beta), static_cast<double>(c)));
      const float eff = fmaxf(1.f - pw, 1e-6f);
      w[threadIdx.x] = (1.f - beta) / eff;
    } else {
      w[threadIdx.x] = 1.f;
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    if (use_weights) {
      float tot = 0.f;
      for (int c = 0; c < C; ++c) tot += w[c];
      class_w[threadIdx.x] = w[threadIdx.x] / (tot + 1e-8f) * static_cast<float>(C);
    } else {
      class_w[threadIdx.x] = 1.f;
    }
  }
}

struct RowSoftmax {
  float z;        // clamped logit of this lane's class (or -inf for lanes >= C)
  float logp;     // log-softmax
  float p;
  bool inside;    // |logit| <= 10 -> clamp passes gradient
};

__device__ __forceinline__ RowSoftmax row_softmax(const float* __restrict__ logits, int C, int lane) {
  RowSoftmax r;
  const float raw = (lane < C) ? logits[lane] : 0.f;
  r.inside = (raw >= -10.f && raw <= 10.f);
  r.z = (lane < C) ? fminf(fmaxf(raw, -10.f), 10.f) : -INFINITY;
  const float mx = warp_max(r.z);
  const float ex = (lane < C) ? expf(r.z - mx) : 0.f;
  const float lse = mx + logf(warp_sum(ex));
  r.logp = r.z - lse;
  r.p = (lane < C) ? expf(r.logp) : 0.f;
  return r;
}

__device__ __forceinline__ int row_argmax(const float* __restrict__ logits, int C, int lane) {
  // first index of the maximum, as torch.argmax
  float v = (lane < C) ? logits[lane] : -INFINITY;
  int idx = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  return idx;
}

// The lane's share of one embedding row (columns lane, lane + 32, ...), read ONCE and clamped (prototypes.py:18): the
// class loop below then runs on registers.  kEmbRegs * 32 columns fit (D <= 512, the reference's proj_dim); wider
// rows take the generic loop.
constexpr int kEmbRegs = 16;
__device__ __forceinline__ void load_emb_lane(const LossArgs& a, int row, int lane, float (&raw)[kEmbRegs], float (&e)[kEmbRegs]) {
#pragma unroll
  for (int k = 0; k < kEmbRegs; ++k) {
    const int d = lane + 32 * k;
    raw[k] = (d < a.D) ? ld_dyn(a.emb, static_cast<size_t>(row) * a.D + d, a.emb_f32) : 0.f;
    e[k] = fminf(fmaxf(raw[k], -10.f), 10.f);
  }
}
// squared distance of the row to prototype c, same summation order as the generic loop (lane-strided, then warp_sum)
__device__ __forceinline__ float sqdist_lane(const LossArgs& a, const float* __restrict__ protos, const float (&e)[kEmbRegs],
                                             int c, int lane) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kEmbRegs; ++k) {
    const int d = lane + 32 * k;
    if (d < a.D) {
      const float df = e[k] - protos[static_cast<size_t>(c) * a.D + d];
      s = fmaf(df, df, s);
    }
  }
  return warp_sum(s);
}

__global__ void __launch_bounds__(256)
loss_rows_kernel(LossArgs a, int stage_protos) {
  // the prototypes are parameters (nothing in this step writes them): copied to shared memory BEFORE
  // griddepcontrol.wait, while the preceding kernel still runs -- by now they are cold in L2
  extern __shared__ float sproto[];
  const float* protos = a.protos;
  if (stage_protos && a.emb != nullptr) {
    for (int i = threadIdx.x; i < a.C * a.D; i += blockDim.x) sproto[i] = __ldg(a.protos + i);
    protos = sproto;
  }
  pdl_sync();
  __shared__ float acc[LS_COUNT];
  if (threadIdx.x < LS_COUNT) acc[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < a.B) {
    long long yl = a.labels[row];
    const int y = static_cast<int>(yl < 0 ? 0 : (yl > a.C - 1 ? a.C - 1 : yl));   // losses.py:15 clamp
    const float* lg = a.logits + static_cast<size_t>(row) * a.C;
    const RowSoftmax r = row_softmax(lg, a.C, lane);
    // label-smoothed CE
    const float q = (lane == y) ? 1.f - a.smoothing : a.smoothing / static_cast<float>(a.C - 1);
    float ce = (lane < a.C) ? -q * r.logp : 0.f;
    ce = warp_sum(ce);
    // focal
    const float logp_y = __shfl_sync(0xffffffffu, r.logp, y);
    const float p_y = __shfl_sync(0xffffffffu, r.p, y);
    const float pt = fminf(fmaxf(p_y, 1e-6f), 1.f);
    const float fw = powf(1.f - pt, a.gamma);
    const float focal = fw * (-a.class_w[y] * logp_y);
    const int am = row_argmax(lg, a.C, lane);
    const float correct = (static_cast<long long>(am) == yl) ? 1.f : 0.f;
    float pos = 0.f, neg = 0.f;
    if (a.emb != nullptr) {
      // squared distances to every prototype; lane c keeps class c
      float mysq = 0.f;
      if (a.D <= 32 * kEmbRegs) {
        float raw[kEmbRegs], ev[kEmbRegs];
        load_emb_lane(a, row, lane, raw, ev);
        for (int c = 0; c < a.C; ++c) {
          const float s = sqdist_lane(a, protos, ev, c, lane);
          if (lane == c) mysq = s;
        }
      } else {
        for (int c = 0; c < a.C; ++c) {
          float s = 0.f;
          for (int d = lane; d < a.D; d += 32) {
            float e = ld_dyn(a.emb, static_cast<size_t>(row) * a.D + d, a.emb_f32);
            e = fminf(fmaxf(e, -10.f), 10.f);
            const float df = e - protos[static_cast<size_t>(c) * a.D + d];
            s = fmaf(df, df, s);
          }
          s = warp_sum(s);
          if (lane == c) mysq = s;
        }
      }
      pos = sqrtf(__shfl_sync(0xffffffffu, mysq, y));
      float nd = (lane < a.C) ? ((lane == y) ? 10.f : fminf(sqrtf(mysq + 1e-6f), 10.f)) : INFINITY;
      const float mn = -warp_max(-nd);                         // min over classes
      const float ex = (lane < a.C) ? expf(-(nd - mn)) : 0.f;
      neg = -(-mn + logf(warp_sum(ex)));                       // -logsumexp(-nd)
    }
    if (lane == 0) {
      atomicAdd(&acc[LS_CE], ce);
      atomicAdd(&acc[LS_FOCAL], focal);
      atomicAdd(&acc[LS_CORRECT], correct);
      if (a.unc != nullptr) atomicAdd(&acc[LS_UNC], a.unc[row]);
      atomicAdd(&acc[LS_POS], pos);
      atomicAdd(&acc[LS_NEG], neg);
    }
  }
  __syncthreads();
  if (threadIdx.x < LS_COUNT && acc[threadIdx.x] != 0.f) atomicAdd(a.sums + threadIdx.x, acc[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
loss_bwd_kernel(LossArgs a, const float* __restrict__ gscale, int stage_dprotos) {
  // prototype gradients of the CTA's 8 samples are summed in shared memory first ([C, D] floats, when that fits): one
  // global atomic per (class, column) and CTA instead of one per sample -- B-way contention on C*D addresses otherwise.
  // The second [C, D] block holds the prototypes themselves, copied before griddepcontrol.wait (see loss_rows_kernel).
  extern __shared__ float sdp[];
  const bool staged = stage_dprotos != 0 && a.dprotos != nullptr && a.emb != nullptr && a.demb != nullptr;
  const float* protos = a.protos;
  if (stage_dprotos != 0 && a.emb != nullptr) {
    float* sproto = sdp + a.C * a.D;
    for (int i = threadIdx.x; i < a.C * a.D; i += blockDim.x) sproto[i] = __ldg(a.protos + i);
    protos = sproto;
  }
  pdl_sync();
  if (staged) {
    for (int i = threadIdx.x; i < a.C * a.D; i += blockDim.x) sdp[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < a.B) {
    const float gs = gscale ? gscale[0] : 1.f;
    const float invB = 1.f / static_cast<float>(a.B_global);
    const bool ok_ce = isfinite(a.sums[LS_CE] * invB);
    const bool ok_focal = isfinite(a.sums[LS_FOCAL] * invB);
    const bool ok_proto = isfinite(a.sums[LS_POS] * invB + a.margin - a.sums[LS_NEG] * invB);
    long long yl = a.labels[row];
    const int y = static_cast<int>(yl < 0 ? 0 : (yl > a.C - 1 ? a.C - 1 : yl));
    const float* lg = a.logits + static_cast<size_t>(row) * a.C;
    const RowSoftmax r = row_softmax(lg, a.C, lane);
    if (a.dlogits != nullptr) {
      float g = 0.f;
      if (lane < a.C) {
        const float delta = (lane == y) ? 1.f : 0.f;
        if (ok_ce && a.w_ce != 0.f) {
          const float q = (lane == y) ? 1.f - a.smoothing : a.smoothing / static_cast<float>(a.C - 1);
          g += a.w_ce * (r.p - q);
        }
      }
      const float logp_y = __shfl_sync(0xffffffffu, r.logp, y);
      const float p_y = __shfl_sync(0xffffffffu, r.p, y);
      if (lane < a.C && ok_focal && a.w_focal != 0.f) {
        const float delta = (lane == y) ? 1.f : 0.f;
        const bool in_range = (p_y >= 1e-6f && p_y <= 1.f);
        const float pt = fminf(fmaxf(p_y, 1e-6f), 1.f);
        const float wy = a.class_w[y];
        const float fw = powf(1.f - pt, a.gamma);
        const float dfw_dpt = in_range ? -a.gamma * powf(1.f - pt, a.gamma - 1.f) : 0.f;
        const float dpt_dz = p_y * (delta - r.p);
        const float ce_w = -wy * logp_y;
        const float dce_dz = wy * (r.p - delta);
        g += a.w_focal * (dfw_dpt * dpt_dz * ce_w + fw * dce_dz);
      }
      if (lane < a.C) a.dlogits[static_cast<size_t>(row) * a.C + lane] = r.inside ? g * invB * gs : 0.f;
    }
    if (a.dunc != nullptr && lane == 0) {
      const float mean_correct = a.sums[LS_CORRECT] * invB;
      a.dunc[row] = a.w_unc * mean_correct * invB * gs;
    }
    if (a.emb != nullptr && a.demb != nullptr) {
      const bool on = ok_proto && a.w_proto != 0.f;
      // recompute distances (the lane's share of the row stays in registers for the gradient loop below)
      const bool in_regs = a.D <= 32 * kEmbRegs;
      float rawv[kEmbRegs], ev[kEmbRegs];
      float mysq = 0.f;
      if (in_regs) {
        load_emb_lane(a, row, lane, rawv, ev);
        for (int c = 0; c < a.C; ++c) {
          const float s = sqdist_lane(a, protos, ev, c, lane);
          if (lane == c) mysq = s;
        }
      } else {
        for (int c = 0; c < a.C; ++c) {
          float s = 0.f;
          for (int d = lane; d < a.D; d += 32) {
            float e = ld_dyn(a.emb, static_cast<size_t>(row) * a.D + d, a.emb_f32);
            e = fminf(fmaxf(e, -10.f), 10.f);
            const float df = e - protos[static_cast<size_t>(c) * a.D + d];
            s = fmaf(df, df, s);
          }
          s = warp_sum(s);
          if (lane == c) mysq = s;
        }
      }
      const float pos = sqrtf(__shfl_sync(0xffffffffu, mysq, y));
      const float dist = sqrtf(mysq + 1e-6f);
      const float nd = (lane < a.C) ? ((lane == y) ? 10.f : fminf(dist, 10.f)) : INFINITY;
      const float mn = -warp_max(-nd);
      const float ex = (lane < a.C) ? expf(-(nd - mn)) : 0.f;
      const float sm = ex / warp_sum(ex);                          // softmax(-nd)
      // coefficient of (e - P_c) for each class: own class +1/pos ; others -s_c/d_c when d_c <= 10
      float coef = 0.f;
      if (lane < a.C) {
        if (lane == y) coef = pos > 0.f ? 1.f / pos : 0.f;
        else if (dist <= 10.f) coef = -sm / dist;
      }
      const float k = on ? a.w_proto * invB * gs : 0.f;
      auto grad_col = [&](int d, float raw) {
        const float e = fminf(fmaxf(raw, -10.f), 10.f);
        const bool inside = (raw >= -10.f && raw <= 10.f);
        float ge = 0.f;
        for (int c = 0; c < a.C; ++c) {
          const float cc = __shfl_sync(0xffffffffu, coef, c);
          const float df = e - protos[static_cast<size_t>(c) * a.D + d];
          const float t = cc * df * k;
          ge += t;
          if (a.dprotos != nullptr && t != 0.f) {
            if (staged) atomicAdd(&sdp[c * a.D + d], -t);
            else atomicAdd(a.dprotos + static_cast<size_t>(c) * a.D + d, -t);
          }
        }
        st_dyn(a.demb, static_cast<size_t>(row) * a.D + d, a.demb_f32, inside ? ge : 0.f);
      };
      if (in_regs) {
        // (the shuffles inside grad_col need the whole warp: D is a multiple of 32 here or the tail lanes idle together)
#pragma unroll
        for (int kk = 0; kk < kEmbRegs; ++kk) {
          if (32 * kk < a.D) {                               // warp-uniform
            const int d = lane + 32 * kk;
            if (d < a.D) grad_col(d, rawv[kk]);
            else for (int c = 0; c < a.C; ++c) (void)__shfl_sync(0xffffffffu, coef, c);
          }
        }
      } else {
        for (int d = lane; d < a.D; d += 32) grad_col(d, ld_dyn(a.emb, static_cast<size_t>(row) * a.D + d, a.emb_f32));
      }
    }
  }  // row < B
  if (staged) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.C * a.D; i += blockDim.x) {
      const float v = sdp[i];
      if (v != 0.f) atomicAdd(a.dprotos + i, v);
    }
  }
}

}  // namespace

int loss_fwd(const LossArgs& a, cudaStream_t s) {
  SER_REQUIRE(a.C >= 2 && a.C <= kMaxC, "loss: 2 <= num_classes <= 32");
  SER_REQUIRE(a.B > 0, "loss: empty batch");
  ProfScope prof("loss_fwd", 0.0, 4.0 * a.B * (a.C + (a.emb ? a.D : 0)), s);
  SER_CUDA_CHECK(launch_pdl(loss_prep_kernel, dim3(1), dim3(256), 0, s, a.labels, a.counts, a.B, a.C, a.beta, a.focal_use_weights, a.class_w, a.sums));
  SER_LAUNCH_CHECK();
  const size_t pbytes = sizeof(float) * static_cast<size_t>(a.C) * (a.emb ? a.D : 0);
  const int stage_p = (pbytes > 0 && pbytes <= 20 * 1024) ? 1 : 0;
  SER_CUDA_CHECK(launch_pdl(loss_rows_kernel, dim3(ceil_div(a.B, 8)), dim3(256), stage_p ? pbytes : 0, s, a, stage_p));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int loss_bwd_scaled(const LossArgs& a, const float* gscale, cudaStream_t s) {
  SER_REQUIRE(a.C >= 2 && a.C <= kMaxC, "loss: 2 <= num_classes <= 32");
  ProfScope prof("loss_bwd", 0.0, 4.0 * a.B * (2.0 * a.C + (a.emb ? 2.0 * a.D : 0)), s);
  const size_t stage_bytes = sizeof(float) * static_cast<size_t>(a.C) * (a.emb ? a.D : 0);
  const int stage = (a.dprotos != nullptr && stage_bytes > 0 && stage_bytes <= 20 * 1024) ? 1 : 0;
  SER_CUDA_CHECK(launch_pdl(loss_bwd_kernel, dim3(ceil_div(a.B, 8)), dim3(256), stage ? 2 * stage_bytes : 0, s, a, gscale, stage));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int loss_bwd(const LossArgs& a, cudaStream_t s) { return loss_bwd_scaled(a, nullptr, s); }

namespace {
// terms = {ce, focal, unc_loss, proto, total, accuracy}; non-finite terms are replaced by 0 as the
// reference does (losses.py:28-29,62-63; prototypes.py:51-52)
__global__ void loss_finalize_kernel(const float* __restrict__ sums, long long B_global, float margin, float w_ce,
                                     float w_focal, float w_unc, float w_proto, int have_proto,
                                     float* __restrict__ terms) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float invB = 1.f / static_cast<float>(B_global);
  float ce = sums[LS_CE] * invB;
  float focal = sums[LS_FOCAL] * invB;
  const float acc = sums[LS_CORRECT] * invB;
  const float unc = sums[LS_UNC] * invB * acc;
  float proto = have_proto ? sums[LS_POS] * invB + margin - sums[LS_NEG] * invB : 0.f;
  if (!isfinite(ce)) ce = 0.f;
  if (!isfinite(focal)) focal = 0.f;
  if (!isfinite(proto)) proto = 0.f;
  terms[0] = ce; terms[1] = focal; terms[2] = unc; terms[3] = proto;
  terms[4] = w_ce * ce + w_focal * focal + w_unc * unc + w_proto * proto;
  terms[5] = acc;
}
}  // namespace

int loss_finalize(const float* sums, long long B_global, float margin, float w_ce, float w_focal, float w_unc,
                  float w_proto, int have_proto, float* terms, cudaStream_t s) {
  SER_CUDA_CHECK(launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), 0, s, sums, B_global, margin, w_ce, w_focal, w_unc, w_proto, have_proto, terms));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

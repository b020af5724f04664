// Masked multi-head cross attention core (QK^T -> masked softmax -> PV), forward and backward.
// Semantics follow nn.MultiheadAttention's math path as used by cross_attention.py:41,49
// (torch/nn/functional.py:6609-6645): q is scaled by 1/sqrt(dh) after its bias, key padding is an
// additive -inf, softmax over keys, attention weights themselves are never materialised (the
// reference discards them).  A sample whose keys are ALL padded produces NaN for every query.
//
// This file holds the streaming-softmax CUDA-core implementation (fp32 math, fp32 or bf16 storage):
// one thread owns one query row (dh = 32 values in registers), K/V tiles are staged in shared memory.
// The score matrix never touches HBM; only the per-row log-sum-exp is saved for the backward pass.
#include "kernels.cuh"
#include <stdlib.h>
#include "prof.cuh"

namespace ser {

namespace {

constexpr int DH = 32;        // head dim of the reference configuration (256 / 8)
constexpr int KT = 64;        // keys (or queries, in the dK/dV kernel) per shared-memory tile
constexpr int NT = 128;       // threads per block

template <typename T>
__device__ __forceinline__ void load_row32(const T* p, float (&v)[DH]) {
#pragma unroll
  for (int i = 0; i < DH; i += 8) {
    float t[8];
    load8(p + i, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i + j] = t[j];
  }
}
template <typename T>
__device__ __forceinline__ void store_row32(T* p, const float (&v)[DH]) {
#pragma unroll
  for (int i = 0; i < DH; i += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[i + j];
    store8(p + i, t);
  }
}

// cooperative load of `rows` x DH elements (row stride ld) into smem [KT][DH] as fp32, zero padded
template <typename T>
__device__ __forceinline__ void stage_tile(const T* __restrict__ g, long long ld, int rows, float (*sm)[DH]) {
  for (int e = threadIdx.x; e < KT * (DH / 8); e += blockDim.x) {
    const int r = e / (DH / 8), c = (e % (DH / 8)) * 8;
    float t[8];
    if (r < rows) load8(g + static_cast<size_t>(r) * ld + c, t);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[r][c + j] = t[j];
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
attn_fwd_kernel(const T* __restrict__ Q, long long ldq, const T* __restrict__ K, long long ldk,
                const T* __restrict__ V, long long ldv, const float* __restrict__ kmask, T* __restrict__ O,
                long long ldo, float* __restrict__ lse, int H, int Tq, int Tk, float scale, DropSpec drop) {
  pdl_sync();
  __shared__ float sK[KT][DH];
  __shared__ float sV[KT][DH];
  __shared__ float sMask[KT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * NT + threadIdx.x;
  const bool active = qi < Tq;
  DropKey dkey{0u, 1u};
  if (drop.on()) dkey = drop_key(drop);
  const unsigned drow = ((static_cast<unsigned>(b) * H + h) * Tq + qi) * static_cast<unsigned>((Tk + 1) / 2);

  float q[DH], acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  if (active) {
    load_row32(Q + (static_cast<size_t>(b) * Tq + qi) * ldq + h * DH, q);
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] *= scale;
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < Tk; k0 += KT) {
    const int rows = min(KT, Tk - k0);
    __syncthreads();
    stage_tile(K + (static_cast<size_t>(b) * Tk + k0) * ldk + h * DH, ldk, rows, sK);
    stage_tile(V + (static_cast<size_t>(b) * Tk + k0) * ldv + h * DH, ldv, rows, sV);
    for (int j = threadIdx.x; j < KT; j += blockDim.x)
      sMask[j] = (j < rows && (kmask == nullptr || kmask[static_cast<size_t>(b) * Tk + k0 + j] != 0.f)) ? 1.f : 0.f;
    __syncthreads();
    for (int j0 = 0; j0 < rows; j0 += 8) {
      float s[8];
      float gmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(&sK[j][d]);
          dot = fmaf(q[d], kv.x, dot); dot = fmaf(q[d + 1], kv.y, dot);
          dot = fmaf(q[d + 2], kv.z, dot); dot = fmaf(q[d + 3], kv.w, dot);
        }
        s[jj] = (j < rows && sMask[j] != 0.f) ? dot : -INFINITY;
        gmax = fmaxf(gmax, s[jj]);
      }
      const float m_new = fmaxf(m, gmax);
      if (m_new == -INFINITY) continue;              // nothing but padded keys so far
      const float corr = __expf(m - m_new);          // m = -inf -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float p = __expf(s[jj] - m_new);             // masked -> exp(-inf) = 0
        l += p;                                      // the softmax normaliser sees every key; dropout acts on the weights
        const int j = j0 + jj;
        if (drop.on()) p *= drop_one(dkey, drow + ((k0 + j) >> 1), (k0 + j) & 1, drop.thr, drop.scale);
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(&sV[j][d]);
          acc[d] = fmaf(p, vv.x, acc[d]); acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
          acc[d + 2] = fmaf(p, vv.z, acc[d + 2]); acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
        }
      }
      m = m_new;
    }
  }
  if (active) {
    const float inv = (l > 0.f) ? 1.f / l : __int_as_float(0x7fc00000);   // all keys padded -> NaN (reference)
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] *= inv;
    store_row32(O + (static_cast<size_t>(b) * Tq + qi) * ldo + h * DH, acc);
    if (lse != nullptr) lse[(static_cast<size_t>(b) * H + h) * Tq + qi] = (l > 0.f) ? m + logf(l) : __int_as_float(0x7fc00000);
  }
}

// dQ: one thread per query.  Also emits delta = rowsum(dO * O) for the dK/dV kernel.
template <typename T>
__global__ void __launch_bounds__(NT)
attn_bwd_dq_kernel(const T* __restrict__ Q, long long ldq, const T* __restrict__ K, long long ldk,
                   const T* __restrict__ V, long long ldv, const float* __restrict__ kmask,
                   const T* __restrict__ O, long long ldo, const T* __restrict__ dO, long long lddo,
                   const float* __restrict__ lse, float* __restrict__ delta, T* __restrict__ dQ, long long lddq,
                   int H, int Tq, int Tk, float scale, DropSpec drop) {
  pdl_sync();
  __shared__ float sK[KT][DH];
  __shared__ float sV[KT][DH];
  __shared__ float sMask[KT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * NT + threadIdx.x;
  const bool active = qi < Tq;
  DropKey dkey{0u, 1u};
  if (drop.on()) dkey = drop_key(drop);
  const unsigned drow = ((static_cast<unsigned>(b) * H + h) * Tq + qi) * static_cast<unsigned>((Tk + 1) / 2);
  float q[DH], go[DH], dq[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; go[d] = 0.f; dq[d] = 0.f; }
  float row_lse = 0.f, dl = 0.f;
  if (active) {
    load_row32(Q + (static_cast<size_t>(b) * Tq + qi) * ldq + h * DH, q);
    load_row32(dO + (static_cast<size_t>(b) * Tq + qi) * lddo + h * DH, go);
    float o[DH];
    load_row32(O + (static_cast<size_t>(b) * Tq + qi) * ldo + h * DH, o);
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] *= scale; dl = fmaf(go[d], o[d], dl); }
    row_lse = lse[(static_cast<size_t>(b) * H + h) * Tq + qi];
    delta[(static_cast<size_t>(b) * H + h) * Tq + qi] = dl;
  }
  for (int k0 = 0; k0 < Tk; k0 += KT) {
    const int rows = min(KT, Tk - k0);
    __syncthreads();
    stage_tile(K + (static_cast<size_t>(b) * Tk + k0) * ldk + h * DH, ldk, rows, sK);
    stage_tile(V + (static_cast<size_t>(b) * Tk + k0) * ldv + h * DH, ldv, rows, sV);
    for (int j = threadIdx.x; j < KT; j += blockDim.x)
      sMask[j] = (j < rows && (kmask == nullptr || kmask[static_cast<size_t>(b) * Tk + k0 + j] != 0.f)) ? 1.f : 0.f;
    __syncthreads();
    for (int j = 0; j < rows; ++j) {
      if (sMask[j] == 0.f) continue;
      float dot = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sK[j][d]);
        const float4 vv = *reinterpret_cast<const float4*>(&sV[j][d]);
        dot = fmaf(q[d], kv.x, dot); dot = fmaf(q[d + 1], kv.y, dot);
        dot = fmaf(q[d + 2], kv.z, dot); dot = fmaf(q[d + 3], kv.w, dot);
        dp = fmaf(go[d], vv.x, dp); dp = fmaf(go[d + 1], vv.y, dp);
        dp = fmaf(go[d + 2], vv.z, dp); dp = fmaf(go[d + 3], vv.w, dp);
      }
      const float p = __expf(dot - row_lse);
      if (drop.on()) dp *= drop_one(dkey, drow + ((k0 + j) >> 1), (k0 + j) & 1, drop.thr, drop.scale);
      const float ds = p * (dp - dl);
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sK[j][d]);
        dq[d] = fmaf(ds, kv.x, dq[d]); dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]);
        dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int d = 0; d < DH; ++d) dq[d] *= scale;
    store_row32(dQ + (static_cast<size_t>(b) * Tq + qi) * lddq + h * DH, dq);
  }
}

// dK, dV: one thread per key; queries (Q, dO, lse, delta) are staged tile by tile.
template <typename T>
__global__ void __launch_bounds__(NT)
attn_bwd_dkv_kernel(const T* __restrict__ Q, long long ldq, const T* __restrict__ K, long long ldk,
                    const T* __restrict__ V, long long ldv, const float* __restrict__ kmask,
                    const T* __restrict__ dO, long long lddo, const float* __restrict__ lse,
                    const float* __restrict__ delta, T* __restrict__ dK, long long lddk, T* __restrict__ dV,
                    long long lddv, int H, int Tq, int Tk, float scale, DropSpec drop) {
  pdl_sync();
  __shared__ float sQ[KT][DH];
  __shared__ float sG[KT][DH];
  __shared__ float sLse[KT];
  __shared__ float sDel[KT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int kj = blockIdx.x * NT + threadIdx.x;
  const bool active = kj < Tk;
  DropKey dkey{0u, 1u};
  if (drop.on()) dkey = drop_key(drop);
  const unsigned half_tk = static_cast<unsigned>((Tk + 1) / 2);
  float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  bool valid = false;
  if (active) {
    load_row32(K + (static_cast<size_t>(b) * Tk + kj) * ldk + h * DH, k);
    load_row32(V + (static_cast<size_t>(b) * Tk + kj) * ldv + h * DH, v);
    valid = (kmask == nullptr) || (kmask[static_cast<size_t>(b) * Tk + kj] != 0.f);
  }
  for (int q0 = 0; q0 < Tq; q0 += KT) {
    const int rows = min(KT, Tq - q0);
    __syncthreads();
    stage_tile(Q + (static_cast<size_t>(b) * Tq + q0) * ldq + h * DH, ldq, rows, sQ);
    stage_tile(dO + (static_cast<size_t>(b) * Tq + q0) * lddo + h * DH, lddo, rows, sG);
    for (int j = threadIdx.x; j < KT; j += blockDim.x) {
      sLse[j] = (j < rows) ? lse[(static_cast<size_t>(b) * H + h) * Tq + q0 + j] : 0.f;
      sDel[j] = (j < rows) ? delta[(static_cast<size_t>(b) * H + h) * Tq + q0 + j] : 0.f;
    }
    __syncthreads();
    if (valid) {
      for (int i = 0; i < rows; ++i) {
        float dot = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(&sQ[i][d]);
          const float4 gv = *reinterpret_cast<const float4*>(&sG[i][d]);
          dot = fmaf(qv.x, k[d], dot); dot = fmaf(qv.y, k[d + 1], dot);
          dot = fmaf(qv.z, k[d + 2], dot); dot = fmaf(qv.w, k[d + 3], dot);
          dp = fmaf(gv.x, v[d], dp); dp = fmaf(gv.y, v[d + 1], dp);
          dp = fmaf(gv.z, v[d + 2], dp); dp = fmaf(gv.w, v[d + 3], dp);
        }
        const float p = __expf(dot * scale - sLse[i]);
        float pm = p;                                  // dropped weight: feeds dV; dP is masked the same way
        if (drop.on()) {
          const unsigned row = (static_cast<unsigned>(b) * H + h) * Tq + q0 + i;
          const float mk = drop_one(dkey, row * half_tk + (kj >> 1), kj & 1, drop.thr, drop.scale);
          pm *= mk; dp *= mk;
        }
        const float ds = p * (dp - sDel[i]) * scale;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(&sQ[i][d]);
          const float4 gv = *reinterpret_cast<const float4*>(&sG[i][d]);
          dk[d] = fmaf(ds, qv.x, dk[d]); dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]);
          dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
          dv[d] = fmaf(pm, gv.x, dv[d]); dv[d + 1] = fmaf(pm, gv.y, dv[d + 1]);
          dv[d + 2] = fmaf(pm, gv.z, dv[d + 2]); dv[d + 3] = fmaf(pm, gv.w, dv[d + 3]);
        }
      }
    }
  }
  if (active) {
    store_row32(dK + (static_cast<size_t>(b) * Tk + kj) * lddk + h * DH, dk);
    store_row32(dV + (static_cast<size_t>(b) * Tk + kj) * lddv + h * DH, dv);
  }
}

template <typename T>
int fwd_impl(const AttnArgs& a, cudaStream_t s) {
  const double fl = 4.0 * a.B * a.H * static_cast<double>(a.Tq) * a.Tk * a.dh;
  const double by = sizeof(T) * static_cast<double>(a.B) * a.H * a.dh * (2.0 * a.Tq + 2.0 * a.Tk);
  ProfScope prof("attention_fwd", fl, by, s);
  dim3 grid(ceil_div(a.Tq, NT), a.H, a.B);
  SER_CUDA_CHECK(launch_pdl(attn_fwd_kernel<T>, dim3(grid), dim3(NT), 0, s, reinterpret_cast<const T*>(a.Q), a.ldq, reinterpret_cast<const T*>(a.K),
                                         a.ldk, reinterpret_cast<const T*>(a.V), a.ldv, a.kmask,
                                         reinterpret_cast<T*>(a.O), a.ldo, a.lse, a.H, a.Tq, a.Tk, a.scale, a.drop));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

template <typename T>
int bwd_impl(const AttnArgs& a, cudaStream_t s) {
  const double fl = 10.0 * a.B * a.H * static_cast<double>(a.Tq) * a.Tk * a.dh;
  const double by = sizeof(T) * static_cast<double>(a.B) * a.H * a.dh * (4.0 * a.Tq + 4.0 * a.Tk);
  ProfScope prof("attention_bwd", fl, by, s);
  dim3 gq(ceil_div(a.Tq, NT), a.H, a.B);
  SER_CUDA_CHECK(launch_pdl(attn_bwd_dq_kernel<T>, dim3(gq), dim3(NT), 0, s, reinterpret_cast<const T*>(a.Q), a.ldq, reinterpret_cast<const T*>(a.K), a.ldk,
      reinterpret_cast<const T*>(a.V), a.ldv, a.kmask, reinterpret_cast<const T*>(a.O), a.ldo,
      reinterpret_cast<const T*>(a.dO), a.lddo, a.lse, a.delta, reinterpret_cast<T*>(a.dQ), a.lddq, a.H, a.Tq, a.Tk,
      a.scale, a.drop));
  SER_LAUNCH_CHECK();
  dim3 gk(ceil_div(a.Tk, NT), a.H, a.B);
  SER_CUDA_CHECK(launch_pdl(attn_bwd_dkv_kernel<T>, dim3(gk), dim3(NT), 0, s, reinterpret_cast<const T*>(a.Q), a.ldq, reinterpret_cast<const T*>(a.K), a.ldk,
      reinterpret_cast<const T*>(a.V), a.ldv, a.kmask, reinterpret_cast<const T*>(a.dO), a.lddo, a.lse, a.delta,
      reinterpret_cast<T*>(a.dK), a.lddk, reinterpret_cast<T*>(a.dV), a.lddv, a.H, a.Tq, a.Tk, a.scale, a.drop));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int check(const AttnArgs& a) {
  SER_REQUIRE(a.dh == DH, "attention: head dim must be 32 (shared_dim 256 / 8 heads)");
  SER_REQUIRE(a.B > 0 && a.H > 0 && a.Tq > 0 && a.Tk > 0, "attention: empty problem");
  SER_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0,
              "attention: leading dimensions must be multiples of 8");
  SER_REQUIRE(!a.drop.on() || static_cast<long long>(a.B) * a.H * a.Tq * ((a.Tk + 1) / 2) < (1LL << 32),
              "attention: dropout site too large for the 32-bit pair index");
  return SER_OK;
}

}  // namespace

// bf16 tier: which tensor-core path.  The tcgen05 kernels (attention_tc5.cu) win where a CTA has several streamed tiles
// to amortise its set-up (tensor-memory allocation, barrier initialisation, the first TMA round trip): measured on a
// B200 with dropout 0.1 (tools/attn_check.py --time), forward 332 / 287 us vs 486 / 470 us for the mma.sync kernels at
// (Tq, Tk) = (1500, 256) / (256, 1500) and 52 vs 60 us at (250, 64); they lose where half of a 128-row tile is empty and
// the loop is short: forward 56 vs 50 us at (64, 250), backward 186 / 157 us vs 140 / 130 us at the two cfg2 shapes
// (backward at the long shapes: 969 / 933 us vs 1070 / 1070 us).  impl = 2 / 1 (or SER_ATTN_FWD / SER_ATTN_BWD = 2 / 1)
// forces one path.
static bool use_tc5(const AttnArgs& a, bool backward) {
  static const int env_f = getenv("SER_ATTN_FWD") ? atoi(getenv("SER_ATTN_FWD")) : 0;
  static const int env_b = getenv("SER_ATTN_BWD") ? atoi(getenv("SER_ATTN_BWD")) : 0;
  const int forced = a.impl != 0 ? a.impl : (backward ? env_b : env_f);
  if (forced == 1 || !attention_tc5_supported(a)) return false;
  if (forced == 2) return true;
  const long long cells = static_cast<long long>(a.Tq) * a.Tk;
  if (backward) return cells >= 131072;
  return a.Tq > 64 || cells >= 131072;
}

// A/B switch: SER_ATTN_KEEPBITS=0 makes the backward kernels re-hash the dropout decisions instead of reading the bits
// the forward kernel stored
static bool keep_bits_enabled() {
  static const bool on = !(getenv("SER_ATTN_KEEPBITS") != nullptr && atoi(getenv("SER_ATTN_KEEPBITS")) == 0);
  return on;
}

int attention_fwd(const AttnArgs& a, cudaStream_t s) {
  SER_TRY(check(a));
  // bf16 tier: tensor-core kernels (attention_tc.cu); fp32 tier: the CUDA-core kernels of this file
  if (a.dtype == DT_F32) return fwd_impl<float>(a, s);
  if (a.impl == 2 && !attention_tc5_supported(a)) {
    set_last_error(__FILE__, __LINE__, "attention: the tcgen05 kernels need head dim 32, an even head count and 16-byte aligned operands");
    return SER_ERR_UNSUPPORTED;
  }
  if (!use_tc5(a, false)) return attention_fwd_tc(a, s);
  AttnArgs b = a;
  if (!keep_bits_enabled()) b.keep_bits = nullptr;
  return attention_fwd_tc5(b, s);
}

int attention_bwd(const AttnArgs& a, cudaStream_t s) {
  SER_TRY(check(a));
  SER_REQUIRE(a.delta != nullptr && a.lse != nullptr, "attention_bwd: lse / delta buffers required");
  if (a.dtype == DT_F32) return bwd_impl<float>(a, s);
  if (a.impl == 2 && !attention_tc5_supported(a)) {
    set_last_error(__FILE__, __LINE__, "attention: the tcgen05 kernels need head dim 32, an even head count and 16-byte aligned operands");
    return SER_ERR_UNSUPPORTED;
  }
  if (!use_tc5(a, true)) return attention_bwd_tc(a, s);
  AttnArgs b = a;
  // the keep bits exist only if the FORWARD of this problem ran on the tcgen05 kernel (same shape, same impl / switches)
  if (!keep_bits_enabled() || !use_tc5(a, false)) b.keep_bits = nullptr;
  return attention_bwd_tc5(b, s);
}

}  // namespace ser

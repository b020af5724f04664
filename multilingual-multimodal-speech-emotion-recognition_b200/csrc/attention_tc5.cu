// Masked multi-head cross attention on the 5th-generation tensor cores (bf16 tier): S = Q K^T and O = P V are
// tcgen05.mma instructions with the accumulators in tensor memory, Q / K / V tiles arrive through TMA.
// Semantics are those of attention.cu / attention_tc.cu (nn.MultiheadAttention's math path as used by
// src/models/cross_attention.py:41,49; torch/nn/functional.py:6609-6645): additive -inf key padding, softmax over keys,
// dropout on the attention weights, weights never materialised in HBM, only the per-row log-sum-exp is saved; a sample
// whose keys are ALL padded yields NaN for every query.
//
// Shape of the problem: head dim 32, 8 heads packed along the feature axis of the projection buffers (row pitch 768
// elements), Tq x Tk in {250 x 64, 64 x 250, 1500 x 256, 256 x 1500} per (sample, head).  One CTA owns 128 query rows of
// one sample and a PAIR of heads: a 128-byte row of the Q / K / V buffers is exactly two heads, so one SWIZZLE_128B TMA
// box per operand serves both, and head h of the pair is addressed by advancing the shared-memory descriptor by 64 h
// bytes inside the swizzle atom (the same arithmetic a GEMM uses for its K = 16 steps).  Warp roles:
//   warps 0-3   softmax of head 0 of the pair (thread = query row = TMEM lane), warps 4-7 the same for head 1
//   lane 0 of warps 0 / 4 additionally issues that head's tcgen05.mma (and, for head 0, the TMA loads of the 64-key
//   K / V tiles, 2 stages) in program order -- there is no dedicated control warp
// Per 64-key tile and head:  S[128 x 64] = Q_h K_h^T (two K = 16 MMAs) -> the softmax warps read their S row from tensor
// memory (tcgen05.ld, 64 fp32 registers), update the running maximum / sum, write P (bf16, dropout applied) into a
// 128-byte-swizzled shared tile that is the A operand of O_part[128 x 64] = P V_pair (four K = 16 MMAs; V is the
// MN-major B operand, N covers both heads and each head reads its own 32 columns) -> the softmax warps fold O_part into
// their fp32 register accumulator with the usual rescaling.  The two heads are independent pipelines, so while one
// head's warps do their exponentials the other head's MMAs are in flight; two CTAs share an SM (256 TMEM columns each).
//
// The attention core is bound by the exponentials and the surrounding fp32 instructions (64 FLOP of tensor work per
// exp at head dim 32), not by the tensor pipe: what tcgen05 buys is that the tensor work and its operand traffic leave
// the instruction stream of the softmax warps entirely (the mma.sync kernel spent its issue slots on ldmatrix + HMMA).
#include "kernels.cuh"
#include "prof.cuh"
#include "tc5.cuh"
#include <stdlib.h>

namespace ser {

namespace {

using namespace tc5;
typedef __nv_bfloat16 bf16;

constexpr int A5_DH = 32;
constexpr int A5_ROWS = 128;              // rows owned by a CTA (= TMEM lanes)
constexpr int A5_KT = 64;                 // streamed rows per tile (keys in fwd / dQ, queries in dK/dV)
constexpr int A5_THREADS = 256;           // 2 heads x 4 softmax warps (thread = row)
constexpr int A5_TILE_OWN = A5_ROWS * 128;   // bytes of a 128-row x 128-byte tile
constexpr int A5_TILE_STR = A5_KT * 128;     // bytes of a 64-row x 128-byte tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {             // one MUFU.EX2; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// K-major operand tile (rows of 128 bytes, SWIZZLE_128B): descriptor of the K = 16 slice starting `byte_off` into a row
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, uint32_t byte_off) {
  return make_smem_desc(tile + byte_off, 0u, 1024u);
}
// MN-major operand tile (rows = contraction index, 128 bytes = 64 MN elements per row): K = 16 slice number k
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int k) {
  return make_smem_desc(tile + static_cast<uint32_t>(k) * (UMMA_K * 128u), A5_KT * 128u, 1024u);
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
struct FwdBars {
  uint64_t q_full;
  uint64_t kv_full[2], kv_empty[2];       // K / V stage landed (TMA) / released (both heads' P V retired: count 2)
  uint64_t s_full[2];                     // per head: S accumulator written
  uint64_t o_full[2];                     // per head: O_part accumulator written
  uint32_t tmem_slot;
};

// named barrier of one head's four softmax warps (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int h) {
  if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// 256 threads, no dedicated control warp: a ninth warp would sit on one of the four sub-partitions and its register
// allocation alone would keep a second CTA off the SM.  The first lane of each head's first warp issues that head's
// MMAs (and, for head 0, the TMA loads) in program order; the four warps of a head meet at a named barrier once their
// S reads and P writes of a tile are done.
template <bool DROP>
__global__ void __launch_bounds__(A5_THREADS, 2)
attn5_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const float* __restrict__ kmask, bf16* __restrict__ O,
                 long long ldo, float* __restrict__ lse, int H, int Tq, int Tk, float scale, DropSpec drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // [128 x 128 B]   Q rows, both heads of the pair
  uint8_t* sK = sQ + A5_TILE_OWN;                      // 2 x [64 x 128 B]
  uint8_t* sV = sK + 2 * A5_TILE_STR;                  // 2 x [64 x 128 B]
  uint8_t* sP = sV + 2 * A5_TILE_STR;                  // 2 heads x [128 x 128 B] bf16 probabilities (A operand of P V)
  FwdBars* bars = reinterpret_cast<FwdBars*>(sP + 2 * A5_TILE_OWN);
  float* kbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [ntiles * 64]: 0 valid key, -inf otherwise
  int* tfull = reinterpret_cast<int*>(kbias + ((Tk + A5_KT - 1) / A5_KT) * A5_KT);    // [ntiles]: 1 = every key of the tile valid

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, hp = blockIdx.y, q0 = blockIdx.x * A5_ROWS;
  const int ntiles = (Tk + A5_KT - 1) / A5_KT;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
      mbar_init(&bars->q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->kv_full[i], 1); mbar_init(&bars->kv_empty[i], 2);
        mbar_init(&bars->s_full[i], 1); mbar_init(&bars->o_full[i], 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int t = warp; t < ntiles; t += A5_THREADS / 32) {     // one warp per tile: two keys per lane
    bool ok2 = true;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = t * A5_KT + lane + 32 * u;
      const bool ok = j < Tk && (kmask == nullptr || kmask[static_cast<size_t>(b) * Tk + j] != 0.f);
      kbias[j] = ok ? 0.f : -INFINITY;
      ok2 = ok2 && ok;
    }
    const bool all = __all_sync(0xffffffffu, ok2);
    if (lane == 0) tfull[t] = all ? 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  // tensor-memory columns: S of head h at [64 h, 64 h + 64); O_part of head h at [128 + 64 h, 128 + 64 h + 64)
  // (the P V product is 64 columns wide = both heads of the pair; head h reads columns 32 h .. 32 h + 31 of its own)

  const int h = warp >> 2, quarter = warp & 3;
  const bool leader = (quarter == 0 && lane == 0);        // issues this head's MMAs
  const int col0 = hp * 2 * A5_DH;
  constexpr uint32_t idesc_s = make_idesc<A5_KT, 0, 0, A5_ROWS>();        // S: A = Q K-major, B = K K-major, N = 64 keys
  constexpr uint32_t idesc_pv = make_idesc<2 * A5_DH, 0, 1, A5_ROWS>();   // P V: A = P K-major, B = V MN-major, N = 64
  const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP + h * A5_TILE_OWN);
  auto load_kv = [&](int t) {                               // head 0's leader only
    const int st = t & 1;
    mbar_expect_tx(&bars->kv_full[st], 2 * A5_TILE_STR);
    tma_load_3d(&tmK, &bars->kv_full[st], sK + st * A5_TILE_STR, col0, b * Tk + t * A5_KT, 0);
    tma_load_3d(&tmV, &bars->kv_full[st], sV + st * A5_TILE_STR, col0, b * Tk + t * A5_KT, 0);
  };
  auto issue_s = [&](int t) {                               // S_h(t) = Q_h K_h(t)^T into this head's S columns
    const int st = t & 1;
    mbar_wait(&bars->kv_full[st], (t >> 1) & 1);
    tc_fence_after();
    const uint32_t aK = smem_u32(sK + st * A5_TILE_STR);
#pragma unroll
    for (int k = 0; k < A5_DH / UMMA_K; ++k)
      tc_mma_bf16(tmem_base + h * A5_KT, desc_kmajor(aQ, h * 64 + k * 32), desc_kmajor(aK, h * 64 + k * 32), idesc_s,
                  k > 0 ? 1u : 0u);
    tc_commit(&bars->s_full[h]);
  };
  if (leader) {
    if (h == 0) {
      mbar_expect_tx(&bars->q_full, A5_TILE_OWN);
      tma_load_3d(&tmQ, &bars->q_full, sQ, col0, b * Tq + q0, 0);
      load_kv(0);
      if (ntiles > 1) load_kv(1);
    }
    mbar_wait(&bars->q_full, 0);
    issue_s(0);
  }
  __syncwarp();

  const int row = quarter * 32 + lane;                      // row of the CTA tile = TMEM lane
  const int qi = q0 + row;
  const bool active = (q0 + quarter * 32) < Tq;             // warp-uniform: warps past the end of the sample only keep the barriers moving
  const int head = hp * 2 + h;
  const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(h * A5_KT);
  const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(128 + h * 64 + h * A5_DH);
  const float c = scale * kLog2e;
  const uint32_t p_row = aP + static_cast<uint32_t>(row) * 128u;
  const uint32_t swz = static_cast<uint32_t>(row & 7);
  DropKey dkey{0u, 1u};
  unsigned drow = 0u;
  if (DROP) {
    dkey = drop_key(drop);
    drow = ((static_cast<unsigned>(b) * H + head) * Tq + qi) * static_cast<unsigned>((Tk + 1) / 2);
  }
  float o[A5_DH];
#pragma unroll
  for (int j = 0; j < A5_DH; ++j) o[j] = 0.f;
  float m = -INFINITY, l = 0.f;                             // running max (log2 units, scale included) and sum

  for (int t = 0; t < ntiles; ++t) {
    if (t > 0) {
      // the previous tile's P V (issued before this tile's Q K^T, so it retires first): its product is relative to the
      // previous running maximum, like o
      mbar_wait(&bars->o_full[h], (t - 1) & 1);
      tc_fence_after();
      if (leader && h == 0 && t + 1 < ntiles) {             // the stage of tile t-1 is free once BOTH heads' P V retired
        mbar_wait(&bars->kv_empty[(t + 1) & 1], ((t - 1) >> 1) & 1);
        load_kv(t + 1);
      }
      __syncwarp();
      if (active) {
        uint32_t ov[A5_DH];
        tmem_ld32_issue(t_o, ov);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < A5_DH; ++j) o[j] += __uint_as_float(ov[j]);
      }
    }
    mbar_wait(&bars->s_full[h], t & 1);
    tc_fence_after();
    if (active) {
      const bool full = tfull[t] != 0;                      // block-uniform: every key of this tile is valid
      float x[A5_KT];                                       // scaled, masked scores of this row (log2 units)
      {
        uint32_t sv[A5_KT];
        tmem_ld32_issue(t_s, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld32_issue(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int j = 0; j < A5_KT; ++j) x[j] = __uint_as_float(sv[j]) * c;
        } else {
          const float4* kb = reinterpret_cast<const float4*>(kbias + t * A5_KT);
#pragma unroll
          for (int j = 0; j < A5_KT / 4; ++j) {
            const float4 bb = kb[j];
            x[4 * j] = fmaf(__uint_as_float(sv[4 * j]), c, bb.x);
            x[4 * j + 1] = fmaf(__uint_as_float(sv[4 * j + 1]), c, bb.y);
            x[4 * j + 2] = fmaf(__uint_as_float(sv[4 * j + 2]), c, bb.z);
            x[4 * j + 3] = fmaf(__uint_as_float(sv[4 * j + 3]), c, bb.w);
          }
        }
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < A5_KT; j += 4) tmax = fmaxf(tmax, fmaxf(fmaxf(x[j], x[j + 1]), fmaxf(x[j + 2], x[j + 3])));
      const float mn = fmaxf(m, tmax);
      const float mu = (mn == -INFINITY) ? 0.f : mn;        // nothing but padded keys so far: keep exp2 finite
      const float corr = ex2(m - mu);                       // m = -inf -> 0
      m = mn;
#pragma unroll
      for (int j = 0; j < A5_DH; ++j) o[j] *= corr;
      float ls = 0.f;
#pragma unroll
      for (int j8 = 0; j8 < A5_KT / 8; ++j8) {              // 8 keys = one 16-byte chunk of the P row
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = 8 * j8 + 2 * u;
          float p0 = ex2(x[j] - mu), p1 = ex2(x[j + 1] - mu);
          ls += p0 + p1;                                    // the normaliser sees every key; dropout acts on the weights
          if (DROP) {
            const float2 mk = drop_pair(dkey, drow + static_cast<unsigned>(t * (A5_KT / 2) + (j >> 1)), drop.thr, drop.scale);
            p0 *= mk.x; p1 *= mk.y;
          }
          w[u] = pack_bf16x2(p0, p1);
        }
        sts128(p_row + ((static_cast<uint32_t>(j8) ^ swz) << 4), w[0], w[1], w[2], w[3]);
      }
      l = fmaf(l, corr, ls);
    }
    fence_async_smem();                                     // generic-proxy writes of P -> visible to the MMA's async proxy
    tc_fence_before();                                      // this thread's S / O_part reads precede the barrier
    group_sync(h);
    if (leader) {
      tc_fence_after();
      const uint32_t aV = smem_u32(sV + (t & 1) * A5_TILE_STR);
#pragma unroll
      for (int k = 0; k < A5_KT / UMMA_K; ++k)
        tc_mma_bf16(tmem_base + 128 + h * 64, desc_kmajor(aP, k * 32), desc_mnmajor(aV, k), idesc_pv, k > 0 ? 1u : 0u);
      tc_commit(&bars->o_full[h]);
      tc_commit(&bars->kv_empty[t & 1]);                    // (one of the two arrivals that release this K / V stage)
      if (t + 1 < ntiles) issue_s(t + 1);
    }
    __syncwarp();
  }
  // last tile's product
  mbar_wait(&bars->o_full[h], (ntiles - 1) & 1);
  tc_fence_after();
  if (active) {
    uint32_t ov[A5_DH];
    tmem_ld32_issue(t_o, ov);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < A5_DH; ++j) o[j] += __uint_as_float(ov[j]);
    if (qi < Tq) {
      const float nan = __int_as_float(0x7fc00000);
      const float inv = (l > 0.f) ? 1.f / l : nan;          // all keys padded -> NaN (reference behaviour)
      bf16* dst = O + (static_cast<size_t>(b) * Tq + qi) * ldo + head * A5_DH;
#pragma unroll
      for (int j = 0; j < A5_DH / 8; ++j) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * j] * inv, o[8 * j + 1] * inv);
        v.y = pack_bf16x2(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
        v.z = pack_bf16x2(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
        v.w = pack_bf16x2(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
        *reinterpret_cast<uint4*>(dst + 8 * j) = v;
      }
      if (lse != nullptr) lse[(static_cast<size_t>(b) * H + head) * Tq + qi] = (l > 0.f) ? m * kLn2 + logf(l) : nan;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace

// can the tcgen05 kernels take this problem?  (head dim 32, an even number of heads, TMA-addressable operands)
bool attention_tc5_supported(const AttnArgs& a) {
  static const bool disabled = (getenv("SER_ATTN_LEGACY") != nullptr);      // A/B switch: mma.sync kernels instead
  auto ok_ptr = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return !disabled && a.dtype == DT_BF16 && a.dh == A5_DH && a.H % 2 == 0 && a.H >= 2 && a.Tq >= 1 && a.Tk >= 1 &&
         a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0 && ok_ptr(a.Q) && ok_ptr(a.K) && ok_ptr(a.V) &&
         ok_ptr(a.O) && a.Tk <= 8192;
}

int attention_fwd_tc5(const AttnArgs& a, cudaStream_t s) {
  const double fl = 4.0 * a.B * a.H * static_cast<double>(a.Tq) * a.Tk * a.dh;
  const double by = 2.0 * static_cast<double>(a.B) * a.H * a.dh * (2.0 * a.Tq + 2.0 * a.Tk);
  ProfScope prof("attention_fwd", fl, by, s);
  CUtensorMap tmQ, tmK, tmV;
  const int HD = a.H * a.dh;
  SER_TRY(make_tmap(&tmQ, a.Q, 0, static_cast<long long>(a.B) * a.Tq, HD, a.ldq, A5_ROWS, 64, 1, 0));
  SER_TRY(make_tmap(&tmK, a.K, 0, static_cast<long long>(a.B) * a.Tk, HD, a.ldk, A5_KT, 64, 1, 0));
  SER_TRY(make_tmap(&tmV, a.V, 0, static_cast<long long>(a.B) * a.Tk, HD, a.ldv, A5_KT, 64, 1, 0));
  const int ntiles = ceil_div(a.Tk, A5_KT);
  const int smem = 1024 + A5_TILE_OWN + 4 * A5_TILE_STR + 2 * A5_TILE_OWN + 256 + ntiles * (A5_KT + 1) * static_cast<int>(sizeof(float));
  auto* kern = a.drop.on() ? attn5_fwd_kernel<true> : attn5_fwd_kernel<false>;
  static bool configured[2] = {false, false};
  if (!configured[a.drop.on() ? 1 : 0]) {
    SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        1024 + 3 * A5_TILE_OWN + 4 * A5_TILE_STR + 256 + (8192 + 128) * 4 + 64));
    // two CTAs per SM need ~170 KB of shared memory: ask for the largest carve-out (the default heuristic picked 102 KB)
    SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured[a.drop.on() ? 1 : 0] = true;
  }
  dim3 grid(ceil_div(a.Tq, A5_ROWS), a.H / 2, a.B);
  kern<<<grid, A5_THREADS, smem, s>>>(tmQ, tmK, tmV, a.kmask, reinterpret_cast<bf16*>(a.O), a.ldo, a.lse, a.H, a.Tq, a.Tk,
                                      a.scale, a.drop);
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int attention_bwd_tc5(const AttnArgs& a, cudaStream_t s) { return attention_bwd_tc(a, s); }   // (tcgen05 backward: below, WIP)

}  // namespace ser

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

__device__ __forceinline__ void group_sync_n(int h, int n) {
  if (h == 0) asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory");
}

// 256 threads, no dedicated control warp: a ninth warp would sit on one of the four sub-partitions and its register
// allocation alone would keep a second CTA off the SM.  The first lane of each head's first warp issues that head's
// MMAs (and, for head 0, the TMA loads) in program order; the four warps of a head meet at a named barrier once their
// S reads and P writes of a tile are done.
template <bool DROP>
__global__ void __launch_bounds__(A5_THREADS, 2)
attn5_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const float* __restrict__ kmask, bf16* __restrict__ O,
                 long long ldo, float* __restrict__ lse, int H, int Tq, int Tk, float scale, DropSpec drop, int narrow,
                 unsigned* __restrict__ keep_bits) {
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
  pdl_sync();
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
  // P V: A = P K-major, B = V MN-major.  narrow = 0: N = 64 (both heads' V columns, each head keeps its 32);
  // narrow = 1: N = 32, the B descriptor starts 64 h bytes into the 128-byte atom (this head's V columns only)
  const uint32_t idesc_pv = narrow ? make_idesc<A5_DH, 0, 1, A5_ROWS>() : make_idesc<2 * A5_DH, 0, 1, A5_ROWS>();
  const uint32_t v_off = narrow ? static_cast<uint32_t>(h * 64) : 0u;
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
  const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                       static_cast<uint32_t>(128 + h * 64 + (narrow ? 0 : h * A5_DH));
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
          for (int j = 0; j < A5_KT; ++j) x[j] = __uint_as_float(sv[j]);
        } else {                                            // padded keys: -inf in raw score units
          const float4* kb = reinterpret_cast<const float4*>(kbias + t * A5_KT);
#pragma unroll
          for (int j = 0; j < A5_KT / 4; ++j) {
            const float4 bb = kb[j];
            x[4 * j] = __uint_as_float(sv[4 * j]) + bb.x;
            x[4 * j + 1] = __uint_as_float(sv[4 * j + 1]) + bb.y;
            x[4 * j + 2] = __uint_as_float(sv[4 * j + 2]) + bb.z;
            x[4 * j + 3] = __uint_as_float(sv[4 * j + 3]) + bb.w;
          }
        }
      }
      // x holds RAW masked scores; the softmax scale rides in the exponent's FFMA: p = 2^(x c - max)
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < A5_KT; j += 4) tmax = fmaxf(tmax, fmaxf(fmaxf(x[j], x[j + 1]), fmaxf(x[j + 2], x[j + 3])));
      tmax *= c;                                            // c > 0
      const float mn = fmaxf(m, tmax);
      const float mu = (mn == -INFINITY) ? 0.f : mn;        // nothing but padded keys so far: keep exp2 finite
      const float corr = ex2(m - mu);                       // m = -inf -> 0
      m = mn;
#pragma unroll
      for (int j = 0; j < A5_DH; ++j) o[j] *= corr;
      float ls = 0.f;
      unsigned kw = 0u;                                     // keep decisions of 32 keys (bit j%32), stored every 32 keys
#pragma unroll
      for (int j8 = 0; j8 < A5_KT / 8; ++j8) {              // 8 keys = one 16-byte chunk of the P row
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = 8 * j8 + 2 * u;
          float p0 = ex2(fmaf(x[j], c, -mu)), p1 = ex2(fmaf(x[j + 1], c, -mu));
          ls += p0 + p1;                                    // the normaliser sees every key; dropout acts on the weights
          if (DROP) {
            const unsigned bb = drop_bits(dkey, drow + static_cast<unsigned>(t * (A5_KT / 2) + (j >> 1)));
            const bool k0 = (bb & 0xffffu) >= drop.thr, k1 = (bb >> 16) >= drop.thr;
            p0 *= k0 ? drop.scale : 0.f;
            p1 *= k1 ? drop.scale : 0.f;
            if (k0) kw |= 1u << (j & 31);                    // (compile-time positions: one predicated OR per decision)
            if (k1) kw |= 2u << (j & 31);
          }
          w[u] = pack_bf16x2(p0, p1);
        }
        sts128(p_row + ((static_cast<uint32_t>(j8) ^ swz) << 4), w[0], w[1], w[2], w[3]);
        if (DROP && (j8 & 3) == 3) {                        // 32 decisions complete: the backward kernels read them instead of re-hashing
          const int wi = 2 * t + (j8 >> 2), W = (Tk + 31) >> 5;
          if (keep_bits != nullptr && qi < Tq && wi < W)
            keep_bits[((static_cast<size_t>(b) * H + head) * Tq + qi) * W + wi] = kw;
          kw = 0u;
        }
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
        tc_mma_bf16(tmem_base + 128 + h * 64, desc_kmajor(aP, k * 32), desc_mnmajor(aV + v_off, k), idesc_pv, k > 0 ? 1u : 0u);
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

// ------------------------------------------------------------------------------------------------------------
// backward.  Two kernels, as in attention_tc.cu: dQ (CTA = 128 query rows x head pair, streams 32-key tiles) and
// dK / dV (CTA = 128 key rows x head pair, streams 32-query tiles).  Per streamed tile and head, two products land in
// tensor memory -- the scores and dP = dO V^T (or their transposes) -- the softmax warps turn them into P and
// dS = P (mask dP - delta) with the saved log-sum-exp, write those as bf16 into ONE swizzled shared tile (P in bytes
// 0..63 of a row, dS in bytes 64..127: two K = 32 operands addressed by descriptor offsets), and the gradient
// products accumulate in tensor memory over all tiles (N = 32: the MN-major B descriptor starts 64 h bytes into the
// 128-byte atom, i.e. at this head's columns).  Tensor memory per head: 32 + 32 + 32 (+ 32) columns -> 256 per CTA,
// two CTAs per SM.  The thread that issues a head's MMAs rotates over its four warps from tile to tile.
// ------------------------------------------------------------------------------------------------------------
constexpr int B5_KT = 32;                          // streamed rows per tile
constexpr int B5_TILE_STR = B5_KT * 128;           // bytes of a 32-row x 128-byte tile
constexpr int B5_NST = 4;                          // stages of the streamed operand pair (dQ kernel)
constexpr int B5_NST_KV = 2;                       // ... (dK / dV kernel)

template <int NST>
struct BwdBars {
  uint64_t own_full;                      // the CTA's own two 128-row tiles landed
  uint64_t str_full[NST], str_empty[NST]; // streamed tile pair landed / released (both heads' accumulate MMAs retired)
  uint64_t s_full[2];                     // per head: score + dP products of the current tile written
  uint64_t acc_full[2];                   // per head: the accumulators are final
  uint32_t tmem_slot;
};

template <bool DROP>
__global__ void __launch_bounds__(A5_THREADS, 2)
attn5_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const float* __restrict__ kmask, const bf16* __restrict__ O, long long ldo,
                    const bf16* __restrict__ dO, long long lddo, const float* __restrict__ lse,
                    float* __restrict__ delta, bf16* __restrict__ dQ, long long lddq, int H, int Tq, int Tk, float scale,
                    DropSpec drop, const unsigned* __restrict__ keep_bits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // [128 x 128 B] Q rows (both heads)
  uint8_t* sG = sQ + A5_TILE_OWN;                      // [128 x 128 B] dO rows
  uint8_t* sK = sG + A5_TILE_OWN;                      // NST x [32 x 128 B]
  uint8_t* sV = sK + B5_NST * B5_TILE_STR;             // NST x [32 x 128 B]
  uint8_t* sD = sV + B5_NST * B5_TILE_STR;             // 2 heads x [128 x 128 B]: dS in bytes 0..63 of a row
  BwdBars<B5_NST>* bars = reinterpret_cast<BwdBars<B5_NST>*>(sD + 2 * A5_TILE_OWN);
  float* kbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [ntiles * 32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, hp = blockIdx.y, q0 = blockIdx.x * A5_ROWS;
  const int ntiles = (Tk + B5_KT - 1) / B5_KT;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
      mbar_init(&bars->own_full, 1);
      for (int i = 0; i < B5_NST; ++i) { mbar_init(&bars->str_full[i], 1); mbar_init(&bars->str_empty[i], 2); }
      for (int i = 0; i < 2; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->acc_full[i], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();
  for (int j = threadIdx.x; j < ntiles * B5_KT; j += A5_THREADS) {
    const bool ok = j < Tk && (kmask == nullptr || kmask[static_cast<size_t>(b) * Tk + j] != 0.f);
    kbias[j] = ok ? 0.f : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  // tensor-memory columns: S_h at [32 h, +32), dP_h at [64 + 32 h, +32), dQ_h at [128 + 32 h, +32)

  const int h = warp >> 2, quarter = warp & 3;
  const int col0 = hp * 2 * A5_DH;
  constexpr uint32_t idesc_s = make_idesc<B5_KT, 0, 0, A5_ROWS>();       // [128 x 32] = A(K-major) B(K-major)^T, K = d
  constexpr uint32_t idesc_acc = make_idesc<A5_DH, 0, 1, A5_ROWS>();     // [128 x 32] += A(K-major) B(MN-major), K = tile rows
  const uint32_t aQ = smem_u32(sQ), aG = smem_u32(sG), aD = smem_u32(sD + h * A5_TILE_OWN);
  auto load_tile = [&](int t) {
    const int st = t % B5_NST;
    mbar_expect_tx(&bars->str_full[st], 2 * B5_TILE_STR);
    tma_load_3d(&tmK, &bars->str_full[st], sK + st * B5_TILE_STR, col0, b * Tk + t * B5_KT, 0);
    tma_load_3d(&tmV, &bars->str_full[st], sV + st * B5_TILE_STR, col0, b * Tk + t * B5_KT, 0);
  };
  auto issue_products = [&](int t) {                        // S_h(t) = Q_h K_h^T, dP_h(t) = dO_h V_h^T
    const int st = t % B5_NST;
    mbar_wait(&bars->str_full[st], (t / B5_NST) & 1);
    tc_fence_after();
    const uint32_t aK = smem_u32(sK + st * B5_TILE_STR), aV = smem_u32(sV + st * B5_TILE_STR);
#pragma unroll
    for (int k = 0; k < A5_DH / UMMA_K; ++k)
      tc_mma_bf16(tmem_base + h * 32, desc_kmajor(aQ, h * 64 + k * 32), desc_kmajor(aK, h * 64 + k * 32), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < A5_DH / UMMA_K; ++k)
      tc_mma_bf16(tmem_base + 64 + h * 32, desc_kmajor(aG, h * 64 + k * 32), desc_kmajor(aV, h * 64 + k * 32), idesc_s, k > 0 ? 1u : 0u);
    tc_commit(&bars->s_full[h]);
  };
  if (quarter == 3 && lane == 0) {                          // the leader "before tile 0"
    if (h == 0) {
      mbar_expect_tx(&bars->own_full, 2 * A5_TILE_OWN);
      tma_load_3d(&tmQ, &bars->own_full, sQ, col0, b * Tq + q0, 0);
      tma_load_3d(&tmG, &bars->own_full, sG, col0, b * Tq + q0, 0);
      for (int t = 0; t < B5_NST && t < ntiles; ++t) load_tile(t);
    }
    mbar_wait(&bars->own_full, 0);
    issue_products(0);
  }
  __syncwarp();

  const int row = quarter * 32 + lane;
  const int qi = q0 + row;
  const bool active = (q0 + quarter * 32) < Tq;
  const int head = hp * 2 + h;
  const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  const float c = scale * kLog2e;
  const uint32_t d_row = aD + static_cast<uint32_t>(row) * 128u;
  const uint32_t swz = static_cast<uint32_t>(row & 7);
  DropKey dkey{0u, 1u};
  unsigned drow = 0u;
  if (DROP) {
    dkey = drop_key(drop);
    drow = ((static_cast<unsigned>(b) * H + head) * Tq + qi) * static_cast<unsigned>((Tk + 1) / 2);
  }
  // per-row constants: -lse in log2 units (+inf log-sum-exp for rows past the sample -> P = 0) and delta = sum(dO * O)
  float nl = -INFINITY, dl = 0.f;
  if (qi < Tq) {
    const size_t r = static_cast<size_t>(b) * Tq + qi;
    nl = -lse[(static_cast<size_t>(b) * H + head) * Tq + qi] * kLog2e;
    float g[8], ov[8];
#pragma unroll
    for (int i = 0; i < A5_DH / 8; ++i) {
      load8(dO + r * lddo + head * A5_DH + 8 * i, g);
      load8(O + r * ldo + head * A5_DH + 8 * i, ov);
#pragma unroll
      for (int k = 0; k < 8; ++k) dl = fmaf(g[k], ov[k], dl);
    }
    delta[(static_cast<size_t>(b) * H + head) * Tq + qi] = dl;
  }

  // keep bits written by the forward kernel: one word per (row, 32-key tile), requested one tile ahead
  const unsigned* krow = nullptr;
  unsigned kcur = 0u;
  if (DROP && keep_bits != nullptr && qi < Tq) {
    krow = keep_bits + ((static_cast<size_t>(b) * H + head) * Tq + qi) * ((Tk + 31) >> 5);
    kcur = __ldg(krow);
  }
  for (int t = 0; t < ntiles; ++t) {
    unsigned knext = 0u;
    if (DROP && krow != nullptr && t + 1 < ntiles) knext = __ldg(krow + t + 1);
    mbar_wait(&bars->s_full[h], t & 1);                    // (also: the previous tile's dQ MMAs have retired)
    tc_fence_after();
    if (h == 0 && quarter == (t & 3) && lane == 0 && t >= 1 && t - 1 + B5_NST < ntiles) {
      mbar_wait(&bars->str_empty[(t - 1) % B5_NST], ((t - 1) / B5_NST) & 1);
      load_tile(t - 1 + B5_NST);
    }
    __syncwarp();
    if (active) {
      uint32_t sv[B5_KT], dv[B5_KT];
      tmem_ld32_issue(lane_base + h * 32, sv);
      tmem_ld32_issue(lane_base + 64 + h * 32, dv);
      tmem_ld_wait();
      const float4* kb = reinterpret_cast<const float4*>(kbias + t * B5_KT);
#pragma unroll
      for (int j8 = 0; j8 < B5_KT / 8; ++j8) {
        uint32_t w[4];
        const float4 b0 = kb[2 * j8], b1 = kb[2 * j8 + 1];
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = 8 * j8 + 2 * u;
          const float p0 = ex2(fmaf(__uint_as_float(sv[j]), c, bb[2 * u] + nl));
          const float p1 = ex2(fmaf(__uint_as_float(sv[j + 1]), c, bb[2 * u + 1] + nl));
          float g0 = __uint_as_float(dv[j]), g1 = __uint_as_float(dv[j + 1]);
          if (DROP) {                                       // dP = mask * (dO V^T)
            if (keep_bits != nullptr) {
              g0 = ((kcur >> j) & 1u) ? g0 * drop.scale : 0.f;
              g1 = ((kcur >> (j + 1)) & 1u) ? g1 * drop.scale : 0.f;
            } else {
              const float2 mk = drop_pair(dkey, drow + static_cast<unsigned>(t * (B5_KT / 2) + (j >> 1)), drop.thr, drop.scale);
              g0 *= mk.x; g1 *= mk.y;
            }
          }
          w[u] = pack_bf16x2(p0 * (g0 - dl), p1 * (g1 - dl));
        }
        sts128(d_row + ((static_cast<uint32_t>(j8) ^ swz) << 4), w[0], w[1], w[2], w[3]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(h);
    if (quarter == (t & 3) && lane == 0) {
      tc_fence_after();
      const uint32_t aK = smem_u32(sK + (t % B5_NST) * B5_TILE_STR) + static_cast<uint32_t>(h * 64);
#pragma unroll
      for (int k = 0; k < B5_KT / UMMA_K; ++k)             // dQ_h += dS_h K_h   (K tile as the MN-major operand)
        tc_mma_bf16(tmem_base + 128 + h * 32, desc_kmajor(aD, k * 32), desc_mnmajor(aK, k), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
      tc_commit(&bars->str_empty[t % B5_NST]);
      if (t + 1 < ntiles) issue_products(t + 1);
      else tc_commit(&bars->acc_full[h]);
    }
    __syncwarp();
    kcur = knext;
  }
  mbar_wait(&bars->acc_full[h], 0);
  tc_fence_after();
  if (active) {
    uint32_t acc[A5_DH];
    tmem_ld32_issue(lane_base + 128 + h * 32, acc);
    tmem_ld_wait();
    if (qi < Tq) {
      bf16* dst = dQ + (static_cast<size_t>(b) * Tq + qi) * lddq + head * A5_DH;
#pragma unroll
      for (int j = 0; j < A5_DH / 8; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(acc[8 * j]) * scale, __uint_as_float(acc[8 * j + 1]) * scale);
        v.y = pack_bf16x2(__uint_as_float(acc[8 * j + 2]) * scale, __uint_as_float(acc[8 * j + 3]) * scale);
        v.z = pack_bf16x2(__uint_as_float(acc[8 * j + 4]) * scale, __uint_as_float(acc[8 * j + 5]) * scale);
        v.w = pack_bf16x2(__uint_as_float(acc[8 * j + 6]) * scale, __uint_as_float(acc[8 * j + 7]) * scale);
        *reinterpret_cast<uint4*>(dst + 8 * j) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// SPLIT = 2 (experiment, off by default -- see the launcher): two threads per key row, each owning 16 of a streamed
// tile's 32 query columns (the warps w and w + 4 of a head share a tensor-memory lane quarter, like the two epilogue
// warps of gemm_tc.cu): 32 resident warps per SM instead of 16.
template <bool DROP, int SPLIT>
__global__ void __launch_bounds__(A5_THREADS * SPLIT, 2)
attn5_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                     const float* __restrict__ kmask, const float* __restrict__ lse, const float* __restrict__ delta,
                     bf16* __restrict__ dK, long long lddk, bf16* __restrict__ dV, long long lddv, int H, int Tq, int Tk,
                     float scale, DropSpec drop, const unsigned* __restrict__ keep_bits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                                  // [128 x 128 B] K rows (both heads)
  uint8_t* sV = sK + A5_TILE_OWN;                      // [128 x 128 B] V rows
  uint8_t* sQ = sV + A5_TILE_OWN;                      // NST x [32 x 128 B]
  uint8_t* sG = sQ + B5_NST_KV * B5_TILE_STR;          // NST x [32 x 128 B] dO rows
  uint8_t* sD = sG + B5_NST_KV * B5_TILE_STR;          // 2 heads x [128 x 128 B]: P^T in bytes 0..63 of a row, dS^T in 64..127
  BwdBars<B5_NST_KV>* bars = reinterpret_cast<BwdBars<B5_NST_KV>*>(sD + 2 * A5_TILE_OWN);
  // per streamed tile and head: -lse in log2 units (-inf past the sample: P = 0) and delta of its 32 queries, double
  // buffered: [buffer][head][0: -lse, 1: delta][32]
  float* sLD = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, hp = blockIdx.y, k0 = blockIdx.x * A5_ROWS;
  const int ntiles = (Tq + B5_KT - 1) / B5_KT;
  // (warp 1 of each head fetches them for the NEXT tile before the head's barrier of the current one)
  auto fetch_ld = [&](int t, int hh) {
    const int q = t * B5_KT + lane;
    const size_t idx = (static_cast<size_t>(b) * H + hp * 2 + hh) * Tq + q;
    float* dst = sLD + ((t & 1) * 2 + hh) * 64;
    dst[lane] = (q < Tq) ? -lse[idx] * kLog2e : -INFINITY;
    dst[32 + lane] = (q < Tq) ? delta[idx] : 0.f;
  };
  pdl_sync();
  constexpr int WPH = 4 * SPLIT;                       // warps per head
  constexpr int CW = B5_KT / SPLIT;                    // query columns of a tile per thread
  if ((warp % WPH) == 1) fetch_ld(0, warp / WPH);

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
      mbar_init(&bars->own_full, 1);
      for (int i = 0; i < B5_NST_KV; ++i) { mbar_init(&bars->str_full[i], 1); mbar_init(&bars->str_empty[i], 2); }
      for (int i = 0; i < 2; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->acc_full[i], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  // tensor-memory columns: S^T_h at [32 h, +32), dP^T_h at [64 + 32 h, +32), dV_h at [128 + 32 h, +32), dK_h at [192 + 32 h, +32)

  const int h = warp / WPH, quarter = warp & 3, half = (warp % WPH) >> 2;   // half: which CW columns of a tile
  const int c0 = half * CW;
  const int col0 = hp * 2 * A5_DH;
  constexpr uint32_t idesc_s = make_idesc<B5_KT, 0, 0, A5_ROWS>();
  constexpr uint32_t idesc_acc = make_idesc<A5_DH, 0, 1, A5_ROWS>();
  const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aD = smem_u32(sD + h * A5_TILE_OWN);
  auto load_tile = [&](int t) {
    const int st = t % B5_NST_KV;
    mbar_expect_tx(&bars->str_full[st], 2 * B5_TILE_STR);
    tma_load_3d(&tmQ, &bars->str_full[st], sQ + st * B5_TILE_STR, col0, b * Tq + t * B5_KT, 0);
    tma_load_3d(&tmG, &bars->str_full[st], sG + st * B5_TILE_STR, col0, b * Tq + t * B5_KT, 0);
  };
  auto issue_products = [&](int t) {                        // S^T_h(t) = K_h Q_h^T, dP^T_h(t) = V_h dO_h^T
    const int st = t % B5_NST_KV;
    mbar_wait(&bars->str_full[st], (t / B5_NST_KV) & 1);
    tc_fence_after();
    const uint32_t aQ = smem_u32(sQ + st * B5_TILE_STR), aG = smem_u32(sG + st * B5_TILE_STR);
#pragma unroll
    for (int k = 0; k < A5_DH / UMMA_K; ++k)
      tc_mma_bf16(tmem_base + h * 32, desc_kmajor(aK, h * 64 + k * 32), desc_kmajor(aQ, h * 64 + k * 32), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < A5_DH / UMMA_K; ++k)
      tc_mma_bf16(tmem_base + 64 + h * 32, desc_kmajor(aV, h * 64 + k * 32), desc_kmajor(aG, h * 64 + k * 32), idesc_s, k > 0 ? 1u : 0u);
    tc_commit(&bars->s_full[h]);
  };
  if (quarter == 3 && half == 0 && lane == 0) {
    if (h == 0) {
      mbar_expect_tx(&bars->own_full, 2 * A5_TILE_OWN);
      tma_load_3d(&tmK, &bars->own_full, sK, col0, b * Tk + k0, 0);
      tma_load_3d(&tmV, &bars->own_full, sV, col0, b * Tk + k0, 0);
      for (int t = 0; t < B5_NST_KV && t < ntiles; ++t) load_tile(t);
    }
    mbar_wait(&bars->own_full, 0);
    issue_products(0);
  }
  __syncwarp();

  const int row = quarter * 32 + lane;
  const int kj = k0 + row;
  const bool active = (k0 + quarter * 32) < Tk;
  const int head = hp * 2 + h;
  const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  const float c = scale * kLog2e;
  const uint32_t d_row = aD + static_cast<uint32_t>(row) * 128u;
  const uint32_t swz = static_cast<uint32_t>(row & 7);
  const float kb = (kj < Tk && (kmask == nullptr || kmask[static_cast<size_t>(b) * Tk + kj] != 0.f)) ? 0.f : -INFINITY;
  DropKey dkey{0u, 1u};
  unsigned half_tk = 0u, dcol = 0u, dhalf = 0u, drow0 = 0u;
  if (DROP) {
    dkey = drop_key(drop);
    half_tk = static_cast<unsigned>((Tk + 1) / 2);
    dcol = static_cast<unsigned>(kj) >> 1; dhalf = static_cast<unsigned>(kj) & 1u;
    drow0 = (static_cast<unsigned>(b) * H + head) * Tq;
  }

  // keep bits written by the forward kernel: word (query, kw) holds the decisions of keys 32 kw .. 32 kw + 31 -- exactly
  // this warp's 32 key rows.  Lane l fetches the word of query qt + l (one tile ahead); a 32 x 32 bit transpose across the
  // warp (5 butterfly steps) then leaves lane l with its own key's decisions for the tile's 32 queries.
  const int kwords = (Tk + 31) >> 5;
  const unsigned* kbase = nullptr;
  unsigned kraw = 0u;
  if (DROP && keep_bits != nullptr && active) {
    kbase = keep_bits + (static_cast<size_t>(b) * H + head) * Tq * kwords + ((k0 + quarter * 32) >> 5);
    if (lane < Tq) kraw = __ldg(kbase + static_cast<size_t>(lane) * kwords);
  }
  for (int t = 0; t < ntiles; ++t) {
    unsigned knext = 0u, kcol = 0u;
    if (DROP && kbase != nullptr) {
      const int qn = (t + 1) * B5_KT + lane;
      if (qn < Tq) knext = __ldg(kbase + static_cast<size_t>(qn) * kwords);
      kcol = kraw;
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        const unsigned m = sft == 16 ? 0x0000FFFFu : sft == 8 ? 0x00FF00FFu : sft == 4 ? 0x0F0F0F0Fu : sft == 2 ? 0x33333333u : 0x55555555u;
        const unsigned pr = __shfl_xor_sync(0xffffffffu, kcol, sft);
        kcol = (lane & sft) ? (((pr & ~m) >> sft) | (kcol & ~m)) : ((kcol & m) | ((pr & m) << sft));
      }
    }
    mbar_wait(&bars->s_full[h], t & 1);
    tc_fence_after();
    if (h == 0 && half == 0 && quarter == (t & 3) && lane == 0 && t >= 1 && t - 1 + B5_NST_KV < ntiles) {
      mbar_wait(&bars->str_empty[(t - 1) % B5_NST_KV], ((t - 1) / B5_NST_KV) & 1);
      load_tile(t - 1 + B5_NST_KV);
    }
    __syncwarp();
    if (active) {
      const int qt = t * B5_KT;
      const bool partial = qt + B5_KT > Tq;                 // the last tile: columns past the sample carry other rows' data
      uint32_t sv[CW], dv[CW];
      if (SPLIT == 1) {
        tmem_ld32_issue(lane_base + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld32_issue(lane_base + 64 + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[0]));
      } else {
        tmem_ld16_issue(lane_base + h * 32 + c0, *reinterpret_cast<uint32_t(*)[16]>(&sv[0]));
        tmem_ld16_issue(lane_base + 64 + h * 32 + c0, *reinterpret_cast<uint32_t(*)[16]>(&dv[0]));
      }
      tmem_ld_wait();
      const float4* ld4 = reinterpret_cast<const float4*>(sLD + ((t & 1) * 2 + h) * 64) + c0 / 4;
#pragma unroll
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        const float4 l0 = ld4[2 * j8], l1 = ld4[2 * j8 + 1], e0 = ld4[8 + 2 * j8], e1 = ld4[8 + 2 * j8 + 1];
        const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float dq_[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        uint32_t wp[4], wd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float pv[2], dsv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 8 * j8 + 2 * u + e;
            float p = ex2(fmaf(__uint_as_float(sv[j]), c, kb) + lq[2 * u + e]);
            float g = __uint_as_float(dv[j]);
            if (partial && !(qt + c0 + j < Tq)) { p = 0.f; g = 0.f; }      // never let a foreign row's NaN through 0 * NaN
            float w = p;
            if (DROP) {
              if (keep_bits != nullptr) {
                const bool kp = ((kcol >> (c0 + j)) & 1u) != 0u;
                w = kp ? w * drop.scale : 0.f;
                g = kp ? g * drop.scale : 0.f;
              } else {
                const float mk = drop_one(dkey, (drow0 + static_cast<unsigned>(qt + c0 + j)) * half_tk + dcol, dhalf, drop.thr, drop.scale);
                w *= mk; g *= mk;
              }
            }
            pv[e] = w;
            dsv[e] = p * (g - dq_[2 * u + e]);
          }
          wp[u] = pack_bf16x2(pv[0], pv[1]);
          wd[u] = pack_bf16x2(dsv[0], dsv[1]);
        }
        sts128(d_row + ((static_cast<uint32_t>(c0 / 8 + j8) ^ swz) << 4), wp[0], wp[1], wp[2], wp[3]);
        sts128(d_row + ((static_cast<uint32_t>(4 + c0 / 8 + j8) ^ swz) << 4), wd[0], wd[1], wd[2], wd[3]);
      }
      // The last tile's rows past the sample are another sample's dO rows (TMA boxes do not know about samples): their
      // P^T / dS^T columns are zero, but 0 x NaN inside the tensor core is NaN, and a fully padded neighbour hands back
      // NaN gradients.  This head's 64 bytes of those rows are cleared before they become the B operand of dV (the
      // products that read them as dP^T have retired: s_full).
      if (partial && quarter == 0 && half == 0 && qt + lane >= Tq) {
        const uint32_t g_row = smem_u32(sG + (t % B5_NST_KV) * B5_TILE_STR) + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
        for (int cix = 0; cix < 4; ++cix)
          sts128(g_row + ((static_cast<uint32_t>(4 * h + cix) ^ static_cast<uint32_t>(lane & 7)) << 4), 0u, 0u, 0u, 0u);
      }
    }
    if (quarter == 1 && half == 0 && t + 1 < ntiles) fetch_ld(t + 1, h);
    fence_async_smem();
    tc_fence_before();
    group_sync_n(h, 128 * SPLIT);
    if (half == 0 && quarter == (t & 3) && lane == 0) {
      tc_fence_after();
      const int st = t % B5_NST_KV;
      const uint32_t aG = smem_u32(sG + st * B5_TILE_STR) + static_cast<uint32_t>(h * 64);
      const uint32_t aQ = smem_u32(sQ + st * B5_TILE_STR) + static_cast<uint32_t>(h * 64);
#pragma unroll
      for (int k = 0; k < B5_KT / UMMA_K; ++k)             // dV_h += P^T_h dO_h
        tc_mma_bf16(tmem_base + 128 + h * 32, desc_kmajor(aD, k * 32), desc_mnmajor(aG, k), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < B5_KT / UMMA_K; ++k)             // dK_h += dS^T_h Q_h
        tc_mma_bf16(tmem_base + 192 + h * 32, desc_kmajor(aD, 64 + k * 32), desc_mnmajor(aQ, k), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
      tc_commit(&bars->str_empty[st]);
      if (t + 1 < ntiles) issue_products(t + 1);
      else tc_commit(&bars->acc_full[h]);
    }
    __syncwarp();
    kraw = knext;
  }
  mbar_wait(&bars->acc_full[h], 0);
  tc_fence_after();
  if (active) {
    constexpr int EW = A5_DH / SPLIT;                  // accumulator columns per thread
    uint32_t av[EW], ak[EW];
    if (SPLIT == 1) {
      tmem_ld32_issue(lane_base + 128 + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&av[0]));
      tmem_ld32_issue(lane_base + 192 + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&ak[0]));
    } else {
      tmem_ld16_issue(lane_base + 128 + h * 32 + half * EW, *reinterpret_cast<uint32_t(*)[16]>(&av[0]));
      tmem_ld16_issue(lane_base + 192 + h * 32 + half * EW, *reinterpret_cast<uint32_t(*)[16]>(&ak[0]));
    }
    tmem_ld_wait();
    if (kj < Tk) {
      const size_t r = static_cast<size_t>(b) * Tk + kj;
      bf16* pk = dK + r * lddk + head * A5_DH + half * EW;
      bf16* pv = dV + r * lddv + head * A5_DH + half * EW;
#pragma unroll
      for (int j = 0; j < EW / 8; ++j) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(ak[8 * j]) * scale, __uint_as_float(ak[8 * j + 1]) * scale);
        v.y = pack_bf16x2(__uint_as_float(ak[8 * j + 2]) * scale, __uint_as_float(ak[8 * j + 3]) * scale);
        v.z = pack_bf16x2(__uint_as_float(ak[8 * j + 4]) * scale, __uint_as_float(ak[8 * j + 5]) * scale);
        v.w = pack_bf16x2(__uint_as_float(ak[8 * j + 6]) * scale, __uint_as_float(ak[8 * j + 7]) * scale);
        *reinterpret_cast<uint4*>(pk + 8 * j) = v;
        v.x = pack_bf16x2(__uint_as_float(av[8 * j]), __uint_as_float(av[8 * j + 1]));
        v.y = pack_bf16x2(__uint_as_float(av[8 * j + 2]), __uint_as_float(av[8 * j + 3]));
        v.z = pack_bf16x2(__uint_as_float(av[8 * j + 4]), __uint_as_float(av[8 * j + 5]));
        v.w = pack_bf16x2(__uint_as_float(av[8 * j + 6]), __uint_as_float(av[8 * j + 7]));
        *reinterpret_cast<uint4*>(pv + 8 * j) = v;
      }
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
  static const int narrow = getenv("SER_ATTN_NARROW") ? atoi(getenv("SER_ATTN_NARROW")) : 1;     // A/B switch
  SER_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(A5_THREADS), smem, s, tmQ, tmK, tmV, a.kmask, reinterpret_cast<bf16*>(a.O), a.ldo, a.lse, a.H, a.Tq, a.Tk,
                                      a.scale, a.drop, narrow, a.keep_bits));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

int attention_bwd_tc5(const AttnArgs& a, cudaStream_t s) {
  const double fl = 10.0 * a.B * a.H * static_cast<double>(a.Tq) * a.Tk * a.dh;
  const double by = 2.0 * static_cast<double>(a.B) * a.H * a.dh * (4.0 * a.Tq + 4.0 * a.Tk);
  ProfScope prof("attention_bwd", fl, by, s);
  auto ok_ptr = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SER_REQUIRE(ok_ptr(a.dO) && ok_ptr(a.dQ) && ok_ptr(a.dK) && ok_ptr(a.dV) && a.lddo % 8 == 0 && a.lddq % 8 == 0 &&
                  a.lddk % 8 == 0 && a.lddv % 8 == 0,
              "attention_bwd_tc5: gradient buffers must be 16-byte aligned with 8-element aligned leading dimensions");
  const int HD = a.H * a.dh;
  const long long Mq = static_cast<long long>(a.B) * a.Tq, Mk = static_cast<long long>(a.B) * a.Tk;
  const bool drop = a.drop.on();
  {
    CUtensorMap tmQ, tmG, tmK, tmV;
    SER_TRY(make_tmap(&tmQ, a.Q, 0, Mq, HD, a.ldq, A5_ROWS, 64, 1, 0));
    SER_TRY(make_tmap(&tmG, a.dO, 0, Mq, HD, a.lddo, A5_ROWS, 64, 1, 0));
    SER_TRY(make_tmap(&tmK, a.K, 0, Mk, HD, a.ldk, B5_KT, 64, 1, 0));
    SER_TRY(make_tmap(&tmV, a.V, 0, Mk, HD, a.ldv, B5_KT, 64, 1, 0));
    const int ntiles = ceil_div(a.Tk, B5_KT);
    const int smem = 1024 + 4 * A5_TILE_OWN + 2 * B5_NST * B5_TILE_STR + 256 + ntiles * B5_KT * static_cast<int>(sizeof(float));
    auto* kern = drop ? attn5_bwd_dq_kernel<true> : attn5_bwd_dq_kernel<false>;
    static bool configured[2] = {false, false};
    if (!configured[drop ? 1 : 0]) {
      SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          1024 + 4 * A5_TILE_OWN + 2 * B5_NST * B5_TILE_STR + 256 + (8192 + 32) * 4));
      SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      configured[drop ? 1 : 0] = true;
    }
    dim3 grid(ceil_div(a.Tq, A5_ROWS), a.H / 2, a.B);
    SER_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(A5_THREADS), smem, s, tmQ, tmG, tmK, tmV, a.kmask, reinterpret_cast<const bf16*>(a.O), a.ldo,
                                        reinterpret_cast<const bf16*>(a.dO), a.lddo, a.lse, a.delta,
                                        reinterpret_cast<bf16*>(a.dQ), a.lddq, a.H, a.Tq, a.Tk, a.scale, a.drop,
                                        static_cast<const unsigned*>(a.keep_bits)));
    SER_LAUNCH_CHECK();
  }
  {
    CUtensorMap tmK, tmV, tmQ, tmG;
    SER_TRY(make_tmap(&tmK, a.K, 0, Mk, HD, a.ldk, A5_ROWS, 64, 1, 0));
    SER_TRY(make_tmap(&tmV, a.V, 0, Mk, HD, a.ldv, A5_ROWS, 64, 1, 0));
    SER_TRY(make_tmap(&tmQ, a.Q, 0, Mq, HD, a.ldq, B5_KT, 64, 1, 0));
    SER_TRY(make_tmap(&tmG, a.dO, 0, Mq, HD, a.lddo, B5_KT, 64, 1, 0));
    const int smem = 1024 + 4 * A5_TILE_OWN + 2 * B5_NST_KV * B5_TILE_STR + 256 + 1024;
    // A/B switch.  Measured (fwd + bwd, dropout 0.1, stored keep bits): (1500, 256) 1196 us with one thread per key row,
    // 1298 us with two; (256, 1500) 1141 vs 1272 us -- the second set of warps pays a second bit transpose, a 256-thread
    // named barrier per tile and 72 bytes of spills at the 64-register cap, and the MMA / barrier chain per tile, not the
    // element-wise latency, is what the kernel waits on.  One thread per row stays the default.
    static const int split = (getenv("SER_ATTN_DKV_SPLIT") && atoi(getenv("SER_ATTN_DKV_SPLIT")) == 2) ? 2 : 1;
    auto* kern = split == 2 ? (drop ? attn5_bwd_dkv_kernel<true, 2> : attn5_bwd_dkv_kernel<false, 2>)
                            : (drop ? attn5_bwd_dkv_kernel<true, 1> : attn5_bwd_dkv_kernel<false, 1>);
    static bool configured[2] = {false, false};
    if (!configured[drop ? 1 : 0]) {
      SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      configured[drop ? 1 : 0] = true;
    }
    dim3 grid(ceil_div(a.Tk, A5_ROWS), a.H / 2, a.B);
    SER_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(A5_THREADS * split), smem, s, tmK, tmV, tmQ, tmG, a.kmask, a.lse, a.delta, reinterpret_cast<bf16*>(a.dK), a.lddk,
                                        reinterpret_cast<bf16*>(a.dV), a.lddv, a.H, a.Tq, a.Tk, a.scale, a.drop,
                                        static_cast<const unsigned*>(a.keep_bits)));
    SER_LAUNCH_CHECK();
  }
  return SER_OK;
}

}  // namespace ser

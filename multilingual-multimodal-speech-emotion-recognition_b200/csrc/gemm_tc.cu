// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// One persistent CTA per SM walks output tiles of 128 x BN (BN = 128 or 256).  Warp roles:
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled shared-memory stages)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, M=128, N=BN, K=16)
//   warps 2..9  : epilogue (two warps per TMEM lane quarter, alternating chunks).  Every lane owns one output row (tcgen05.ld 32x32b).  The row is finished in
//                 registers -- alpha, bias, activation, gate derivative, residual -- packed, written into a
//                 128B-swizzled staging tile (32 rows x 128 B per warp) and leaves the SM as ONE TMA store
//                 (or TMA reduce-add for split-K / accumulate).  Residual / gate operands arrive the same way:
//                 a TMA load of the matching 32 x 128 B box, issued one chunk ahead.  No per-thread global
//                 address arithmetic, no row predicates (TMA clips the M tail), fully coalesced traffic.
// The accumulator is double buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.  Both operands may be K-major or MN-major (nn.Linear weights are
// consumed in place for forward, dX and dW GEMMs: no transposed copies are ever materialised).
//
// Bias gradients ride along (RS variants, dW GEMMs): db = column sums of dY = row sums of the MN-major A operand.
// CTAs of output column tile 0 issue, per k-block, four extra N = 16 MMAs of the A tile against a constant tile of
// ones into 16 spare TMEM columns; the epilogue adds column 0 of that accumulator to rowsum[] with one atomic per
// row.  dY is not read a second time and the 18 colsum launches of a training step disappear.  (RS kernels use a
// single accumulator buffer: split-K sizes these problems to one work item per CTA anyway.)
//
// Serves: adapters, q/k/v + MHA in/out projections, out_a/out_t, pooling scorer, fusion projections,
// the classifier heads (reference: src/models/audio_encoder.py:19-21, cross_attention.py:38-51,
// pooling.py:9-13, fusion.py:8-16, classifier.py:73-129) in the bf16 tier.
#include "common.cuh"
#include "prof.cuh"
#include "tc5.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <mutex>

namespace ser {

namespace {

constexpr int BM = 128;       // UMMA M (one TMEM lane per output row)
constexpr int BK = 64;        // 64 bf16 = one 128-byte swizzle atom along the contraction axis
constexpr int kEpiWarps = 8;          // two per TMEM lane quarter: they interleave the 128-byte chunks of a tile
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemTotal = 232448;                    // 227 KB opt-in maximum per CTA
constexpr int kStgTile = 32 * 128;                    // one staged chunk: 32 rows x 128 B (SWIZZLE_128B)
constexpr int kBarBytes = 1024;                       // mbarriers, TMEM slot, and (offset 512) the 512-byte tile of ones
constexpr int kMaxStages = 8;

// Shared memory is carved at run time: an epilogue warp needs a 4 KB output staging tile and, only when a residual /
// gate operand is streamed in, a 4 KB source tile.  GEMMs without a source operand (all dW GEMMs, most forward ones)
// turn the 32 KB saved into one more pipeline stage: the main loop is bound by the LATENCY of a stage refill (tensor
// pipe 50 % active with 3 stages of 48 KB in flight -- a 2-CTA multicast that halves the bytes per CTA changes
// nothing), so depth is what buys throughput.
template <int BN> struct TileCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;      // double-buffered accumulator
  static constexpr int stg_bytes(bool with_src) { return kEpiWarps * (with_src ? 2 : 1) * kStgTile; }
  // bytes of one pipeline stage in ONE CTA: a CTA of a cta_group::2 pair holds its own A rows and HALF of the B tile
  static constexpr int stage_bytes(bool pair) { return kABytes + (pair ? kBBytes / 2 : kBBytes); }
  // 3 / 4 stages (BN = 256, with / without source), 5 / 6 (BN = 128); pairs: 5 / 6 (BN = 256), 6 / 8 (BN = 128)
  static constexpr int stages(bool with_src, bool pair) {
    return (kSmemTotal - 1024 /*align*/ - stg_bytes(with_src) - kBarBytes) / stage_bytes(pair) < kMaxStages
               ? (kSmemTotal - 1024 - stg_bytes(with_src) - kBarBytes) / stage_bytes(pair) : kMaxStages;
  }
};

enum : int { SRC_NONE = 0, SRC_RESIDUAL = 1, SRC_GATE = 2 };

struct TcEpilogue {
  const float* bias;
  int c_f32;         // output (and source) element type: fp32 or bf16
  int src;           // SRC_*: a second [M,N] operand streamed through TMA (same dtype as the output)
  int gate_mode;
  int act;
  int atomic;        // accumulate with TMA reduce-add (split-K or C +=); fp32 output only
  float alpha;
  long long bias_stride;   // batch stride of bias (elements)
  float* rowsum;     // RS kernels: rowsum[m] += sum_k op(A)[m, k]  (pre-zeroed by the host side)
  long long rs_stride;   // batch stride of rowsum (elements)
};

using namespace tc5;

// ------------------------------------------------------------------------------------------------
// epilogue of one 128 x BN tile for one warp (32 rows).  OUT_F32: 32 fp32 columns per 128-byte chunk,
// otherwise 64 bf16 columns.  SRC: residual / gate operand of the same dtype, streamed by TMA.
// ------------------------------------------------------------------------------------------------
struct EpiWarp {
  uint32_t base;              // shared address of this warp's staging tiles: output, then (if streamed) source
  __device__ __forceinline__ uint32_t cst() const { return base; }
  __device__ __forceinline__ uint32_t sst() const { return base + kStgTile; }
  uint64_t* src_full;         // mbarrier of the source tile
  uint32_t sphase;            // parity of its next completion
  int half;                   // 0 / 1: which of the two warps of this lane quarter (takes chunks half, half+2, ...)
};

template <int BN, bool OUT_F32, int SRC>
__device__ __forceinline__ void epilogue_tile(const TcEpilogue& ep, const CUtensorMap* tmC, const CUtensorMap* tmS,
                                              EpiWarp& w, uint32_t taddr0, int lane, int n0, int mw, int bz,
                                              bool lead_split, uint64_t* tfull, uint32_t tfull_parity) {
  constexpr int CW = OUT_F32 ? 32 : 64;          // output columns per 128-byte chunk
  constexpr int NCH = BN / CW;
  const bool with_src = (SRC != SRC_NONE) && (SRC == SRC_GATE || lead_split);
  const uint32_t swz = static_cast<uint32_t>(lane & 7);
  const uint32_t rowoff = static_cast<uint32_t>(lane) * 128u;
  // Eight epilogue warps: the two warps of a lane quarter take alternating chunks, so while one waits on tensor
  // memory, its source tile or the store engine, the other issues -- the epilogue of the K = 256 shapes was bound by
  // the latency of a single warp per scheduler (ncu: 0.25 eligible warps, 23 % issue-slot use).
  // The first source chunk of this warp is requested before the accumulator is even complete.
  if (with_src && lane == 0 && w.half < NCH) {
    mbar_expect_tx(w.src_full, kStgTile);
    tma_load_3d_raw(tmS, w.src_full, w.sst(), n0 + w.half * CW, mw, bz);
  }
  mbar_wait(tfull, tfull_parity);
  tc_fence_after();
#pragma unroll 1
  for (int c = w.half; c < NCH; c += 2) {
    uint32_t raw[CW];
    tmem_ld32_issue(taddr0 + c * CW, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
    if (!OUT_F32) tmem_ld32_issue(taddr0 + c * CW + 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[CW - 32]));
    // the output tile of this warp's previous chunk must have been read by its TMA store before it is overwritten
    if (lane == 0) tma_wait_group_read<0>();
    __syncwarp();
    tmem_ld_wait();
    float v[CW];
    const float alpha = ep.alpha;
    if (ep.bias != nullptr && lead_split) {
      const float4* bp = reinterpret_cast<const float4*>(ep.bias + bz * ep.bias_stride + n0 + c * CW);
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 b = __ldg(bp + j);
        v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), alpha, b.x);
        v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), alpha, b.y);
        v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), alpha, b.z);
        v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), alpha, b.w);
      }
    } else if (alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(raw[j]) * alpha;
    } else {                               // the common dgrad / wgrad case: no bias, no scale -> no arithmetic at all
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(raw[j]);
    }
    if (ep.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (ep.act == ACT_TANH) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = tanhf(v[j]);
    } else if (ep.act == ACT_SIGMOID) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
    }
    if (SRC != SRC_NONE && with_src) {
      mbar_wait(w.src_full, w.sphase);
      w.sphase ^= 1u;
      const uint32_t sbase = w.sst() + rowoff;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                       // 8 x 16 bytes of this lane's source row
        uint32_t x0, x1, x2, x3;
        lds128(sbase + ((static_cast<uint32_t>(j) ^ swz) << 4), x0, x1, x2, x3);
        if (OUT_F32) {
          const float s0 = __uint_as_float(x0), s1 = __uint_as_float(x1), s2 = __uint_as_float(x2), s3 = __uint_as_float(x3);
          if (SRC == SRC_RESIDUAL) {
            v[4 * j] += s0; v[4 * j + 1] += s1; v[4 * j + 2] += s2; v[4 * j + 3] += s3;
          } else {
            v[4 * j] = apply_gate(v[4 * j], s0, ep.gate_mode); v[4 * j + 1] = apply_gate(v[4 * j + 1], s1, ep.gate_mode);
            v[4 * j + 2] = apply_gate(v[4 * j + 2], s2, ep.gate_mode); v[4 * j + 3] = apply_gate(v[4 * j + 3], s3, ep.gate_mode);
          }
        } else {
          const uint32_t xs[4] = {x0, x1, x2, x3};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (SRC == SRC_RESIDUAL) {
              v[8 * j + 2 * t] += bf_lo(xs[t]); v[8 * j + 2 * t + 1] += bf_hi(xs[t]);
            } else if (ep.gate_mode == GATE_RELU) {
              // d relu: keep where the saved activation is > 0, decided on the packed bf16 bits (sign clear and
              // magnitude non-zero) without unpacking
              const uint32_t w = xs[t];
              if ((w & 0x00008000u) != 0u || (w & 0x00007fffu) == 0u) v[8 * j + 2 * t] = 0.f;
              if ((w & 0x80000000u) != 0u || (w & 0x7fff0000u) == 0u) v[8 * j + 2 * t + 1] = 0.f;
            } else {
              v[8 * j + 2 * t] = apply_gate(v[8 * j + 2 * t], bf_lo(xs[t]), ep.gate_mode);
              v[8 * j + 2 * t + 1] = apply_gate(v[8 * j + 2 * t + 1], bf_hi(xs[t]), ep.gate_mode);
            }
          }
        }
      }
      // source tile consumed by every lane: request this warp's next chunk into it
      __syncwarp();
      if (c + 2 < NCH && lane == 0) {
        mbar_expect_tx(w.src_full, kStgTile);
        tma_load_3d_raw(tmS, w.src_full, w.sst(), n0 + (c + 2) * CW, mw, bz);
      }
    }
    // finished row -> swizzled staging tile -> one TMA store per warp and chunk
    const uint32_t cbase = w.cst() + rowoff;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = cbase + ((static_cast<uint32_t>(j) ^ swz) << 4);
      if (OUT_F32) {
        sts128(dst, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
               __float_as_uint(v[4 * j + 3]));
      } else {
        sts128(dst, pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]),
               pack2(v[8 * j + 6], v[8 * j + 7]));
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (ep.atomic) tma_reduce_add_3d(tmC, w.cst(), n0 + c * CW, mw, bz);
      else tma_store_3d(tmC, w.cst(), n0 + c * CW, mw, bz);
      tma_commit_group();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int kOnesN = 16;                       // N of the row-sum MMA (the smallest N an M = 128 UMMA takes)
__device__ __forceinline__ constexpr uint32_t make_idesc_ones(int amaj, int mm) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(amaj) << 15) | (0u << 16)   // B (ones): K-major
         | (static_cast<uint32_t>(kOnesN >> 3) << 17) | (static_cast<uint32_t>(mm >> 4) << 24);
}

// PAIR: launched as 2-CTA clusters driving ONE cta_group::2 MMA per k-step: the pair owns a 256 x BN output tile,
// each CTA stores its own 128 A rows and HALF of the B tile (so a stage is 16 + BN/16 KB instead of 16 + BN/8 KB:
// less shared-memory traffic per MMA cycle and a deeper pipeline), the leader CTA's elected thread issues the MMAs, and
// each CTA's tensor memory receives its 128 accumulator rows.  All TMA bytes of a stage are counted on the LEADER's
// full barrier; tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the follower's
// epilogue warps release the accumulator on the leader's barrier.
template <int BN, int AMAJ, int BMAJ, bool RS, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS,
               const TcEpilogue ep, const int M, const int N, const int K, const int splits, const int batch,
               const int nstages) {
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the shared address space visible to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stg_stride = (ep.src != SRC_NONE ? 2 : 1) * kStgTile;     // bytes of staging per epilogue warp
  constexpr int kBStage = PAIR ? Cfg::kBBytes / 2 : Cfg::kBBytes;       // bytes of B per stage held by this CTA
  constexpr int kStage = Cfg::kABytes + kBStage;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nstages * Cfg::kABytes;
  uint8_t* smem_stg = smem + nstages * kStage;                       // 1024-byte aligned: 8 or 16 staging tiles of 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + kEpiWarps * stg_stride);
  uint64_t* full_bar = bars;                         // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;           // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint64_t* src_bar = tempty_bar + 2;                // [kEpiWarps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(src_bar + kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = N / BN;
  const int kblocks = (K + BK - 1) / BK;
  // work units: PAIR -> one unit = the pair of row tiles (2u, 2u+1) handled by the two CTAs of a cluster in lockstep
  const int crank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_units = PAIR ? m_tiles / 2 : m_tiles;
  const int total_tiles = m_units * n_tiles * splits * batch;
  const int unit0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], PAIR ? 2 * kEpiWarps : kEpiWarps); }
    for (int s = 0; s < kEpiWarps; ++s) mbar_init(&src_bar[s], 1);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    if (ep.src != SRC_NONE) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmS)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {       // both CTAs of the pair, same warp id
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(Cfg::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(Cfg::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // RS: the B operand of the row-sum MMA is a 16 x 16 block of bf16 ones in the un-swizzled core-matrix layout
  // (four 128-byte core matrices = 512 bytes; all ones, so the LBO / SBO order is immaterial)
  uint8_t* ones_tile = reinterpret_cast<uint8_t*>(bars) + 512;
  if (RS && warp >= 2) {
    const int i = threadIdx.x - 64;
    if (i < 512 / 16) sts128(smem_u32(ones_tile) + 16u * i, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (PAIR) cluster_sync();       // the peer's barriers exist before anything is multicast to them
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();                     // everything above is independent of the previous kernel's output

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int nt = tile % n_tiles;
        const int rest = tile / n_tiles;
        const int mt = PAIR ? 2 * (rest % m_units) + crank : rest % m_units;
        const int rest2 = rest / m_units;
        const int sp = rest2 % splits;
        const int bz = rest2 / splits;
        const int kb0 = static_cast<int>((static_cast<long long>(sp) * kblocks) / splits);
        const int kb1 = static_cast<int>((static_cast<long long>(sp + 1) * kblocks) / splits);
        const int m0 = mt * BM, n0 = nt * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * kBStage;
          const int k0 = kb * BK;
          if (PAIR) {
            // the leader's full barrier counts the bytes of both CTAs of the pair
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStage);
            if (AMAJ == 0) {
              tma_load_3d_pair(&tmA, lbar, sa, k0, m0, bz);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_3d_pair(&tmA, lbar, sa + j * (BK * 128), m0 + 64 * j, k0, bz);
            }
            if (BMAJ == 0) {
              tma_load_3d_pair(&tmB, lbar, sb, k0, n0 + crank * (BN / 2), bz);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_3d_pair(&tmB, lbar, sb + j * (BK * 128), n0 + 64 * (crank * (BN / 128) + j), k0, bz);
            }
            if (++stage == nstages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (AMAJ == 0) {
            tma_load_3d(&tmA, &full_bar[stage], sa, k0, m0, bz);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_3d(&tmA, &full_bar[stage], sa + j * (BK * 128), m0 + 64 * j, k0, bz);
          }
          if (BMAJ == 0) {
            tma_load_3d(&tmB, &full_bar[stage], sb, k0, n0, bz);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(&tmB, &full_bar[stage], sb + j * (BK * 128), n0 + 64 * j, k0, bz);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && crank == 0) {            // (pair: only the leader CTA issues MMAs)
      constexpr uint32_t idesc = make_idesc<BN, AMAJ, BMAJ, PAIR ? 2 * BM : BM>();
      constexpr uint32_t a_lbo = (AMAJ == 0) ? 0u : BK * 128u;
      constexpr uint32_t b_lbo = (BMAJ == 0) ? 0u : BK * 128u;
      constexpr uint32_t a_kstep = (AMAJ == 0) ? UMMA_K * 2u : UMMA_K * 128u;   // bytes per K=16 step
      constexpr uint32_t b_kstep = (BMAJ == 0) ? UMMA_K * 2u : UMMA_K * 128u;
      constexpr uint32_t idesc_ones = make_idesc_ones(AMAJ, PAIR ? 2 * BM : BM);
      // SWIZZLE_NONE descriptor: core matrices 128 B apart along K (LBO) and 256 B apart along N (SBO)
      const uint64_t ones_desc = (make_smem_desc(smem_u32(ones_tile), 128u, 256u) & ~(7ull << 61));
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const bool rs_tile = RS && (tile % n_tiles == 0);     // one column tile per (row tile, split) sums the rows
        const int rest = tile / n_tiles;
        const int sp = (rest / m_units) % splits;
        const int kb0 = static_cast<int>((static_cast<long long>(sp) * kblocks) / splits);
        const int kb1 = static_cast<int>((static_cast<long long>(sp + 1) * kblocks) / splits);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t sb = smem_u32(smem_b + stage * kBStage);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            if (PAIR) tc_mma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else tc_mma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (rs_tile) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
              if (PAIR) tc_mma_bf16_pair(tmem_base + BN, adesc, ones_desc, idesc_ones, (kb > kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16(tmem_base + BN, adesc, ones_desc, idesc_ones, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          // frees the smem stage once these MMAs retire (pair: in both CTAs)
          if (PAIR) tc_commit_pair(&empty_bar[stage]);
          else tc_commit(&empty_bar[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_pair(&tfull_bar[acc]);     // accumulator complete -> both CTAs' epilogues
        else tc_commit(&tfull_bar[acc]);             // accumulator complete -> epilogue
        if (RS) acc_phase ^= 1;              // single accumulator buffer (columns [BN, BN+16) hold the row sums)
        else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (8 warps)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access (warp id mod 4)
    const int ew = warp - 2;                 // 0..7
    EpiWarp w;
    w.base = smem_u32(smem_stg) + static_cast<uint32_t>(ew * stg_stride);
    w.src_full = src_bar + ew;
    w.sphase = 0;
    w.half = ew >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = unit0; tile < total_tiles; tile += unit_step) {
      const int nt = tile % n_tiles;
      const int rest = tile / n_tiles;
      const int mt = PAIR ? 2 * (rest % m_units) + crank : rest % m_units;
      const int rest2 = rest / m_units;
      const int sp = rest2 % splits;
      const int bz = rest2 / splits;
      const int mw = mt * BM + q * 32;       // first row of this warp
      const int n0 = nt * BN;
      const bool lead = (sp == 0);
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      uint64_t* tf = &tfull_bar[acc];
      if (ep.c_f32) {
        if (ep.src == SRC_NONE) epilogue_tile<BN, true, SRC_NONE>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
        else if (ep.src == SRC_RESIDUAL) epilogue_tile<BN, true, SRC_RESIDUAL>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
        else epilogue_tile<BN, true, SRC_GATE>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
      } else {
        if (ep.src == SRC_NONE) epilogue_tile<BN, false, SRC_NONE>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
        else if (ep.src == SRC_RESIDUAL) epilogue_tile<BN, false, SRC_RESIDUAL>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
        else epilogue_tile<BN, false, SRC_GATE>(ep, &tmC, &tmS, w, taddr0, lane, n0, mw, bz, lead, tf, acc_phase);
      }
      if (RS && nt == 0 && w.half == 0) {
        // column 0 of the ones-product: this lane's row sum over the k-range of this split
        const float rsum = __uint_as_float(tmem_ld1(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN));
        tmem_ld_wait();
        const int row = mw + lane;
        if (row < M) atomicAdd(ep.rowsum + bz * ep.rs_stride + row, rsum);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));     // the leader's MMA thread waits for both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (RS) acc_phase ^= 1;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    // staged tiles must outlive the TMA stores that read them
    if (lane == 0) tma_wait_group_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();       // neither CTA frees tensor memory or leaves while the pair's MMAs / arrivals may be in flight
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

}  // namespace

// 2-D matrix stored row-major as [rows, cols] with leading dimension ld (elements), optionally batched;
// the box is [box_rows, box_cols] with box_cols * esize <= 128 bytes (one swizzle atom).
int make_tmap(CUtensorMap* tm, const void* base, int f32, long long rows, long long cols, long long ld, int box_rows,
              int box_cols, int batch, long long batch_stride) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable"); return SER_ERR_CUDA; }
  if (batch <= 1) { batch = 1; batch_stride = rows * ld; }
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * es, static_cast<cuuint64_t>(batch_stride) * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d base=%p f32=%d "
             "batch=%d stride=%lld", (int)r, rows, cols, ld, box_rows, box_cols, base, f32, batch, batch_stride);
    set_last_error(__FILE__, __LINE__, msg);
    return SER_ERR_CUDA;
  }
  return SER_OK;
}

namespace {

template <int BN, int AMAJ, int BMAJ, bool RS = false, bool PAIR = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS,
           const TcEpilogue& ep, int M, int N, int K, int splits, int batch, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, AMAJ, BMAJ, RS, PAIR>;
  if (!configured) {
    SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    configured = true;
  }
  const bool with_src = ep.src != SRC_NONE;
  static const int stage_cap = getenv("SER_GEMM_STAGES") ? atoi(getenv("SER_GEMM_STAGES")) : kMaxStages;   // A/B switch
  int nstages = Cfg::stages(with_src, PAIR);
  if (stage_cap >= 2 && nstages > stage_cap) nstages = stage_cap;
  const int smem_bytes = nstages * Cfg::stage_bytes(PAIR) + 1024 + Cfg::stg_bytes(with_src) + kBarBytes;
  const int m_tiles = ceil_div(M, BM), n_tiles = N / BN;
  if (!PAIR) {
    const long long total = static_cast<long long>(m_tiles) * n_tiles * splits * batch;
    const int grid = static_cast<int>(total < gemm_grid_sms() ? total : gemm_grid_sms());
    SER_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmC, tmS, ep, M, N, K, splits,
                              batch, nstages));
  } else {
    const long long units = static_cast<long long>(m_tiles / 2) * n_tiles * splits * batch;
    const long long max_clusters = gemm_grid_sms() / 2;
    const int clusters = static_cast<int>(units < max_clusters ? units : max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmS, ep, M, N, K, splits, batch, nstages));
  }
  SER_LAUNCH_CHECK();
  return SER_OK;
}

template <int BN>
int dispatch_major(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                   const CUtensorMap& tmS, const TcEpilogue& ep, int splits, bool mc, cudaStream_t stream) {
  const int nb = a.batch > 1 ? a.batch : 1;
  if (mc) {
    if (!a.a_trans && !a.b_trans) return launch<BN, 0, 0, false, true>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
    if (!a.a_trans && a.b_trans) return launch<BN, 0, 1, false, true>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
    if (a.a_trans && a.b_trans && ep.rowsum != nullptr)
      return launch<BN, 1, 1, true, true>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
    if (a.a_trans && a.b_trans) return launch<BN, 1, 1, false, true>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
  }
  if (!a.a_trans && !a.b_trans) return launch<BN, 0, 0>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
  if (!a.a_trans && a.b_trans) return launch<BN, 0, 1>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
  if (a.a_trans && a.b_trans && ep.rowsum != nullptr)
    return launch<BN, 1, 1, true>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
  if (a.a_trans && a.b_trans) return launch<BN, 1, 1>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
  return launch<BN, 1, 0>(tmA, tmB, tmC, tmS, ep, a.M, a.N, a.K, splits, nb, stream);
}

}  // namespace

bool gemm_tc_rowsum_ok(const GemmArgs& a) {
  static const bool disabled = (getenv("SER_NO_GEMM_ROWSUM") != nullptr);      // A/B switch: separate colsum launches
  return !disabled && a.dtype == DT_BF16 && a.a_trans && a.b_trans && a.R == nullptr && a.gate_mode == GATE_NONE &&
         !a.accumulate;
}

int gemm_tc_bf16(const GemmArgs& a, cudaStream_t stream) {
  SER_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm_tc: empty problem");
  SER_REQUIRE(a.N % 128 == 0, "gemm_tc: N must be a multiple of 128");
  SER_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm_tc: leading dimensions must be multiples of 8 elements");
  SER_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
              "gemm_tc: operands must be 16-byte aligned");
  SER_REQUIRE(a.ldc % 8 == 0, "gemm_tc: ldc must be a multiple of 8");
  SER_REQUIRE(a.gate_mode == GATE_NONE || (a.G != nullptr && !a.g_f32 && a.ldg % 8 == 0),
              "gemm_tc: the gate operand must be bf16 with an 8-element aligned leading dimension");
  SER_REQUIRE(a.R == nullptr || a.ldr % 8 == 0, "gemm_tc: ldr must be a multiple of 8");
  int BN = (a.N % 256 == 0) ? 256 : 128;
  static const int bn_env = getenv("SER_GEMM_BN") ? atoi(getenv("SER_GEMM_BN")) : 0;      // A/B switch: force 128-wide tiles
  if (bn_env == 128) BN = 128;
  // (measured with it: 768 <- 256 + residual, 64000 rows, cold operands: 56.7 us with 128-wide tiles vs 51.4 us with 256)
  // small problems: narrower tiles put twice as many SMs to work and halve each CTA's serial epilogue
  // (long contractions are split along K instead and keep the wider, shared-memory-friendlier tile)
  if (BN == 256 && ceil_div(a.K, BK) < 64 &&
      2LL * ceil_div(a.M, BM) * (a.N / 256) * (a.batch > 1 ? a.batch : 1) <= device_sm_count()) BN = 128;

  const int m_tiles = ceil_div(a.M, BM), n_tiles = a.N / BN, kblocks = ceil_div(a.K, BK);
  int splits = a.splits;
  const bool linear = (a.act == ACT_NONE && a.gate_mode == GATE_NONE && a.c_f32);
  if (splits <= 0) {
    splits = 1;
    const int tiles = m_tiles * n_tiles;
    const int sms = device_sm_count();
    const int tiles_all = tiles * (a.batch > 1 ? a.batch : 1);
    // split-K pays only when the contraction is long (token-dimension wgrads); tiny problems stay unsplit:
    // the fp32-atomic epilogue costs more than the serial K loop it saves
    if (linear && tiles_all * 2 <= sms && kblocks >= 64) {
      splits = sms / tiles_all;
      const int max_by_k = kblocks / 8;        // keep >= 8 k-blocks per split
      if (splits > max_by_k) splits = max_by_k;
      if (splits < 1) splits = 1;
    }
  }
  if (!linear) splits = 1;
  if (splits > kblocks) splits = kblocks;

  SER_REQUIRE(a.batch <= 1 || (a.strideA % 8 == 0 && a.strideB % 8 == 0 && a.strideC % 8 == 0),
              "gemm_tc: batch strides must be multiples of 8");
  SER_REQUIRE((reinterpret_cast<uintptr_t>(a.C) & 15) == 0, "gemm_tc: output must be 16-byte aligned");
  // cta_group::2 pairs (256 x BN tiles): row-tile pairs of problems large enough to be bound by the main loop
  static const int mc_env = getenv("SER_GEMM_PAIR") ? atoi(getenv("SER_GEMM_PAIR")) : -1;     // A/B switch: 0 off, 1 force
  bool mc = (m_tiles % 2 == 0) && !(a.a_trans && !a.b_trans);
  if (mc_env == 0) mc = false;
  else if (mc_env != 1)
    // BN = 128 tiles stay single (the pooling scorer 64000x128x768: 29.5 us paired vs 27.5 us single); grids that
    // cannot fill the SMs stay single (pairing halves the number of independent CTAs); short contractions (K = 256)
    // are epilogue / HBM bound (42.7 vs 33.6 us paired)
    mc = mc && BN == 256 &&
         static_cast<long long>(m_tiles) * n_tiles * splits * (a.batch > 1 ? a.batch : 1) >= device_sm_count() - 20 &&
         kblocks / splits >= 8;
  CUtensorMap tmA, tmB, tmC, tmS;
  if (!a.a_trans) SER_TRY(make_tmap(&tmA, a.A, 0, a.M, a.K, a.lda, BM, BK, a.batch, a.strideA));
  else            SER_TRY(make_tmap(&tmA, a.A, 0, a.K, a.M, a.lda, BK, 64, a.batch, a.strideA));
  if (!a.b_trans) SER_TRY(make_tmap(&tmB, a.B, 0, a.N, a.K, a.ldb, mc ? BN / 2 : BN, BK, a.batch, a.strideB));
  else            SER_TRY(make_tmap(&tmB, a.B, 0, a.K, a.N, a.ldb, BK, 64, a.batch, a.strideB));

  // epilogue operands travel as 32-row x 128-byte TMA boxes: 64 bf16 or 32 fp32 columns
  const int cw = a.c_f32 ? 32 : 64;
  SER_TRY(make_tmap(&tmC, a.C, a.c_f32, a.M, a.N, a.ldc, 32, cw, a.batch, a.strideC));
  TcEpilogue ep;
  ep.c_f32 = a.c_f32;
  ep.bias = a.bias;
  ep.gate_mode = a.gate_mode;
  ep.act = a.act;
  ep.alpha = a.alpha;
  ep.src = SRC_NONE;
  ep.rowsum = a.rowsum;
  ep.rs_stride = a.strideRS;
  ep.bias_stride = a.strideBias;
  SER_REQUIRE(a.strideBias % 4 == 0, "gemm_tc: the bias batch stride must be a multiple of 4 elements");
  if (a.rowsum != nullptr) {
    SER_REQUIRE(gemm_tc_rowsum_ok(a), "gemm_tc: rowsum rides on plain dW-type GEMMs only (both operands MN-major)");
    const int nb = a.batch > 1 ? a.batch : 1;
    if (a.out_zeroed) { /* caller's buffer is zero-filled */ }
    else if (nb == 1 || a.strideRS == a.M) SER_CUDA_CHECK(cudaMemsetAsync(a.rowsum, 0, sizeof(float) * a.M * nb, stream));
    else SER_CUDA_CHECK(cudaMemset2DAsync(a.rowsum, a.strideRS * sizeof(float), 0, a.M * sizeof(float), nb, stream));
  }
  const void* R = a.R;
  int accumulate = a.accumulate;
  if (splits > 1 && R != nullptr && R == a.C) {
    // in-place residual with split-K: C already holds R, so every split simply accumulates into it
    R = nullptr;
    accumulate = 1;
  }
  if (R != nullptr && a.gate_mode != GATE_NONE) {
    set_last_error(__FILE__, __LINE__, "gemm_tc: residual and gate operands cannot be combined");
    return SER_ERR_UNSUPPORTED;
  }
  if (R != nullptr) {
    SER_REQUIRE(a.r_f32 == a.c_f32, "gemm_tc: the residual must have the output's element type");
    SER_REQUIRE((reinterpret_cast<uintptr_t>(R) & 15) == 0 && (a.batch <= 1 || a.strideR % 8 == 0),
                "gemm_tc: residual must be 16-byte aligned");
    SER_TRY(make_tmap(&tmS, R, a.r_f32, a.M, a.N, a.ldr, 32, cw, a.batch, a.strideR));
    ep.src = SRC_RESIDUAL;
  } else if (a.gate_mode != GATE_NONE) {
    SER_REQUIRE(!a.c_f32, "gemm_tc: a gated GEMM writes bf16 (the gate operand's element type)");
    SER_REQUIRE((reinterpret_cast<uintptr_t>(a.G) & 15) == 0 && (a.batch <= 1 || a.strideG % 8 == 0),
                "gemm_tc: gate operand must be 16-byte aligned");
    SER_TRY(make_tmap(&tmS, a.G, 0, a.M, a.N, a.ldg, 32, cw, a.batch, a.strideG));
    ep.src = SRC_GATE;
  } else {
    tmS = tmC;
  }
  ep.atomic = (splits > 1 || accumulate) ? 1 : 0;
  if (ep.atomic) {
    SER_REQUIRE(a.c_f32, "gemm_tc: accumulate / split-K needs an fp32 output");
    if (!accumulate && !a.out_zeroed) {
      for (int b = 0; b < (a.batch > 1 ? a.batch : 1); ++b)
        SER_CUDA_CHECK(cudaMemset2DAsync(reinterpret_cast<float*>(a.C) + b * a.strideC, a.ldc * sizeof(float), 0,
                                         a.N * sizeof(float), a.M, stream));
    }
  }
  const double gflops = 2.0 * a.M * a.N * a.K * (a.batch > 1 ? a.batch : 1);
  const double gesz = (a.dtype == DT_F32) ? 4.0 : 2.0;
  const double gbytes = (static_cast<double>(a.M) * a.K + static_cast<double>(a.N) * a.K) * gesz +
                        static_cast<double>(a.M) * a.N * ((a.c_f32 ? 4.0 : 2.0) + (a.R ? (a.r_f32 ? 4.0 : 2.0) : 0.0) +
                                                         (a.G ? (a.g_f32 ? 4.0 : 2.0) : 0.0));
  char pname[96];
  if (prof_enabled())
    snprintf(pname, sizeof(pname), "%s:%dx%dx%d:s%d", a.a_trans ? "gemm_tc_wgrad" : (a.b_trans ? "gemm_tc_dgrad" : "gemm_tc_fwd"),
             a.M, a.N, a.K, splits);
  ProfScope prof(pname, gflops, gbytes, stream);
  if (BN == 256) return dispatch_major<256>(a, tmA, tmB, tmC, tmS, ep, splits, mc, stream);
  return dispatch_major<128>(a, tmA, tmB, tmC, tmS, ep, splits, mc, stream);
}

}  // namespace ser

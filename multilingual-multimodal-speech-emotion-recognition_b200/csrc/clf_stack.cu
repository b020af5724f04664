// The 35-block residual stack of AdvancedOpenMaxClassifier as ONE kernel per direction (bf16 tier).
//
// Reference: src/models/classifier.py:207-212 (outer LayerNorm -> block -> residual from the outer-LN output)
// and :79-89 (block = LayerNorm -> Linear -> ReLU -> Linear).  Per block i:
//     y = LN_o(h_i);  n = LN_i(y);  u = relu(W1 n + b1);  h_{i+1} = y + W2 u + b2
// 70 strictly sequential [B,512]x[512,512] GEMMs separated by row-wise LayerNorms: launched layer by layer this
// is pure launch / fill latency (0.13 GFLOP per launch).  Here a thread-block CLUSTER of 8 CTAs owns 128 batch
// rows for the whole stack:
//   * CTA j of the cluster owns output columns [64j, 64j+64) of every GEMM: its 64x512 weight slice streams through
//     an 8-stage TMA ring (always one GEMM ahead), the accumulator is a 128x64 fp32 tile in TMEM,
//     tcgen05.mma M=128 N=64 K=16 issued by one thread.
//   * every epilogue thread owns one row x 64 columns IN REGISTERS across the block (the residual y never leaves
//     the register file); LayerNorm statistics are combined across the 8 column slices through distributed
//     shared memory (st.shared::cluster + barrier.cluster), Chan-style (mean, M2) so the variance is two-pass exact.
//   * the full-width GEMM operand (n, u, and in backward dh, da) is exchanged through the SAME global buffers the
//     backward pass needs anyway (saved activations / weight-gradient operands): each CTA stores its bf16 slice,
//     a cluster barrier orders it, and every CTA pulls the 128x512 operand back as eight 128B-swizzled TMA boxes.
// Backward mirrors it with the weights read MN-major in place (dX = dY W), recomputing x-hat from the saved h and
// statistics, and reduces the four LayerNorm parameter gradients over rows with a register transpose-reduce.
// The batched weight-gradient GEMMs stay outside (head_modules.cu).
#include "kernels.cuh"
#include "prof.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace ser {

namespace {

constexpr int CS = 8;                 // CTAs per cluster = column slices
constexpr int PD = 512;               // base_dim of the stack
constexpr int NS = PD / CS;           // 64 output columns per CTA
constexpr int RM = 128;               // maximum rows per cluster (UMMA M = 128); small batches run M = 64 clusters
constexpr int KBLK = PD / 64;         // 8 k-blocks of 64
constexpr int A_BYTES = RM * 128;     // one k-block of the A operand at M = 128: 128 rows x 128 B (shared-memory carve-up)
constexpr int W_BYTES = NS * 128;     // one weight stage: 64 x 64 bf16
constexpr int kThreads = 192;
constexpr float kEps = 1e-5f;

constexpr int OFF_A = 0;
constexpr int OFF_W = OFF_A + KBLK * A_BYTES;                 // 131072
constexpr int OFF_STATS = OFF_W + KBLK * W_BYTES;             // 196608: float2 [2][CS][RM]
constexpr int OFF_COLACC = OFF_STATS + 2 * CS * RM * 8;       // 212992: float [4][NS]
constexpr int OFF_PAR = OFF_COLACC + 4 * NS * 4;              // 214016: float [2][6][NS] per-layer parameter slices
constexpr int OFF_BARS = OFF_PAR + 2 * 6 * NS * 4;            // 217088
constexpr int kSmemBytes = OFF_BARS + 256 + 1024;

// ---------------------------------------------------------------------------------------------------------
// PTX
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// 1-D bulk copies: shared -> global, and global -> the same shared offset of every CTA in `mask`
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_but_one() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// shared-memory matrix descriptor (sm_100), SWIZZLE_128B; see gemm_tc.cu
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
template <int BMAJ> __device__ __forceinline__ uint32_t make_idesc(int rm) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BMAJ) << 16) |
         (static_cast<uint32_t>(NS >> 3) << 17) | (static_cast<uint32_t>(rm >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------
// shared pieces of the two kernels
// ---------------------------------------------------------------------------------------------------------
struct Ctx {
  uint32_t sA, sW, sStats, sColacc, sPar;   // shared-window addresses
  uint32_t a_full, w_full, w_empty, acc_full;   // first barrier of each family (8 bytes apart)
  uint32_t tmem;
  uint32_t rank;                         // column slice of this CTA
  int row0;                              // first row of the cluster
  int rm;                                // rows per cluster: 128 (UMMA M = 128) or 64 (UMMA M = 64)
  uint32_t a_bytes;                      // rm * 128: one k-block of the A operand / one exchange block
};

__device__ __forceinline__ uint32_t bar_at(uint32_t base, int i) { return base + 8u * static_cast<uint32_t>(i); }

// MMA issuer: one GEMM = 8 k-blocks x 4 tcgen05.mma (M = 128 or 64, N = 64, K = 16) into the 64-column accumulator.
// The issuing thread is the critical path of a GEMM phase (an N = 64 MMA is accepted every ~46 cycles, measured), so
// the loop carries nothing but the MMAs: one operand barrier, two weight barriers (k-blocks 0-3 / 4-7), ONE commit.
template <int BMAJ>
__device__ __forceinline__ void issue_gemm(const Ctx& c, uint32_t parity, long long* tl = nullptr) {
  const uint32_t idesc = make_idesc<BMAJ>(c.rm);
  constexpr uint32_t b_kstep = (BMAJ == 0) ? 32u : 16u * 128u;     // bytes per K = 16 step inside a weight stage
  constexpr uint32_t b_lbo = (BMAJ == 0) ? 0u : 64u * 128u;
  // descriptors differ only in their start-address field (bits 0-13, 16-byte units): build the two bases once and
  // step them with one add per MMA -- the issue loop is a single thread's dependent instruction stream
  const uint64_t a_base = make_smem_desc(c.sA, 0u, 1024u);
  const uint64_t b_base = make_smem_desc(c.sW, b_lbo, 1024u);
  const uint64_t a_kb = static_cast<uint64_t>(c.a_bytes >> 4);     // descriptor units per k-block of A
  mbar_wait(bar_at(c.a_full, 0), parity);                          // the whole operand arrives as one transfer
  if (tl != nullptr) tl[1] = clock64();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    mbar_wait(bar_at(c.w_full, half), parity);
    tc_fence_after();
    if (tl != nullptr && half == 1) tl[2] = clock64();
#pragma unroll
    for (int kq = 0; kq < KBLK / 2; ++kq) {
      const int kb = half * (KBLK / 2) + kq;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = a_base + static_cast<uint64_t>(kb) * a_kb + static_cast<uint64_t>((k * 32u) >> 4);
        const uint64_t bdesc = b_base + static_cast<uint64_t>((kb * W_BYTES + k * b_kstep) >> 4);
        tc_mma_bf16(c.tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
    }
  }
  tc_commit(c.acc_full);                   // accumulator complete; also tells the producer the weight ring is free
  if (tl != nullptr) tl[3] = clock64();
}

// Exchange buffer (global, L2 resident): [cluster][slot = GEMM parity][k-block = producing CTA][128 rows x 128 B], each
// 16 KB block already in the 128B-swizzled K-major image the tensor core reads.  Every CTA arms all eight k-block
// barriers, then fetches ONE block -- the slice it published itself -- as a single contiguous bulk copy multicast
// to the whole cluster (a contiguous 16 KB request moves several times faster than 128 strided 128-byte rows).
__device__ __forceinline__ char* xchg_block(const Ctx& c, char* xchg, int g) {
  return xchg + ((static_cast<size_t>(blockIdx.x / CS) * 2 + (g & 1)) * CS + c.rank) * c.a_bytes;
}
// Measured on B200: incoming bulk transfers are delivered to a CTA one after the other, ~600 cycles each whether they
// carry 8 or 16 KB, so eight per-slice multicasts take ~4800 cycles.  The eight blocks of a slot are contiguous in the
// exchange buffer, so CTA 0 fetches the WHOLE operand as one bulk copy and multicasts it to the cluster.
__device__ __forceinline__ void load_a(const Ctx& c, char* xchg, int g) {
  fence_proxy_async_all();
  mbar_expect_tx(bar_at(c.a_full, 0), KBLK * c.a_bytes);
  if (c.rank == 0) {
    char* slot = xchg + (static_cast<size_t>(blockIdx.x / CS) * 2 + (g & 1)) * CS * c.a_bytes;
    bulk_load_mc(c.sA, slot, KBLK * c.a_bytes, bar_at(c.a_full, 0), static_cast<uint16_t>((1u << CS) - 1u));
  }
}
// producer: weight slice of one GEMM.  FWD: rows = output features of this CTA, columns = k; BWD (MN-major):
// rows = k (output features of the forward Linear), columns = this CTA's input features.
// The 64 KB ring holds exactly one GEMM's slice, so the next GEMM's weights are requested the moment the running
// GEMM's accumulator is complete (`acc_parity` >= 0: wait for that commit) and land during the epilogue / exchange
// phase that follows, when nothing else is filling shared memory.
template <int BMAJ>
__device__ __forceinline__ void load_w(const Ctx& c, const CUtensorMap* tm, int layer, int acc_parity) {
  if (acc_parity >= 0) mbar_wait(c.acc_full, static_cast<uint32_t>(acc_parity));
  for (int half = 0; half < 2; ++half) {
    mbar_expect_tx(bar_at(c.w_full, half), (KBLK / 2) * W_BYTES);
    for (int kq = 0; kq < KBLK / 2; ++kq) {
      const int kb = half * (KBLK / 2) + kq;
      if (BMAJ == 0) tma_load_3d(tm, bar_at(c.w_full, half), c.sW + kb * W_BYTES, kb * 64, static_cast<int>(c.rank) * NS, layer);
      else           tma_load_3d(tm, bar_at(c.w_full, half), c.sW + kb * W_BYTES, static_cast<int>(c.rank) * NS, kb * 64, layer);
    }
  }
}

// all-reduce of two per-row partial values over the 8 column slices of the cluster (one cluster barrier)
__device__ __forceinline__ void exchange2(const Ctx& c, int buf, int rl, bool active, float a, float b, float (&oa)[CS],
                                          float (&ob)[CS]) {
  const uint32_t slot = c.sStats + static_cast<uint32_t>(((buf * CS + static_cast<int>(c.rank)) * RM + rl) * 8);
  if (active) {
#pragma unroll
    for (int t = 0; t < CS; ++t) st_cluster_f2(map_to_cta(slot, t), a, b);
  }
  cluster_sync_all();
#pragma unroll
  for (int s = 0; s < CS; ++s) {
    const float2 v = ld_shared_f2(c.sStats + static_cast<uint32_t>(((buf * CS + s) * RM + rl) * 8));
    oa[s] = v.x; ob[s] = v.y;
  }
}

// LayerNorm statistics of a 512-wide row held as 8 slices of 64: local two-pass (mean, M2), Chan combination
__device__ __forceinline__ void row_stats(const Ctx& c, int buf, int rl, bool active, const float (&v)[NS], float& mean,
                                          float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NS; ++k) s += v[k];
  const float m = s * (1.f / NS);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NS; ++k) { const float d = v[k] - m; q = fmaf(d, d, q); }
  float ms[CS], qs[CS];
  exchange2(c, buf, rl, active, m, q, ms, qs);
  float mu = 0.f;
#pragma unroll
  for (int t = 0; t < CS; ++t) mu += ms[t];
  mu *= (1.f / CS);
  float m2 = 0.f;
#pragma unroll
  for (int t = 0; t < CS; ++t) { const float d = ms[t] - mu; m2 += qs[t] + static_cast<float>(NS) * d * d; }
  mean = mu;
  rstd = rsqrtf(m2 * (1.f / PD) + kEps);
}

__device__ __forceinline__ void load_vec64(const float* __restrict__ p, float (&v)[NS]) {
#pragma unroll
  for (int k = 0; k < NS / 4; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + k);
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
}
__device__ __forceinline__ void store_f32_64(float* p, const float (&v)[NS]) {
#pragma unroll
  for (int k = 0; k < NS / 4; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
// One row-slice (64 bf16 = 128 B per thread) of the next GEMM operand: written ONCE into this CTA's own k-block of
// the A buffer (swizzled), from where one thread sends it (a) as a contiguous 16 KB bulk store to the exchange buffer
// and (b) as a TMA tensor store to the saved-activation buffer [L,B,512] the backward pass reads (rows >= B are
// clipped by the tensor map).  Both stores are complete before the caller enters the cluster barrier.
__device__ __forceinline__ void publish_slice(const Ctx& c, const float (&v)[NS], int rl, bool active, int et, char* xblock,
                                              const CUtensorMap* tm_save, int col0, int layer) {
  const uint32_t tile = c.sA + c.rank * c.a_bytes;
  const uint32_t base = tile + static_cast<uint32_t>(rl) * 128u;
  const uint32_t swz = static_cast<uint32_t>(rl & 7);
  // earlier bulk stores may still be READING the staging tiles (the saved-activation store of the previous slice,
  // the fp32 stream stores): drain their reads before the tile is rewritten
  if (et == 0) bulk_wait_read_all();
  named_bar_sync(3, 128);
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(base + ((static_cast<uint32_t>(j) ^ swz) << 4), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
             pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
  }
  fence_async_smem();
  named_bar_sync(3, 128);
  if (et == 0) {
    // only the exchange block is on the critical path: its own bulk group, waited for completion; the
    // saved-activation store rides in a second group that is merely drained before the next rewrite
    bulk_store(xblock, tile, c.a_bytes);
    bulk_commit();
    tma_store_3d(tm_save, tile, col0, c.row0, layer);
    bulk_commit();
    bulk_wait_but_one();
  }
}
// fp32 row-slice (64 floats = 2 x 128 B per thread) -> two swizzled tiles in the idle k-blocks next to our own ->
// two TMA tensor stores (fire and forget; drained by the next publish_slice before any peer can overwrite the tiles)
__device__ __forceinline__ void store_slice_f32(const Ctx& c, const float (&v)[NS], int rl, bool active, int et,
                                                const CUtensorMap* tm, int col0, int layer) {
  const uint32_t swz = static_cast<uint32_t>(rl & 7);
  if (active) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t base = c.sA + ((c.rank + 1u + t) & 7u) * c.a_bytes + static_cast<uint32_t>(rl) * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(base + ((static_cast<uint32_t>(j) ^ swz) << 4), __float_as_uint(v[32 * t + 4 * j]),
               __float_as_uint(v[32 * t + 4 * j + 1]), __float_as_uint(v[32 * t + 4 * j + 2]), __float_as_uint(v[32 * t + 4 * j + 3]));
    }
  }
  fence_async_smem();
  named_bar_sync(3, 128);
  if (et == 0) {
    tma_store_3d(tm, c.sA + ((c.rank + 1u) & 7u) * c.a_bytes, col0, c.row0, layer);
    tma_store_3d(tm, c.sA + ((c.rank + 2u) & 7u) * c.a_bytes, col0 + 32, c.row0, layer);
    bulk_commit();
  }
}

// Per-layer parameter slices (64 floats each) are copied to shared memory one block AHEAD with cp.async, so the
// row owners never wait on an L2 round trip between two cluster barriers; readers get them as broadcast LDS.
__device__ __forceinline__ void lds_vec(uint32_t addr, float* v, int n) {
  for (int k = 0; k < n / 4; ++k)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * k]), "=f"(v[4 * k + 1]), "=f"(v[4 * k + 2]), "=f"(v[4 * k + 3]) : "r"(addr + 16u * k) : "memory");
}
__device__ __forceinline__ uint32_t par_addr(const Ctx& c, int buf, int vec) { return c.sPar + static_cast<uint32_t>(((buf * 6 + vec) * NS) * 4); }

struct StackParams {
  int B, L;
  int rm;                                          // rows per cluster (128 or 64)
  const float* pv[6]; long long ps[6]; int npv;   // per-layer parameter vectors staged in shared memory (layer-0 pointer, stride)
  long long* dbg;                                  // optional clock64 timeline of one block (SER_CLF_TIMELINE)
  char* xchg;                                      // exchange buffer, ceil(B/128) * 2 * 128 KB
  // per-layer parameter vectors: pointer of layer 0 + element stride between layers
  const float* b1; const float* b2; const float* lni_g; const float* lni_b; long long s_blk;   // block params share one stride
  const float* lno_g; const float* lno_b; long long s_lno;
  float* h;                 // [(L+1), B, 512] fp32 residual stream (h[0] is the input)
  __nv_bfloat16* n; __nv_bfloat16* r;      // [L, B, 512] exchange + saved operands
  float* stats_o; float* stats_i;          // [L, B, 2]
  // backward
  const float* dh_in; float* dh_out;       // [B, 512] fp32: dL/dh_L in, dL/dh_0 out
  __nv_bfloat16* dhn; __nv_bfloat16* dr;   // [L, B, 512] exchange + weight-gradient operands
  float* dlni_g; float* dlni_b; float* dlno_g; float* dlno_b;    // layer 0 pointers (strides as the parameters)
  DropSpec drop;                           // block[3] / block[5] dropout (site = DS_CLF_BLOCK0 + 2 layer + {0, 1}); off by default
};

// keep decisions of the thread's 64-column slice of row `row` at dropout site `site` ([B, 512] site: 256 pairs per
// row), bit k = keep column col0 + k.  Data independent: computed while the thread waits for the tensor cores.
__device__ __forceinline__ unsigned keep_bits16(const DropKey& key, unsigned pair0, unsigned thr) {
  unsigned m = 0u;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const unsigned b = drop_bits(key, pair0 + k);
    m |= (((b & 0xffffu) >= thr) ? 1u : 0u) << (2 * k);
    m |= (((b >> 16) >= thr) ? 1u : 0u) << (2 * k + 1);
  }
  return m;
}
// `split`: M = 64 clusters leave lanes 16..31 of every row-owner warp without a row; they hash the upper 32 columns of
// their partner lane's row (same `row` value by construction) and the halves are swapped with one shuffle.
__device__ __forceinline__ unsigned long long drop_slice_bits(const DropSpec& d, unsigned long long seed, unsigned site,
                                                             int row, int col0, bool split, int lane) {
  const DropKey key = drop_key_of(seed, site);      // (the seed is read from global memory once per kernel)
  const unsigned pair0 = static_cast<unsigned>(row) * (PD / 2) + static_cast<unsigned>(col0 >> 1);
  unsigned lo, hi;
  if (split) {
    const unsigned half = static_cast<unsigned>(lane) >> 4;
    const unsigned mine = keep_bits16(key, pair0 + half * (NS / 4), d.thr);
    const unsigned other = __shfl_xor_sync(0xffffffffu, mine, 16);
    lo = half ? other : mine;
    hi = half ? mine : other;
  } else {
    lo = keep_bits16(key, pair0, d.thr);
    hi = keep_bits16(key, pair0 + NS / 4, d.thr);
  }
  return (static_cast<unsigned long long>(hi) << 32) | lo;
}
__device__ __forceinline__ void drop_slice_apply(unsigned long long bits, float scale, float (&v)[NS]) {
  const unsigned lo = static_cast<unsigned>(bits), hi = static_cast<unsigned>(bits >> 32);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    v[k] = ((lo >> k) & 1u) ? v[k] * scale : 0.f;
    v[32 + k] = ((hi >> k) & 1u) ? v[32 + k] * scale : 0.f;
  }
}

__device__ __forceinline__ void par_prefetch(const Ctx& c, const StackParams& p, int layer, int buf, int et, int col0, bool on) {
  if (on && et < p.npv * 16) {
    const int v = et >> 4, ch = et & 15;
    const float* src = p.pv[v] + layer * p.ps[v] + col0 + ch * 4;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(par_addr(c, buf, v) + 16u * ch), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// start of a block: recycle the older buffer for the layer after this one, then make this layer's slices visible
__device__ __forceinline__ void par_advance(const Ctx& c, const StackParams& p, int next_layer, bool have_next, int next_buf, int et, int col0) {
  named_bar_sync(2, 128);
  par_prefetch(c, p, next_layer, next_buf, et, col0, have_next);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  named_bar_sync(2, 128);
}
#define SER_TL(k) do { if (tl) p.dbg[k] = clock64(); } while (0)

__device__ __forceinline__ void setup(Ctx& c, uint8_t* smem, int warp, int lane, int rm) {
  c.rm = rm;
  c.a_bytes = static_cast<uint32_t>(rm) * 128u;
  c.sA = smem_u32(smem + OFF_A);
  c.sW = smem_u32(smem + OFF_W);
  c.sStats = smem_u32(smem + OFF_STATS);
  c.sColacc = smem_u32(smem + OFF_COLACC);
  c.sPar = smem_u32(smem + OFF_PAR);
  const uint32_t bars = smem_u32(smem + OFF_BARS);
  c.a_full = bars; c.w_full = bars + 64; c.w_empty = bars + 128; c.acc_full = bars + 192;
  const uint32_t tmem_slot = bars + 200;
  c.rank = cluster_rank();
  c.row0 = static_cast<int>(blockIdx.x / CS) * rm;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < KBLK; ++i) { mbar_init(bar_at(c.a_full, i), 1); mbar_init(bar_at(c.w_full, i), 1); mbar_init(bar_at(c.w_empty, i), 1); }
    mbar_init(c.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem) : "r"(tmem_slot) : "memory");
  pdl_sync();                    // barriers, tensor memory and the cluster are set up while the previous kernel drains
  cluster_sync_all();            // every CTA's barriers / stats buffers exist before any remote traffic
}

__device__ __forceinline__ void teardown(const Ctx& c, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(64) : "memory");
  }
  cluster_sync_all();            // no CTA leaves while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kThreads, 1)
clf_stack_fwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                     const __grid_constant__ CUtensorMap tmN, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmH, const StackParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Ctx c;
  setup(c, smem, warp, lane, p.rm);
  const int L = p.L;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) load_w<0>(c, &tmW1, 0, -1);
    for (int g = 0; g < 2 * L; ++g) {
      __syncwarp();
      if ((g & 1) == 0) { cluster_sync_all(); cluster_sync_all(); }     // the two LayerNorm statistics rounds
      cluster_sync_all();                                               // operand slices of GEMM g published
      if (lane == 0) {
        load_a(c, p.xchg, g);
        if (g + 1 < 2 * L) load_w<0>(c, ((g + 1) & 1) ? &tmW2 : &tmW1, (g + 1) >> 1, g & 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    for (int g = 0; g < 2 * L; ++g) {
      __syncwarp();
      if ((g & 1) == 0) { cluster_sync_all(); cluster_sync_all(); }
      cluster_sync_all();
      if (lane == 0) {
        long long* tl = (p.dbg != nullptr && blockIdx.x == 0 && g == 2 * (L / 2)) ? p.dbg + 10 : nullptr;
        if (tl != nullptr) tl[0] = clock64();
        tc_fence_after();
        issue_gemm<0>(c, static_cast<uint32_t>(g & 1), tl);
      }
    }
  } else {
    // ------------------------------------------------------------------ row owners (4 warps x 32 rows)
    // UMMA M = 128: accumulator row r sits in TMEM lane r.  M = 64: row r sits in lane 32*(r/16) + r%16 (probed on
    // B200, tools/probe_tmem_layout.py), i.e. the first 16 lanes of every warp own rows and the other 16 idle.
    const int q = warp & 3;
    const bool active = (c.rm == RM) || (lane < 16);
    const int rl = (c.rm == RM) ? q * 32 + lane : q * 16 + (lane & 15);
    const int row = c.row0 + rl;
    const bool valid = active && row < p.B;
    const int col0 = static_cast<int>(c.rank) * NS;
    const uint32_t taddr = c.tmem + (static_cast<uint32_t>(q * 32) << 16);
    const unsigned long long dseed = DROP ? __ldg(p.drop.seed) : 0ull;
    float hv[NS];
    if (valid) load_vec64(p.h + static_cast<size_t>(row) * PD + col0, hv);
    else {
#pragma unroll
      for (int k = 0; k < NS; ++k) hv[k] = 0.f;
    }
    // parameter vectors in shared memory: 0 gamma_o, 1 beta_o, 2 gamma_i, 3 beta_i, 4 b1, 5 b2
    const int et = threadIdx.x - 64;
    par_prefetch(c, p, 0, 0, et, col0, true);
    for (int i = 0; i < L; ++i) {
      const bool tl = (p.dbg != nullptr) && (blockIdx.x == 0) && (et == 0) && (i == L / 2);
      const int pb = i & 1;
      par_advance(c, p, i + 1, i + 1 < L, pb ^ 1, et, col0);
      float par[NS];
      SER_TL(0);
      // ---- y = LN_outer(h)
      float mu, rs;
      row_stats(c, 0, rl, active, hv, mu, rs);
      SER_TL(1);
      if (valid && c.rank == 0) *reinterpret_cast<float2*>(p.stats_o + (static_cast<size_t>(i) * p.B + row) * 2) = make_float2(mu, rs);
      lds_vec(par_addr(c, pb, 0), par, NS);
#pragma unroll
      for (int k = 0; k < NS; ++k) hv[k] = (hv[k] - mu) * rs * par[k];
      lds_vec(par_addr(c, pb, 1), par, NS);
#pragma unroll
      for (int k = 0; k < NS; ++k) hv[k] += par[k];                    // hv now holds y (kept for the residual)
      // ---- n = LN_inner(y)
      row_stats(c, 1, rl, active, hv, mu, rs);
      SER_TL(2);
      if (valid && c.rank == 0) *reinterpret_cast<float2*>(p.stats_i + (static_cast<size_t>(i) * p.B + row) * 2) = make_float2(mu, rs);
      float nv[NS];
      lds_vec(par_addr(c, pb, 2), par, NS);
#pragma unroll
      for (int k = 0; k < NS; ++k) nv[k] = (hv[k] - mu) * rs * par[k];
      lds_vec(par_addr(c, pb, 3), par, NS);
#pragma unroll
      for (int k = 0; k < NS; ++k) nv[k] += par[k];
      publish_slice(c, nv, rl, active, et, xchg_block(c, p.xchg, 2 * i), &tmN, col0, i);
      tc_fence_before();
      SER_TL(3);
      // the dropout decisions are data independent: hashed inside the cluster barrier, while this SM's MMA warp is
      // idle too (hashing during the GEMM would compete with the single MMA-issuing thread for issue slots)
      unsigned long long keep = 0ull;
      cluster_arrive();
      if (DROP) keep = drop_slice_bits(p.drop, dseed, DS_CLF_BLOCK0 + 2 * i, row, col0, c.rm != RM, lane);
      cluster_wait();
      SER_TL(4);
      // ---- u = relu(W1 n + b1)
      mbar_wait(c.acc_full, 0u);
      SER_TL(5);
      tc_fence_after();
      tmem_ld32(taddr, nv);
      tmem_ld32(taddr + 32, nv + 32);
      lds_vec(par_addr(c, pb, 4), par, NS);
#pragma unroll
      for (int k = 0; k < NS; ++k) nv[k] = fmaxf(nv[k] + par[k], 0.f);
      if (DROP) drop_slice_apply(keep, p.drop.scale, nv);                      // block[3]; r is saved post-dropout
      publish_slice(c, nv, rl, active, et, xchg_block(c, p.xchg, 2 * i + 1), &tmR, col0, i);
      tc_fence_before();
      SER_TL(6);
      cluster_arrive();
      if (DROP) keep = drop_slice_bits(p.drop, dseed, DS_CLF_BLOCK0 + 2 * i + 1, row, col0, c.rm != RM, lane);
      cluster_wait();
      SER_TL(7);
      // ---- h_next = y + W2 u + b2
      mbar_wait(c.acc_full, 1u);
      SER_TL(8);
      tc_fence_after();
      tmem_ld32(taddr, nv);
      tmem_ld32(taddr + 32, nv + 32);
      lds_vec(par_addr(c, pb, 5), par, NS);
      if (DROP) {                                                              // block[5]: y + dropout(W2 u + b2)
#pragma unroll
        for (int k = 0; k < NS; ++k) nv[k] += par[k];
        drop_slice_apply(keep, p.drop.scale, nv);
#pragma unroll
        for (int k = 0; k < NS; ++k) hv[k] += nv[k];
      } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) hv[k] += nv[k] + par[k];
      }
      store_slice_f32(c, hv, rl, active, et, &tmH, col0, i + 1);
      tc_fence_before();
      SER_TL(9);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (et == 0) bulk_wait_all();
  }
  teardown(c, warp);
}

// ---------------------------------------------------------------------------------------------------------
// backward (dX chain + LayerNorm parameter gradients)
// ---------------------------------------------------------------------------------------------------------
// sum over the 32 lanes of v[0..31] (one value per column): lane l ends up with the total of column l
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
__device__ __forceinline__ void colacc_add(const Ctx& c, int which, int col, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(c.sColacc + static_cast<uint32_t>((which * NS + col) * 4)), "f"(v) : "memory");
}

template <bool DROP>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kThreads, 1)
clf_stack_bwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                     const __grid_constant__ CUtensorMap tmDhn, const __grid_constant__ CUtensorMap tmDr,
                     const StackParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Ctx c;
  setup(c, smem, warp, lane, p.rm);
  const int L = p.L;

  // GEMM g (g = 0 .. 2L-1): even = du = dh W2 of block L-1-g/2, odd = dn = da W1 of the same block
  if (warp == 0) {
    if (lane == 0) load_w<1>(c, &tmW2, L - 1, -1);
    for (int g = 0; g < 2 * L; ++g) {
      __syncwarp();
      cluster_sync_all();                                               // operand slices of GEMM g published
      if (lane == 0) {
        load_a(c, p.xchg, g);
        if (g + 1 < 2 * L) {
          const int nl = L - 1 - ((g + 1) >> 1);
          load_w<1>(c, ((g + 1) & 1) ? &tmW1 : &tmW2, nl, g & 1);
        }
      }
      __syncwarp();
      if (g & 1) { cluster_sync_all(); cluster_sync_all(); }            // the two LayerNorm-backward reduction rounds
    }
  } else if (warp == 1) {
    for (int g = 0; g < 2 * L; ++g) {
      __syncwarp();
      cluster_sync_all();
      if (lane == 0) { tc_fence_after(); issue_gemm<1>(c, static_cast<uint32_t>(g & 1)); }
      __syncwarp();
      if (g & 1) { cluster_sync_all(); cluster_sync_all(); }
    }
  } else {
    // UMMA M = 128: accumulator row r sits in TMEM lane r.  M = 64: row r sits in lane 32*(r/16) + r%16 (probed on
    // B200, tools/probe_tmem_layout.py), i.e. the first 16 lanes of every warp own rows and the other 16 idle.
    const int q = warp & 3;
    const bool active = (c.rm == RM) || (lane < 16);
    const int rl = (c.rm == RM) ? q * 32 + lane : q * 16 + (lane & 15);
    const int row = c.row0 + rl;
    const bool valid = active && row < p.B;
    const int col0 = static_cast<int>(c.rank) * NS;
    const uint32_t taddr = c.tmem + (static_cast<uint32_t>(q * 32) << 16);
    const size_t BP = static_cast<size_t>(p.B) * PD;
    const int et = threadIdx.x - 64;                 // 0..127 among the row owners
    // zero the column accumulators (each thread owns two of the 256 entries)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(c.sColacc + et * 4), "f"(0.f) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(c.sColacc + (et + 128) * 4), "f"(0.f) : "memory");
    named_bar_sync(1, 128);
    const unsigned long long dseed = DROP ? __ldg(p.drop.seed) : 0ull;
    float gv[NS];
    if (valid) load_vec64(p.dh_in + static_cast<size_t>(row) * PD + col0, gv);
    else {
#pragma unroll
      for (int k = 0; k < NS; ++k) gv[k] = 0.f;
    }
    // parameter vectors in shared memory: 0 gamma_o, 1 beta_o, 2 gamma_i
    par_prefetch(c, p, L - 1, (L - 1) & 1, et, col0, true);
    unsigned long long keep = 0ull;             // block[5] keep bits of the block about to be processed
    if (DROP) keep = drop_slice_bits(p.drop, dseed, DS_CLF_BLOCK0 + 2 * (L - 1) + 1, row, col0, c.rm != RM, lane);
    for (int i = L - 1; i >= 0; --i) {
      const bool tl = (p.dbg != nullptr) && (blockIdx.x == 0) && (et == 0) && (i == L / 2);
      const int pb = i & 1;
      par_advance(c, p, i - 1, i > 0, pb ^ 1, et, col0);
      SER_TL(16);
      // ---- publish dh_{i+1} (bf16): A operand of du = dh W2 and of the batched dW2 GEMM
      if (DROP) {
        // block[5]: the branch (GEMM operand and the saved dW2 / db2 operand) sees mask * dh; gv keeps the skip path
        float gm[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) gm[k] = gv[k];
        drop_slice_apply(keep, p.drop.scale, gm);
        publish_slice(c, gm, rl, active, et, xchg_block(c, p.xchg, 2 * (L - 1 - i)), &tmDhn, col0, i);
      } else {
        publish_slice(c, gv, rl, active, et, xchg_block(c, p.xchg, 2 * (L - 1 - i)), &tmDhn, col0, i);
      }
      tc_fence_before();
      SER_TL(17);
      cluster_arrive();
      if (DROP && i > 0) keep = drop_slice_bits(p.drop, dseed, DS_CLF_BLOCK0 + 2 * (i - 1) + 1, row, col0, c.rm != RM, lane);   // next block's mask
      cluster_wait();
      SER_TL(18);
      // operands of the epilogues below, requested while the GEMM runs (row-per-thread loads: 32 L1 wavefronts per
      // instruction -- they must not sit in front of the publish in the load/store pipe, measured +4500 cycles)
      float xo[NS];
      float2 so = make_float2(0.f, 0.f), si = make_float2(0.f, 0.f);
      uint4 um[NS / 8];
      if (valid) {
        load_vec64(p.h + i * BP + static_cast<size_t>(row) * PD + col0, xo);
        so = *reinterpret_cast<const float2*>(p.stats_o + (static_cast<size_t>(i) * p.B + row) * 2);
        si = *reinterpret_cast<const float2*>(p.stats_i + (static_cast<size_t>(i) * p.B + row) * 2);
        const uint4* up = reinterpret_cast<const uint4*>(p.r + i * BP + static_cast<size_t>(row) * PD + col0);
#pragma unroll
        for (int k = 0; k < NS / 8; ++k) um[k] = up[k];
      } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) xo[k] = 0.f;
#pragma unroll
        for (int k = 0; k < NS / 8; ++k) um[k] = make_uint4(0u, 0u, 0u, 0u);
      }
      // ---- da = du * (u > 0)   (the loads above stay in flight across the wait: nothing consumes them yet)
      mbar_wait(c.acc_full, 0u);
      SER_TL(19);
      tc_fence_after();
      {
        float du[NS];
        tmem_ld32(taddr, du);
        tmem_ld32(taddr + 32, du + 32);
#pragma unroll
        for (int k = 0; k < NS / 8; ++k) {
          const uint32_t wds[4] = {um[k].x, um[k].y, um[k].z, um[k].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            // relu output is >= 0: "u > 0" <=> any magnitude bit set in the bf16 half
            if ((wds[t] & 0x00007fffu) == 0u) du[8 * k + 2 * t] = 0.f;
            if ((wds[t] & 0x7fff0000u) == 0u) du[8 * k + 2 * t + 1] = 0.f;
          }
        }
        if (DROP) {          // block[3]: r is saved post-dropout, so the test above is also the keep mask; kept units carry 1/(1-p)
#pragma unroll
          for (int k = 0; k < NS; ++k) du[k] *= p.drop.scale;
        }
        publish_slice(c, du, rl, active, et, xchg_block(c, p.xchg, 2 * (L - 1 - i) + 1), &tmDr, col0, i);
      }
      tc_fence_before();
      SER_TL(20);
      cluster_sync_all();
      SER_TL(21);
      // ---- dn = da W1, then the two LayerNorm backward steps on the register-resident row slice
      mbar_wait(c.acc_full, 1u);
      SER_TL(22);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < NS; ++k) xo[k] = (xo[k] - so.x) * so.y;        // x-hat of the outer LayerNorm
      const uint32_t go_p = par_addr(c, pb, 0), bo_p = par_addr(c, pb, 1), gi_p = par_addr(c, pb, 2);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float dn[32], t0[32], t1[32], go[32], bo[32], gi[32];
        tmem_ld32(taddr + hh * 32, dn);
        if (!active) {                       // idle lanes of an M = 64 tile read untouched tensor memory
#pragma unroll
          for (int k = 0; k < 32; ++k) dn[k] = 0.f;
        }
        lds_vec(go_p + hh * 128, go, 32); lds_vec(bo_p + hh * 128, bo, 32); lds_vec(gi_p + hh * 128, gi, 32);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int kk = hh * 32 + k;
          const float xi = (fmaf(xo[kk], go[k], bo[k]) - si.x) * si.y;
          const float a = dn[k] * gi[k];
          s1 += a; s2 = fmaf(a, xi, s2);
          t0[k] = dn[k] * xi; t1[k] = dn[k];
        }
        const float cg = transpose_reduce32(t0, lane);
        const float cb = transpose_reduce32(t1, lane);
        colacc_add(c, 0, hh * 32 + lane, cg);
        colacc_add(c, 1, hh * 32 + lane, cb);
      }
      {
        float as[CS], bs[CS];
        exchange2(c, 0, rl, active, s1, s2, as, bs);
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int t = 0; t < CS; ++t) { s1 += as[t]; s2 += bs[t]; }
        s1 *= (1.f / PD); s2 *= (1.f / PD);
      }
      SER_TL(23);
      float u1 = 0.f, u2 = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float dn[32], t0[32], t1[32], go[32], bo[32], gi[32];
        tmem_ld32(taddr + hh * 32, dn);
        if (!active) {                       // idle lanes of an M = 64 tile read untouched tensor memory
#pragma unroll
          for (int k = 0; k < 32; ++k) dn[k] = 0.f;
        }
        lds_vec(go_p + hh * 128, go, 32); lds_vec(bo_p + hh * 128, bo, 32); lds_vec(gi_p + hh * 128, gi, 32);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int kk = hh * 32 + k;
          const float xi = (fmaf(xo[kk], go[k], bo[k]) - si.x) * si.y;
          const float a = dn[k] * gi[k];
          const float dy = fmaf(si.y, a - s1 - xi * s2, gv[kk]);       // + skip gradient (residual from y)
          gv[kk] = dy;
          const float a2 = dy * go[k];
          u1 += a2; u2 = fmaf(a2, xo[kk], u2);
          t0[k] = dy * xo[kk]; t1[k] = dy;
        }
        const float cg = transpose_reduce32(t0, lane);
        const float cb = transpose_reduce32(t1, lane);
        colacc_add(c, 2, hh * 32 + lane, cg);
        colacc_add(c, 3, hh * 32 + lane, cb);
      }
      tc_fence_before();
      {
        float as[CS], bs[CS];
        exchange2(c, 1, rl, active, u1, u2, as, bs);
        u1 = 0.f; u2 = 0.f;
#pragma unroll
        for (int t = 0; t < CS; ++t) { u1 += as[t]; u2 += bs[t]; }
        u1 *= (1.f / PD); u2 *= (1.f / PD);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float go[32];
        lds_vec(go_p + hh * 128, go, 32);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int kk = hh * 32 + k;
          gv[kk] = so.y * (gv[kk] * go[k] - u1 - xo[kk] * u2);
        }
      }
      SER_TL(24);
      // ---- flush this block's LayerNorm parameter gradients (summed over the cluster's 128 rows)
      named_bar_sync(1, 128);
      {
        float v0, v1;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(c.sColacc + et * 4) : "memory");
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(c.sColacc + (et + 128) * 4) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(c.sColacc + et * 4), "f"(0.f) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(c.sColacc + (et + 128) * 4), "f"(0.f) : "memory");
        // entries: [0] dgamma_inner, [1] dbeta_inner, [2] dgamma_outer, [3] dbeta_outer, 64 columns each
        const int w0 = et >> 6, cc = et & 63;            // et: 0..127 -> which 0/1 ; et+128 -> which 2/3
        float* d0 = (w0 == 0 ? p.dlni_g : p.dlni_b) + i * p.s_blk + col0 + cc;
        float* d1 = (w0 == 0 ? p.dlno_g : p.dlno_b) + i * p.s_lno + col0 + cc;
        atomicAdd(d0, v0);
        atomicAdd(d1, v1);
      }
      named_bar_sync(1, 128);
      SER_TL(25);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (et == 0) bulk_wait_all();
    if (valid) store_f32_64(p.dh_out + static_cast<size_t>(row) * PD + col0, gv);
  }
  teardown(c, warp);
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}
// [layers][rows][512] bf16 or fp32 (row pitch 512, layer pitch `layer_stride` elements), box = box_rows x 128 bytes
int make_map(CUtensorMap* tm, const void* base, long long rows, long long layer_stride, int layers, int box_rows,
             int f32 = 0) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) { set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable"); return SER_ERR_CUDA; }
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(PD), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(layers)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(PD) * es, static_cast<cuuint64_t>(layer_stride) * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error(__FILE__, __LINE__, "clf_stack: cuTensorMapEncodeTiled failed"); return SER_ERR_CUDA; }
  return SER_OK;
}

template <typename K>
int configure(K kern) {
  SER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return SER_OK;
}

}  // namespace

// rows per cluster: the GEMM phases are bound by shared-memory fill (the 128 x 512 operand all-gather), so small
// batches run UMMA M = 64 clusters -- twice as many SMs pull half as much each; large batches keep M = 128
static int rows_per_cluster(int B) {
  static const char* force = getenv("SER_CLF_RM");
  if (force != nullptr) return atoi(force) == 64 ? 64 : RM;
  return (ceil_div(B, 64) * CS <= device_sm_count()) ? 64 : RM;
}

bool clf_stack_supported(int dtype, int P, int L, const ClfStackArgs& a) {
  // the exchange buffer lives in the (otherwise unused) saved-y buffer of the descriptor: [L, B, 512] fp32
  const int rm = rows_per_cluster(a.B);
  const size_t need = static_cast<size_t>(ceil_div(a.B, rm)) * 2 * CS * rm * 128;
  const size_t have = static_cast<size_t>(L) * a.B * PD * sizeof(float);
  return dtype == DT_BF16 && P == PD && L >= 1 && a.s_blk > 0 && a.s_lno > 0 && a.s_w1 > 0 && a.s_w2 > 0 &&
         (a.s_w1 % 8 == 0) && (a.s_w2 % 8 == 0) && a.xchg != nullptr && have >= need &&
         (reinterpret_cast<uintptr_t>(a.xchg) & 127) == 0;
}

static long long* timeline_buffer() {
  static long long* buf = nullptr;
  static bool init = false;
  if (!init) {
    init = true;
    if (getenv("SER_CLF_TIMELINE") != nullptr && cudaMalloc(&buf, 32 * sizeof(long long)) == cudaSuccess)
      cudaMemset(buf, 0, 32 * sizeof(long long));
    else
      buf = nullptr;
  }
  return buf;
}

static StackParams to_params(const ClfStackArgs& a, bool backward) {
  StackParams p{};
  p.B = a.B; p.L = a.L;
  p.rm = rows_per_cluster(a.B);
  p.b1 = a.b1; p.b2 = a.b2; p.lni_g = a.lni_g; p.lni_b = a.lni_b; p.s_blk = a.s_blk;
  p.lno_g = a.lno_g; p.lno_b = a.lno_b; p.s_lno = a.s_lno;
  p.h = a.h; p.n = reinterpret_cast<__nv_bfloat16*>(a.n); p.r = reinterpret_cast<__nv_bfloat16*>(a.r);
  p.stats_o = a.stats_o; p.stats_i = a.stats_i;
  p.dh_in = a.dh_in; p.dh_out = a.dh_out;
  p.dhn = reinterpret_cast<__nv_bfloat16*>(a.dhn); p.dr = reinterpret_cast<__nv_bfloat16*>(a.dr);
  p.dlni_g = a.dlni_g; p.dlni_b = a.dlni_b; p.dlno_g = a.dlno_g; p.dlno_b = a.dlno_b;
  const float* pv[6] = {a.lno_g, a.lno_b, a.lni_g, a.lni_b, a.b1, a.b2};
  const long long ps[6] = {a.s_lno, a.s_lno, a.s_blk, a.s_blk, a.s_blk, a.s_blk};
  for (int i = 0; i < 6; ++i) { p.pv[i] = pv[i]; p.ps[i] = ps[i]; }
  p.npv = backward ? 3 : 6;
  p.dbg = timeline_buffer();
  p.xchg = reinterpret_cast<char*>(a.xchg);
  p.drop = a.drop;
  return p;
}

// SER_CLF_TIMELINE=1: one row-owner thread stamps clock64() at the phase boundaries of the middle block; the
// host prints the deltas after the launch (debug aid: synchronises the stream)
static void timeline_report(const char* what, cudaStream_t s) {
  long long* d = timeline_buffer();
  if (d == nullptr) return;
  long long h[32];
  cudaStreamSynchronize(s);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  fprintf(stderr, "[clf timeline %s]", what);
  const int lo = (what[0] == 'f') ? 0 : 16, hi = (what[0] == 'f') ? 9 : 25;
  for (int k = lo + 1; k <= hi; ++k) fprintf(stderr, " t%d-t%d=%lld", k, k - 1, h[k] - h[k - 1]);
  fprintf(stderr, " | block=%lld cycles", h[hi] - h[lo]);
  if (what[0] == 'f')
    fprintf(stderr, " | mma(GEMM1): barrier->a_full[0]=%lld a_full[0]->a_full[7]=%lld ->issued=%lld ; epilogue saw barrier at %+lld, acc_full at %+lld (vs mma barrier exit)",
            h[11] - h[10], h[12] - h[11], h[13] - h[12], h[4] - h[10], h[5] - h[10]);
  fprintf(stderr, "\n");
}

int clf_stack_fwd(const ClfStackArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    SER_TRY(configure(clf_stack_fwd_kernel<false>));
    SER_TRY(configure(clf_stack_fwd_kernel<true>));
    configured = true;
  }
  CUtensorMap tmW1, tmW2, tmN, tmR, tmH;
  SER_TRY(make_map(&tmW1, a.w1, PD, a.s_w1, a.L, NS));
  SER_TRY(make_map(&tmW2, a.w2, PD, a.s_w2, a.L, NS));
  const int rm = rows_per_cluster(a.B);
  SER_TRY(make_map(&tmN, a.n, a.B, static_cast<long long>(a.B) * PD, a.L, rm));
  SER_TRY(make_map(&tmR, a.r, a.B, static_cast<long long>(a.B) * PD, a.L, rm));
  SER_TRY(make_map(&tmH, a.h, a.B, static_cast<long long>(a.B) * PD, a.L + 1, rm, 1));
  const int clusters = ceil_div(a.B, rm);
  // algorithmic work: 2 GEMMs per block; bytes: weights once per cluster + the fp32 stream and bf16 operands
  ProfScope prof("clf_stack_fwd", 4.0 * a.B * PD * PD * a.L,
                 static_cast<double>(a.L) * (2.0 * PD * PD * 2 * clusters + a.B * PD * (4.0 + 2.0 + 2.0)), s);
  auto* kern = a.drop.on() ? clf_stack_fwd_kernel<true> : clf_stack_fwd_kernel<false>;
  SER_CUDA_CHECK(launch_pdl(kern, dim3(clusters * CS), dim3(kThreads), kSmemBytes, s, tmW1, tmW2, tmN, tmR, tmH, to_params(a, false)));
  SER_LAUNCH_CHECK();
  timeline_report("fwd", s);
  return SER_OK;
}

int clf_stack_bwd(const ClfStackArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    SER_TRY(configure(clf_stack_bwd_kernel<false>));
    SER_TRY(configure(clf_stack_bwd_kernel<true>));
    configured = true;
  }
  CUtensorMap tmW1, tmW2, tmDhn, tmDr;
  SER_TRY(make_map(&tmW1, a.w1, PD, a.s_w1, a.L, 64));
  SER_TRY(make_map(&tmW2, a.w2, PD, a.s_w2, a.L, 64));
  const int rm = rows_per_cluster(a.B);
  SER_TRY(make_map(&tmDhn, a.dhn, a.B, static_cast<long long>(a.B) * PD, a.L, rm));
  SER_TRY(make_map(&tmDr, a.dr, a.B, static_cast<long long>(a.B) * PD, a.L, rm));
  const int clusters = ceil_div(a.B, rm);
  ProfScope prof("clf_stack_bwd", 4.0 * a.B * PD * PD * a.L,
                 static_cast<double>(a.L) * (2.0 * PD * PD * 2 * clusters + a.B * PD * (4.0 + 2.0 + 2.0 + 2.0)), s);
  auto* kern = a.drop.on() ? clf_stack_bwd_kernel<true> : clf_stack_bwd_kernel<false>;
  SER_CUDA_CHECK(launch_pdl(kern, dim3(clusters * CS), dim3(kThreads), kSmemBytes, s, tmW1, tmW2, tmDhn, tmDr, to_params(a, true)));
  SER_LAUNCH_CHECK();
  timeline_report("bwd", s);
  return SER_OK;
}

}  // namespace ser

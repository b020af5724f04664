// Hardware probe (debug aid, not part of the C-ABI in include/ser_head.h): where does tcgen05.mma cta_group::1 with
// M = 64 put accumulator row r in tensor memory?  D[r][n] = (r + 1) * (n + 1) is computed with one K = 16 MMA and
// all 128 TMEM lanes x 64 columns are dumped, so the lane that holds row r can be read off the values.
#include "common.cuh"   // -I <package>/csrc (tools/probe_tmem_layout.py builds this file)
#include <cuda.h>

namespace ser {
namespace {

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {      // K-major, SWIZZLE_128B, SBO = 1024
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

template <int M>
__global__ void __launch_bounds__(128) probe_kernel(float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // [128 rows][64 k] swizzled
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + 16384);   // [64 n][64 k] swizzled
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < (16384 + 8192) / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.f);
  __syncthreads();
  // element (r, k) of a K-major SW128 tile: r * 128 B + ((k / 8) ^ (r & 7)) * 16 B + (k % 8) * 2 B
  if (t < M) A[(t * 128 + ((0 ^ (t & 7)) * 16)) / 2] = __float2bfloat16(static_cast<float>(t + 1));
  if (t < 64) Bm[(t * 128 + ((0 ^ (t & 7)) * 16)) / 2] = __float2bfloat16(static_cast<float>(t + 1));
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // pre-fill all 128 lanes x 64 columns with a sentinel so untouched lanes are visible
  {
    const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < 64; ++c) {
      const uint32_t v = __float_as_uint(-7.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (t == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(64 >> 3) << 17) |
                               (static_cast<uint32_t>(M >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem), "l"(desc_k(s32(A))), "l"(desc_k(s32(Bm))), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
  }
  {
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
    } while (!done);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(ta + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[(warp * 32 + lane) * 64 + c] = __uint_as_float(v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}


// MMA timing probe: `reps` back-to-back tcgen05.mma (K = 16) of shape M x N into the same accumulator, one commit,
// clock64 from first issue to observed completion.  Operands are whatever is in shared memory (zeros).
template <int M>
__global__ void __launch_bounds__(128) mma_time_kernel(long long* __restrict__ out, int N, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (t == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>(M >> 4) << 24);
    const uint64_t ad = desc_k(s32(smem)), bd = desc_k(s32(smem + 16384));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(r > 0 ? 1u : 0u) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
    } while (!done);
    const long long t2 = clock64();
    out[0] = t1 - t0;      // issue time
    out[1] = t2 - t0;      // until completion observed
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// Tensor-memory load / store throughput probe: `nw` warps of one CTA (block of 512 threads) each issue `reps`
// back-to-back tcgen05.ld (or st) of 32 lanes x 32 columns (4 KB per instruction) and one wait; clock64 per warp.
// mode 0: ld 32x32b.x32, 1: st 32x32b.x32, 2: ld 32x32b.x64, 3: ld 16x256b.x8 (also 4 KB)
__global__ void __launch_bounds__(512) ldtm_time_kernel(long long* __restrict__ out, int nw, int reps, int mode) {
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  if (warp < nw) {
    const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = i;
    const long long t0 = clock64();
    for (int k = 0; k < reps; ++k) {
      const uint32_t a = base + static_cast<uint32_t>((k * 64) & 255);
      if (mode == 0) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(a));
      } else if (mode == 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                     "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                     "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                     ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                       "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                       "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
      } else if (mode == 3) {
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(a));
      }
    }
    if (mode == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
    if ((t & 31) == 0) out[warp] = t1 - t0;
  }
  if (acc == 0x12345678u) out[63] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace
}  // namespace ser

extern "C" int ser_debug_ldtm_time(long long* out, int nw, int reps, int mode, int ctas, void* stream) {
  ser::ldtm_time_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, nw, reps, mode);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// out: [128 lanes][64 columns] fp32 (device).  m = 64 or 128.
extern "C" int ser_debug_probe_tmem_layout(float* out, int m, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int smem = 16384 + 8192 + 64 + 1024;
  if (m == 64) ser::probe_kernel<64><<<1, 128, smem, s>>>(out);
  else ser::probe_kernel<128><<<1, 128, smem, s>>>(out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// out[0] = cycles to issue, out[1] = cycles until the commit is observed, for `reps` MMAs of shape m x n x 16
extern "C" int ser_debug_mma_time(long long* out, int m, int n, int reps, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int smem = 16384 + 32768 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(ser::mma_time_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(ser::mma_time_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cfg = true;
  }
  if (m == 64) ser::mma_time_kernel<64><<<1, 128, smem, s>>>(out, n, reps);
  else ser::mma_time_kernel<128><<<1, 128, smem, s>>>(out, n, reps);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

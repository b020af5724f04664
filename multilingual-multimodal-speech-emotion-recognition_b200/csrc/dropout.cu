// Stand-alone dropout passes over dense [rows, cols] tensors (the sites whose producing kernel does not apply the
// mask itself) and the mask export used by the parity tests.  Mask definition: dropout.cuh.
#include "kernels.cuh"
#include "dropout.cuh"
#include "prof.cuh"

namespace ser {

namespace {

// 8 consecutive elements (4 column pairs) per thread and iteration
template <typename T>
__global__ void __launch_bounds__(256)
dropout_apply_kernel(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ res, long long n8, DropSpec d) {
  pdl_sync();
  const DropKey key = drop_key(d);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8], r[8];
    load8(in + i * 8, v);
    if (res != nullptr) load8(res + i * 8, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 m = drop_pair(key, static_cast<unsigned>(i * 4 + k), d.thr, d.scale);
      v[2 * k] *= m.x; v[2 * k + 1] *= m.y;
    }
    if (res != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    }
    store8(out + i * 8, v);
  }
}

// scalar tail / small tensors
template <typename T>
__global__ void dropout_apply_tail_kernel(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ res,
                                          long long begin, long long n, DropSpec d) {
  pdl_sync();
  const DropKey key = drop_key(d);
  const long long i = begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float v = to_f32(in[i]) * drop_one(key, static_cast<unsigned>(i >> 1), static_cast<unsigned>(i & 1), d.thr, d.scale);
  if (res != nullptr) v += to_f32(res[i]);
  out[i] = from_f32<T>(v);
}

__global__ void dropout_mask_kernel(float* __restrict__ out, long long rows, int cols, DropSpec d) {
  pdl_sync();
  const DropKey key = d.on() ? drop_key(d) : DropKey{0u, 1u};      // p = 0: no seed to read, all ones
  const long long n = rows * cols;
  const unsigned half_cols = static_cast<unsigned>((cols + 1) / 2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned r = static_cast<unsigned>(i / cols), c = static_cast<unsigned>(i % cols);
    out[i] = d.on() ? drop_one(key, r * half_cols + (c >> 1), c & 1u, d.thr, d.scale) : 1.f;
  }
}

template <typename T>
int apply_impl(const void* in, void* out, const void* res, long long n, const DropSpec& d, cudaStream_t s) {
  const T* pi = reinterpret_cast<const T*>(in);
  T* po = reinterpret_cast<T*>(out);
  const T* pr = reinterpret_cast<const T*>(res);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 31) == 0;
  const long long n8 = aligned ? n / 8 : 0;
  if (n8 > 0) {
    const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (n8 + 255) / 256));
    SER_CUDA_CHECK(launch_pdl(dropout_apply_kernel<T>, dim3(blocks), dim3(256), 0, s, pi, po, pr, n8, d));
    SER_LAUNCH_CHECK();
  }
  if (n8 * 8 < n) {
    const long long rem = n - n8 * 8;
    SER_CUDA_CHECK(launch_pdl(dropout_apply_tail_kernel<T>, dim3(ceil_div(rem, 256)), dim3(256), 0, s, pi, po, pr, n8 * 8, n, d));
    SER_LAUNCH_CHECK();
  }
  return SER_OK;
}

}  // namespace

int dropout_apply(const void* in, void* out, const void* res, int f32, long long rows, int cols, const DropSpec& d,
                  cudaStream_t s) {
  SER_REQUIRE(d.on(), "dropout_apply: dropout is off");
  SER_REQUIRE(cols % 2 == 0, "dropout: dense sites need an even number of columns");
  const long long n = rows * cols;
  SER_REQUIRE(n > 0 && n / 2 < (1LL << 32), "dropout: site too large for the 32-bit pair index");
  ProfScope prof("dropout", static_cast<double>(n), (f32 ? 4.0 : 2.0) * n * (res ? 3.0 : 2.0), s);
  return f32 ? apply_impl<float>(in, out, res, n, d, s) : apply_impl<__nv_bfloat16>(in, out, res, n, d, s);
}

int dropout_mask(const DropSpec& d, long long rows, int cols, float* out, cudaStream_t s) {
  SER_REQUIRE(rows > 0 && cols > 0 && out != nullptr, "dropout_mask: empty site");
  SER_REQUIRE(rows * ((cols + 1) / 2) < (1LL << 32), "dropout: site too large for the 32-bit pair index");
  const long long n = rows * cols;
  const int blocks = static_cast<int>(min(static_cast<long long>(device_sm_count()) * 8, (n + 255) / 256));
  SER_CUDA_CHECK(launch_pdl(dropout_mask_kernel, dim3(blocks), dim3(256), 0, s, out, rows, cols, d));
  SER_LAUNCH_CHECK();
  return SER_OK;
}

}  // namespace ser

// Counter-based dropout masks shared by every kernel that implements an nn.Dropout of the path
// (cross_attention.py:18,25,43,51; fusion.py:9,12; classifier.py:83,85,109,127,195).
//
// A mask is a pure function of (seed, site, row, col): no state, nothing materialised, so the forward kernel, the
// backward kernels and the test-side mask export (ser_dropout_mask) all regenerate the same decisions.  `seed` is read
// from DEVICE memory (one uint64), so a CUDA graph replays with fresh masks when the host side bumps the seed in-graph.
// One 32-bit hash decides two neighbouring columns (16 bits each): element (row, col) of a [rows, cols] site uses
//   bits = mix((row * ceil(cols/2) + col/2) * key.mul + key.add),  draw = col odd ? bits >> 16 : bits & 0xffff,
//   keep = draw >= round(p * 65536),   value = keep ? x / (1 - p) : 0          (torch.nn.functional.dropout semantics)
// The stream differs from torch's Philox stream (it has to: SURVEY.md section 8(c)); the distribution is the same.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ser_head.h"

namespace ser {

// site ids: one per nn.Dropout instance of a module (classifier block l uses DS_CLF_BLOCK0 + 2 l + {0: hidden, 1: out})
enum : unsigned {
  DS_XA_PROB_A = SER_DS_XA_PROB_A, DS_XA_PROB_T = SER_DS_XA_PROB_T, DS_XA_RES_A = SER_DS_XA_RES_A, DS_XA_RES_T = SER_DS_XA_RES_T,
  DS_FUS_A = SER_DS_FUS_A, DS_FUS_T = SER_DS_FUS_T,
  DS_FEAT = SER_DS_FEAT, DS_CLF_IN = SER_DS_CLF_IN, DS_CLF_OUT = SER_DS_CLF_OUT, DS_CLF_UNC = SER_DS_CLF_UNC,
  DS_CLF_BLOCK0 = SER_DS_CLF_BLOCK0,
};

struct DropSpec {
  const unsigned long long* seed = nullptr;   // device pointer
  unsigned thr = 0;                           // keep iff 16-bit draw >= thr; 0 = dropout off
  float scale = 1.f;                          // 1 / (1 - p)
  unsigned site = 0;
  __host__ __device__ bool on() const { return thr != 0; }
};

inline DropSpec make_drop(const unsigned long long* seed, float p, unsigned site) {
  DropSpec d;
  if (seed == nullptr || !(p > 0.f)) return d;
  long t = lrintf(p * 65536.f);
  if (t < 1) t = 1;
  if (t > 65536) t = 65536;
  d.seed = seed; d.site = site;
  d.thr = static_cast<unsigned>(t);
  d.scale = (p >= 1.f) ? 0.f : 1.f / (1.f - p);
  return d;
}
inline DropSpec with_site(DropSpec d, unsigned site) { d.site = site; return d; }

struct DropKey { unsigned add, mul; };

__device__ __forceinline__ unsigned drop_mix(unsigned x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
// key of (seed value, site); kernels that visit many sites load the seed once and derive the keys from the register
__device__ __forceinline__ DropKey drop_key_of(unsigned long long s, unsigned site) {
  DropKey k;
  const unsigned a = drop_mix(static_cast<unsigned>(s) + 0x9E3779B9u * (site + 1u));
  k.mul = drop_mix(static_cast<unsigned>(s >> 32) ^ a ^ 0x7F4A7C15u) | 1u;
  k.add = a * k.mul;        // (pair + a) * mul = pair * mul + add
  return k;
}
__device__ __forceinline__ DropKey drop_key(const DropSpec& d) { return drop_key_of(__ldg(d.seed), d.site); }
// the two 16-bit draws of column pair `pair` (= row * ceil(cols/2) + col/2): keyed odd multiply-add (one IMAD: the
// key's offset is pre-multiplied) followed by the full murmur3 finalizer.  A cheaper one-multiply finalizer was tried
// and rejected: its per-row / per-column keep rates are over-dispersed (z-score std 1.15-2.0 instead of 1.0,
// tests/test_gpu_parity.py::test_dropout_mask_statistics_and_determinism) and it bought no measurable time.
__device__ __forceinline__ unsigned drop_bits(const DropKey& k, unsigned pair) { return drop_mix(pair * k.mul + k.add); }
// multipliers (0 or scale) of the even / odd column of a pair
__device__ __forceinline__ float2 drop_pair(const DropKey& k, unsigned pair, unsigned thr, float scale) {
  const unsigned b = drop_bits(k, pair);
  return make_float2((b & 0xffffu) >= thr ? scale : 0.f, (b >> 16) >= thr ? scale : 0.f);
}
// multiplier of a single element, `half` = col & 1
__device__ __forceinline__ float drop_one(const DropKey& k, unsigned pair, unsigned half, unsigned thr, float scale) {
  const unsigned b = drop_bits(k, pair);
  return ((half ? (b >> 16) : (b & 0xffffu)) >= thr) ? scale : 0.f;
}

// out = in * mask (+ res) over a dense [rows, cols] site (cols even); in place allowed; f32 != 0: fp32 storage else bf16
int dropout_apply(const void* in, void* out, const void* res, int f32, long long rows, int cols, const DropSpec& d,
                  cudaStream_t s);
// out[r, c] = mask multiplier (0 or 1/(1-p)) as fp32: what the kernels apply at (site, r, c); tests feed it to the oracle
int dropout_mask(const DropSpec& d, long long rows, int cols, float* out, cudaStream_t s);

}  // namespace ser

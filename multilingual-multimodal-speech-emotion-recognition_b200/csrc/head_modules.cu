// Module-level orchestration behind the C-ABI: each function enqueues the kernel sequence of one
// reference module (forward or backward) on the caller's stream.  The same code drives both tiers:
// `dtype` picks the CUDA-core fp32 GEMM or the tcgen05 bf16 GEMM and the storage type of activations.
#include "kernels.cuh"
#include "modules.cuh"
#include <stdlib.h>
#include "../../include/ser_head.h"

namespace ser {

namespace {

inline size_t esize(int dtype) { return dtype == DT_F32 ? 4 : 2; }
inline int is_f32(int dtype) { return dtype == DT_F32 ? 1 : 0; }
inline const void* off(const void* p, long long elems, int dtype) {
  return reinterpret_cast<const char*>(p) + elems * static_cast<long long>(esize(dtype));
}
inline void* off(void* p, long long elems, int dtype) {
  return reinterpret_cast<char*>(p) + elems * static_cast<long long>(esize(dtype));
}

// bump allocator over the caller-provided workspace
struct Arena {
  char* base; size_t cap; size_t used = 0; bool ok = true;
  Arena(void* p, size_t n) : base(reinterpret_cast<char*>(p)), cap(n) {}
  void* take(size_t bytes) {
    const size_t a = (used + 255) & ~size_t(255);
    if (base == nullptr || a + bytes > cap) { ok = false; return nullptr; }
    used = a + bytes;
    return base + a;
  }
};
inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// C[M,Nout] = act(A[M,Kin] W[Nout,Kin]^T + bias) (+ R)
GemmArgs fwd_args(int dt, int M, int Nout, int Kin, const void* A, long long lda, const void* W, long long ldw,
                  const float* bias, void* C, long long ldc, int c_f32, int act, const void* R, long long ldr, int r_f32) {
  GemmArgs g;
  g.dtype = dt; g.M = M; g.N = Nout; g.K = Kin;
  g.A = A; g.lda = lda; g.a_trans = 0;
  g.B = W; g.ldb = ldw; g.b_trans = 0;
  g.C = C; g.ldc = ldc; g.c_f32 = c_f32;
  g.bias = bias; g.act = act;
  g.R = R; g.ldr = ldr; g.r_f32 = r_f32;
  return g;
}
int linear_fwd(int dt, int M, int Nout, int Kin, const void* A, long long lda, const void* W, long long ldw,
               const float* bias, void* C, long long ldc, int c_f32, int act, const void* R, long long ldr, int r_f32,
               cudaStream_t s) {
  return gemm(fwd_args(dt, M, Nout, Kin, A, lda, W, ldw, bias, C, ldc, c_f32, act, R, ldr, r_f32), s);
}

// dX[M,Kin] = gate'(dY[M,Nout] W[Nout,Kin]) (+ R)
GemmArgs dgrad_args(int dt, int M, int Nout, int Kin, const void* dY, long long lddy, const void* W, long long ldw,
                    void* dX, long long lddx, int dx_f32, const void* G, long long ldg, int g_f32, int gate_mode,
                    const void* R, long long ldr, int r_f32, float alpha = 1.f) {
  GemmArgs g;
  g.alpha = alpha;
  g.dtype = dt; g.M = M; g.N = Kin; g.K = Nout;
  g.A = dY; g.lda = lddy; g.a_trans = 0;
  g.B = W; g.ldb = ldw; g.b_trans = 1;
  g.C = dX; g.ldc = lddx; g.c_f32 = dx_f32;
  g.G = G; g.ldg = ldg; g.g_f32 = g_f32; g.gate_mode = gate_mode;
  g.R = R; g.ldr = ldr; g.r_f32 = r_f32;
  return g;
}
int linear_dgrad(int dt, int M, int Nout, int Kin, const void* dY, long long lddy, const void* W, long long ldw,
                 void* dX, long long lddx, int dx_f32, const void* G, long long ldg, int g_f32, int gate_mode,
                 const void* R, long long ldr, int r_f32, cudaStream_t s, float alpha = 1.f) {
  return gemm(dgrad_args(dt, M, Nout, Kin, dY, lddy, W, ldw, dX, lddx, dx_f32, G, ldg, g_f32, gate_mode, R, ldr, r_f32, alpha), s);
}

// dW[Nout,Kin] (fp32) = dY[M,Nout]^T X[M,Kin];  db[Nout] (optional) = column sums of dY, produced by the same GEMM in
// the bf16 tier (gemm_tc.cu row-sum MMAs) and by a colsum launch otherwise
GemmArgs wgrad_args(int dt, int M, int Nout, int Kin, const void* dY, long long lddy, const void* X, long long ldx,
                    float* dW, long long lddw, float* db = nullptr, int zeroed = 0) {
  GemmArgs g;
  g.rowsum = db;
  g.out_zeroed = zeroed;          // dW / db already hold zeros (descriptor field grads_zeroed): no memsets
  g.dtype = dt; g.M = Nout; g.N = Kin; g.K = M;
  g.A = dY; g.lda = lddy; g.a_trans = 1;
  g.B = X; g.ldb = ldx; g.b_trans = 1;
  g.C = dW; g.ldc = lddw; g.c_f32 = 1;
  return g;
}
int linear_wgrad(int dt, int M, int Nout, int Kin, const void* dY, long long lddy, const void* X, long long ldx,
                 float* dW, long long lddw, cudaStream_t s, float* db = nullptr, int zeroed = 0) {
  return gemm(wgrad_args(dt, M, Nout, Kin, dY, lddy, X, ldx, dW, lddw, db, zeroed), s);
}

// The same GEMM for the audio-side and the text-side operands of a module: ONE batched launch (batch = 2) when every
// operand's text-side pointer sits at a positive, 8-element aligned distance from its audio-side twin (the Python
// binding allocates the twin activations as halves of one tensor, and a module's parameters / gradients are laid out
// symmetrically in its flat buffer); two launches otherwise.
int gemm_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t s) {
  auto dist = [](const void* p1, const void* p0, long long es, long long* out) -> bool {
    if ((p0 == nullptr) != (p1 == nullptr)) return false;
    if (p0 == nullptr) { *out = 0; return true; }
    const long long bytes = reinterpret_cast<const char*>(p1) - reinterpret_cast<const char*>(p0);
    if (bytes <= 0 || bytes % (8 * es) != 0) return false;
    *out = bytes / es;
    return true;
  };
  static const bool disabled = (getenv("SER_NO_PAIR_BATCH") != nullptr);      // A/B switch
  const long long ea = g0.dtype == DT_F32 ? 4 : 2;
  GemmArgs g = g0;
  bool ok = !disabled && g0.M == g1.M && g0.N == g1.N && g0.K == g1.K && g0.lda == g1.lda && g0.ldb == g1.ldb &&
            g0.ldc == g1.ldc && g0.ldr == g1.ldr && g0.ldg == g1.ldg && g0.batch <= 1 && g1.batch <= 1;
  ok = ok && dist(g1.A, g0.A, ea, &g.strideA) && dist(g1.B, g0.B, ea, &g.strideB) &&
       dist(g1.C, g0.C, g0.c_f32 ? 4 : 2, &g.strideC) && dist(g1.R, g0.R, g0.r_f32 ? 4 : 2, &g.strideR) &&
       dist(g1.G, g0.G, g0.g_f32 ? 4 : 2, &g.strideG) && dist(g1.bias, g0.bias, 4, &g.strideBias) &&
       dist(g1.rowsum, g0.rowsum, 4, &g.strideRS);
  if (ok) {
    g.batch = 2;
    return gemm(g, s);
  }
  SER_TRY(gemm(g0, s));
  return gemm(g1, s);
}

// zero a parameter-gradient region unless the caller handed in zero-filled buffers
#define SER_ZERO_UNLESS(zeroed, ptr, bytes)                                   \
  do {                                                                        \
    if (!(zeroed)) SER_TRY(zero_async((ptr), (bytes), s));                    \
  } while (0)

}  // namespace

// =================================================================================================
// a1 adapter
// =================================================================================================
int adapter_fwd(const ser_adapter_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  SER_REQUIRE(d.M > 0 && d.x && d.h && d.y, "adapter_fwd: null tensor");
  SER_TRY(linear_fwd(dt, d.M, d.S, d.D, d.x, d.D, d.w1, d.D, d.b1, d.h, d.S, f, ACT_RELU, nullptr, 0, f, s));
  SER_TRY(linear_fwd(dt, d.M, d.D, d.S, d.h, d.S, d.w2, d.S, d.b2, d.y, d.D, f, ACT_NONE,
                     d.add_residual ? d.x : nullptr, d.D, f, s));
  return SER_OK;
}

int adapter_bwd(const ser_adapter_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  SER_REQUIRE(d.M > 0 && d.dy && d.dh && d.h && d.x, "adapter_bwd: null tensor");
  SER_TRY(linear_wgrad(dt, d.M, d.D, d.S, d.dy, d.D, d.h, d.S, d.dw2, d.S, s, d.db2, d.grads_zeroed));
  SER_TRY(linear_dgrad(dt, d.M, d.D, d.S, d.dy, d.D, d.w2, d.S, d.dh, d.S, f, d.h, d.S, f, GATE_RELU, nullptr, 0, f, s));
  SER_TRY(linear_wgrad(dt, d.M, d.S, d.D, d.dh, d.S, d.x, d.D, d.dw1, d.D, s, d.db1, d.grads_zeroed));
  if (d.dx != nullptr)
    SER_TRY(linear_dgrad(dt, d.M, d.S, d.D, d.dh, d.S, d.w1, d.D, d.dx, d.D, f, nullptr, 0, f, GATE_NONE,
                         d.add_residual ? d.dy : nullptr, d.D, f, s));
  return SER_OK;
}

// =================================================================================================
// a2 cross-modal attention
// =================================================================================================

// ------------------------------------------------------------------------------------------------
// bf16 tier with folded Linear chains (fold.cu): per modality ONE token-level GEMM produces [Q' | K' | V'] and ONE
// produces z = a + out(out_proj(ctx)); the 8 token-level 256->256 GEMMs of the unfolded path (and their dgrad / wgrad /
// bias-gradient launches) become weight-sized GEMMs.  fold_w / fold_b layouts:
//   fold_w (act dtype): Wbd_a [3S,3S] | Wbd_t | Wc_a [3S,D] | Wc_t | Wz_a [D,S] | Wz_t
//   fold_b (fp32)     : bc_a [3S] | bc_t | bz_a [D] | bz_t
// ------------------------------------------------------------------------------------------------
namespace {
struct FoldBufs {
  void* wbd_a; void* wbd_t; void* wc_a; void* wc_t; void* wz_a; void* wz_t;
  float* bc_a; float* bc_t; float* bz_a; float* bz_t;
};
FoldBufs fold_bufs(const ser_xattn_desc& d) {
  const long long S = d.S, D = d.D, S3 = 3 * S;
  FoldBufs f;
  char* w = reinterpret_cast<char*>(d.fold_w);
  const size_t e = 2;
  f.wbd_a = w; f.wbd_t = w + S3 * S3 * e;
  f.wc_a = w + 2 * S3 * S3 * e; f.wc_t = w + (2 * S3 * S3 + S3 * D) * e;
  f.wz_a = w + (2 * S3 * S3 + 2 * S3 * D) * e; f.wz_t = w + (2 * S3 * S3 + 2 * S3 * D + D * S) * e;
  f.bc_a = d.fold_b; f.bc_t = d.fold_b + S3; f.bz_a = d.fold_b + 2 * S3; f.bz_t = d.fold_b + 2 * S3 + D;
  return f;
}
// C[M,N] (c_f32 ? fp32 : bf16) = op(A) op(B)^T, plain bf16 tensor-core GEMM on weight-sized operands
int small_gemm(int M, int N, int K, const void* A, long long lda, int a_trans, const void* B, long long ldb, int b_trans,
               void* C, long long ldc, int c_f32, cudaStream_t s) {
  GemmArgs g;
  g.dtype = DT_BF16; g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_trans = a_trans;
  g.B = B; g.ldb = ldb; g.b_trans = b_trans;
  g.C = C; g.ldc = ldc; g.c_f32 = c_f32;
  g.splits = 1;
  return gemm(g, s);
}
// the same weight-sized GEMM for both modalities: one batched launch when every operand's audio -> text distance is a
// positive multiple of 8 elements (it is: parameters, gradients and scratch are laid out symmetrically), else two
int small_gemm2(int M, int N, int K, const void* A0, const void* A1, long long lda, int a_trans, const void* B0,
                const void* B1, long long ldb, int b_trans, void* C0, void* C1, long long ldc, int c_f32, cudaStream_t s) {
  auto dist = [](const void* p1, const void* p0, long long es) {
    const long long bytes = reinterpret_cast<const char*>(p1) - reinterpret_cast<const char*>(p0);
    return (bytes > 0 && bytes % (8 * es) == 0) ? bytes / es : -1;
  };
  const long long sa = dist(A1, A0, 2), sb = dist(B1, B0, 2), sc = dist(C1, C0, c_f32 ? 4 : 2);
  if (sa > 0 && sb > 0 && sc > 0) {
    GemmArgs g;
    g.dtype = DT_BF16; g.M = M; g.N = N; g.K = K;
    g.A = A0; g.lda = lda; g.a_trans = a_trans; g.B = B0; g.ldb = ldb; g.b_trans = b_trans;
    g.C = C0; g.ldc = ldc; g.c_f32 = c_f32; g.splits = 1;
    g.batch = 2; g.strideA = sa; g.strideB = sb; g.strideC = sc;
    return gemm(g, s);
  }
  SER_TRY(small_gemm(M, N, K, A0, lda, a_trans, B0, ldb, b_trans, C0, ldc, c_f32, s));
  return small_gemm(M, N, K, A1, lda, a_trans, B1, ldb, b_trans, C1, ldc, c_f32, s);
}
bool xattn_folded(const ser_xattn_desc& d) {
  return xattn_fold_enabled(d.dtype, d.D, d.S) && (d.Dt == 0 || d.Dt == d.D) && d.fold_w != nullptr && d.fold_b != nullptr;
}
}  // namespace

// The ONE place that decides whether the Linear chains of cross attention are folded (exported as ser_xattn_folded):
// the binding sizes qkv_* / o_* / fold_* from this answer, so the SER_NO_FOLD switch can never make the library write
// full-size activations into placeholder buffers.
bool xattn_fold_enabled(int dtype, int D, int S) {
  static const bool disabled = (getenv("SER_NO_FOLD") != nullptr);     // A/B switch
  return !disabled && dtype == DT_BF16 && D == 3 * S;
}

static int xattn_fwd_folded(const ser_xattn_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = 0;
  const int S = d.S, D = d.D, S3 = 3 * d.S;
  const int Ma = d.B * d.Ta, Mt = d.B * d.Tt;
  const FoldBufs fb = fold_bufs(d);
  if (!d.reuse_text) {      // (reuse_text: the folded weights of an earlier call with these buffers are still valid)
    // weight-only part: block-diagonal in-projections, folded weights and biases
    SER_TRY(fold_assemble(d.win_a, d.win_t, fb.wbd_a, fb.wbd_t, S, s));
    SER_TRY(small_gemm2(S3, D, S3, fb.wbd_a, fb.wbd_t, S3, 0, d.wqkv_a, d.wqkv_t, D, 1, fb.wc_a, fb.wc_t, D, 0, s));   // Wc = Wbd Wqkv
    SER_TRY(small_gemm2(D, S, S, d.wout_a, d.wout_t, S, 0, d.wo_a, d.wo_t, S, 1, fb.wz_a, fb.wz_t, S, 0, s));           // Wz = Wout Wo
    FoldBiasArgs ba{};
    ba.S = S; ba.D = D; ba.win_a = d.win_a; ba.win_t = d.win_t; ba.bin_a = d.bin_a; ba.bin_t = d.bin_t;
    ba.bqkv_a = d.bqkv_a; ba.bqkv_t = d.bqkv_t; ba.wout_a = d.wout_a; ba.wout_t = d.wout_t; ba.bo_a = d.bo_a; ba.bo_t = d.bo_t;
    ba.bout_a = d.bout_a; ba.bout_t = d.bout_t; ba.bc_a = fb.bc_a; ba.bc_t = fb.bc_t; ba.bz_a = fb.bz_a; ba.bz_t = fb.bz_t;
    SER_TRY(fold_bias_fwd(ba, s));
  }
  // token-level part
  SER_TRY(linear_fwd(dt, Ma, S3, D, d.a, D, fb.wc_a, D, fb.bc_a, d.p_a, S3, f, ACT_NONE, nullptr, 0, f, s));
  if (!d.reuse_text) SER_TRY(linear_fwd(dt, Mt, S3, D, d.t, D, fb.wc_t, D, fb.bc_t, d.p_t, S3, f, ACT_NONE, nullptr, 0, f, s));
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  AttnArgs at{};
  at.dtype = dt; at.B = d.B; at.H = d.H; at.dh = S / d.H;
  at.scale = 1.0f / sqrtf(static_cast<float>(at.dh));
  at.Tq = d.Ta; at.Tk = d.Tt;
  at.Q = d.p_a; at.ldq = S3;
  at.K = off(d.p_t, S, dt); at.ldk = S3;
  at.V = off(d.p_t, 2 * S, dt); at.ldv = S3;
  at.kmask = d.t_mask; at.O = d.ctx_a; at.ldo = S; at.lse = d.lse_a;
  at.drop = with_site(drop, DS_XA_PROB_A); at.keep_bits = d.keep_a;
  SER_TRY(attention_fwd(at, s));
  at.Tq = d.Tt; at.Tk = d.Ta;
  at.Q = d.p_t;
  at.K = off(d.p_a, S, dt);
  at.V = off(d.p_a, 2 * S, dt);
  at.kmask = d.a_mask; at.O = d.ctx_t; at.lse = d.lse_t;
  at.drop = with_site(drop, DS_XA_PROB_T); at.keep_bits = d.keep_t;
  SER_TRY(attention_fwd(at, s));
  // z = x + dropout(ctx Wz^T + bz): with dropout on, mask and residual move from the GEMM epilogue into the
  // LayerNorm kernel's prologue (which also saves z for the backward)
  const DropSpec drop_ra = with_site(drop, DS_XA_RES_A), drop_rt = with_site(drop, DS_XA_RES_T);
  SER_TRY(linear_fwd(dt, Ma, D, S, d.ctx_a, S, fb.wz_a, S, fb.bz_a, d.z_a, D, f, ACT_NONE, drop.on() ? nullptr : d.a, D, f, s));
  SER_TRY(layernorm_fwd(d.z_a, f, d.enh_a, f, nullptr, f, d.ln_a_g, d.ln_a_b, d.stats_a, Ma, D, 0, s,
                        drop.on() ? d.a : nullptr, drop.on() ? d.z_a : nullptr, &drop_ra));
  SER_TRY(linear_fwd(dt, Mt, D, S, d.ctx_t, S, fb.wz_t, S, fb.bz_t, d.z_t, D, f, ACT_NONE, drop.on() ? nullptr : d.t, D, f, s));
  SER_TRY(layernorm_fwd(d.z_t, f, d.enh_t, f, nullptr, f, d.ln_t_g, d.ln_t_b, d.stats_t, Mt, D, 0, s,
                        drop.on() ? d.t : nullptr, drop.on() ? d.z_t : nullptr, &drop_rt));
  return SER_OK;
}

static size_t xattn_fold_ws_bytes(int D, int S) {
  const size_t S3 = 3 * static_cast<size_t>(S), d = D, s = S;
  // dbz, dbc (fp32) ; dWz32, dWz16 ; dWc32 (reused as dWbd32), dWc16   -- all x 2 modalities
  return pad256(2 * d * 4) + pad256(2 * S3 * 4) + pad256(2 * d * s * 4) + pad256(2 * d * s * 2) + pad256(2 * S3 * d * 4) +
         pad256(2 * S3 * d * 2) + 4096;
}

static int xattn_bwd_folded(const ser_xattn_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = 0;
  const int S = d.S, D = d.D, S3 = 3 * d.S;
  const int Ma = d.B * d.Ta, Mt = d.B * d.Tt;
  const size_t e = 2;
  const FoldBufs fb = fold_bufs(d);
  Arena ws(d.ws, d.ws_bytes);
  void* dz_a = ws.take(static_cast<size_t>(Ma) * D * e);
  void* dz_t = ws.take(static_cast<size_t>(Mt) * D * e);
  void* dctx_a = ws.take(static_cast<size_t>(Ma) * S * e);
  void* dctx_t = ws.take(static_cast<size_t>(Mt) * S * e);
  void* dp_a = ws.take(static_cast<size_t>(Ma) * S3 * e);
  void* dp_t = ws.take(static_cast<size_t>(Mt) * S3 * e);
  float* delta_a = reinterpret_cast<float*>(ws.take(static_cast<size_t>(d.B) * d.H * d.Ta * sizeof(float)));
  float* delta_t = reinterpret_cast<float*>(ws.take(static_cast<size_t>(d.B) * d.H * d.Tt * sizeof(float)));
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  // with dropout on, the branch gradient is mask * dz (its own buffer: the skip path keeps the unmasked dz)
  void* dzm_a = drop.on() ? ws.take(static_cast<size_t>(Ma) * D * e) : dz_a;
  void* dzm_t = drop.on() ? ws.take(static_cast<size_t>(Mt) * D * e) : dz_t;
  // weight-sized scratch, [audio | text] contiguous per kind so one cast / one batched GEMM serves both modalities
  float* dbz[2]; float* dbc[2]; float* dwz32[2]; void* dwz16[2]; float* dwc32[2]; void* dwc16[2];
  {
    const size_t nz = static_cast<size_t>(D) * S, nc = static_cast<size_t>(S3) * D;
    float* pbz = reinterpret_cast<float*>(ws.take(2 * static_cast<size_t>(D) * 4));
    float* pbc = reinterpret_cast<float*>(ws.take(2 * static_cast<size_t>(S3) * 4));
    float* pz32 = reinterpret_cast<float*>(ws.take(2 * nz * 4));
    char* pz16 = reinterpret_cast<char*>(ws.take(2 * nz * 2));
    float* pc32 = reinterpret_cast<float*>(ws.take(2 * nc * 4));
    char* pc16 = reinterpret_cast<char*>(ws.take(2 * nc * 2));
    for (int m = 0; m < 2 && ws.ok; ++m) {
      dbz[m] = pbz + m * D; dbc[m] = pbc + m * S3;
      dwz32[m] = pz32 + m * nz; dwz16[m] = pz16 + m * nz * 2;
      dwc32[m] = pc32 + m * nc; dwc16[m] = pc16 + m * nc * 2;
    }
  }
  if (!ws.ok) { set_last_error(__FILE__, __LINE__, "xattn_bwd: workspace too small"); return SER_ERR_WORKSPACE; }
  // the weight-sized scratch (split-K / row-sum targets dbz, dbc, dWz, dWc) is zeroed with ONE memset
  SER_TRY(zero_async(dbz[0], reinterpret_cast<char*>(dwc16[0]) - reinterpret_cast<char*>(dbz[0]), s));

  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_a_g, sizeof(float) * D);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_a_b, sizeof(float) * D);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_t_g, sizeof(float) * D);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_t_b, sizeof(float) * D);
  // with dropout on, the LayerNorm backward also writes the branch gradient mask * dz (the skip path keeps dz)
  const DropSpec drop_ra = with_site(drop, DS_XA_RES_A), drop_rt = with_site(drop, DS_XA_RES_T);
  SER_TRY(layernorm_bwd(d.d_enh_a, f, d.z_a, f, d.stats_a, d.ln_a_g, d.ln_a_b, nullptr, f, dz_a, f, nullptr, f,
                        d.dln_a_g, d.dln_a_b, Ma, D, 0, s, drop.on() ? dzm_a : nullptr, &drop_ra));
  SER_TRY(layernorm_bwd(d.d_enh_t, f, d.z_t, f, d.stats_t, d.ln_t_g, d.ln_t_b, nullptr, f, dz_t, f, nullptr, f,
                        d.dln_t_g, d.dln_t_b, Mt, D, 0, s, drop.on() ? dzm_t : nullptr, &drop_rt));
  // z = ctx Wz^T + bz + residual
  struct SideZ { int M; void* dz; const void* ctx; void* dctx; const void* wz; const void* wout; const void* wo;
                 float* dwout; float* dwo; int m; };
  const SideZ sz[2] = {
      {Ma, dzm_a, d.ctx_a, dctx_a, fb.wz_a, d.wout_a, d.wo_a, d.dwout_a, d.dwo_a, 0},
      {Mt, dzm_t, d.ctx_t, dctx_t, fb.wz_t, d.wout_t, d.wo_t, d.dwout_t, d.dwo_t, 1},
  };
  for (const SideZ& z : sz) {
    SER_TRY(linear_wgrad(dt, z.M, D, S, z.dz, D, z.ctx, S, dwz32[z.m], S, s, dbz[z.m], 1));
    SER_TRY(linear_dgrad(dt, z.M, D, S, z.dz, D, z.wz, S, z.dctx, S, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f, s));
  }
  SER_TRY(cast_any(dwz32[0], 1, dwz16[0], 0, 2LL * D * S, s));
  // Wz = Wout Wo :  dWout = dWz Wo^T ,  dWo = Wout^T dWz
  SER_TRY(small_gemm2(D, S, S, dwz16[0], dwz16[1], S, 0, d.wo_a, d.wo_t, S, 0, d.dwout_a, d.dwout_t, S, 1, s));
  SER_TRY(small_gemm2(S, S, D, d.wout_a, d.wout_t, S, 1, dwz16[0], dwz16[1], S, 1, d.dwo_a, d.dwo_t, S, 1, s));
  // attention core backward
  AttnArgs at{};
  at.dtype = dt; at.B = d.B; at.H = d.H; at.dh = S / d.H;
  at.scale = 1.0f / sqrtf(static_cast<float>(at.dh));
  at.ldq = at.ldk = at.ldv = S3; at.ldo = S; at.lddo = S; at.lddq = at.lddk = at.lddv = S3;
  at.Tq = d.Ta; at.Tk = d.Tt;
  at.Q = d.p_a; at.K = off(d.p_t, S, dt); at.V = off(d.p_t, 2 * S, dt);
  at.kmask = d.t_mask; at.O = d.ctx_a; at.lse = d.lse_a; at.dO = dctx_a; at.delta = delta_a;
  at.dQ = dp_a; at.dK = off(dp_t, S, dt); at.dV = off(dp_t, 2 * S, dt);
  at.drop = with_site(drop, DS_XA_PROB_A); at.keep_bits = d.keep_a;
  SER_TRY(attention_bwd(at, s));
  at.Tq = d.Tt; at.Tk = d.Ta;
  at.Q = d.p_t; at.K = off(d.p_a, S, dt); at.V = off(d.p_a, 2 * S, dt);
  at.kmask = d.a_mask; at.O = d.ctx_t; at.lse = d.lse_t; at.dO = dctx_t; at.delta = delta_t;
  at.dQ = dp_t; at.dK = off(dp_a, S, dt); at.dV = off(dp_a, 2 * S, dt);
  at.drop = with_site(drop, DS_XA_PROB_T); at.keep_bits = d.keep_t;
  SER_TRY(attention_bwd(at, s));
  // p = x Wc^T + bc
  struct SideP { int M; void* dp; const void* x; void* dx; const void* dz; const void* wc; const void* wbd; const void* wqkv;
                 float* dwqkv; int m; };
  const SideP sp[2] = {
      {Ma, dp_a, d.a, d.da, dz_a, fb.wc_a, fb.wbd_a, d.wqkv_a, d.dwqkv_a, 0},
      {Mt, dp_t, d.t, d.dt, dz_t, fb.wc_t, fb.wbd_t, d.wqkv_t, d.dwqkv_t, 1},
  };
  for (const SideP& p : sp) {
    SER_TRY(linear_wgrad(dt, p.M, S3, D, p.dp, S3, p.x, D, dwc32[p.m], D, s, dbc[p.m], 1));
    SER_TRY(linear_dgrad(dt, p.M, S3, D, p.dp, S3, p.wc, D, p.dx, D, f, nullptr, 0, f, GATE_NONE, p.dz, D, f, s));
  }
  SER_TRY(cast_any(dwc32[0], 1, dwc16[0], 0, 2LL * S3 * D, s));
  // Wc = Wbd Wqkv :  dWqkv = Wbd^T dWc (exact: off-diagonal blocks of Wbd are zero) ,  dWbd = dWc Wqkv^T
  SER_TRY(small_gemm2(S3, D, S3, fb.wbd_a, fb.wbd_t, S3, 1, dwc16[0], dwc16[1], D, 1, d.dwqkv_a, d.dwqkv_t, D, 1, s));
  SER_TRY(small_gemm2(S3, S3, D, dwc16[0], dwc16[1], D, 0, d.wqkv_a, d.wqkv_t, D, 0, dwc32[0], dwc32[1], S3, 1, s));   // dWbd32 reuses dWc32
  FoldBwdArgs g{};
  g.S = S; g.D = D; g.win_a = d.win_a; g.win_t = d.win_t; g.wout_a = d.wout_a; g.wout_t = d.wout_t;
  g.bqkv_a = d.bqkv_a; g.bqkv_t = d.bqkv_t; g.bo_a = d.bo_a; g.bo_t = d.bo_t;
  g.dwbd_a = dwc32[0]; g.dwbd_t = dwc32[1]; g.dbc_a = dbc[0]; g.dbc_t = dbc[1]; g.dbz_a = dbz[0]; g.dbz_t = dbz[1];
  g.dwin_a = d.dwin_a; g.dwin_t = d.dwin_t; g.dbin_a = d.dbin_a; g.dbin_t = d.dbin_t; g.dbqkv_a = d.dbqkv_a; g.dbqkv_t = d.dbqkv_t;
  g.dbo_a = d.dbo_a; g.dbo_t = d.dbo_t; g.dwout_a = d.dwout_a; g.dwout_t = d.dwout_t; g.dbout_a = d.dbout_a; g.dbout_t = d.dbout_t;
  SER_TRY(fold_bwd_glue(g, s));
  return SER_OK;
}

int xattn_fwd(const ser_xattn_desc& d, cudaStream_t s) {
  if (xattn_folded(d)) return xattn_fwd_folded(d, s);
  const int dt = d.dtype, f = is_f32(dt);
  const int S = d.S, D = d.D, S3 = 3 * d.S;
  const int Dt = d.Dt > 0 ? d.Dt : d.D;              // CrossModalAttention(audio_dim != text_dim): text-side width
  const int Ma = d.B * d.Ta, Mt = d.B * d.Tt;
  SER_REQUIRE(S % d.H == 0, "xattn: shared_dim must be divisible by num_heads");   // cross_attention.py:12
  SER_REQUIRE(Ma > 0 && Mt > 0, "xattn: empty input");
  SER_REQUIRE(d.qkv_a && d.qkv_t && d.o_a && d.o_t, "xattn: the unfolded path needs full-size qkv_* / o_* buffers");
  // outer projections, one packed GEMM per modality (cross_attention.py:38-40,46-48)
  SER_TRY(linear_fwd(dt, Ma, S3, D, d.a, D, d.wqkv_a, D, d.bqkv_a, d.qkv_a, S3, f, ACT_NONE, nullptr, 0, f, s));
  if (!d.reuse_text) SER_TRY(linear_fwd(dt, Mt, S3, Dt, d.t, Dt, d.wqkv_t, Dt, d.bqkv_t, d.qkv_t, S3, f, ACT_NONE, nullptr, 0, f, s));
  // MHA in-projections (torch/nn/functional.py:5798 chunking): attn_a takes (qa, kt, vt), attn_t takes (qt, ka, va)
  struct InProj { const void* src; int M; int scol; const void* w; const float* b; int wrow; void* dst; int dcol; };
  const InProj ip[6] = {
      {d.qkv_a, Ma, 0,     d.win_a, d.bin_a, 0,     d.p_a, 0},
      {d.qkv_t, Mt, S,     d.win_a, d.bin_a, S,     d.p_t, S},
      {d.qkv_t, Mt, 2 * S, d.win_a, d.bin_a, 2 * S, d.p_t, 2 * S},
      {d.qkv_t, Mt, 0,     d.win_t, d.bin_t, 0,     d.p_t, 0},
      {d.qkv_a, Ma, S,     d.win_t, d.bin_t, S,     d.p_a, S},
      {d.qkv_a, Ma, 2 * S, d.win_t, d.bin_t, 2 * S, d.p_a, 2 * S},
  };
  for (const InProj& p : ip) {
    if (d.reuse_text && p.dst == d.p_t) continue;          // text-side projections of an earlier view
    SER_TRY(linear_fwd(dt, p.M, S, S, off(p.src, p.scol, dt), S3, off(p.w, static_cast<long long>(p.wrow) * S, dt), S,
                       p.b + p.wrow, off(p.dst, p.dcol, dt), S3, f, ACT_NONE, nullptr, 0, f, s));
  }
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  AttnArgs at{};
  at.dtype = dt; at.B = d.B; at.H = d.H; at.dh = S / d.H;
  at.scale = 1.0f / sqrtf(static_cast<float>(at.dh));
  // A <- T
  at.Tq = d.Ta; at.Tk = d.Tt;
  at.Q = d.p_a; at.ldq = S3;
  at.K = off(d.p_t, S, dt); at.ldk = S3;
  at.V = off(d.p_t, 2 * S, dt); at.ldv = S3;
  at.kmask = d.t_mask; at.O = d.ctx_a; at.ldo = S; at.lse = d.lse_a;
  at.drop = with_site(drop, DS_XA_PROB_A); at.keep_bits = d.keep_a;
  SER_TRY(attention_fwd(at, s));
  // T <- A
  at.Tq = d.Tt; at.Tk = d.Ta;
  at.Q = d.p_t;
  at.K = off(d.p_a, S, dt);
  at.V = off(d.p_a, 2 * S, dt);
  at.kmask = d.a_mask; at.O = d.ctx_t; at.lse = d.lse_t;
  at.drop = with_site(drop, DS_XA_PROB_T); at.keep_bits = d.keep_t;
  SER_TRY(attention_fwd(at, s));
  // out_proj, out_a / out_t, dropout, + residual, LayerNorm (cross_attention.py:42-43,50-51)
  SER_TRY(linear_fwd(dt, Ma, S, S, d.ctx_a, S, d.wo_a, S, d.bo_a, d.o_a, S, f, ACT_NONE, nullptr, 0, f, s));
  const DropSpec drop_ra = with_site(drop, DS_XA_RES_A), drop_rt = with_site(drop, DS_XA_RES_T);
  SER_TRY(linear_fwd(dt, Ma, D, S, d.o_a, S, d.wout_a, S, d.bout_a, d.z_a, D, f, ACT_NONE, drop.on() ? nullptr : d.a, D, f, s));
  SER_TRY(layernorm_fwd(d.z_a, f, d.enh_a, f, nullptr, f, d.ln_a_g, d.ln_a_b, d.stats_a, Ma, D, 0, s,
                        drop.on() ? d.a : nullptr, drop.on() ? d.z_a : nullptr, &drop_ra));
  SER_TRY(linear_fwd(dt, Mt, S, S, d.ctx_t, S, d.wo_t, S, d.bo_t, d.o_t, S, f, ACT_NONE, nullptr, 0, f, s));
  SER_TRY(linear_fwd(dt, Mt, Dt, S, d.o_t, S, d.wout_t, S, d.bout_t, d.z_t, Dt, f, ACT_NONE, drop.on() ? nullptr : d.t, Dt, f, s));
  SER_TRY(layernorm_fwd(d.z_t, f, d.enh_t, f, nullptr, f, d.ln_t_g, d.ln_t_b, d.stats_t, Mt, Dt, 0, s,
                        drop.on() ? d.t : nullptr, drop.on() ? d.z_t : nullptr, &drop_rt));
  return SER_OK;
}

size_t xattn_bwd_ws_bytes(int dtype, int B, int Ta, int Tt, int D, int S, int H) {
  const size_t e = esize(dtype);
  size_t tot = 0;
  for (int T : {Ta, Tt}) {
    const size_t M = static_cast<size_t>(B) * T;
    tot += 2 * pad256(M * D * e);        // dz, mask * dz (dropout)
    tot += 2 * pad256(M * S * e);        // do, dctx
    tot += 2 * pad256(M * 3 * S * e);    // dp, dqkv
    tot += pad256(static_cast<size_t>(B) * H * T * sizeof(float));   // delta
  }
  return tot + 4096 + xattn_fold_ws_bytes(D, S);
}

int xattn_bwd(const ser_xattn_desc& d, cudaStream_t s) {
  if (xattn_folded(d)) return xattn_bwd_folded(d, s);
  SER_REQUIRE(d.qkv_a && d.qkv_t && d.o_a && d.o_t, "xattn: the unfolded path needs full-size qkv_* / o_* buffers");
  const int dt = d.dtype, f = is_f32(dt);
  const int S = d.S, D = d.D, S3 = 3 * d.S;
  const int Dt = d.Dt > 0 ? d.Dt : d.D;              // text-side width (ser_head.h: Dt)
  const int Ma = d.B * d.Ta, Mt = d.B * d.Tt;
  const size_t e = esize(dt);
  Arena ws(d.ws, d.ws_bytes);
  void* dz_a = ws.take(static_cast<size_t>(Ma) * D * e);
  void* dz_t = ws.take(static_cast<size_t>(Mt) * Dt * e);
  void* do_a = ws.take(static_cast<size_t>(Ma) * S * e);
  void* do_t = ws.take(static_cast<size_t>(Mt) * S * e);
  void* dctx_a = ws.take(static_cast<size_t>(Ma) * S * e);
  void* dctx_t = ws.take(static_cast<size_t>(Mt) * S * e);
  void* dp_a = ws.take(static_cast<size_t>(Ma) * S3 * e);
  void* dp_t = ws.take(static_cast<size_t>(Mt) * S3 * e);
  void* dqkv_a = ws.take(static_cast<size_t>(Ma) * S3 * e);
  void* dqkv_t = ws.take(static_cast<size_t>(Mt) * S3 * e);
  float* delta_a = reinterpret_cast<float*>(ws.take(static_cast<size_t>(d.B) * d.H * d.Ta * sizeof(float)));
  float* delta_t = reinterpret_cast<float*>(ws.take(static_cast<size_t>(d.B) * d.H * d.Tt * sizeof(float)));
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  void* dzm_a = drop.on() ? ws.take(static_cast<size_t>(Ma) * D * e) : dz_a;
  void* dzm_t = drop.on() ? ws.take(static_cast<size_t>(Mt) * Dt * e) : dz_t;
  if (!ws.ok) { set_last_error(__FILE__, __LINE__, "xattn_bwd: workspace too small"); return SER_ERR_WORKSPACE; }

  // LayerNorm backward (parameter gradients accumulate with atomics -> zero first)
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_a_g, sizeof(float) * D);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_a_b, sizeof(float) * D);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_t_g, sizeof(float) * Dt);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_t_b, sizeof(float) * Dt);
  const DropSpec drop_ra = with_site(drop, DS_XA_RES_A), drop_rt = with_site(drop, DS_XA_RES_T);
  SER_TRY(layernorm_bwd(d.d_enh_a, f, d.z_a, f, d.stats_a, d.ln_a_g, d.ln_a_b, nullptr, f, dz_a, f, nullptr, f,
                        d.dln_a_g, d.dln_a_b, Ma, D, 0, s, drop.on() ? dzm_a : nullptr, &drop_ra));
  SER_TRY(layernorm_bwd(d.d_enh_t, f, d.z_t, f, d.stats_t, d.ln_t_g, d.ln_t_b, nullptr, f, dz_t, f, nullptr, f,
                        d.dln_t_g, d.dln_t_b, Mt, Dt, 0, s, drop.on() ? dzm_t : nullptr, &drop_rt));
  // out_a / out_t and out_proj
  struct Side { int M; int D; void* dz; const void* o; const void* ctx; void* dob; void* dctx; const void* wout; const void* wo;
                float* dwout; float* dbout; float* dwo; float* dbo; };
  // (branch gradient = mask * dz from the LayerNorm backward; the skip path -- last two GEMMs below -- keeps dz)
  const Side sides[2] = {
      {Ma, D, dzm_a, d.o_a, d.ctx_a, do_a, dctx_a, d.wout_a, d.wo_a, d.dwout_a, d.dbout_a, d.dwo_a, d.dbo_a},
      {Mt, Dt, dzm_t, d.o_t, d.ctx_t, do_t, dctx_t, d.wout_t, d.wo_t, d.dwout_t, d.dbout_t, d.dwo_t, d.dbo_t},
  };
  for (const Side& sd : sides) {
    SER_TRY(linear_wgrad(dt, sd.M, sd.D, S, sd.dz, sd.D, sd.o, S, sd.dwout, S, s, sd.dbout, d.grads_zeroed));
    SER_TRY(linear_dgrad(dt, sd.M, sd.D, S, sd.dz, sd.D, sd.wout, S, sd.dob, S, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f, s));
    SER_TRY(linear_wgrad(dt, sd.M, S, S, sd.dob, S, sd.ctx, S, sd.dwo, S, s, sd.dbo, d.grads_zeroed));
    SER_TRY(linear_dgrad(dt, sd.M, S, S, sd.dob, S, sd.wo, S, sd.dctx, S, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f, s));
  }
  // attention core backward
  AttnArgs at{};
  at.dtype = dt; at.B = d.B; at.H = d.H; at.dh = S / d.H;
  at.scale = 1.0f / sqrtf(static_cast<float>(at.dh));
  at.ldq = at.ldk = at.ldv = S3; at.ldo = S; at.lddo = S; at.lddq = at.lddk = at.lddv = S3;
  // A <- T
  at.Tq = d.Ta; at.Tk = d.Tt;
  at.Q = d.p_a; at.K = off(d.p_t, S, dt); at.V = off(d.p_t, 2 * S, dt);
  at.kmask = d.t_mask; at.O = d.ctx_a; at.lse = d.lse_a; at.dO = dctx_a; at.delta = delta_a;
  at.dQ = dp_a; at.dK = off(dp_t, S, dt); at.dV = off(dp_t, 2 * S, dt);
  at.drop = with_site(drop, DS_XA_PROB_A); at.keep_bits = d.keep_a;
  SER_TRY(attention_bwd(at, s));
  // T <- A
  at.Tq = d.Tt; at.Tk = d.Ta;
  at.Q = d.p_t; at.K = off(d.p_a, S, dt); at.V = off(d.p_a, 2 * S, dt);
  at.kmask = d.a_mask; at.O = d.ctx_t; at.lse = d.lse_t; at.dO = dctx_t; at.delta = delta_t;
  at.dQ = dp_t; at.dK = off(dp_a, S, dt); at.dV = off(dp_a, 2 * S, dt);
  at.drop = with_site(drop, DS_XA_PROB_T); at.keep_bits = d.keep_t;
  SER_TRY(attention_bwd(at, s));
  // MHA in-projection backward
  struct InProjB { const void* dp; const void* src; int M; int col; const void* w; float* dw; float* db; int wrow; void* dsrc; };
  const InProjB ip[6] = {
      {dp_a, d.qkv_a, Ma, 0,     d.win_a, d.dwin_a, d.dbin_a, 0,     dqkv_a},
      {dp_t, d.qkv_t, Mt, S,     d.win_a, d.dwin_a, d.dbin_a, S,     dqkv_t},
      {dp_t, d.qkv_t, Mt, 2 * S, d.win_a, d.dwin_a, d.dbin_a, 2 * S, dqkv_t},
      {dp_t, d.qkv_t, Mt, 0,     d.win_t, d.dwin_t, d.dbin_t, 0,     dqkv_t},
      {dp_a, d.qkv_a, Ma, S,     d.win_t, d.dwin_t, d.dbin_t, S,     dqkv_a},
      {dp_a, d.qkv_a, Ma, 2 * S, d.win_t, d.dwin_t, d.dbin_t, 2 * S, dqkv_a},
  };
  for (const InProjB& p : ip) {
    const void* g = off(p.dp, p.col, dt);
    SER_TRY(linear_wgrad(dt, p.M, S, S, g, S3, off(p.src, p.col, dt), S3, p.dw + static_cast<long long>(p.wrow) * S, S, s, p.db + p.wrow, d.grads_zeroed));
    SER_TRY(linear_dgrad(dt, p.M, S, S, g, S3, off(p.w, static_cast<long long>(p.wrow) * S, dt), S,
                         off(p.dsrc, p.col, dt), S3, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f, s));
  }
  // outer projections + residual path
  SER_TRY(linear_wgrad(dt, Ma, S3, D, dqkv_a, S3, d.a, D, d.dwqkv_a, D, s, d.dbqkv_a, d.grads_zeroed));
  SER_TRY(linear_dgrad(dt, Ma, S3, D, dqkv_a, S3, d.wqkv_a, D, d.da, D, f, nullptr, 0, f, GATE_NONE, dz_a, D, f, s));
  SER_TRY(linear_wgrad(dt, Mt, S3, Dt, dqkv_t, S3, d.t, Dt, d.dwqkv_t, Dt, s, d.dbqkv_t, d.grads_zeroed));
  SER_TRY(linear_dgrad(dt, Mt, S3, Dt, dqkv_t, S3, d.wqkv_t, Dt, d.dt, Dt, f, nullptr, 0, f, GATE_NONE, dz_t, Dt, f, s));
  return SER_OK;
}

// =================================================================================================
// a3 attentive statistics pooling
// =================================================================================================
static AspArgs to_asp(const ser_asp_desc& d) {
  AspArgs a{};
  a.dtype = d.dtype; a.B = d.B; a.T = d.T; a.D = d.D; a.Hd = d.Hd;
  a.x = d.x; a.u = d.u; a.w2 = d.w2; a.b2 = d.b2; a.mask = d.mask; a.e = d.e; a.alpha = d.alpha;
  a.out = d.out; a.out_f32 = d.out_f32; a.dout = d.dout; a.dout_f32 = d.dout_f32; a.dx = d.dx;
  a.dalpha = d.dalpha; a.dpre = d.dpre; a.dw2 = d.dw2; a.db2 = d.db2;
  return a;
}

int asp_module_fwd(const ser_asp_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int M = d.B * d.T;
  SER_REQUIRE(M > 0 && d.x && d.u && d.e && d.alpha && d.out, "asp_fwd: null tensor");
  SER_TRY(linear_fwd(dt, M, d.Hd, d.D, d.x, d.D, d.w1, d.D, d.b1, d.u, d.Hd, f, ACT_TANH, nullptr, 0, f, s));
  return asp_fwd(to_asp(d), s);
}

int asp_module_bwd(const ser_asp_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int M = d.B * d.T;
  SER_REQUIRE(d.dout && d.dx && d.dpre && d.dalpha, "asp_bwd: null tensor");
  SER_ZERO_UNLESS(d.grads_zeroed, d.dw2, sizeof(float) * d.Hd);
  SER_ZERO_UNLESS(d.grads_zeroed, d.db2, sizeof(float));
  SER_TRY(asp_bwd(to_asp(d), s));
  SER_TRY(linear_wgrad(dt, M, d.Hd, d.D, d.dpre, d.Hd, d.x, d.D, d.dw1, d.D, s, d.db1, d.grads_zeroed));
  SER_TRY(linear_dgrad(dt, M, d.Hd, d.D, d.dpre, d.Hd, d.w1, d.D, d.dx, d.D, f, nullptr, 0, f, GATE_NONE, d.dx, d.D, f, s));
  return SER_OK;
}

// =================================================================================================
// a4 gated fusion
// =================================================================================================
static MixArgs to_mix(const ser_fusion_desc& d) {
  MixArgs m{};
  m.dtype = d.dtype; m.B = d.B; m.P = d.P; m.G = d.G;
  m.pa = d.pa; m.pt = d.pt; m.ga = d.ga; m.gt = d.gt;
  m.wga = d.wg2a; m.bga = d.bg2a; m.wgt = d.wg2t; m.bgt = d.bg2t;
  m.gates = d.gates; m.fused = d.fused; m.dfused = d.dfused;
  m.dwga = d.dwg2a; m.dbga = d.dbg2a; m.dwgt = d.dwg2t; m.dbgt = d.dbg2t;
  return m;
}

int fusion_fwd(const ser_fusion_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int B = d.B, P = d.P, G = d.G, Din = d.Din;
  const int Dint = d.Din_t > 0 ? d.Din_t : d.Din;      // FusionLayer(audio_dim != text_dim): the pair below becomes two launches
  SER_REQUIRE(B > 0 && d.av && d.tv && d.fused, "fusion_fwd: null tensor");
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);       // proj_a[2] / proj_t[2] (fusion.py:9,12)
  // the two modality branches run the same three GEMMs: each pair is one batched launch when the operands are twins
  SER_TRY(gemm_pair(fwd_args(dt, B, P, Din, d.av, Din, d.w1a, Din, d.b1a, d.ha, P, f, ACT_RELU, nullptr, 0, f),
                    fwd_args(dt, B, P, Dint, d.tv, Dint, d.w1t, Dint, d.b1t, d.ht, P, f, ACT_RELU, nullptr, 0, f), s));
  if (drop.on()) {
    SER_TRY(dropout_apply(d.ha, d.ha, nullptr, f, B, P, with_site(drop, DS_FUS_A), s));
    SER_TRY(dropout_apply(d.ht, d.ht, nullptr, f, B, P, with_site(drop, DS_FUS_T), s));
  }
  SER_TRY(gemm_pair(fwd_args(dt, B, P, P, d.ha, P, d.w2a, P, d.b2a, d.pa, P, f, ACT_NONE, nullptr, 0, f),
                    fwd_args(dt, B, P, P, d.ht, P, d.w2t, P, d.b2t, d.pt, P, f, ACT_NONE, nullptr, 0, f), s));
  SER_TRY(gemm_pair(fwd_args(dt, B, G, P, d.pa, P, d.wg1a, P, d.bg1a, d.ga, G, f, ACT_RELU, nullptr, 0, f),
                    fwd_args(dt, B, G, P, d.pt, P, d.wg1t, P, d.bg1t, d.gt, G, f, ACT_RELU, nullptr, 0, f), s));
  return fusion_mix_fwd(to_mix(d), s);
}

size_t fusion_bwd_ws_bytes(int dtype, int B, int Din, int P, int G) {
  (void)Din;
  const size_t e = esize(dtype);
  return 4 * pad256(static_cast<size_t>(B) * P * e) + 2 * pad256(static_cast<size_t>(B) * G * e) + 4096;
}

int fusion_bwd(const ser_fusion_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int B = d.B, P = d.P, G = d.G, Din = d.Din;
  const int Dint = d.Din_t > 0 ? d.Din_t : d.Din;
  const size_t e = esize(dt);
  Arena ws(d.ws, d.ws_bytes);
  void* dpa = ws.take(static_cast<size_t>(B) * P * e);
  void* dpt = ws.take(static_cast<size_t>(B) * P * e);
  void* dha = ws.take(static_cast<size_t>(B) * P * e);
  void* dht = ws.take(static_cast<size_t>(B) * P * e);
  void* dga = ws.take(static_cast<size_t>(B) * G * e);
  void* dgt = ws.take(static_cast<size_t>(B) * G * e);
  if (!ws.ok) { set_last_error(__FILE__, __LINE__, "fusion_bwd: workspace too small"); return SER_ERR_WORKSPACE; }
  SER_ZERO_UNLESS(d.grads_zeroed, d.dwg2a, sizeof(float) * G);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dwg2t, sizeof(float) * G);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dbg2a, sizeof(float));
  SER_ZERO_UNLESS(d.grads_zeroed, d.dbg2t, sizeof(float));
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  const float hscale = drop.on() ? drop.scale : 1.f;
  MixArgs m = to_mix(d);
  m.dpa = dpa; m.dpt = dpt; m.dga = dga; m.dgt = dgt;
  SER_TRY(fusion_mix_bwd(m, s));
  struct Side { void* dp; void* dg; void* dh; const void* p; const void* h; const void* v; const void* wg1; const void* w2;
                const void* w1; float* dwg1; float* dbg1; float* dw2; float* db2; float* dw1; float* db1; void* dv; };
  const Side sides[2] = {
      {dpa, dga, dha, d.pa, d.ha, d.av, d.wg1a, d.w2a, d.w1a, d.dwg1a, d.dbg1a, d.dw2a, d.db2a, d.dw1a, d.db1a, d.dav},
      {dpt, dgt, dht, d.pt, d.ht, d.tv, d.wg1t, d.w2t, d.w1t, d.dwg1t, d.dbg1t, d.dw2t, d.db2t, d.dw1t, d.db1t, d.dtv},
  };
  {
    const Side& A_ = sides[0];
    const Side& T_ = sides[1];
    const int z = d.grads_zeroed;
    // the three weight-gradient pairs run on the side branch; the dX chain (gate dgrad -> proj[3] dgrad -> proj[0] dgrad)
    // stays on the caller's stream as three consecutive launches
    SideBranch sb(s);
    SER_TRY(sb.fork());
    SER_TRY(gemm_pair(wgrad_args(dt, B, G, P, A_.dg, G, A_.p, P, A_.dwg1, P, A_.dbg1, z),
                      wgrad_args(dt, B, G, P, T_.dg, G, T_.p, P, T_.dwg1, P, T_.dbg1, z), sb.side()));
    SER_TRY(gemm_pair(dgrad_args(dt, B, G, P, A_.dg, G, A_.wg1, P, A_.dp, P, f, nullptr, 0, f, GATE_NONE, A_.dp, P, f),
                      dgrad_args(dt, B, G, P, T_.dg, G, T_.wg1, P, T_.dp, P, f, nullptr, 0, f, GATE_NONE, T_.dp, P, f), s));
    SER_TRY(sb.fork());
    SER_TRY(gemm_pair(wgrad_args(dt, B, P, P, A_.dp, P, A_.h, P, A_.dw2, P, A_.db2, z),
                      wgrad_args(dt, B, P, P, T_.dp, P, T_.h, P, T_.dw2, P, T_.db2, z), sb.side()));
    // h is saved post-dropout: h > 0 exactly where the unit was kept and the ReLU open; the kept units carry 1/(1-p)
    SER_TRY(gemm_pair(dgrad_args(dt, B, P, P, A_.dp, P, A_.w2, P, A_.dh, P, f, A_.h, P, f, GATE_RELU, nullptr, 0, f, hscale),
                      dgrad_args(dt, B, P, P, T_.dp, P, T_.w2, P, T_.dh, P, f, T_.h, P, f, GATE_RELU, nullptr, 0, f, hscale), s));
    SER_TRY(sb.fork());
    SER_TRY(gemm_pair(wgrad_args(dt, B, P, Din, A_.dh, P, A_.v, Din, A_.dw1, Din, A_.db1, z),
                      wgrad_args(dt, B, P, Dint, T_.dh, P, T_.v, Dint, T_.dw1, Dint, T_.db1, z), sb.side()));
    SER_TRY(gemm_pair(dgrad_args(dt, B, P, Din, A_.dh, P, A_.w1, Din, A_.dv, Din, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f),
                      dgrad_args(dt, B, P, Dint, T_.dh, P, T_.w1, Dint, T_.dv, Dint, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f), s));
    SER_TRY(sb.join());
  }
  return SER_OK;
}

// =================================================================================================
// a5 classifier stack + heads
// =================================================================================================

// Arguments of the fused stack kernels (clf_stack.cu) when the per-layer pointers are uniformly strided
// (they are: every module keeps its parameters in one flat buffer, _params.FlatParams).
static bool stack_args(const ser_clf_desc& d, ClfStackArgs& a) {
  const int L = d.L;
  a = ClfStackArgs{};
  a.B = d.B; a.L = L;
  if (d.dtype != DT_BF16 || L < 2) return false;
  static const bool disabled = (getenv("SER_NO_FUSED_CLF") != nullptr);     // A/B switch: per-layer launches instead
  if (disabled) return false;
  auto estride = [](const void* p1, const void* p0, size_t es) {
    return static_cast<long long>((reinterpret_cast<const char*>(p1) - reinterpret_cast<const char*>(p0)) / static_cast<long long>(es));
  };
  a.w1 = d.w1[0]; a.s_w1 = estride(d.w1[1], d.w1[0], 2);
  a.w2 = d.w2[0]; a.s_w2 = estride(d.w2[1], d.w2[0], 2);
  a.b1 = d.b1[0]; a.b2 = d.b2[0]; a.lni_g = d.lni_g[0]; a.lni_b = d.lni_b[0];
  a.s_blk = estride(d.b1[1], d.b1[0], 4);
  a.lno_g = d.lno_g[0]; a.lno_b = d.lno_b[0]; a.s_lno = estride(d.lno_g[1], d.lno_g[0], 4);
  for (int i = 1; i < L; ++i) {
    if (estride(d.w1[i], d.w1[i - 1], 2) != a.s_w1 || estride(d.w2[i], d.w2[i - 1], 2) != a.s_w2) return false;
    if (estride(d.b1[i], d.b1[i - 1], 4) != a.s_blk || estride(d.b2[i], d.b2[i - 1], 4) != a.s_blk ||
        estride(d.lni_g[i], d.lni_g[i - 1], 4) != a.s_blk || estride(d.lni_b[i], d.lni_b[i - 1], 4) != a.s_blk) return false;
    if (estride(d.lno_g[i], d.lno_g[i - 1], 4) != a.s_lno || estride(d.lno_b[i], d.lno_b[i - 1], 4) != a.s_lno) return false;
  }
  a.h = d.h; a.n = d.n; a.r = d.r; a.stats_o = d.stats_o; a.stats_i = d.stats_i;
  a.xchg = d.y;       // the fused kernels keep y in registers; its [L,B,512] fp32 buffer serves as the exchange scratch
  return clf_stack_supported(d.dtype, d.P, L, a);
}
static bool stack_grad_args(const ser_clf_desc& d, ClfStackArgs& a) {
  const int L = d.L;
  auto estride = [](const float* p1, const float* p0) { return static_cast<long long>(p1 - p0); };
  a.dlni_g = d.dlni_g[0]; a.dlni_b = d.dlni_b[0]; a.dlno_g = d.dlno_g[0]; a.dlno_b = d.dlno_b[0];
  for (int i = 1; i < L; ++i) {
    if (estride(d.dlni_g[i], d.dlni_g[i - 1]) != a.s_blk || estride(d.dlni_b[i], d.dlni_b[i - 1]) != a.s_blk) return false;
    if (estride(d.dlno_g[i], d.dlno_g[i - 1]) != a.s_lno || estride(d.dlno_b[i], d.dlno_b[i - 1]) != a.s_lno) return false;
  }
  return true;
}

int clf_fwd(const ser_clf_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int B = d.B, P = d.P, F = d.F, L = d.L;
  const size_t BP = static_cast<size_t>(B) * P;
  SER_REQUIRE(B > 0 && d.x && d.h && d.y && d.n && d.r && d.logits, "clf_fwd: null tensor");
  // input_projection: Linear -> LayerNorm -> ReLU (classifier.py:105-110)
  const int Pin = d.Pin > 0 ? d.Pin : P;             // DeepClassifier(input_dim != base_dim): only this GEMM sees input_dim
  SER_TRY(linear_fwd(dt, B, P, Pin, d.x, Pin, d.w_in, Pin, d.b_in, d.p0, P, 1, ACT_NONE, nullptr, 0, 1, s));
  SER_TRY(layernorm_fwd(d.p0, 1, d.h, 1, nullptr, 1, d.ln_in_g, d.ln_in_b, d.stats0, B, P, 1, s));
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  if (drop.on()) SER_TRY(dropout_apply(d.h, d.h, nullptr, 1, B, P, with_site(drop, DS_CLF_IN), s));   // input_projection[3]
  ClfStackArgs sa;
  const bool fused = stack_args(d, sa);
  sa.drop = drop;
  if (fused) SER_TRY(clf_stack_fwd(sa, s));          // all L blocks in one cluster kernel (clf_stack.cu)
  for (int i = 0; i < L && !fused; ++i) {
    float* hi = d.h + i * BP;
    float* hn = d.h + (i + 1) * BP;
    float* yi = d.y + i * BP;
    void* ni = off(d.n, static_cast<long long>(i) * BP, dt);
    void* ri = off(d.r, static_cast<long long>(i) * BP, dt);
    // outer LayerNorm, then the block's own LayerNorm (classifier.py:207-212, :79-80), one fused pass
    SER_TRY(layernorm2_fwd(hi, yi, ni, f, d.lno_g[i], d.lno_b[i], d.lni_g[i], d.lni_b[i],
                           d.stats_o + static_cast<size_t>(i) * B * 2, d.stats_i + static_cast<size_t>(i) * B * 2, B, P, s));
    SER_TRY(linear_fwd(dt, B, P, P, ni, P, d.w1[i], P, d.b1[i], ri, P, f, ACT_RELU, nullptr, 0, 1, s));
    if (drop.on()) SER_TRY(dropout_apply(ri, ri, nullptr, f, B, P, with_site(drop, DS_CLF_BLOCK0 + 2 * i), s));   // block[3]
    // residual from the OUTER-LN output y; with dropout on (block[5]) it is added by the mask pass instead
    SER_TRY(linear_fwd(dt, B, P, P, ri, P, d.w2[i], P, d.b2[i], hn, P, 1, ACT_NONE, drop.on() ? nullptr : yi, P, 1, s));
    if (drop.on()) SER_TRY(dropout_apply(hn, hn, yi, 1, B, P, with_site(drop, DS_CLF_BLOCK0 + 2 * i + 1), s));
  }
  const float* hL = d.h + static_cast<size_t>(L) * BP;
  SER_TRY(cast_any(hL, 1, d.h_last, f, static_cast<long long>(BP), s));
  // output_projection[0..2]: Linear(512->256) -> LN -> ReLU (classifier.py:215-218)
  if (F % 128 == 0 || dt == DT_F32) {
    SER_TRY(linear_fwd(dt, B, F, P, d.h_last, P, d.w_out, P, d.b_out, d.q, F, 1, ACT_NONE, nullptr, 0, 1, s));
  } else {
    set_last_error(__FILE__, __LINE__, "clf: base_dim/2 must be a multiple of 128 in the bf16 tier");
    return SER_ERR_UNSUPPORTED;
  }
  SER_TRY(layernorm_fwd(d.q, 1, d.f, 1, nullptr, 1, d.ln_out_g, d.ln_out_b, d.stats_q, B, F, 1, s));
  if (drop.on()) SER_TRY(dropout_apply(d.f, d.f, nullptr, 1, B, F, with_site(drop, DS_CLF_OUT), s));   // output_projection[3]
  // heads run in fp32 on the master weights (tiny: C x 256, 64 x 256), one fused launch (heads.cu)
  SER_TRY(heads_fwd(d.f, d.w_c, d.b_c, d.w_u1, d.b_u1, d.w_u2, d.b_u2, d.logits, d.u1, d.unc, B, F, d.C, d.U,
                    with_site(drop, DS_CLF_UNC), s));
  return SER_OK;
}

size_t clf_bwd_ws_bytes(int dtype, int B, int P, int F, int C, int U) {
  (void)C;
  const size_t e = esize(dtype);
  const size_t b = static_cast<size_t>(B);
  const size_t Lmax = 64;          // upper bound on the number of residual blocks covered by this query
  return pad256(b * F * 4) + pad256(b * F * e) + 3 * pad256(b * P * 4) + pad256(b * P * e) + pad256(b * U * 4) +
         pad256(b * 4) + 2 * pad256(Lmax * b * P * e) + 8192;
}

int clf_bwd(const ser_clf_desc& d, cudaStream_t s) {
  const int dt = d.dtype, f = is_f32(dt);
  const int B = d.B, P = d.P, F = d.F, L = d.L, C = d.C, U = d.U;
  const size_t BP = static_cast<size_t>(B) * P;
  const size_t e = esize(dt);
  SER_REQUIRE(d.dlogits != nullptr || d.dunc != nullptr, "clf_bwd: no incoming gradient");
  Arena ws(d.ws, d.ws_bytes);
  SER_REQUIRE(L <= 64, "clf_bwd: at most 64 residual blocks");
  float* df = reinterpret_cast<float*>(ws.take(static_cast<size_t>(B) * F * 4));
  void* dq = ws.take(static_cast<size_t>(B) * F * e);
  float* dh32 = reinterpret_cast<float*>(ws.take(BP * 4));      // gradient of the fp32 residual stream
  float* dh32b = reinterpret_cast<float*>(ws.take(BP * 4));     // ping-pong partner
  float* dn32 = reinterpret_cast<float*>(ws.take(BP * 4));
  void* dp0 = ws.take(BP * e);
  float* du1 = reinterpret_cast<float*>(ws.take(static_cast<size_t>(B) * U * 4));
  float* dsg = reinterpret_cast<float*>(ws.take(static_cast<size_t>(B) * 4));
  // per-block GEMM operands kept for ONE batched weight-gradient launch per weight family after the loop:
  void* dhn_all = ws.take(static_cast<size_t>(L) * BP * e);     // [L,B,P] act: gradient at each block's output
  void* dr_all = ws.take(static_cast<size_t>(L) * BP * e);      // [L,B,P] act: gradient at relu(W1 n + b1)
  if (!ws.ok) { set_last_error(__FILE__, __LINE__, "clf_bwd: workspace too small"); return SER_ERR_WORKSPACE; }

  // ---- heads (fp32): row-wise gradients + every head parameter gradient in two launches (heads.cu) ----
  const DropSpec drop = make_drop(d.drop_seed, d.p_drop, 0);
  const float hscale = drop.on() ? drop.scale : 1.f;
  // weight gradients are leaves of the backward graph: they run on the side branch (common.cuh SideBranch) beside the
  // dX chain -- the heads' and the output projection's beside the 35-block stack kernel (which occupies 32 SMs), the
  // stack's own batched weight gradients beside the input projection's backward
  SideBranch sb(s);
  SER_TRY(heads_bwd(d.dlogits, d.dunc, d.unc, d.u1, d.f, d.w_c, d.w_u1, d.w_u2, df, du1, dsg, d.dw_c, d.db_c, d.dw_u1,
                    d.db_u1, d.dw_u2, d.db_u2, B, F, C, U, with_site(drop, DS_CLF_UNC), s, &sb));
  if (drop.on()) SER_TRY(dropout_apply(df, df, nullptr, 1, B, F, with_site(drop, DS_CLF_OUT), s));
  // ---- output projection: relu(LN(q)) ----
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_out_g, sizeof(float) * F);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_out_b, sizeof(float) * F);
  SER_TRY(layernorm_bwd(df, 1, d.q, 1, d.stats_q, d.ln_out_g, d.ln_out_b, nullptr, 1, dq, f, nullptr, 1, d.dln_out_g,
                        d.dln_out_b, B, F, 1, s));
  SER_TRY(sb.fork());
  SER_TRY(linear_wgrad(dt, B, F, P, dq, F, d.h_last, P, d.dw_out, P, sb.side(), d.db_out, d.grads_zeroed));
  SER_TRY(linear_dgrad(dt, B, F, P, dq, F, d.w_out, P, dh32, P, 1, nullptr, 0, 1, GATE_NONE, nullptr, 0, 1, s));
  ClfStackArgs sa;
  bool fused = stack_args(d, sa) && stack_grad_args(d, sa);
  sa.drop = drop;
  if (fused) {
    // the whole dX chain + LayerNorm parameter gradients in one cluster kernel; it leaves dL/dh_0 in dh32b and the
    // per-block GEMM operands (dh_{i+1}, da_i as bf16) in dhn_all / dr_all for the batched weight gradients below
    sa.dh_in = dh32; sa.dh_out = dh32b; sa.dhn = dhn_all; sa.dr = dr_all;
    SER_TRY(clf_stack_bwd(sa, s));
  } else {
    SER_TRY(cast_any(dh32, 1, off(dhn_all, static_cast<long long>(L - 1) * BP, dt), f, static_cast<long long>(BP), s));
  }
  // ---- 35 residual blocks, last to first: only the serial dX chain lives in the loop ----
  // LayerNorm parameter gradients accumulate with atomics: the per-block dlni / dlno buffers must arrive zeroed
  // (they are slices of the caller's zero-initialised gradient buffer).
  float* cur = fused ? dh32b : dh32;
  float* nxt = fused ? dh32 : dh32b;
  for (int i = L - 1; i >= 0 && !fused; --i) {
    const float* hi = d.h + i * BP;
    const void* ri = off(static_cast<const void*>(d.r), static_cast<long long>(i) * BP, dt);
    void* dhn_i = off(dhn_all, static_cast<long long>(i) * BP, dt);
    void* dr_i = off(dr_all, static_cast<long long>(i) * BP, dt);
    // block[5]: the branch sees mask * dh_{i+1} (this act-dtype copy; the fp32 skip path `cur` stays unmasked);
    // block[3]: r is saved post-dropout, so the ReLU gate is also the keep mask and the kept units carry 1/(1-p)
    if (drop.on()) SER_TRY(dropout_apply(dhn_i, dhn_i, nullptr, f, B, P, with_site(drop, DS_CLF_BLOCK0 + 2 * i + 1), s));
    SER_TRY(linear_dgrad(dt, B, P, P, dhn_i, P, d.w2[i], P, dr_i, P, f, ri, P, f, GATE_RELU, nullptr, 0, 1, s, hscale));
    SER_TRY(linear_dgrad(dt, B, P, P, dr_i, P, d.w1[i], P, dn32, P, 1, nullptr, 0, 1, GATE_NONE, nullptr, 0, 1, s));
    // dy = dh_next (skip) + LN_inner'(dn);  dh_i = LN_outer'(dy): fp32 stream + act-dtype copy for block i-1
    void* dh_prev = (i > 0) ? off(dhn_all, static_cast<long long>(i - 1) * BP, dt) : nullptr;
    SER_TRY(layernorm2_bwd(dn32, cur, hi, d.stats_o + static_cast<size_t>(i) * B * 2,
                           d.stats_i + static_cast<size_t>(i) * B * 2, d.lno_g[i], d.lno_b[i], d.lni_g[i], nxt, dh_prev, f,
                           d.dlni_g[i], d.dlni_b[i], d.dlno_g[i], d.dlno_b[i], B, P, s));
    float* t = cur; cur = nxt; nxt = t;
  }
  // ---- weight / bias gradients of all blocks: batched when the per-block buffers are uniformly strided ----
  bool uniform = L > 1;
  long long sw1 = 0, sw2 = 0, sb1 = 0, sb2 = 0;
  if (uniform) {
    sw1 = d.dw1[1] - d.dw1[0]; sw2 = d.dw2[1] - d.dw2[0]; sb1 = d.db1[1] - d.db1[0]; sb2 = d.db2[1] - d.db2[0];
    for (int i = 1; i < L && uniform; ++i)
      uniform = (d.dw1[i] - d.dw1[i - 1] == sw1) && (d.dw2[i] - d.dw2[i - 1] == sw2) &&
                (d.db1[i] - d.db1[i - 1] == sb1) && (d.db2[i] - d.db2[i - 1] == sb2);
    uniform = uniform && sw1 > 0 && sw2 > 0 && sb1 > 0 && sb2 > 0 && (P % 8 == 0);
  }
  SER_TRY(sb.fork());                                   // (dhn_all / dr_all are complete: the stack kernel has been enqueued)
  cudaStream_t sw = sb.side();
  if (uniform) {
    GemmArgs g;
    g.dtype = dt; g.M = P; g.N = P; g.K = B; g.a_trans = 1; g.b_trans = 1; g.lda = P; g.ldb = P; g.ldc = P; g.c_f32 = 1;
    g.batch = L; g.strideA = static_cast<long long>(BP); g.strideB = static_cast<long long>(BP);
    g.out_zeroed = d.grads_zeroed;
    // bias gradients ride on the weight-gradient GEMMs (row sums of the MN-major dY operand)
    g.A = dhn_all; g.B = d.r; g.C = d.dw2[0]; g.strideC = sw2; g.rowsum = d.db2[0]; g.strideRS = sb2;
    SER_TRY(gemm(g, sw));
    g.A = dr_all; g.B = d.n; g.C = d.dw1[0]; g.strideC = sw1; g.rowsum = d.db1[0]; g.strideRS = sb1;
    SER_TRY(gemm(g, sw));
  } else {
    for (int i = 0; i < L; ++i) {
      const void* dhn_i = off(static_cast<const void*>(dhn_all), static_cast<long long>(i) * BP, dt);
      const void* dr_i = off(static_cast<const void*>(dr_all), static_cast<long long>(i) * BP, dt);
      SER_TRY(linear_wgrad(dt, B, P, P, dhn_i, P, off(static_cast<const void*>(d.r), static_cast<long long>(i) * BP, dt), P, d.dw2[i], P, sw, d.db2[i], d.grads_zeroed));
      SER_TRY(linear_wgrad(dt, B, P, P, dr_i, P, off(static_cast<const void*>(d.n), static_cast<long long>(i) * BP, dt), P, d.dw1[i], P, sw, d.db1[i], d.grads_zeroed));
    }
  }
  // ---- input projection: h0 = relu(LN(p0)) ----
  if (drop.on()) SER_TRY(dropout_apply(cur, cur, nullptr, 1, B, P, with_site(drop, DS_CLF_IN), s));
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_in_g, sizeof(float) * P);
  SER_ZERO_UNLESS(d.grads_zeroed, d.dln_in_b, sizeof(float) * P);
  SER_TRY(layernorm_bwd(cur, 1, d.p0, 1, d.stats0, d.ln_in_g, d.ln_in_b, nullptr, 1, dp0, f, nullptr, 1, d.dln_in_g,
                        d.dln_in_b, B, P, 1, s));
  const int Pin = d.Pin > 0 ? d.Pin : P;
  SER_TRY(sb.fork());
  SER_TRY(linear_wgrad(dt, B, P, Pin, dp0, P, d.x, Pin, d.dw_in, Pin, sb.side(), d.db_in, d.grads_zeroed));
  if (d.dx != nullptr)
    SER_TRY(linear_dgrad(dt, B, P, Pin, dp0, P, d.w_in, Pin, d.dx, Pin, f, nullptr, 0, f, GATE_NONE, nullptr, 0, f, s));
  SER_TRY(sb.join());
  return SER_OK;
}

}  // namespace ser

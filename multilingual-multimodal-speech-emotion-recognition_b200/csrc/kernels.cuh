// Internal launch functions shared by the module-level orchestration (head_modules.cu) and api.cu.
// Every function enqueues on `stream`, returns an SER_* code and never synchronises.
#pragma once
#include "common.cuh"
#include "dropout.cuh"

namespace ser {

// ---- elementwise.cu ------------------------------------------------------------------------------
int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t s);
int cast_any(const void* src, int src_f32, void* dst, int dst_f32, long long n, cudaStream_t s);
// zero `bytes` bytes with a kernel launch (falls back to cudaMemsetAsync for unaligned regions)
int zero_async(void* p, size_t bytes, cudaStream_t s);
int unpack_frames(const void* packed, const long long* offsets, void* out, float* mask, int B, int T, int D, int elem_bytes,
                  cudaStream_t s);
// fp32 -> bf16 of n <= 16 buffers in one launch; src / dst / counts are HOST arrays (counts multiples of 8)
int cast_multi(int n, const void* const* src, void* const* dst, const long long* counts, cudaStream_t s);
// out[n] (fp32) = sum_m X[m, n]           X: [M, N] with leading dim ld, fp32 or bf16
int colsum(const void* X, int x_f32, long long ld, int M, int N, float* out, cudaStream_t s);
// y = [relu](LayerNorm(x)); stats[m] = {mean, rstd}.  N % 8 == 0, N <= 1024.  eps = 1e-5.
int layernorm_fwd(const void* x, int x_f32, void* y, int y_f32, void* y2, int y2_f32, const float* gamma,
                  const float* beta, float* stats, int M, int N, int relu, cudaStream_t s, const void* res = nullptr,
                  void* xout = nullptr, const DropSpec* drop = nullptr);
// res / xout / drop (optional): normalise res + dropout(x) instead of x and save that row to xout (x's dtype)
// dx = LN'(dy) [+ add];  dgamma/dbeta accumulate (+=) into fp32 [N] buffers (caller zeroes them).
// With relu != 0 the op was y = relu(LN(x)) and dy is masked by (LN(x) > 0) first.
int layernorm_bwd(const void* dy, int dy_f32, const void* x, int x_f32, const float* stats, const float* gamma,
                  const float* beta, const void* add, int add_f32, void* dx, int dx_f32, void* dx2, int dx2_f32,
                  float* dgamma, float* dbeta, int M, int N, int relu, cudaStream_t s, void* dxm = nullptr,
                  const DropSpec* drop = nullptr);
// dxm / drop (optional): also write mask * dx, the gradient entering a dropout-ed branch (dropout.cuh)
// single-pass variant for bf16 rows of 256 / 512 / 768 columns (layernorm_fused.cu); layernorm_bwd dispatches to it
bool layernorm_bwd_fused_ok(int dy_f32, int x_f32, int dx_f32, const void* add, const void* dx2, const float* dgamma,
                            int M, int N, int relu);
int layernorm_bwd_fused(const void* dy, const void* x, const float* stats, const float* gamma, void* dx, void* dxm,
                        const DropSpec& drop, float* dgamma, float* dbeta, int M, int N, cudaStream_t s);
int colsum_batched(const void* X, int x_f32, long long ld, int M, int N, float* out, int batch, long long strideX,
                   long long strideOut, cudaStream_t s);
// classifier block prologue / its backward: y = LN_outer(h), n = LN_inner(y) in one pass (fp32 stream)
int layernorm2_fwd(const float* h, float* y, void* n, int n_f32, const float* go, const float* bo, const float* gi,
                   const float* bi, float* stats_o, float* stats_i, int M, int N, cudaStream_t s);
int layernorm2_bwd(const float* dn, const float* dskip, const float* h, const float* stats_o, const float* stats_i,
                   const float* go, const float* bo, const float* gi, float* dh, void* dh2, int dh2_f32, float* dgi,
                   float* dbi, float* dgo, float* dbo, int M, int N, cudaStream_t s);
// ds = dy * y * (1 - y)    (sigmoid backward, fp32)
int sigmoid_bwd(const float* dy, const float* y, float* ds, long long n, cudaStream_t s);

// ---- attention.cu --------------------------------------------------------------------------------
// Masked multi-head cross attention, heads packed along the feature axis (head h = columns [h*dh, (h+1)*dh)).
//   Q: [B*Tq, *] ldq,  K/V: [B*Tk, *] ldk/ldv,  kmask: [B, Tk] float (0 = padded key) or NULL
//   O: [B*Tq, H*dh] ldo,  lse: [B, H, Tq] fp32 (log-sum-exp of the scaled masked scores)
struct AttnArgs {
  int dtype;
  int B, H, Tq, Tk, dh;
  const void* Q; long long ldq;
  const void* K; long long ldk;
  const void* V; long long ldv;
  const float* kmask;
  void* O; long long ldo;
  float* lse;
  float scale;
  // backward only
  const void* dO; long long lddo;
  void* dQ; long long lddq;
  void* dK; long long lddk;
  void* dV; long long lddv;
  float* delta;             // [B, H, Tq] scratch: rowsum(dO * O)
  // dropout on the attention weights (nn.MultiheadAttention(dropout=p)): site [B*H*Tq, Tk]; off by default
  DropSpec drop;
  int impl = 0;             // bf16 tier: 0 = choose (tcgen05 kernels when the shape qualifies), 1 = mma.sync, 2 = tcgen05
  // optional keep bits of the attention-weight dropout, [B*H*Tq, ceil(Tk/32)] words (tcgen05 kernels: the forward writes
  // them, the backward kernels read them instead of re-hashing); nullptr = regenerate from the seed
  unsigned* keep_bits = nullptr;
};
int attention_fwd(const AttnArgs& a, cudaStream_t s);
int attention_bwd(const AttnArgs& a, cudaStream_t s);
// bf16 tensor-core variants (attention_tc.cu); attention_fwd / attention_bwd dispatch to them for dtype == bf16
int attention_fwd_tc(const AttnArgs& a, cudaStream_t s);
int attention_bwd_tc(const AttnArgs& a, cudaStream_t s);
// tcgen05 / TMEM / TMA variants (attention_tc5.cu): head dim 32, even head count, TMA-addressable operands
bool attention_tc5_supported(const AttnArgs& a);
int attention_fwd_tc5(const AttnArgs& a, cudaStream_t s);
int attention_bwd_tc5(const AttnArgs& a, cudaStream_t s);

// ---- clf_stack.cu --------------------------------------------------------------------------------
// The whole residual stack (classifier.py:207-212) as one cluster kernel per direction, bf16 tier, base_dim 512.
// Per-layer parameters are addressed as (layer-0 pointer, element stride between layers).
struct ClfStackArgs {
  int B, L;
  const void* w1; long long s_w1; const void* w2; long long s_w2;      // bf16 [512,512] per layer
  const float* b1; const float* b2; const float* lni_g; const float* lni_b; long long s_blk;
  const float* lno_g; const float* lno_b; long long s_lno;
  float* h; void* n; void* r; float* stats_o; float* stats_i;           // saved activations (see ser_clf_desc)
  void* xchg;                                                           // scratch >= ceil(B/128) * 256 KB (the saved-y buffer)
  const float* dh_in; float* dh_out; void* dhn; void* dr;               // backward
  float* dlni_g; float* dlni_b; float* dlno_g; float* dlno_b;
  DropSpec drop;                                                        // block[3] / block[5] dropout, off by default
};
bool clf_stack_supported(int dtype, int P, int L, const ClfStackArgs& a);
int clf_stack_fwd(const ClfStackArgs& a, cudaStream_t s);
int clf_stack_bwd(const ClfStackArgs& a, cudaStream_t s);

// ---- fold.cu -------------------------------------------------------------------------------------
// element-wise glue of the folded Linear chains of CrossModalAttention (bf16 tier; see fold.cu)
int fold_assemble(const void* win_a, const void* win_t, void* wbd_a, void* wbd_t, int S, cudaStream_t s);
struct FoldBiasArgs {
  int S, D;
  const void* win_a; const void* win_t; const float* bin_a; const float* bin_t; const float* bqkv_a; const float* bqkv_t;
  const void* wout_a; const void* wout_t; const float* bo_a; const float* bo_t; const float* bout_a; const float* bout_t;
  float* bc_a; float* bc_t; float* bz_a; float* bz_t;
};
int fold_bias_fwd(const FoldBiasArgs& a, cudaStream_t s);
struct FoldBwdArgs {
  int S, D;
  const void* win_a; const void* win_t; const void* wout_a; const void* wout_t;
  const float* bqkv_a; const float* bqkv_t; const float* bo_a; const float* bo_t;
  const float* dwbd_a; const float* dwbd_t; const float* dbc_a; const float* dbc_t; const float* dbz_a; const float* dbz_t;
  float* dwin_a; float* dwin_t; float* dbin_a; float* dbin_t; float* dbqkv_a; float* dbqkv_t;
  float* dbo_a; float* dbo_t; float* dwout_a; float* dwout_t; float* dbout_a; float* dbout_t;
};
int fold_bwd_glue(const FoldBwdArgs& a, cudaStream_t s);

// ---- heads.cu ------------------------------------------------------------------------------------
// logits + uncertainty head on the fp32 penultimate features (classifier.py:192-198,224,229); u1 / unc may be NULL
int heads_fwd(const float* f, const float* w_c, const float* b_c, const float* w_u1, const float* b_u1,
              const float* w_u2, const float* b_u2, float* logits, float* u1, float* unc, int B, int F, int C, int U,
              const DropSpec& drop, cudaStream_t s);
// dlogits / dunc may be NULL (treated as zero); every parameter gradient is written (not accumulated)
int heads_bwd(const float* dlogits, const float* dunc, const float* unc, const float* u1, const float* f,
              const float* w_c, const float* w_u1, const float* w_u2, float* df, float* du1, float* dsg, float* dw_c,
              float* db_c, float* dw_u1, float* db_u1, float* dw_u2, float* db_u2, int B, int F, int C, int U,
              const DropSpec& drop, cudaStream_t s, SideBranch* sb = nullptr);

// ---- pooling.cu ----------------------------------------------------------------------------------
// Attentive statistics pooling, everything after the 768->128 tanh GEMM (src/models/pooling.py:21-28).
struct AspArgs {
  int dtype;
  int B, T, D, Hd;           // D = 768, Hd = 128
  const void* x;             // [B*T, D]
  const void* u;             // [B*T, Hd] tanh(W1 x + b1)
  const float* w2; const float* b2;   // [Hd], [1]
  const float* mask;         // [B, T] or NULL
  float* e;                  // [B, T] scratch scores
  float* alpha;              // [B, T] softmax weights (saved)
  void* out;                 // [B, 2*D] mean | std
  int out_f32;
  // backward
  const void* dout; int dout_f32;   // [B, 2*D]
  void* dx;                  // [B*T, D]  (statistics part; the scorer GEMM adds its part afterwards)
  float* dalpha;             // [B, T] scratch (zeroed inside)
  void* dpre;                // [B*T, Hd] gradient at the scorer pre-activation
  float* dw2; float* db2;    // [Hd], [1] accumulate (+=)
};
int asp_fwd(const AspArgs& a, cudaStream_t s);
int asp_bwd(const AspArgs& a, cudaStream_t s);

// ---- fusion mix (fusion.py:21-25) ----------------------------------------------------------------
struct MixArgs {
  int dtype; int B, P, G;    // P = 512, G = 256
  const void* pa; const void* pt;        // [B, P]
  const void* ga; const void* gt;        // [B, G] relu(gate hidden)
  const float* wga; const float* bga;    // [G], [1]
  const float* wgt; const float* bgt;
  float* gates;              // [B, 2] sigmoid outputs (wa, wt), saved
  void* fused;               // [B, P]
  // backward
  const void* dfused;        // [B, P]
  void* dpa; void* dpt;      // [B, P] direct part (gate path added by the GEMM afterwards)
  void* dga; void* dgt;      // [B, G] gradient at the gate hidden pre-activation (already relu-masked)
  float* dwga; float* dbga; float* dwgt; float* dbgt;   // accumulate (+=)
};
int fusion_mix_fwd(const MixArgs& a, cudaStream_t s);
int fusion_mix_bwd(const MixArgs& a, cudaStream_t s);

// ---- loss.cu -------------------------------------------------------------------------------------
// Training loss terms of src/train.py:151-168 (a9, a10, a8, a11).  sums[] layout:
enum : int {
  LS_CE = 0, LS_FOCAL = 1, LS_UNC = 2 /* sum(unc) */, LS_CORRECT = 3 /* sum(correct) */, LS_POS = 4, LS_NEG = 5,
  LS_COUNT = 8
};
struct LossArgs {
  int B, C, D;                           // D = prototype dim (512)
  long long B_global;                    // divisor of every mean (== B on one GPU)
  const float* logits;                   // [B, C] fp32
  const float* unc;                      // [B] fp32 or NULL
  const void* emb; int emb_f32;          // [B, D] or NULL (prototype term off)
  const float* protos;                   // [C, D]
  const long long* labels;               // [B] int64
  const float* counts;                   // [C] GLOBAL per-class counts as float, or NULL -> computed from labels
  float smoothing, beta, gamma, margin;
  int focal_use_weights;
  float* class_w;                        // [C] scratch/out: normalised class weights
  float* sums;                           // [LS_COUNT] raw local sums (fwd out; bwd in, possibly all-reduced)
  // backward: loss = w_ce*ce + w_focal*focal + w_unc*mean(unc)*mean(correct) + w_proto*proto, scaled by gscale
  float w_ce, w_focal, w_unc, w_proto;
  float* dlogits;                        // [B, C]
  float* dunc;                           // [B]
  void* demb; int demb_f32;              // [B, D]
  float* dprotos;                        // [C, D] accumulate (+=)
};
int loss_fwd(const LossArgs& a, cudaStream_t s);
int loss_bwd(const LossArgs& a, cudaStream_t s);
// same, with every gradient multiplied by the device scalar *gscale (autograd's grad_output); NULL = 1
int loss_bwd_scaled(const LossArgs& a, const float* gscale, cudaStream_t s);
// terms[6] = {ce, focal, unc_loss, proto, weighted total, accuracy}; non-finite terms -> 0
int loss_finalize(const float* sums, long long B_global, float margin, float w_ce, float w_focal, float w_unc,
                  float w_proto, int have_proto, float* terms, cudaStream_t s);

// ---- supcon.cu -----------------------------------------------------------------------------------
// SupConLoss (losses.py:67-88): loss[0] = scalar; backward recomputes the similarities; gscale = device scalar or NULL
size_t supcon_ws_bytes(int B, int D);
int supcon_fwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, float* loss, void* ws,
               size_t ws_bytes, cudaStream_t s);
int supcon_bwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, const float* gscale,
               void* df, int df_f32, void* ws, size_t ws_bytes, cudaStream_t s);

// ---- optim.cu ------------------------------------------------------------------------------------
// torch.optim.AdamW step over n fp32 tensors (host pointer tables); gscale: optional device scalar multiplied into g
// AMP (optional): grad_scale divides the gradients, found_inf != 0 skips the update, step_dev = device step count
int adamw_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* counts,
                float lr, float beta1, float beta2, float eps, float weight_decay, int step, const float* gscale,
                cudaStream_t s, const float* grad_scale = nullptr, const float* found_inf = nullptr,
                const float* step_dev = nullptr);
// coef[0] = min(1, max_norm / (||g||_2 + 1e-6)) over all tensors (clip_grad_norm_), norm_out[0] = ||g||_2 (optional)
int grad_clip_coef(int n, const float* const* g, const long long* counts, float max_norm, float* scratch, float* coef,
                   float* norm_out, cudaStream_t s);

// ---- eval.cu -------------------------------------------------------------------------------------
// OpenMax re-scaling (classifier.py:240-275): logits_out = logits * (u > 0.3 ? 1 - 0.8u : 1)
int late_ood(const float* logits, const void* feats, int feats_f32, const float* prototypes, const float* covariances,
             const float* temperature, const float* mix, float* distances, float* scores, int B, int C, int D,
             cudaStream_t s);
int openmax_fwd(const float* feats, const float* logits, const float* act_vecs, const float* w_alpha,
                const float* w_beta, const float* w_tau, float* out, int B, int C, int F, cudaStream_t s);
// TTA view mean + temperature + softmax / argmax / energy (eval.py:186-206, utils.py:12-14)
int eval_post(const float* logits_views, int V, int B, int C, float temperature, float* mean_logits, float* probs,
              long long* preds, float* energy, cudaStream_t s);
// err[t] = mean_b | max softmax(logits/T_t) - [argmax == label] |   (eval.py:48-67)
int temperature_sweep(const float* logits, const long long* labels, int B, int C, const float* temps, int nT,
                      float* err, cudaStream_t s);

}  // namespace ser

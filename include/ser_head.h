/*
 * ser_head.h -- C-ABI of the B200-native fusion head (libser_head.so).
 *
 * The reference (kananmittal/Multilingual-Multimodal-Speech-Emotion-Recognition) has no FFI or
 * operator registry: the boundary of the hot path is the Python class surface of
 * src/models/{cross_attention,pooling,fusion,classifier,prototypes,losses}.py plus the `adapter`
 * attribute of the two encoders.  Each entry point below is what the
 * torch.autograd.Function behind one of those classes calls; the reference interface it replaces is
 * cited as file:line (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - return value 0 = ok, non-zero = error (see SER_ERR_*); ser_last_error() gives the message;
 *   - no allocation inside: scratch comes from the `ws` buffer whose size ser_*_ws_bytes() reports;
 *   - `dtype` selects the tier: SER_F32 (CUDA-core fp32, 1e-4 parity tier) or SER_BF16 (tcgen05
 *     tensor cores, bf16 operands / fp32 accumulate).  Parameters and parameter gradients are always
 *     fp32 (master copies owned by the Python nn.Modules); "act" tensors have the tier's dtype;
 *   - matrices are row-major and dense unless a leading dimension is given.
 */
#ifndef SER_HEAD_H_
#define SER_HEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SER_F32 0
#define SER_BF16 1

#define SER_OK 0
#define SER_ERR_CUDA 1
#define SER_ERR_ARG 2
#define SER_ERR_UNSUPPORTED 3
#define SER_ERR_WORKSPACE 4

/* ---- library ------------------------------------------------------------------------------- */
int ser_version(void);
const char* ser_last_error(void);
int ser_sm_count(void);
/* leave `n` SMs out of the persistent GEMM grids (for the NCCL kernels of a data-parallel step); 0 = use them all */
int ser_set_reserved_sms(int n);
/* number of CUDA kernels this library has launched so far in this process                       */
long long ser_launch_count(void);
/* sizeof() of descriptor `id` as compiled (0 gemm, 1 adapter, 2 xattn, 3 asp, 4 fusion, 5 clf, 6 loss, 7 featfuse, 8 attn):
 * lets a foreign-language binding verify its struct layout at load time                           */
int ser_desc_size(int id);

/* ---- opt-in profiler: CUDA-event timing per kernel family on the launching stream ------------- */
int ser_prof_enable(int on);
/* device-synchronises; writes "<family> <launches> <total_ms> <alg_flops> <alg_bytes>\n" per family       */
int ser_prof_report(char* buf, int cap);
/* keeps `stream` busy for `microseconds` (one spinning thread).  The profiling pass enqueues it in front of a step of
 * eager launches so that the host runs ahead of the device and the per-kernel event intervals measure kernels that run
 * back to back, as they do inside the replayed CUDA graph, instead of the host's launch latency between them.  */
int ser_prof_stall(double microseconds, void* stream);
/* n profiled launches of an empty kernel, recorded as family "prof_null": the floor of one event interval          */
int ser_prof_null(int n, void* stream);

/* ---- generic fused GEMM (building block; also exported for tests and micro-benchmarks) -------
 * C[M,N] = epilogue(alpha * op(A) op(B)^T), see csrc/common.cuh GemmArgs.  Replaces every nn.Linear
 * call on the path (e.g. cross_attention.py:38-40, classifier.py:81-85).                         */
typedef struct ser_gemm_desc {
  int dtype;              /* SER_F32 | SER_BF16 : dtype of A and B                               */
  int M, N, K;
  const void* A; long long lda; int a_trans;   /* a_trans=0: [M,K] row-major; 1: stored [K,M]    */
  const void* B; long long ldb; int b_trans;   /* b_trans=0: [N,K] row-major (nn.Linear weight)  */
  void* C; long long ldc; int c_f32;
  const float* bias;                           /* [N] fp32 or NULL                               */
  const void* R; long long ldr; int r_f32;     /* residual added after the activation, or NULL   */
  const void* G; long long ldg; int g_f32; int gate_mode; /* 0 none, 1 relu'(G), 2 tanh'(G)       */
  int act;                                     /* 0 none, 1 relu, 2 tanh, 3 sigmoid              */
  int accumulate;                              /* C += result (fp32 C only)                      */
  float alpha;
  int splits;                                  /* split-K factor, 0 = auto                       */
  float* rowsum;                               /* optional, a_trans = 1 only: rowsum[m] = sum_k op(A)[m,k] -- the bias
                                                  gradient when A = dY stored [tokens, features]; overwritten; NULL = off */
} ser_gemm_desc;
int ser_gemm(const ser_gemm_desc* d, void* stream);


/* ---- utilities ------------------------------------------------------------------------------ */
/* fp32 -> bf16 parameter copy (the Python modules keep fp32 masters; bf16 tier consumes copies)    */
int ser_cast(const void* src, int src_f32, void* dst, int dst_f32, long long n, void* stream);
/* the same for up to 16 buffers in ONE launch, fp32 -> bf16 (all modules of a head before its forward);
 * src / dst / counts are HOST arrays of length n, counts multiples of 8 elements                      */
int ser_cast_multi(int n, const void* const* src, void* const* dst, const long long* counts, void* stream);

/* packed valid frames -> zero-padded batch + mask: out[b,t,:] = t < len_b ? packed[offsets[b] + t, :] : 0 and
 * mask[b,t] = t < len_b (mask may be NULL), len_b = offsets[b+1] - offsets[b]; offsets = B + 1 int64 values in device
 * memory.  Replaces the zero-padding of the encoders' batch assembly (src/models/audio_encoder.py:140-163,
 * src/models/text_encoder.py:75-78) when the hidden states arrive from the host: only valid frames cross PCIe.   */
int ser_unpack_frames(const void* packed, const long long* offsets, void* out, float* mask, int B, int T, int D,
                      int elem_bytes, void* stream);

/* row/column primitives behind the individually callable children of the classifier
 * (nn.LayerNorm / bias-gradient column sums; src/train.py:221-236 calls those children one by one)  */
int ser_layernorm_fwd(const void* x, int x_f32, void* y, int y_f32, const float* gamma, const float* beta,
                      float* stats, int M, int N, int relu, void* stream);
int ser_layernorm_bwd(const void* dy, int dy_f32, const void* x, int x_f32, const float* stats, const float* gamma,
                      const float* beta, void* dx, int dx_f32, float* dgamma, float* dbeta, int M, int N, int relu,
                      void* stream);            /* dgamma / dbeta are overwritten */
int ser_colsum(const void* X, int x_f32, long long ld, int M, int N, float* out, void* stream);

/* ---- dropout ------------------------------------------------------------------------------------
 * Every nn.Dropout of the path (cross_attention.py:18,25,43,51; fusion.py:9,12; classifier.py:83,85,109,127,195) is
 * applied inside the kernels from a counter-based mask: a pure function of (*drop_seed, site, row, col), see
 * csrc/dropout.cuh.  A module descriptor switches it on with p_drop > 0 and a non-NULL drop_seed (DEVICE pointer to
 * one uint64; forward and backward of the same step must see the same value).  The random stream is not torch's
 * Philox stream; keep probability and the 1/(1-p) scaling are torch.nn.functional.dropout's.
 * ser_dropout_mask writes the multipliers (0 or 1/(1-p)) the kernels apply at `site` for a [rows, cols] tensor:
 * the parity tests hand them to the CPU oracle.  Sites: attention weights [B*H*Tq, Tk]; everything else [rows, dim]. */
#define SER_DS_XA_PROB_A 1   /* attn_a weights, rows = (b*H + h)*Ta + i, cols = Tt */
#define SER_DS_XA_PROB_T 2
#define SER_DS_XA_RES_A 3    /* self.dropout(a_out) [B*Ta, D] */
#define SER_DS_XA_RES_T 4
#define SER_DS_FUS_A 5       /* proj_a[2] [B, P] */
#define SER_DS_FUS_T 6
#define SER_DS_CLF_IN 7      /* input_projection[3] [B, P] */
#define SER_DS_CLF_OUT 8     /* output_projection[3] [B, F] */
#define SER_DS_CLF_UNC 9     /* uncertainty_head[2] [B, U] */
#define SER_DS_FEAT 10       /* utterance-feature fusion output [B*T, D] (ser_featfuse_*) */
#define SER_DS_CLF_BLOCK0 16 /* residual block l: 16 + 2l = block[3] (after ReLU), 16 + 2l + 1 = block[5] */
int ser_dropout_mask(const unsigned long long* seed, int site, float p, long long rows, int cols, float* out,
                     void* stream);

/* ---- a1: bottleneck adapter  (src/models/audio_encoder.py:19-21,112; text_encoder.py:17-19,57) --
 * y = x + W2 relu(W1 x + b1) + b2.   "act" = tensor of the tier's dtype; weights w* are act-dtype
 * copies of the nn.Linear weights ([out,in] row-major); biases and all gradients are fp32.         */
typedef struct ser_adapter_desc {
  int dtype; int M; int D; int S;          /* tokens, 768, 256                                    */
  int add_residual;                        /* 1: y = x + branch(x) (fused); 0: y = branch(x) only  */
  const void* x;                           /* [M,D] act                                           */
  const void* w1; const float* b1;         /* [S,D], [S]                                          */
  const void* w2; const float* b2;         /* [D,S], [D]                                          */
  void* h;                                 /* [M,S] act, saved for backward                       */
  void* y;                                 /* [M,D] act, output                                   */
  /* backward */
  const void* dy;                          /* [M,D] act                                           */
  void* dh;                                /* [M,S] act scratch                                   */
  void* dx;                                /* [M,D] act or NULL (frozen-encoder input)            */
  float* dw1; float* db1; float* dw2; float* db2;   /* overwritten                                */
  int grads_zeroed;                        /* 1: the caller hands in zero-filled parameter-gradient buffers (one
                                              allocation per module in the Python binding), so the library skips its own
                                              zeroing of split-K / atomically accumulated outputs; same field in the
                                              descriptors below                                                       */
} ser_adapter_desc;
int ser_adapter_fwd(const ser_adapter_desc* d, void* stream);
int ser_adapter_bwd(const ser_adapter_desc* d, void* stream);

/* ---- a2: CrossModalAttention.forward  (src/models/cross_attention.py:32-53) ---------------------
 * Both directions.  Weight packing (done once per optimiser step by the Python module):
 *   wqkv_a = rows [q_a ; k_a ; v_a] ([3S,D]),  wqkv_t = [q_t ; k_t ; v_t];
 *   win_a / win_t = attn_a / attn_t .in_proj_weight ([3S,S], rows q|k|v); wo_* = out_proj; wout_* = out_a/out_t.
 * Saved activations: qkv_* [M,3S]; p_a = [Qa' | Ka' | Va'] where Qa' feeds attn_a and Ka',Va' feed attn_t
 * (p_t likewise); ctx_* [M,S]; lse_* [B,H,T] fp32; o_* [M,S]; z_* [M,D] pre-LayerNorm; stats_* [M,2].
 * Masks are float, 0 = padded (cross_attention.py:34-35), or NULL.                                 */
typedef struct ser_xattn_desc {
  int dtype; int B, Ta, Tt; int D, S, H;
  const void* a; const void* t;
  const float* a_mask; const float* t_mask;
  const void* wqkv_a; const float* bqkv_a; const void* wqkv_t; const float* bqkv_t;
  const void* win_a; const float* bin_a; const void* win_t; const float* bin_t;
  const void* wo_a; const float* bo_a; const void* wo_t; const float* bo_t;
  const void* wout_a; const float* bout_a; const void* wout_t; const float* bout_t;
  const float* ln_a_g; const float* ln_a_b; const float* ln_t_g; const float* ln_t_b;
  void* qkv_a; void* qkv_t; void* p_a; void* p_t; void* ctx_a; void* ctx_t;
  float* lse_a; float* lse_t; void* o_a; void* o_t; void* z_a; void* z_t; float* stats_a; float* stats_t;
  void* enh_a; void* enh_t;                /* outputs [B*Ta,D], [B*Tt,D]                          */
  /* bf16 tier, optional (NULL = unfolded path): storage for the folded Linear chains (outer q/k/v o MHA
   * in-projection, MHA out_proj o out_a/out_t), written by fwd and read by bwd:
   * fold_w [2*9*S*S + 2*3*S*D + 2*D*S] act elements, fold_b [2*3*S + 2*D] fp32.  With folding on, qkv_* and
   * o_* are not touched.                                                                             */
  void* fold_w; float* fold_b;
  float p_drop; const unsigned long long* drop_seed;   /* attention-weight + residual dropout; 0 / NULL = off */
  /* optional, bf16 tier with dropout on: keep decisions of the attention-weight dropout, one bit per (b, h, query, key),
   * keep_a [B*H*Ta, ceil(Tt/32)] (audio queries) and keep_t [B*H*Tt, ceil(Ta/32)] uint32 words.  The forward kernel
   * writes them while it applies the mask, the two backward kernels read them instead of re-hashing every element
   * (the hash was > 50 % of their instruction stream).  NULL = regenerate from the seed.                              */
  unsigned int* keep_a; unsigned int* keep_t;
  /* forward only, inference: non-zero = p_t (and qkv_t, or fold_w / fold_b on the folded path) still hold the
   * projections of THIS text batch under THESE weights from an earlier ser_xattn_fwd call with the same buffers --
   * the text-side projection GEMMs and the weight folding are skipped.  Test-time augmentation evaluates V audio views
   * of one utterance against the same text (src/eval.py:186-190): 4 of 5 calls take this path.                  */
  int reuse_text;
  /* backward */
  const void* d_enh_a; const void* d_enh_t;
  void* da; void* dt;                      /* [M,D] act gradients w.r.t. the inputs               */
  float* dwqkv_a; float* dbqkv_a; float* dwqkv_t; float* dbqkv_t;
  float* dwin_a; float* dbin_a; float* dwin_t; float* dbin_t;
  float* dwo_a; float* dbo_a; float* dwo_t; float* dbo_t;
  float* dwout_a; float* dbout_a; float* dwout_t; float* dbout_t;
  float* dln_a_g; float* dln_a_b; float* dln_t_g; float* dln_t_b;
  void* ws; size_t ws_bytes;               /* backward scratch, ser_xattn_bwd_ws_bytes()          */
  int grads_zeroed;
  /* CrossModalAttention(audio_dim != text_dim) (cross_attention.py:7-30): width of the TEXT sequence and of every
   * text-side tensor (t, wqkv_t [3S,Dt], wout_t [Dt,S], ln_t_*, z_t, enh_t, dt, ...); 0 = D.  D is then the audio
   * width.  Unequal widths take the unfolded path (the audio / text GEMM twins are no longer one batched launch);
   * size the backward scratch with ser_xattn_bwd_ws_bytes(max(D, Dt)).                                        */
  int Dt;
} ser_xattn_desc;
size_t ser_xattn_bwd_ws_bytes(int dtype, int B, int Ta, int Tt, int D, int S, int H);
/* 1 when ser_xattn_fwd / _bwd take the folded path for this (dtype, D, S) PROVIDED fold_w / fold_b are given; the
 * caller sizes its buffers from this answer: folded -> fold_w / fold_b required, qkv_* / o_* may be NULL;
 * not folded -> qkv_* / o_* must be full size ([M,3S] / [M,S]), fold_* are ignored                       */
int ser_xattn_folded(int dtype, int D, int S);
int ser_xattn_fwd(const ser_xattn_desc* d, void* stream);
int ser_xattn_bwd(const ser_xattn_desc* d, void* stream);

/* ---- attention core of a2 (building block, like ser_gemm; also what the unit tests and micro-benchmarks call) ------
 * Masked multi-head attention of nn.MultiheadAttention's math path (src/models/cross_attention.py:18,25,41,49;
 * torch/nn/functional.py:6609-6645) on already projected, head-packed operands: head h = columns [h*dh, (h+1)*dh).
 *   Q [B*Tq, H*dh] (row pitch ldq), K / V [B*Tk, H*dh]; kmask [B,Tk] float, 0 = padded key, or NULL;
 *   O [B*Tq, H*dh]; lse [B,H,Tq] fp32 = log-sum-exp of the scaled masked scores (saved for the backward);
 *   dropout on the attention weights: site rows = (b*H + h)*Tq + i, cols = Tk.
 * Backward: dQ / dK / dV from dO (delta [B,H,Tq] fp32 scratch).  impl: 0 = choose, 1 = mma.sync kernels,
 * 2 = tcgen05 / TMEM / TMA kernels (bf16 tier, dh = 32, even H; SER_ERR_UNSUPPORTED otherwise).                  */
typedef struct ser_attn_desc {
  int dtype; int B, H, Tq, Tk, dh;
  const void* Q; long long ldq; const void* K; long long ldk; const void* V; long long ldv;
  const float* kmask;
  void* O; long long ldo; float* lse;
  float scale;                             /* 1 / sqrt(dh) in the reference                        */
  float p_drop; const unsigned long long* drop_seed; int drop_site;
  int impl;
  /* backward */
  const void* dO; long long lddo;
  void* dQ; long long lddq; void* dK; long long lddk; void* dV; long long lddv;
  float* delta;
  /* optional (tcgen05 kernels, dropout on): keep bits [B*H*Tq, ceil(Tk/32)] uint32, bit k%32 of word k/32 of row
   * (b*H + h)*Tq + i = "weight (i, k) is kept"; written by the forward, read by the backward; NULL = re-hash       */
  unsigned int* keep_bits;
} ser_attn_desc;
int ser_attention_fwd(const ser_attn_desc* d, void* stream);
int ser_attention_bwd(const ser_attn_desc* d, void* stream);

/* ---- f1: per-utterance feature fusion between the adapter and cross attention ------------------
 * (AudioEncoder.quality_fusion / conditioning_fusion / combined_fusion,
 *  src/models/audio_encoder.py:29-52 applied :114-138, and TextEncoder.asr_fusion, text_encoder.py:26-30,60-73)
 *   y[u,t,:] = dropout(relu(W [x[u,t,:] ; f[u,:]] + b)),   W [D, D+F] = nn.Linear(hid + F, hid).weight (fp32 master)
 * The F features are constant over the frames of utterance u, so the concatenation is never built:
 *   c[u,:] = W[:, D:] f[u] + b   (B x D, fp32),     y = dropout(relu(x W[:, :D]^T + c[u]))
 * wx is the packed copy of W[:, :D] in the tier's dtype (nn.Linear's row pitch D+F is not TMA-addressable); forward
 * writes it, backward reads it.  y is stored post-dropout, so y > 0 is both the ReLU gate and the keep mask.       */
typedef struct ser_featfuse_desc {
  int dtype; int B; int T; int D; int F;   /* utterances, frames per utterance, 768, 8 / 12 / 20   */
  const void* x;                           /* [B*T, D] act                                         */
  const float* feats;                      /* [B, F] fp32                                          */
  const float* w; const float* b;          /* [D, D+F], [D] fp32 masters                           */
  void* wx;                                /* [D, D] act: packed W[:, :D]                          */
  float* c;                                /* [B, D] fp32 scratch (forward)                        */
  void* y;                                 /* [B*T, D] act: output, saved for backward             */
  float p_drop; const unsigned long long* drop_seed;   /* site SER_DS_FEAT, rows = B*T, cols = D   */
  /* backward */
  const void* dy;                          /* [B*T, D] act                                         */
  void* dz;                                /* [B*T, D] act scratch                                 */
  void* dx;                                /* [B*T, D] act or NULL (frozen-encoder input)          */
  float* dc;                               /* [B, D] fp32 scratch                                  */
  float* dwx;                              /* [D, D] fp32 scratch (zero-filled if grads_zeroed)    */
  float* dw; float* db;                    /* [D, D+F], [D]: overwritten                           */
  int grads_zeroed;
} ser_featfuse_desc;
int ser_featfuse_fwd(const ser_featfuse_desc* d, void* stream);
int ser_featfuse_bwd(const ser_featfuse_desc* d, void* stream);

/* ---- a3: AttentiveStatsPooling.forward  (src/models/pooling.py:15-28) -------------------------- */
typedef struct ser_asp_desc {
  int dtype; int B, T, D, Hd;              /* D = 768, Hd = 128                                   */
  const void* x; const float* mask;        /* [B*T,D] act; [B,T] float (0 = pad) or NULL          */
  const void* w1; const float* b1;         /* attention.0: [Hd,D], [Hd]                           */
  const float* w2; const float* b2;        /* attention.2: [Hd], [1] fp32                         */
  void* u;                                 /* [B*T,Hd] act: tanh(W1 x + b1), saved                */
  float* e; float* alpha;                  /* [B,T] scratch scores; [B,T] softmax weights, saved  */
  void* out; int out_f32;                  /* [B,2D] mean | std                                   */
  /* backward */
  const void* dout; int dout_f32;          /* [B,2D]                                              */
  void* dx;                                /* [B*T,D] act                                         */
  void* dpre; float* dalpha;               /* [B*T,Hd] act scratch; [B,T] fp32 scratch            */
  float* dw1; float* db1; float* dw2; float* db2;   /* overwritten                                */
  int grads_zeroed;
} ser_asp_desc;
int ser_asp_fwd(const ser_asp_desc* d, void* stream);
int ser_asp_bwd(const ser_asp_desc* d, void* stream);

/* ---- a4: FusionLayer.forward  (src/models/fusion.py:18-25) ------------------------------------- */
typedef struct ser_fusion_desc {
  int dtype; int B, Din, P, G;             /* Din = 1536, P = 512, G = 256                        */
  const void* av; const void* tv;          /* [B,Din] act                                         */
  const void* w1a; const float* b1a; const void* w2a; const float* b2a;   /* proj_a.0, proj_a.3   */
  const void* w1t; const float* b1t; const void* w2t; const float* b2t;
  const void* wg1a; const float* bg1a; const float* wg2a; const float* bg2a;  /* gate_a.0 (act), gate_a.2 (fp32 [G],[1]) */
  const void* wg1t; const float* bg1t; const float* wg2t; const float* bg2t;
  void* ha; void* ht; void* pa; void* pt; void* ga; void* gt;   /* saved: [B,P],[B,P],[B,G] act   */
  float* gates;                            /* [B,2] sigmoid gate values, saved                    */
  void* fused;                             /* [B,P] act, output                                   */
  float p_drop; const unsigned long long* drop_seed;   /* proj_a[2] / proj_t[2]; 0 / NULL = off; ha / ht are saved post-dropout */
  /* backward */
  const void* dfused;                      /* [B,P] act                                           */
  void* dav; void* dtv;                    /* [B,Din] act                                         */
  float* dw1a; float* db1a; float* dw2a; float* db2a; float* dw1t; float* db1t; float* dw2t; float* db2t;
  float* dwg1a; float* dbg1a; float* dwg2a; float* dbg2a; float* dwg1t; float* dbg1t; float* dwg2t; float* dbg2t;
  void* ws; size_t ws_bytes;               /* ser_fusion_bwd_ws_bytes()                           */
  int grads_zeroed;
  int Din_t;                               /* FusionLayer(audio_dim != text_dim) (fusion.py:6-16): width of tv / w1t / dtv; 0 = Din */
} ser_fusion_desc;
size_t ser_fusion_bwd_ws_bytes(int dtype, int B, int Din, int P, int G);
int ser_fusion_fwd(const ser_fusion_desc* d, void* stream);
int ser_fusion_bwd(const ser_fusion_desc* d, void* stream);

/* ---- a5/a6: AdvancedOpenMaxClassifier.forward, training path (src/models/classifier.py:200-238) --
 * 35 x { y = LN_outer(h); h' = y + W2 relu(W1 LN_inner(y) + b1) + b2 }.  Per-block parameters are passed
 * as HOST arrays (length L) of device pointers.  The residual stream is kept in fp32 in both tiers.
 * The anchor-clustering branch is not evaluated: its similarity output is discarded by the caller and its
 * loss is identically zero with exactly-zero gradients (classifier.py:64-68).     */
typedef struct ser_clf_desc {
  int dtype; int B, P, F, C, L, U;         /* 512, 256, classes, 35, 64                           */
  const void* x;                           /* [B,Pin] act (Pin = P unless the last field says otherwise) */
  const void* w_in; const float* b_in; const float* ln_in_g; const float* ln_in_b;
  const void* const* w1; const float* const* b1; const void* const* w2; const float* const* b2;
  const float* const* lno_g; const float* const* lno_b; const float* const* lni_g; const float* const* lni_b;
  const void* w_out; const float* b_out; const float* ln_out_g; const float* ln_out_b;
  const float* w_c; const float* b_c;      /* output_projection.4: [C,F] fp32                     */
  const float* w_u1; const float* b_u1; const float* w_u2; const float* b_u2;   /* uncertainty head */
  /* saved activations */
  float* p0; float* stats0;                /* [B,P], [B,2]                                        */
  float* h;                                /* [(L+1),B,P] fp32 stream                             */
  float* y;                                /* [L,B,P] fp32 outer-LN outputs                       */
  void* n; void* r;                        /* [L,B,P] act                                         */
  float* stats_o; float* stats_i;          /* [L,B,2]                                             */
  void* h_last;                            /* [B,P] act copy of h[L]                              */
  float* q; float* stats_q;                /* [B,F], [B,2]                                        */
  float* f;                                /* [B,F] penultimate features (fp32), output           */
  float* u1;                               /* [B,U]                                               */
  float* logits; float* unc;               /* [B,C], [B,1] outputs; unc may be NULL               */
  float p_drop; const unsigned long long* drop_seed;   /* all classifier dropouts; 0 / NULL = off; h[0], r, f, u1 are saved post-dropout */
  /* backward */
  const float* dlogits; const float* dunc; /* [B,C]; [B,1] or NULL                                */
  void* dx;                                /* [B,Pin] act                                         */
  float* dw_in; float* db_in; float* dln_in_g; float* dln_in_b;
  float* const* dw1; float* const* db1; float* const* dw2; float* const* db2;
  float* const* dlno_g; float* const* dlno_b; float* const* dlni_g; float* const* dlni_b;  /* ACCUMULATED: pass zeroed */
  float* dw_out; float* db_out; float* dln_out_g; float* dln_out_b;
  float* dw_c; float* db_c; float* dw_u1; float* db_u1; float* dw_u2; float* db_u2;
  void* ws; size_t ws_bytes;               /* ser_clf_bwd_ws_bytes()                              */
  int grads_zeroed;
  int Pin;                                 /* width of x and dx (DeepClassifier input_dim, classifier.py:96-105); 0 = P.
                                            * bf16 tier: a multiple of 128 (dx is a tcgen05 GEMM output)          */
} ser_clf_desc;
size_t ser_clf_bwd_ws_bytes(int dtype, int B, int P, int F, int C, int U);
int ser_clf_fwd(const ser_clf_desc* d, void* stream);
int ser_clf_bwd(const ser_clf_desc* d, void* stream);

/* ---- a8-a11: losses  (src/models/losses.py:12-30,41-64; prototypes.py:13-53; train.py:151-168) ----
 * ser_loss_fwd writes raw batch sums (ce, focal, sum unc, sum correct, pos, neg); a data-parallel caller
 * all-reduces them (and passes global class counts / B_global) before ser_loss_finalize / ser_loss_bwd,
 * which evaluate the reference's non-finite guards on the global values.
 * loss = w_ce*CE_smooth + w_focal*CBFocal + w_unc*mean(unc)*mean(correct) + w_proto*proto            */
typedef struct ser_loss_desc {
  int B, C, D; long long B_global;
  const float* logits; const float* unc;   /* [B,C] ; [B] or NULL                                 */
  const void* emb; int emb_f32;            /* [B,D] or NULL                                       */
  const float* protos;                     /* [C,D]                                               */
  const long long* labels;                 /* [B] int64                                           */
  const float* counts;                     /* [C] global class counts (float) or NULL             */
  float smoothing, beta, gamma, margin; int focal_use_weights;
  float* class_w;                          /* [C] scratch                                         */
  float* sums;                             /* [8]                                                 */
  float w_ce, w_focal, w_unc, w_proto;
  float* terms;                            /* [6]: ce, focal, unc_loss, proto, total, accuracy    */
  /* backward */
  const float* gscale;                     /* device scalar multiplied into every gradient, or NULL */
  float* dlogits; float* dunc; void* demb; int demb_f32; float* dprotos;   /* dprotos accumulates  */
} ser_loss_desc;
int ser_loss_fwd(const ser_loss_desc* d, void* stream);
int ser_loss_finalize(const ser_loss_desc* d, void* stream);
int ser_loss_bwd(const ser_loss_desc* d, void* stream);

/* ---- SupConLoss  (src/models/losses.py:67-88) ------------------------------------------------------
 * Supervised contrastive loss over one batch: normalised embeddings, B x B similarities / temperature, row-max shift,
 * positives = same label off the diagonal, loss = -mean_i(sum_pos log_prob / (n_pos + 1e-12)).  Imported and
 * constructed by the reference's scripts (train.py:8,86).  f: [B,D] (f_f32: fp32 else bf16), labels int64 [B],
 * loss: one fp32; backward: df = gscale[0] * dloss/df (gscale NULL = 1).  ws: ser_supcon_ws_bytes(B, D) bytes.       */
size_t ser_supcon_ws_bytes(int B, int D);
int ser_supcon_fwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, float* loss,
                   void* ws, size_t ws_bytes, void* stream);
int ser_supcon_bwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature,
                   const float* gscale, void* df, int df_f32, void* ws, size_t ws_bytes, void* stream);

/* ---- optimizer step (the consumer of the head's gradients) ---------------------
 * ser_adamw_multi: one torch.optim.AdamW step (decoupled weight decay, bias correction with `step` >= 1) over n fp32
 * tensors of one parameter group (src/train.py:72-83,169-177).  p / g / m / v / counts are HOST arrays of length n
 * (device pointers, element counts).  gscale: optional device scalar multiplied into every gradient first -- the
 * coefficient ser_grad_clip_coef leaves in `coef` (torch.nn.utils.clip_grad_norm_, train_crema.py).                    */
int ser_adamw_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                    const long long* counts, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    const float* gscale, void* stream);
/* The same step behind torch.amp.GradScaler.step(optimizer) (src/train.py:88,169-177, the fp16 AMP branch): nothing
 * is read back to the host.  found_inf [1] (device, may be NULL): non-zero -> the whole update is skipped;
 * grad_scale [1] (device, may be NULL): gradients are divided by it (the scaler's loss scale when the caller did not
 * unscale_ first); step_dev [1] (device, required): the bias-correction step count t >= 1 of THIS step as a float --
 * the caller advances it with `step_dev += 1 - found_inf` before the call, so skipped steps do not count.            */
int ser_adamw_multi_amp(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                        const long long* counts, float lr, float beta1, float beta2, float eps, float weight_decay,
                        const float* step_dev, const float* gscale, const float* grad_scale, const float* found_inf,
                        void* stream);
int ser_grad_clip_coef(int n, const float* const* g, const long long* counts, float max_norm, float* scratch,
                       float* coef, float* norm_out, void* stream);

/* ---- a7 / a12: inference post-processing --------------------------------------------------------
 * ser_openmax_fwd : classifier.py:240-275 (Weibull CDF of distances to activation vectors, re-scale logits)
 * ser_eval_post   : eval.py:186-190 (mean over V views), :201-206 (/T, softmax, argmax), utils.py:12-14 (energy)
 * ser_temperature_sweep : eval.py:48-67, err[t] = mean |max prob - correct| for each temperature    */
int ser_openmax_fwd(const float* feats, const float* logits, const float* act_vecs, const float* w_alpha,
                    const float* w_beta, const float* w_tau, float* out, int B, int C, int F, void* stream);
int ser_eval_post(const float* logits_views, int V, int B, int C, float temperature, float* mean_logits,
                  float* probs, long long* preds, float* energy, void* stream);
int ser_temperature_sweep(const float* logits, const long long* labels, int B, int C, const float* temps, int nT,
                          float* err, void* stream);
/* ser_late_ood : late-stage OOD scoring (src/models/dual_gate_ood.py:203-220 energy,
 *   :280-312 diagonal-Mahalanobis prototype distances, :360-383 score mix), one launch, one warp per sample:
 *   energy = -logsumexp(logits / *temperature);  dist[b,c] = sqrt(sum_d (f[b,d] - P[c,d])^2 / (cov[c,d] + 1e-8));
 *   e_norm = sigmoid(-energy), d_norm = exp(-min_c dist), combined = softmax(mix)[0] e_norm + softmax(mix)[1] d_norm.
 *   temperature [1], mix [2], prototypes / covariances [C,D] are DEVICE pointers (nn.Parameters: no host sync);
 *   feats [B,D] fp32 or bf16; distances [B,C]; scores [B,5] = energy, min distance, e_norm, d_norm, combined.        */
int ser_late_ood(const float* logits, const void* feats, int feats_f32, const float* prototypes,
                 const float* covariances, const float* temperature, const float* mix, float* distances,
                 float* scores, int B, int C, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SER_HEAD_H_ */

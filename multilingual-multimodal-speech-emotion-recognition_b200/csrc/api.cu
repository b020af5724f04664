// extern "C" surface of libser_head.so (declared in include/ser_head.h).
#include "common.cuh"
#include "kernels.cuh"
#include "modules.cuh"
#include <stdlib.h>
#include <string.h>

namespace ser {

static thread_local char g_err[512] = "";

void set_last_error(const char* file, int line, const char* msg) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
}
const char* last_error() { return g_err; }

static long long g_launches = 0;
void count_launch() { ++g_launches; }
long long launch_count() { return g_launches; }

// A/B switch: SER_PDL=0 launches every kernel with plain stream serialization
bool pdl_enabled() {
  static const bool on = !(getenv("SER_PDL") != nullptr && atoi(getenv("SER_PDL")) == 0);
  return on;
}

// ---- side branch (common.cuh) ----
namespace {
struct SideState {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[32];
  int next = 0;
  bool ok = false, tried = false;
};
SideState& side_state() {
  static thread_local SideState st[16];
  int dev = 0;
  cudaGetDevice(&dev);
  SideState& s = st[dev & 15];
  if (!s.tried) {
    s.tried = true;
    const char* e = getenv("SER_SIDE_STREAM");
    if (!(e != nullptr && atoi(e) == 0)) {
      bool good = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
      for (int i = 0; i < 32 && good; ++i) good = cudaEventCreateWithFlags(&s.ev[i], cudaEventDisableTiming) == cudaSuccess;
      s.ok = good;
      if (!good) cudaGetLastError();
    }
  }
  return s;
}
cudaEvent_t next_event(SideState& s) { cudaEvent_t e = s.ev[s.next]; s.next = (s.next + 1) & 31; return e; }
}  // namespace

SideBranch::SideBranch(cudaStream_t main_s) : main_stream(main_s) {
  SideState& st = side_state();
  on = st.ok;
  side_stream = st.stream;
}
int SideBranch::fork() {
  if (!on) return SER_OK;
  SideState& st = side_state();
  cudaEvent_t e = next_event(st);
  SER_CUDA_CHECK(cudaEventRecord(e, main_stream));
  SER_CUDA_CHECK(cudaStreamWaitEvent(side_stream, e, 0));
  used = true;
  return SER_OK;
}
int SideBranch::join() {
  if (!on || !used) return SER_OK;
  SideState& st = side_state();
  cudaEvent_t e = next_event(st);
  SER_CUDA_CHECK(cudaEventRecord(e, side_stream));
  SER_CUDA_CHECK(cudaStreamWaitEvent(main_stream, e, 0));
  used = false;
  return SER_OK;
}

}  // namespace ser

extern "C" {

int ser_version(void) { return 100; }

const char* ser_last_error(void) { return ser::last_error(); }

int ser_sm_count(void) { return ser::device_sm_count(); }

int ser_set_reserved_sms(int n) { ser::set_reserved_sms(n); return SER_OK; }

long long ser_launch_count(void) { return ser::launch_count(); }

int ser_desc_size(int id) {
  switch (id) {
    case 0: return static_cast<int>(sizeof(ser_gemm_desc));
    case 1: return static_cast<int>(sizeof(ser_adapter_desc));
    case 2: return static_cast<int>(sizeof(ser_xattn_desc));
    case 3: return static_cast<int>(sizeof(ser_asp_desc));
    case 4: return static_cast<int>(sizeof(ser_fusion_desc));
    case 5: return static_cast<int>(sizeof(ser_clf_desc));
    case 6: return static_cast<int>(sizeof(ser_loss_desc));
    case 7: return static_cast<int>(sizeof(ser_featfuse_desc));
    case 8: return static_cast<int>(sizeof(ser_attn_desc));
    default: return -1;
  }
}

int ser_gemm(const ser_gemm_desc* d, void* stream) {
  if (d == nullptr) { ser::set_last_error(__FILE__, __LINE__, "null descriptor"); return SER_ERR_ARG; }
  ser::GemmArgs a;
  a.dtype = d->dtype; a.M = d->M; a.N = d->N; a.K = d->K;
  a.A = d->A; a.lda = d->lda; a.a_trans = d->a_trans;
  a.B = d->B; a.ldb = d->ldb; a.b_trans = d->b_trans;
  a.C = d->C; a.ldc = d->ldc; a.c_f32 = d->c_f32;
  a.bias = d->bias;
  a.R = d->R; a.ldr = d->ldr; a.r_f32 = d->r_f32;
  a.G = d->G; a.ldg = d->ldg; a.g_f32 = d->g_f32; a.gate_mode = d->gate_mode;
  a.act = d->act; a.accumulate = d->accumulate; a.alpha = d->alpha; a.splits = d->splits;
  a.rowsum = d->rowsum;
  return ser::gemm(a, reinterpret_cast<cudaStream_t>(stream));
}


#define SER_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define SER_NOT_NULL(d)                                                                  \
  if ((d) == nullptr) { ser::set_last_error(__FILE__, __LINE__, "null descriptor"); return SER_ERR_ARG; }

int ser_cast(const void* src, int src_f32, void* dst, int dst_f32, long long n, void* stream) {
  return ser::cast_any(src, src_f32, dst, dst_f32, n, SER_STREAM(stream));
}

int ser_unpack_frames(const void* packed, const long long* offsets, void* out, float* mask, int B, int T, int D,
                      int elem_bytes, void* stream) {
  return ser::unpack_frames(packed, offsets, out, mask, B, T, D, elem_bytes, SER_STREAM(stream));
}

int ser_cast_multi(int n, const void* const* src, void* const* dst, const long long* counts, void* stream) {
  return ser::cast_multi(n, src, dst, counts, SER_STREAM(stream));
}

int ser_layernorm_fwd(const void* x, int x_f32, void* y, int y_f32, const float* gamma, const float* beta,
                      float* stats, int M, int N, int relu, void* stream) {
  return ser::layernorm_fwd(x, x_f32, y, y_f32, nullptr, 1, gamma, beta, stats, M, N, relu, SER_STREAM(stream));
}
int ser_layernorm_bwd(const void* dy, int dy_f32, const void* x, int x_f32, const float* stats, const float* gamma,
                      const float* beta, void* dx, int dx_f32, float* dgamma, float* dbeta, int M, int N, int relu,
                      void* stream) {
  cudaStream_t s = SER_STREAM(stream);
  SER_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * N, s));
  SER_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * N, s));
  return ser::layernorm_bwd(dy, dy_f32, x, x_f32, stats, gamma, beta, nullptr, 1, dx, dx_f32, nullptr, 1, dgamma, dbeta,
                            M, N, relu, s);
}
int ser_colsum(const void* X, int x_f32, long long ld, int M, int N, float* out, void* stream) {
  return ser::colsum(X, x_f32, ld, M, N, out, SER_STREAM(stream));
}

int ser_dropout_mask(const unsigned long long* seed, int site, float p, long long rows, int cols, float* out,
                     void* stream) {
  return ser::dropout_mask(ser::make_drop(seed, p, static_cast<unsigned>(site)), rows, cols, out, SER_STREAM(stream));
}

int ser_adapter_fwd(const ser_adapter_desc* d, void* stream) { SER_NOT_NULL(d); return ser::adapter_fwd(*d, SER_STREAM(stream)); }
int ser_adapter_bwd(const ser_adapter_desc* d, void* stream) { SER_NOT_NULL(d); return ser::adapter_bwd(*d, SER_STREAM(stream)); }

int ser_featfuse_fwd(const ser_featfuse_desc* d, void* stream) { SER_NOT_NULL(d); return ser::featfuse_fwd(*d, SER_STREAM(stream)); }
int ser_featfuse_bwd(const ser_featfuse_desc* d, void* stream) { SER_NOT_NULL(d); return ser::featfuse_bwd(*d, SER_STREAM(stream)); }

size_t ser_xattn_bwd_ws_bytes(int dtype, int B, int Ta, int Tt, int D, int S, int H) {
  return ser::xattn_bwd_ws_bytes(dtype, B, Ta, Tt, D, S, H);
}
static ser::AttnArgs to_attn_args(const ser_attn_desc& d) {
  ser::AttnArgs a{};
  a.dtype = d.dtype; a.B = d.B; a.H = d.H; a.Tq = d.Tq; a.Tk = d.Tk; a.dh = d.dh;
  a.Q = d.Q; a.ldq = d.ldq; a.K = d.K; a.ldk = d.ldk; a.V = d.V; a.ldv = d.ldv; a.kmask = d.kmask;
  a.O = d.O; a.ldo = d.ldo; a.lse = d.lse; a.scale = d.scale;
  a.dO = d.dO; a.lddo = d.lddo; a.dQ = d.dQ; a.lddq = d.lddq; a.dK = d.dK; a.lddk = d.lddk; a.dV = d.dV; a.lddv = d.lddv;
  a.delta = d.delta;
  a.drop = ser::make_drop(d.drop_seed, d.p_drop, static_cast<unsigned>(d.drop_site));
  a.keep_bits = d.keep_bits;
  a.impl = d.impl;
  return a;
}
int ser_attention_fwd(const ser_attn_desc* d, void* stream) { SER_NOT_NULL(d); return ser::attention_fwd(to_attn_args(*d), SER_STREAM(stream)); }
int ser_attention_bwd(const ser_attn_desc* d, void* stream) { SER_NOT_NULL(d); return ser::attention_bwd(to_attn_args(*d), SER_STREAM(stream)); }

int ser_xattn_folded(int dtype, int D, int S) { return ser::xattn_fold_enabled(dtype, D, S) ? 1 : 0; }
int ser_xattn_fwd(const ser_xattn_desc* d, void* stream) { SER_NOT_NULL(d); return ser::xattn_fwd(*d, SER_STREAM(stream)); }
int ser_xattn_bwd(const ser_xattn_desc* d, void* stream) { SER_NOT_NULL(d); return ser::xattn_bwd(*d, SER_STREAM(stream)); }

int ser_asp_fwd(const ser_asp_desc* d, void* stream) { SER_NOT_NULL(d); return ser::asp_module_fwd(*d, SER_STREAM(stream)); }
int ser_asp_bwd(const ser_asp_desc* d, void* stream) { SER_NOT_NULL(d); return ser::asp_module_bwd(*d, SER_STREAM(stream)); }

size_t ser_fusion_bwd_ws_bytes(int dtype, int B, int Din, int P, int G) { return ser::fusion_bwd_ws_bytes(dtype, B, Din, P, G); }
int ser_fusion_fwd(const ser_fusion_desc* d, void* stream) { SER_NOT_NULL(d); return ser::fusion_fwd(*d, SER_STREAM(stream)); }
int ser_fusion_bwd(const ser_fusion_desc* d, void* stream) { SER_NOT_NULL(d); return ser::fusion_bwd(*d, SER_STREAM(stream)); }

size_t ser_clf_bwd_ws_bytes(int dtype, int B, int P, int F, int C, int U) { return ser::clf_bwd_ws_bytes(dtype, B, P, F, C, U); }
int ser_clf_fwd(const ser_clf_desc* d, void* stream) { SER_NOT_NULL(d); return ser::clf_fwd(*d, SER_STREAM(stream)); }
int ser_clf_bwd(const ser_clf_desc* d, void* stream) { SER_NOT_NULL(d); return ser::clf_bwd(*d, SER_STREAM(stream)); }

static ser::LossArgs to_loss_args(const ser_loss_desc& d) {
  ser::LossArgs a{};
  a.B = d.B; a.C = d.C; a.D = d.D; a.B_global = d.B_global > 0 ? d.B_global : d.B;
  a.logits = d.logits; a.unc = d.unc; a.emb = d.emb; a.emb_f32 = d.emb_f32; a.protos = d.protos;
  a.labels = d.labels; a.counts = d.counts;
  a.smoothing = d.smoothing; a.beta = d.beta; a.gamma = d.gamma; a.margin = d.margin;
  a.focal_use_weights = d.focal_use_weights; a.class_w = d.class_w; a.sums = d.sums;
  a.w_ce = d.w_ce; a.w_focal = d.w_focal; a.w_unc = d.w_unc; a.w_proto = d.w_proto;
  a.dlogits = d.dlogits; a.dunc = d.dunc; a.demb = d.demb; a.demb_f32 = d.demb_f32; a.dprotos = d.dprotos;
  return a;
}
int ser_loss_fwd(const ser_loss_desc* d, void* stream) { SER_NOT_NULL(d); return ser::loss_fwd(to_loss_args(*d), SER_STREAM(stream)); }
int ser_loss_finalize(const ser_loss_desc* d, void* stream) {
  SER_NOT_NULL(d);
  const long long bg = d->B_global > 0 ? d->B_global : d->B;
  return ser::loss_finalize(d->sums, bg, d->margin, d->w_ce, d->w_focal, d->w_unc, d->w_proto, d->emb != nullptr,
                            d->terms, SER_STREAM(stream));
}
int ser_loss_bwd(const ser_loss_desc* d, void* stream) {
  SER_NOT_NULL(d);
  return ser::loss_bwd_scaled(to_loss_args(*d), d->gscale, SER_STREAM(stream));
}

size_t ser_supcon_ws_bytes(int B, int D) { return ser::supcon_ws_bytes(B, D); }
int ser_supcon_fwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature, float* loss,
                   void* ws, size_t ws_bytes, void* stream) {
  return ser::supcon_fwd(f, f_f32, labels, B, D, temperature, loss, ws, ws_bytes, SER_STREAM(stream));
}
int ser_supcon_bwd(const void* f, int f_f32, const long long* labels, int B, int D, float temperature,
                   const float* gscale, void* df, int df_f32, void* ws, size_t ws_bytes, void* stream) {
  return ser::supcon_bwd(f, f_f32, labels, B, D, temperature, gscale, df, df_f32, ws, ws_bytes, SER_STREAM(stream));
}

int ser_adamw_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                    const long long* counts, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    const float* gscale, void* stream) {
  return ser::adamw_multi(n, p, g, m, v, counts, lr, beta1, beta2, eps, weight_decay, step, gscale, SER_STREAM(stream));
}
int ser_adamw_multi_amp(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                        const long long* counts, float lr, float beta1, float beta2, float eps, float weight_decay,
                        const float* step_dev, const float* gscale, const float* grad_scale, const float* found_inf,
                        void* stream) {
  if (step_dev == nullptr) { ser::set_last_error(__FILE__, __LINE__, "adamw_amp: step_dev is required"); return SER_ERR_ARG; }
  return ser::adamw_multi(n, p, g, m, v, counts, lr, beta1, beta2, eps, weight_decay, 1, gscale, SER_STREAM(stream),
                          grad_scale, found_inf, step_dev);
}
int ser_grad_clip_coef(int n, const float* const* g, const long long* counts, float max_norm, float* scratch,
                       float* coef, float* norm_out, void* stream) {
  return ser::grad_clip_coef(n, g, counts, max_norm, scratch, coef, norm_out, SER_STREAM(stream));
}

int ser_late_ood(const float* logits, const void* feats, int feats_f32, const float* prototypes,
                 const float* covariances, const float* temperature, const float* mix, float* distances,
                 float* scores, int B, int C, int D, void* stream) {
  return ser::late_ood(logits, feats, feats_f32, prototypes, covariances, temperature, mix, distances, scores, B, C, D,
                       SER_STREAM(stream));
}
int ser_openmax_fwd(const float* feats, const float* logits, const float* act_vecs, const float* w_alpha,
                    const float* w_beta, const float* w_tau, float* out, int B, int C, int F, void* stream) {
  return ser::openmax_fwd(feats, logits, act_vecs, w_alpha, w_beta, w_tau, out, B, C, F, SER_STREAM(stream));
}
int ser_eval_post(const float* logits_views, int V, int B, int C, float temperature, float* mean_logits,
                  float* probs, long long* preds, float* energy, void* stream) {
  return ser::eval_post(logits_views, V, B, C, temperature, mean_logits, probs, preds, energy, SER_STREAM(stream));
}
int ser_temperature_sweep(const float* logits, const long long* labels, int B, int C, const float* temps, int nT,
                          float* err, void* stream) {
  return ser::temperature_sweep(logits, labels, B, C, temps, nT, err, SER_STREAM(stream));
}

}  // extern "C"

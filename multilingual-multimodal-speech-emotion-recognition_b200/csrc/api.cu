// extern "C" surface of libser_head.so (declared in include/ser_head.h).
#include "common.cuh"
#include "../../include/ser_head.h"
#include <string.h>

namespace ser {

static thread_local char g_err[512] = "";

void set_last_error(const char* file, int line, const char* msg) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
}
const char* last_error() { return g_err; }

}  // namespace ser

extern "C" {

int ser_version(void) { return 100; }

const char* ser_last_error(void) { return ser::last_error(); }

int ser_sm_count(void) { return ser::device_sm_count(); }

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
  return ser::gemm(a, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

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
} ser_gemm_desc;
int ser_gemm(const ser_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SER_HEAD_H_ */

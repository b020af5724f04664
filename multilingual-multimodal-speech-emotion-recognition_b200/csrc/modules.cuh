// Module-level entry points implemented in head_modules.cu (called from api.cu).
#pragma once
#include "common.cuh"
#include "../../include/ser_head.h"

namespace ser {
int adapter_fwd(const ser_adapter_desc& d, cudaStream_t s);
int adapter_bwd(const ser_adapter_desc& d, cudaStream_t s);
int featfuse_fwd(const ser_featfuse_desc& d, cudaStream_t s);
int featfuse_bwd(const ser_featfuse_desc& d, cudaStream_t s);
int xattn_fwd(const ser_xattn_desc& d, cudaStream_t s);
int xattn_bwd(const ser_xattn_desc& d, cudaStream_t s);
size_t xattn_bwd_ws_bytes(int dtype, int B, int Ta, int Tt, int D, int S, int H);
bool xattn_fold_enabled(int dtype, int D, int S);
int asp_module_fwd(const ser_asp_desc& d, cudaStream_t s);
int asp_module_bwd(const ser_asp_desc& d, cudaStream_t s);
int fusion_fwd(const ser_fusion_desc& d, cudaStream_t s);
int fusion_bwd(const ser_fusion_desc& d, cudaStream_t s);
size_t fusion_bwd_ws_bytes(int dtype, int B, int Din, int P, int G);
int clf_fwd(const ser_clf_desc& d, cudaStream_t s);
int clf_bwd(const ser_clf_desc& d, cudaStream_t s);
size_t clf_bwd_ws_bytes(int dtype, int B, int P, int F, int C, int U);
}  // namespace ser

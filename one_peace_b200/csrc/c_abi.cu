// extern "C" boundary of libonepeace_b200.so — see include/onepeace_b200.h for the contract.
#include "../../include/onepeace_b200.h"

#include "common.cuh"
#include "gemm.h"
#include "ops.h"

extern "C" {

int opb_abi_version(void) { return 1; }

const char* opb_status_string(int status) {
  switch (status) {
    case OPB_OK: return "ok";
    case OPB_ERR_INVALID: return "invalid argument (shape / alignment / null pointer)";
    case OPB_ERR_CUDA: return "CUDA error at launch";
    case OPB_ERR_UNSUPPORTED: return "unsupported shape";
    default: return "unknown status";
  }
}

int opb_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int epi, void* out,
                  int64_t ldo, const float* bias, const float* colscale, const float* gamma, const float* resid,
                  int64_t ldr, int out_group, int out_group_stride, int out_row_offset, int resid_period,
                  int resid_row_offset, int cta_group, void* stream) {
  if (A == nullptr || B == nullptr || out == nullptr) return OPB_ERR_INVALID;
  opb::GemmEpilogue ep;
  ep.out = out;
  ep.ldo = ldo;
  ep.bias = bias;
  ep.colscale = colscale;
  ep.gamma = gamma;
  ep.resid = resid;
  ep.ldr = ldr;
  ep.out_group = out_group;
  ep.out_group_stride = out_group_stride;
  ep.out_row_offset = out_row_offset;
  ep.resid_period = resid_period;
  ep.resid_row_offset = resid_row_offset;
  return opb::gemm_bf16(A, static_cast<int>(lda), B, static_cast<int>(ldb), M, N, K, epi, ep, cta_group,
                        static_cast<cudaStream_t>(stream));
}

int opb_attention_fwd(const void* qkv, const float* bias, const uint8_t* key_pad, void* out, float* lse, int B,
                      int S, int H, int s_pad, void* stream) {
  if (qkv == nullptr || out == nullptr) return OPB_ERR_INVALID;
  return opb::attention_fwd(qkv, bias, key_pad, out, lse, B, S, H, s_pad, static_cast<cudaStream_t>(stream));
}

int opb_layernorm(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype, int64_t ld_out,
                  const float* gamma, const float* beta, int rows, int dim, float eps, int gelu, int merge_grid_w,
                  void* stream) {
  if (in == nullptr || out == nullptr) return OPB_ERR_INVALID;
  return opb::layernorm(in, in_dtype, ld_in, out, out_dtype, ld_out, gamma, beta, rows, dim, eps, gelu,
                        merge_grid_w, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

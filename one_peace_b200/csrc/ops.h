// Internal C++ declarations of the non-GEMM kernels' host launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

int attention_fwd(const void* qkv, const float* bias, const uint8_t* key_pad, void* out, float* lse, int B, int S,
                  int H, int s_pad, cudaStream_t stream);

int layernorm(const void* in, int in_dtype, long ld_in, void* out, int out_dtype, long ld_out, const float* gamma,
              const float* beta, int rows, int dim, float eps, int gelu, int merge_grid_w, cudaStream_t stream);

}  // namespace opb

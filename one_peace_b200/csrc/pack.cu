// Parameter preprocessing for the fused-LayerNorm GEMM chain, one HBM pass per weight:
//     LN(x) W^T + b  =  rstd * (x (W . diag(g))^T  -  mu * colsum)  +  (W beta + b)
//   Wg[row(n), k]  = bf16(W[n, k] * g[k])            the folded operand the consumer GEMM multiplies un-normalised rows by
//   colsum[row(n)] = sum_k float(Wg[row(n), k])      (of the ROUNDED values: exactly what the tensor cores will sum)
//   bias'[row(n)]  = sum_k W[n, k] * beta[k] + b[n]
// row(n) = n, or the GeGLU tile interleave (transformer_layer.py:54-67: every 256-row GEMM tile holds 128 rows of wi_0
// followed by the matching 128 rows of wi_1).  Runs after every optimizer step (the packs are rebuilt when a parameter
// changes), where the torch formulation cost ~10 elementwise / reduction launches and ~10 passes per weight.
#include "common.cuh"
#include "ops.h"

namespace opb {

namespace {

template <typename TW>
OPB_DEVICE float4 ldw4(const TW* p) {
  if constexpr (sizeof(TW) == 4) {
    return *reinterpret_cast<const float4*>(p);
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 lo = unpack_bf16x2(v.x), hi = unpack_bf16x2(v.y);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}

// one warp per weight row; K % 4 == 0
template <typename TW>
__global__ void __launch_bounds__(256)
ln_fold_kernel(const TW* __restrict__ W, long ldw, const float* __restrict__ g, const float* __restrict__ beta,
               const float* __restrict__ bias_in, int N, int K, int interleave, __nv_bfloat16* __restrict__ out_w, long ldo,
               float* __restrict__ colsum, float* __restrict__ bias_out) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  int row = n;
  if (interleave == 1) row = (n >> 7) * 256 + (n & 127);
  else if (interleave == 2) row = (n >> 7) * 256 + 128 + (n & 127);
  const TW* w = W + n * ldw;
  __nv_bfloat16* o = out_w + row * ldo;
  float cs = 0.f, bs = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 wv = ldw4<TW>(w + k);
    const float4 gv = g != nullptr ? *reinterpret_cast<const float4*>(g + k) : make_float4(1.f, 1.f, 1.f, 1.f);
    const uint32_t p0 = pack_bf16x2(wv.x * gv.x, wv.y * gv.y), p1 = pack_bf16x2(wv.z * gv.z, wv.w * gv.w);
    *reinterpret_cast<uint2*>(o + k) = make_uint2(p0, p1);
    const float2 r0 = unpack_bf16x2(p0), r1 = unpack_bf16x2(p1);
    cs += (r0.x + r0.y) + (r1.x + r1.y);
    if (beta != nullptr) {
      const float4 bv = *reinterpret_cast<const float4*>(beta + k);
      bs += wv.x * bv.x + wv.y * bv.y + wv.z * bv.z + wv.w * bv.w;
    }
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[row] = cs;
    bias_out[row] = bs + (bias_in != nullptr ? bias_in[n] : 0.f);
  }
}

}  // namespace

int ln_fold(const void* W, int w_dtype, long ldw, const float* g, const float* beta, const float* bias_in, int N, int K,
            int interleave, void* out_w, long ldo, float* colsum, float* bias_out, cudaStream_t stream) {
  if (N <= 0 || K <= 0 || (K & 3) || (ldw & 3) || (ldo & 3) || interleave < 0 || interleave > 2) return OPB_ERR_INVALID;
  if (interleave != 0 && (N & 127)) return OPB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out_w) & 7) || (reinterpret_cast<uintptr_t>(g) & 15) ||
      (reinterpret_cast<uintptr_t>(beta) & 15))
    return OPB_ERR_INVALID;
  const unsigned blocks = static_cast<unsigned>((static_cast<long>(N) * 32 + 255) / 256);
  if (w_dtype == 0)
    ln_fold_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(W), ldw, g, beta, bias_in, N, K, interleave,
                                                     reinterpret_cast<__nv_bfloat16*>(out_w), ldo, colsum, bias_out);
  else if (w_dtype == 1)
    ln_fold_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(W), ldw, g, beta, bias_in, N, K,
                                                             interleave, reinterpret_cast<__nv_bfloat16*>(out_w), ldo, colsum,
                                                             bias_out);
  else return OPB_ERR_INVALID;
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

}  // namespace opb

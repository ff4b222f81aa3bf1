// Row gathers / scatters of the pretraining path and the general (sample-dependent, block-diagonal) dense
// relative-position bias.  All HBM-bound: 16-byte accesses, one warp (gather) or one CTA (bias) per row.
//
//   row_gather        out[r] = idx[r] >= 0 ? src[idx[r]] : fill        preserve_ids gathers of the student passes
//                                                                      (adapter/text.py:92-101,135-142: embeddings gathered
//                                                                      by position id; decoder canvas = mask token with the
//                                                                      preserved encoder rows scattered in), modality-major
//                                                                      <-> batch-major row permutations around attention
//   row_scatter_add   dsrc[idx[r]] += dout[r]  (idx[r] >= 0)           adjoint of row_gather
//   relpos_bias_block bias[bb, h, lo+i, lo+j] = table[bucket[p_i, p_j], h]   with p = preserve ids of sample bb (or i):
//                                                                      the gather_features bias gather (text.py:96-101) and
//                                                                      the block-diagonal placement of transformer_encoder.py
//                                                                      :148-158 in one pass
//   relpos_bias_block_bwd   dtable[bucket[p_i, p_j], h] += dbias[bb, h, lo+i, lo+j]
#include "common.cuh"
#include "ops.h"

namespace opb {

namespace {

template <typename T>
OPB_DEVICE float4 ld4g(const T* p) {
  if constexpr (sizeof(T) == 4) {
    return *reinterpret_cast<const float4*>(p);
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 lo = unpack_bf16x2(v.x), hi = unpack_bf16x2(v.y);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
}
template <typename T>
OPB_DEVICE void st4g(T* p, const float4 v) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = v;
  } else {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
}

// one warp per output row; dim % 4 == 0
template <typename TS, typename TO>
__global__ void __launch_bounds__(256)
row_gather_kernel(const TS* __restrict__ src, long ld_src, const int64_t* __restrict__ idx, const float* __restrict__ fill,
                  const float* __restrict__ add, long add_period, TO* __restrict__ out, long ld_out, long rows, int dim) {
  const long row = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long s = idx[row];
  TO* o = out + row * ld_out;
  const float* a = add ? add + (row % add_period) * dim : nullptr;      // broadcast addend (positional table)
  const TS* p = s >= 0 ? src + s * ld_src : nullptr;
  for (int c = lane * 4; c < dim; c += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p != nullptr) v = ld4g<TS>(p + c);
    else if (fill != nullptr) v = *reinterpret_cast<const float4*>(fill + c);
    if (a != nullptr) {
      const float4 w = *reinterpret_cast<const float4*>(a + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    st4g<TO>(o + c, v);
  }
}

template <typename TD>
__global__ void __launch_bounds__(256)
row_scatter_add_kernel(const TD* __restrict__ dout, long ld_dout, const int64_t* __restrict__ idx, float* __restrict__ dsrc,
                       long ld_dsrc, long rows, int dim) {
  const long row = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long s = idx[row];
  if (s < 0) return;
  const TD* p = dout + row * ld_dout;
  float* d = dsrc + s * ld_dsrc;
  for (int c = lane * 4; c < dim; c += 128) atomicAdd(reinterpret_cast<float4*>(d + c), ld4g<TD>(p + c));
}

// grid = (n, Bb); thread j walks the block's columns.  ids == nullptr: position = index.
__global__ void __launch_bounds__(128)
relpos_bias_block_kernel(const float* __restrict__ table, const int64_t* __restrict__ bucket, long ld_bucket,
                         const int64_t* __restrict__ ids, long ids_ld, int n, int lo, float* __restrict__ bias, int S,
                         int s_pad, int H) {
  const int i = blockIdx.x, bb = blockIdx.y;
  const int64_t* id = ids ? ids + bb * ids_ld : nullptr;
  long pi = id ? id[i] : i;
  if (pi < 0) pi = n - 1;                       // padded slot: masked_fill(preserve_ids.eq(-1), size(1) - 1), text.py:148
  float* base = bias + (static_cast<long>(bb) * H * S + (lo + i)) * s_pad + lo;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    long pj = id ? id[j] : j;
    if (pj < 0) pj = n - 1;
    const float* t = table + bucket[pi * ld_bucket + pj] * H;
    for (int h = 0; h < H; ++h) base[static_cast<long>(h) * S * s_pad + j] = t[h];
  }
}

__global__ void __launch_bounds__(128)
relpos_bias_block_bwd_kernel(const float* __restrict__ dbias, const int64_t* __restrict__ bucket, long ld_bucket,
                             const int64_t* __restrict__ ids, long ids_ld, int n, int lo, float* __restrict__ dtable, int S,
                             int s_pad, int H) {
  const int i = blockIdx.x, bb = blockIdx.y;
  const int64_t* id = ids ? ids + bb * ids_ld : nullptr;
  long pi = id ? id[i] : i;
  if (pi < 0) pi = n - 1;
  const float* base = dbias + (static_cast<long>(bb) * H * S + (lo + i)) * s_pad + lo;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    long pj = id ? id[j] : j;
    if (pj < 0) pj = n - 1;
    float* t = dtable + bucket[pi * ld_bucket + pj] * H;
    for (int h = 0; h < H; ++h) {
      const float g = base[static_cast<long>(h) * S * s_pad + j];
      if (g != 0.f) atomicAdd(t + h, g);
    }
  }
}

}  // namespace

int row_gather(const void* src, int src_dtype, long ld_src, const int64_t* idx, const float* fill, const float* add,
               long add_period, void* out, int out_dtype, long ld_out, long rows, int dim, cudaStream_t stream) {
  if (rows <= 0 || dim <= 0 || (dim & 3) || (ld_src & 3) || (ld_out & 3)) return OPB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(fill) & 15) || (reinterpret_cast<uintptr_t>(add) & 15) || (add && add_period <= 0))
    return OPB_ERR_INVALID;
  const unsigned blocks = static_cast<unsigned>((rows * 32 + 255) / 256);
#define OPB_RG(TS, TO)                                                                                              \
  row_gather_kernel<TS, TO><<<blocks, 256, 0, stream>>>(reinterpret_cast<const TS*>(src), ld_src, idx, fill, add,   \
                                                       add_period, reinterpret_cast<TO*>(out), ld_out, rows, dim)
  if (src_dtype == 0 && out_dtype == 0) OPB_RG(float, float);
  else if (src_dtype == 0 && out_dtype == 1) OPB_RG(float, __nv_bfloat16);
  else if (src_dtype == 1 && out_dtype == 0) OPB_RG(__nv_bfloat16, float);
  else if (src_dtype == 1 && out_dtype == 1) OPB_RG(__nv_bfloat16, __nv_bfloat16);
  else return OPB_ERR_INVALID;
#undef OPB_RG
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int row_scatter_add(const void* dout, int dout_dtype, long ld_dout, const int64_t* idx, float* dsrc, long ld_dsrc, long rows,
                    int dim, cudaStream_t stream) {
  if (rows <= 0 || dim <= 0 || (dim & 3) || (ld_dout & 3) || (ld_dsrc & 3)) return OPB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(dout) & 15) || (reinterpret_cast<uintptr_t>(dsrc) & 15)) return OPB_ERR_INVALID;
  const unsigned blocks = static_cast<unsigned>((rows * 32 + 255) / 256);
  if (dout_dtype == 0)
    row_scatter_add_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(dout), ld_dout, idx, dsrc,
                                                             ld_dsrc, rows, dim);
  else if (dout_dtype == 1)
    row_scatter_add_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dout), ld_dout,
                                                                     idx, dsrc, ld_dsrc, rows, dim);
  else return OPB_ERR_INVALID;
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int relpos_bias_block(const float* table, const int64_t* bucket, long ld_bucket, const int64_t* ids, long ids_ld, int Bb,
                      int n, int lo, float* bias, int S, int s_pad, int H, cudaStream_t stream) {
  if (Bb <= 0 || n <= 0 || lo < 0 || lo + n > S || s_pad < S || H <= 0) return OPB_ERR_INVALID;
  relpos_bias_block_kernel<<<dim3(n, Bb), 128, 0, stream>>>(table, bucket, ld_bucket, ids, ids_ld, n, lo, bias, S, s_pad, H);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int relpos_bias_block_bwd(const float* dbias, const int64_t* bucket, long ld_bucket, const int64_t* ids, long ids_ld, int Bb,
                          int n, int lo, float* dtable, int S, int s_pad, int H, cudaStream_t stream) {
  if (Bb <= 0 || n <= 0 || lo < 0 || lo + n > S || s_pad < S || H <= 0) return OPB_ERR_INVALID;
  relpos_bias_block_bwd_kernel<<<dim3(n, Bb), 128, 0, stream>>>(dbias, bucket, ld_bucket, ids, ids_ld, n, lo, dtable, S, s_pad,
                                                                H);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

}  // namespace opb

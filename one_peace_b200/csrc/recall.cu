// Retrieval evaluation (reference: one_peace/metrics/recall.py:22-78): top-k of every row of the similarity matrix and
// Recall@{1,5,10} counters.
//   topk_rows    one warp per row of sim fp32 [R, C]: every lane keeps a sorted private top-K of its strided columns in
//                registers, then K rounds of a warp-wide arg-max merge the 32 lists.  Ties: larger value first, then
//                the smaller column index (torch.topk leaves tie order unspecified).  HBM-bound: one pass over sim.
//   recall_hits  hits[0..2] += [cand_ids[idx[r, p]] == row_ids[r] for some p < 1 / 5 / 10]  (recall.py:39-41,50-52)
#include "common.cuh"
#include "ops.h"

namespace opb {

namespace {

constexpr int kTopK = 10;

__global__ void topk_rows_kernel(const float* __restrict__ sim, long ld, int* __restrict__ idx, float* __restrict__ val,
                                 int R, int C) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* s = sim + row * ld;
  float v[kTopK];
  int ix[kTopK];
#pragma unroll
  for (int i = 0; i < kTopK; ++i) { v[i] = -INFINITY; ix[i] = 0x7fffffff; }
  for (int c = lane; c < C; c += 32) {
    const float x = s[c];
    if (x > v[kTopK - 1]) {                   // strictly greater: among equal values the earlier column stays
      v[kTopK - 1] = x; ix[kTopK - 1] = c;
#pragma unroll
      for (int i = kTopK - 1; i > 0; --i) {
        if (v[i] > v[i - 1]) {
          const float tv = v[i]; v[i] = v[i - 1]; v[i - 1] = tv;
          const int ti = ix[i]; ix[i] = ix[i - 1]; ix[i - 1] = ti;
        }
      }
    }
  }
  for (int k = 0; k < kTopK; ++k) {
    float bv = v[0];
    int bi = ix[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      idx[static_cast<long>(row) * kTopK + k] = bi == 0x7fffffff ? -1 : bi;
      if (val != nullptr) val[static_cast<long>(row) * kTopK + k] = bv;
    }
    if (ix[0] == bi && bi != 0x7fffffff) {    // the winning lane pops its head
#pragma unroll
      for (int i = 0; i < kTopK - 1; ++i) { v[i] = v[i + 1]; ix[i] = ix[i + 1]; }
      v[kTopK - 1] = -INFINITY; ix[kTopK - 1] = 0x7fffffff;
    }
  }
}

__global__ void recall_hits_kernel(const int* __restrict__ idx, const int64_t* __restrict__ cand_ids,
                                   const int64_t* __restrict__ row_ids, int R, int* __restrict__ hits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int first = kTopK;
  if (r < R) {
    const int64_t want = row_ids[r];
    for (int p = kTopK - 1; p >= 0; --p) {
      const int c = idx[static_cast<long>(r) * kTopK + p];
      if (c >= 0 && cand_ids[c] == want) first = p;
    }
  }
  const unsigned m1 = __ballot_sync(0xffffffffu, first < 1), m5 = __ballot_sync(0xffffffffu, first < 5),
                 m10 = __ballot_sync(0xffffffffu, first < 10);
  if ((threadIdx.x & 31) == 0) {
    if (m1) atomicAdd(hits + 0, __popc(m1));
    if (m5) atomicAdd(hits + 1, __popc(m5));
    if (m10) atomicAdd(hits + 2, __popc(m10));
  }
}

}  // namespace

int topk10_rows(const float* sim, long ld, int* idx, float* val, int R, int C, cudaStream_t stream) {
  if (R <= 0 || C <= 0 || ld < C) return OPB_ERR_INVALID;
  topk_rows_kernel<<<static_cast<unsigned>((static_cast<long>(R) * 32 + 255) / 256), 256, 0, stream>>>(sim, ld, idx, val, R, C);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int recall_hits(const int* idx, const int64_t* cand_ids, const int64_t* row_ids, int R, int* hits3, cudaStream_t stream) {
  if (R <= 0) return OPB_ERR_INVALID;
  recall_hits_kernel<<<(R + 255) / 256, 256, 0, stream>>>(idx, cand_ids, row_ids, R, hits3);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

}  // namespace opb

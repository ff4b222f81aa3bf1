/*
 * onepeace_b200.h — C ABI of libonepeace_b200.so (hand-written sm_100a kernels for the ONE-PEACE
 * encoder / contrastive / optimizer hot path).
 *
 * Conventions (SURVEY.md §8b "What a C-ABI replacement exports"):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - no ownership transfer: the caller (PyTorch) allocates inputs, outputs and workspaces;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous;
 *   - return value: OPB_OK (0) or an OPB_ERR_* code; opb_status_string() names it.  The Python host
 *     raises RuntimeError on non-zero, matching the reference's "Python exceptions only" convention;
 *   - re-entrant; no global state except cached driver entry points and per-kernel attributes.
 *
 * Each entry point cites the reference code (under /root/reference/one_peace/) it replaces.
 */
#ifndef ONEPEACE_B200_H_
#define ONEPEACE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPB_OK 0
#define OPB_ERR_INVALID 1      /* bad shape / alignment / null pointer */
#define OPB_ERR_CUDA 2         /* CUDA runtime / driver error at launch */
#define OPB_ERR_UNSUPPORTED 3  /* shape outside what the kernels were built for */

/* dtype tags */
#define OPB_F32 0
#define OPB_BF16 1

/* GEMM epilogues */
#define OPB_EPI_STORE_BF16 0
#define OPB_EPI_GEGLU_BF16 1
#define OPB_EPI_RESID_F32 2
#define OPB_EPI_STORE_F32 3
#define OPB_EPI_GELU_BF16 4

int opb_abi_version(void);
const char* opb_status_string(int status);

/*
 * C = epilogue(A[M,K] . B[N,K]^T), bf16 operands (row pitches lda/ldb in elements), fp32 accumulate in
 * TMEM via tcgen05.mma, TMA-fed.  Replaces the nn.Linear / F.linear calls of
 *   models/transformer/multihead_attention.py:103-105,124 (q/k/v/out projections),
 *   models/transformer/transformer_layer.py:54-67,149-157 (GeGLU wi_0/wi_1, fc2),
 *   models/one_peace/one_peace_retrieval.py:114-117 (*_proj),
 * and, on patchified inputs, the stride==kernel convolutions of models/adapter/image.py:66-75 and the
 * Conv1d stacks of models/adapter/audio.py:46-80,254-311.
 *   epi = OPB_EPI_STORE_BF16 : out_bf16 = (acc + bias[n]) * colscale[n]
 *         OPB_EPI_GELU_BF16  : out_bf16 = gelu((acc + bias[n]) * colscale[n])
 *         OPB_EPI_GEGLU_BF16 : out_bf16[:, t*128+j] = gelu(acc[:, t*256+j]) * acc[:, t*256+128+j]   (N/2 cols)
 *         OPB_EPI_RESID_F32  : out_f32 = resid + gamma[n] * (acc + bias[n])   (transformer_layer.py:70-88)
 *         OPB_EPI_STORE_F32  : out_f32 = acc + bias[n]
 * bias/colscale/gamma/resid may be NULL.  Row remapping: if out_group > 0,
 *   out_row = (m / out_group) * out_group_stride + (m % out_group) + out_row_offset;
 * if resid_period > 0 the residual row is (m % resid_period) + resid_row_offset (broadcast table),
 * otherwise it is out_row.  cta_group: 1, 2 (CTA pair, 256x256 tiles) or 0 = choose.
 */
int opb_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int epi, void* out,
                  int64_t ldo, const float* bias, const float* colscale, const float* gamma, const float* resid,
                  int64_t ldr, int out_group, int out_group_stride, int out_row_offset, int resid_period,
                  int resid_row_offset, int cta_group, void* stream);

/*
 * Fused self-attention: out = softmax_fp32(q k^T + bias[h] (+ -inf on padded keys)) v.
 * Replaces models/transformer/multihead_attention.py:107-115 together with the (B,H,S,S) bias
 * materialisation of models/transformer/transformer_encoder.py:144-162.
 *   qkv  bf16 [B*S, 3*H*64] (q | k | v, q already scaled), out bf16 [B*S, H*64]
 *   bias fp32 [H, S, s_pad] or NULL (s_pad even, >= S);  key_pad uint8 [B, S] (1 = pad) or NULL
 *   lse  fp32 [B, H, S] or NULL (log-sum-exp per query row, kept for the backward pass)
 */
int opb_attention_fwd(const void* qkv, const float* bias, const uint8_t* key_pad, void* out, float* lse, int B,
                      int S, int H, int s_pad, void* stream);

/*
 * Row LayerNorm (torch.nn.LayerNorm semantics; models/components.py:23-26) with optional exact GELU and
 * optional 2x2 pixel-merge scatter (models/adapter/image.py:37-47 LayerNorm2D + the following stride-2
 * conv's patch gather).  in/out dtype tags: OPB_F32 / OPB_BF16; gamma/beta fp32 or both NULL.
 */
int opb_layernorm(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype, int64_t ld_out,
                  const float* gamma, const float* beta, int rows, int dim, float eps, int gelu, int merge_grid_w,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONEPEACE_B200_H_ */

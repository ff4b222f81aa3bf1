// Internal C++ declarations of the non-GEMM kernels' host launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

int attention_fwd(const void* qkv, const float* bias, const uint8_t* key_pad, void* out, float* lse, float* ln_stats,
                  int B, int S, int H, int s_pad, long bias_bstride, cudaStream_t stream);
int attention_tc_fwd(const void* qkv, const float* lut, const float* lut_max, int lut_len, const int* code_row, const int* code_col,
                     const uint8_t* key_pad, void* out, float* lse, float* ln_stats, int B, int S, int H, int seg_split,
                     cudaStream_t stream);
int relpos_lut_build(const float* table, const int* idx, float* lut, int L, int H, cudaStream_t stream);
int ln_stats_finalize(const float* partial, int parts, int rows, int dim, float eps, float* mu, float* rstd,
                      cudaStream_t stream);

struct LnRemap {
  int row_period = 0, row_valid = 0, out_period = 0, out_row_shift = 0;
  int group_in = 0, group_out = 0;
  int accumulate = 0;
  int raw = 0;
  float* mu_out = nullptr;
  float* rstd_out = nullptr;
};
int layernorm(const void* in, int in_dtype, long ld_in, void* out, int out_dtype, long ld_out, const float* gamma,
              const float* beta, int rows, int dim, float eps, int gelu, int merge_grid_w, const LnRemap& rm,
              cudaStream_t stream);
int pack_group_halo(const float* x, long ldx, void* out, int B, int T, int x_period, int x_row_shift, int out_period,
                    int halo, int dim, int group_in, int group_out, cudaStream_t stream);

int text_embed(const int64_t* tokens, const void* table, int table_dtype, const float* pos, const float* cls,
               float* x, uint8_t* pad_mask, int B, int T, int D, int pad_idx, cudaStream_t stream);
int image_patchify4(const void* img, int img_dtype, void* out, int B, int R, cudaStream_t stream);
int cls_row_init(const float* cls, const float* pos0, float* x, long batch_stride, int B, int D, cudaStream_t stream);
int relpos_bias_build(const float* table, const int64_t* bucket, float* bias, int S, int s_pad, int H,
                      long ld_bucket, cudaStream_t stream);
int audio_frame10(const void* wav, int wav_dtype, void* out, int B, long n_samples, long pitch, cudaStream_t stream);
int l2_normalize_rows(const float* x, long ldx, float* y, void* y_bf16, int rows, int D, cudaStream_t stream);
int zero_padded_rows(float* x, const uint8_t* pad_mask, int rows, int D, cudaStream_t stream);

int transpose_bf16(const void* in, long ld_in, void* out, int rows, int cols, cudaStream_t stream);
int split_bf16x3(const float* x, void* out, long rows, int d, int side, cudaStream_t stream);
int split_bf16x3_x4(const float* const* xs, void* const* outs, const long* rows, const int* sides, int d, cudaStream_t stream);
int infonce_lse_gemm(const void* a_local, const void* b_all, const float* scale, int b, int n, int d, int target_offset, float* ws,
                     int n_valid, cudaStream_t stream);
int infonce_merge_reduce(const float* ws_a, const float* ws_b, int b, int n, int n_valid, float eps, int target_offset, float* lse_a,
                         float* lse_b, float* loss_ab, int* am_ab, float* out3, unsigned int* ticket, cudaStream_t stream);
long infonce_ws_floats(int b, int n);
int infonce_rows(const void* a_local, const void* b_all, const float* scale, int b, int n, int d, int target_offset,
                 float eps, float* ws, float* row_lse, float* row_loss, int* row_argmax, int n_valid, cudaStream_t stream);
int infonce_reduce(const float* loss_a, const float* loss_b, const int* am_a, const int* am_b, int b,
                   int target_offset, float* out3, cudaStream_t stream);
int infonce_grad(const void* a_local, const void* b_all, const void* bT_all, const float* scale,
                 const float* row_lse, int b, int n, int d, int k_logits, int target_offset, float eps, void* g_ws,
                 float* ws_gz, float* grad_a, int n_valid, float coef, cudaStream_t stream);
int infonce_dscale(const float* ws_a, const float* ws_b, int b, int n, float* out, cudaStream_t stream);

// One entry per parameter tensor (device-resident table, 64 bytes; mirrored by ctypes in optim/adam_fused.py)
struct AdamTensor {
  void* p;         // parameter (fp32 or bf16)
  const void* g;   // gradient (fp32 or bf16)
  float* m;        // exp_avg
  float* v;        // exp_avg_sq
  float* master;   // optional fp32 master copy (nullptr: up-cast p)
  long numel;
  int group;
  int p_dtype;     // 0 fp32, 1 bf16
  int g_dtype;
  int pad_;
};
constexpr int kAdamMaxGroups = 128;
struct AdamGroups {
  float lr[kAdamMaxGroups];         // lr * lr_scale of the group (base_optimizer.py:8-13)
  float wd[kAdamMaxGroups];
  float bias_corr[kAdamMaxGroups];  // sqrt(1 - b2^t) / (1 - b1^t)
  float beta1, beta2, eps;
};
int adam_multi_step(const void* tensors, const int* chunk_tensor, const long* chunk_off, int n_chunks,
                    const AdamGroups& groups, const float* grad_scale, cudaStream_t stream);
int grad_norm_clip(const void* tensors, const int* chunk_tensor, const long* chunk_off, int n_chunks, float* partial,
                   float multiply_factor, float max_norm, float* out2, cudaStream_t stream);

// ---- backward pass (backward.cu, attention_bwd.cu) ----
long bwd_ws_floats(int dim);
int layernorm_bwd(const void* x, int x_dtype, long ldx, const void* dy, int dy_dtype, long ld_dy, const float* gamma,
                  const float* beta, void* dx, int dx_dtype, long ld_dx, int accumulate, int rows, int dim, float eps,
                  int gelu, int dy_merge_w, float* ws, float* dgamma, float* dbeta, cudaStream_t stream);
int geglu_fwd(const void* gl, void* u, long rows, int F, cudaStream_t stream);
int geglu_bwd(const void* gl, const void* du, void* dgl, long rows, int F, cudaStream_t stream);
int scale_resid_fwd(const float* x, const void* o, const float* gamma, const float* row_scale, float* out, long rows,
                    int n, cudaStream_t stream);
int scale_resid_bwd(const float* dx, const void* o, const float* gamma, const float* row_scale, void* d_o, float* ws,
                    float* dgamma, float* dbias, int rows, int n, int in_period, int in_valid, int in_shift,
                    cudaStream_t stream);
int batch_sum_f32(const float* in, long ld, float* out, int B, long n, int accumulate, cudaStream_t stream);
int l2_normalize_bwd(const float* x, long ldx, const float* dy, long ld_dy, float* dx, void* dx_bf16, int rows, int D,
                     cudaStream_t stream);
int window_gather(const void* in, void* out, int B, int t_in, int t_out, int stride, int kw, int pad, int groups, int cg,
                  cudaStream_t stream);
int window_scatter(const void* dwin, void* dx, int B, int t_in, int t_out, int stride, int kw, int pad, int groups, int cg,
                   cudaStream_t stream);
int text_embed_bwd(const float* dx, const int64_t* tokens, float* dtable, float* dpos, float* dcls, int B, int T, int D,
                   int pad_idx, cudaStream_t stream);
int colsum_bf16(const void* y, long ldy, float* ws, float* out, int rows, int n, cudaStream_t stream);
int attn_delta(const void* d_o, const void* o, float* delta, int B, int S, int H, cudaStream_t stream);
int relpos_bias_bwd(const float* dbias, const int64_t* bucket, float* dtable, int S, int s_pad, int H, long ld_bucket,
                    cudaStream_t stream);
int attention_bwd(const void* qkv, const void* out, const void* d_out, const float* bias, const uint8_t* key_pad,
                  const float* lse, float* delta, void* dqkv, float* dbias, int B, int S, int H, int s_pad,
                  float q_scale, long bias_bstride, cudaStream_t stream);

// tcgen05 form (attention_bwd_tc.cu), S <= 224; `delta` already computed.  attention_bwd dispatches to it.
// bias_t / dbias_t (optional): the batch-shared bias and its gradient as TRANSPOSED tables (relpos_bias_transpose /
// relpos_dbias_fold) — then bias / dbias are ignored.
int attention_bwd_tc(const void* qkv, const void* d_out, const float* bias, const uint8_t* key_pad, const float* lse,
                     const float* delta, void* dqkv, float* dbias, int B, int S, int H, int s_pad, float q_scale,
                     long bias_bstride, const void* bias_t, float* dbias_t, cudaStream_t stream);
// double-buffered form (attention_bwd_tc2.cu): transposed tables only
int attention_bwd_tc2(const void* qkv, const void* d_out, const uint8_t* key_pad, const float* lse, const float* delta, void* dqkv,
                      int B, int S, int H, float q_scale, const void* bias_t, float* dbias_t, cudaStream_t stream);
int relpos_bias_transpose(const float* bias, void* bias_t, int S, int s_pad, int H, cudaStream_t stream);
int relpos_dbias_fold(const float* dbias_t, float* dbias, int S, int s_pad, int H, cudaStream_t stream);
int relpos_dbias_center(float* dbias, int S, int s_pad, int H, cudaStream_t stream);

// ---- pretraining path: row gathers, sample-dependent / block-diagonal dense relative-position bias (gather.cu) ----
int row_gather(const void* src, int src_dtype, long ld_src, const int64_t* idx, const float* fill, const float* add,
               long add_period, void* out, int out_dtype, long ld_out, long rows, int dim, cudaStream_t stream);
int row_scatter_add(const void* dout, int dout_dtype, long ld_dout, const int64_t* idx, float* dsrc, long ld_dsrc, long rows,
                    int dim, cudaStream_t stream);
int relpos_bias_block(const float* table, const int64_t* bucket, long ld_bucket, const int64_t* ids, long ids_ld, int Bb,
                      int n, int lo, float* bias, int S, int s_pad, int H, cudaStream_t stream);
int relpos_bias_block_bwd(const float* dbias, const int64_t* bucket, long ld_bucket, const int64_t* ids, long ids_ld, int Bb,
                          int n, int lo, float* dtable, int S, int s_pad, int H, cudaStream_t stream);

// ---- parameter preprocessing for the fused-LayerNorm GEMM chain (pack.cu) ----
int ln_fold(const void* W, int w_dtype, long ldw, const float* g, const float* beta, const float* bias_in, int N, int K,
            int interleave, void* out_w, long ldo, float* colsum, float* bias_out, cudaStream_t stream);

// ---- retrieval evaluation (recall.cu) ----
int topk10_rows(const float* sim, long ld, int* idx, float* val, int R, int C, cudaStream_t stream);
int recall_hits(const int* idx, const int64_t* cand_ids, const int64_t* row_ids, int R, int* hits3, cudaStream_t stream);

}  // namespace opb

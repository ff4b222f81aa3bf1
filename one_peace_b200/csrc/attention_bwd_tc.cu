// tcgen05 backward of the fused self-attention for sequences of up to 224 tokens (vision S = 197, text, 'vl' = 214); the
// reference differentiates multihead_attention.py:103-115 through torch autograd:
//     S = q k^T + bias (+ -inf on padded keys),  P = softmax(S),  O = P v
//     dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),   dQ = dS K,   dK = dS^T Q,   dbias[h] += sum_b dS[b, h]
// P is recomputed from the forward's log-sum-exp; delta = sum_d dO O comes from attn_delta (backward.cu).
//
// Round-2 replacement of the two mma.sync kernels of attention_bwd.cu (476 us per layer at B = 64, S = 197, H = 24: 14 % of a
// training step, S / dP recomputed twice, 3072 CTAs of 4 warps).  Design:
//   * ONE persistent CTA per SM walks (batch, head) items.  Everything is TRANSPOSED — TMEM lanes are KEYS:
//         S^T = K Q^T  and  dP^T = V dO^T           ([128 keys] x [QH queries] fp32 accumulators, two tcgen05.mma chains)
//     so the two products that contract over queries read their A operand where the soft-max warps leave it:
//         dV = P^T dO    A = P^T  packed bf16 written back over the S^T columns in tensor memory (tcgen05.st, TMEM A operand)
//         dK = dS^T Q    A = dS^T bf16 [key][query] rows in shared memory (K-major)
//         dQ = dS K      A = the SAME shared-memory bytes read as an MN-major operand (M = queries contiguous) — one copy of
//                        dS serves both products; B operands (dO, Q, K) are the TMA-staged tiles read MN-major (like V in P V).
//     A sequence is 1-2 key tiles of 128 x 1-2 query halves of QH <= 112 columns: per sub-step the tensor core produces S^T and
//     dP^T (2 x QH columns), 8 warps turn them into P^T / dS^T, then three short MMA chains accumulate dV / dK (over query
//     halves) and dQ (over key tiles) in tensor memory: 2 QH + 64 + 64 + 2 x 64 <= 480 columns.
//   * the soft-max warps use tcgen05.ld.16x256b (2 keys x 2 adjacent queries per 8-column block and thread), so a packed
//     bf16 pair is one register and lands in TMEM / shared memory without shuffles; per-query vectors (log-sum-exp, delta)
//     are staged once per item in shared memory; the relative-position bias is prefetched (as half2, times log2 e) while
//     the warps wait for the tensor core; dbias goes to the batch-shared fp32 table with red.global.add (as before).
//   * warp 8 = TMA producer (Q + dO per item, K + V per key tile, two stages each: the next item streams in under the
//     current one), warp 9 = MMA issuer.  Every mbarrier wait is watchdog-bounded (common.cuh).
// q in `qkv` is the SCALED query (the QKV GEMM epilogue applies head_dim^-0.5); dQ is multiplied by the same factor so
// `dqkv` is the gradient of the un-scaled projection output, exactly like attention_bwd.cu.
#include "common.cuh"
#include "ops.h"
#include "tmem_frag.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

namespace opb {

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace {

constexpr int kD = 64;                  // head dim
constexpr int kKT = 128;                // keys per tile (TMEM lanes)
constexpr int kSoftWarps = 8;
constexpr int kThreads = 32 * (kSoftWarps + 2);
// TMEM columns
constexpr int kColST = 0, kColDPT = 112, kColDV = 224, kColDK = 288, kColDQ = 352;     // dQ: two 64-column accumulators
// transposed bias tables: [H][kTKeys][kTQ] — key rows, query columns, zero-padded, so that a thread's (key, query pair) words
// sit at compile-time offsets from one pointer per sub-step (no per-element address arithmetic, clamps or predicates)
constexpr int kTKeys = 256, kTQ = 224;

// no "memory" clobber: the table is only ever touched by these reductions, and a clobber would pin every block's shared-memory loads
// behind the previous block's reduction (no overlap between the 8 blocks of a sub-step)
OPB_DEVICE void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b));
}

struct BwdBars {
  uint64_t qdo_full[2], qdo_empty[2], kv_full[2], kv_empty[2];
  uint64_t s_full, p_full, acc_done;
  uint32_t tmem_base;
};

struct BwdArgs {
  const float* bias; float* dbias; long bias_bstride; int s_pad;
  const uint32_t* bias_t; float* dbias_t;      // transposed tables (kTKeys x kTQ per head), see relpos_bias_transpose
  const uint8_t* key_pad;
  const float* lse; const float* delta;
  __nv_bfloat16* dqkv;
  int B, S, H, n_kt, n_qh;
  float q_scale;
  long n_items;
};

// MN-major bf16 operand with several 64-wide MN chunks `chunk_bytes` apart, 8-row k-groups 1024 B apart (gemm_tcgen05.cu)
OPB_DEVICE uint64_t make_sw128_mn_desc_lbo(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

OPB_DEVICE void soft_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftWarps) : "memory"); }

template <int QH>
__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_do, const BwdArgs a) {
  constexpr int NBLK8 = QH / 8;
  constexpr int KQ16 = QH / 16;                   // 16-query K-steps of the dV / dK products
  constexpr uint32_t KV_BYTES = kKT * 128;        // one K or V tile
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(1024) uint8_t bwd_smem_raw[];
  uint8_t* smem = bwd_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int NQP = a.n_qh * QH;                    // query rows staged per item (multiple of 16, <= 224)
  const uint32_t QDO_BYTES = static_cast<uint32_t>(NQP) * 128;
  const uint32_t qdo_stride = (QDO_BYTES + 1023u) & ~1023u;
  uint8_t* sQ = smem;                             // 2 stages
  uint8_t* sDO = sQ + 2 * qdo_stride;             // 2 stages
  uint8_t* sK = sDO + 2 * qdo_stride;             // 2 stages x 16 KB
  uint8_t* sV = sK + 2 * KV_BYTES;                // 2 stages
  uint8_t* sDS = sV + 2 * KV_BYTES;               // dS^T: [2 query chunks of 64][128 keys][128 B], 128-byte swizzle
  float* sL2 = reinterpret_cast<float*>(sDS + 2 * KV_BYTES);      // [2][224] log-sum-exp * log2 e (+inf: dead query)
  float* sDl = sL2 + 2 * 224;                                     // [2][224] delta
  BwdBars* bars = reinterpret_cast<BwdBars*>(sDl + 2 * 224);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S, H = a.H, D = a.H * kD;
  const long w0 = a.n_items * blockIdx.x / gridDim.x, w1 = a.n_items * (blockIdx.x + 1) / gridDim.x;
  const int n = static_cast<int>(w1 - w0);
  const int n_kt = a.n_kt, n_qh = a.n_qh;
  const int sub_per_item = n_kt * n_qh;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qdo_full[i], 1);
      mbar_init(&bars->qdo_empty[i], 1);
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, kSoftWarps);
    mbar_init(&bars->acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<1>(&bars->tmem_base, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kSoftWarps) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int t = 0;                                  // running key-tile counter
      for (int j = 0; j < n; ++j) {
        const long w = w0 + j;
        const int h = static_cast<int>(w % H), b = static_cast<int>(w / H);
        const int sq = j & 1;
        mbar_wait(&bars->qdo_empty[sq], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->qdo_full[sq], 2 * QDO_BYTES);
        tma_load_2d(&tm_q, &bars->qdo_full[sq], sQ + sq * qdo_stride, h * kD, b * S);
        tma_load_2d(&tm_do, &bars->qdo_full[sq], sDO + sq * qdo_stride, h * kD, b * S);
        for (int kt = 0; kt < n_kt; ++kt, ++t) {
          const int sk = t & 1;
          mbar_wait(&bars->kv_empty[sk], ((t >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->kv_full[sk], 2 * KV_BYTES);
          tma_load_2d(&tm_kv, &bars->kv_full[sk], sK + sk * KV_BYTES, D + h * kD, b * S + kt * kKT);
          tma_load_2d(&tm_kv, &bars->kv_full[sk], sV + sk * KV_BYTES, 2 * D + h * kD, b * S + kt * kKT);
        }
      }
    }
  } else if (warp == kSoftWarps + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kKT, QH);                                  // A, B K-major
      constexpr uint32_t idesc_kv = make_idesc_bf16(kKT, kD) | (1u << 16);                    // B MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(kKT, kD) | (1u << 15) | (1u << 16);        // A and B MN-major
      int t = 0, sub = 0;
      for (int j = 0; j < n; ++j) {
        const int sq = j & 1;
        mbar_wait(&bars->qdo_full[sq], (j >> 1) & 1);
        const uint32_t q_base = smem_u32(sQ + sq * qdo_stride), do_base = smem_u32(sDO + sq * qdo_stride);
        for (int kt = 0; kt < n_kt; ++kt, ++t) {
          const int sk = t & 1;
          mbar_wait(&bars->kv_full[sk], (t >> 1) & 1);
          const uint32_t k_base = smem_u32(sK + sk * KV_BYTES), v_base = smem_u32(sV + sk * KV_BYTES);
          int keys16 = (S - kt * kKT + 15) / 16;                  // 16-key K-steps of dQ that hold live keys
          if (keys16 > 8) keys16 = 8;
          for (int qh = 0; qh < n_qh; ++qh, ++sub) {
            tc_fence_after();
            const uint64_t dk_a = make_sw128_kmajor_desc(k_base), dv_a = make_sw128_kmajor_desc(v_base);
            const uint64_t dq_b = make_sw128_kmajor_desc(q_base + qh * QH * 128);
            const uint64_t ddo_b = make_sw128_kmajor_desc(do_base + qh * QH * 128);
#pragma unroll
            for (int kk = 0; kk < kD / 16; ++kk) umma_bf16<1>(tmem_base + kColST, dk_a + 2 * kk, dq_b + 2 * kk, idesc_s, kk != 0);
#pragma unroll
            for (int kk = 0; kk < kD / 16; ++kk) umma_bf16<1>(tmem_base + kColDPT, dv_a + 2 * kk, ddo_b + 2 * kk, idesc_s, kk != 0);
            umma_commit<1>(&bars->s_full);
            // P^T (tensor memory) and dS^T (shared memory) of this sub-step
            mbar_wait(&bars->p_full, sub & 1);
            tc_fence_after();
            const uint32_t ds_base = smem_u32(sDS);
#pragma unroll
            for (int kk = 0; kk < KQ16; ++kk) {
              // dV (+)= P^T dO: A = 16 queries = 8 packed TMEM columns; B = dO rows (queries) as an MN-major operand
              const uint64_t b_do = make_sw128_mn_desc64(do_base + (qh * QH + 16 * kk) * 128);
              umma_bf16_ts(tmem_base + kColDV, tmem_base + kColST + 8 * kk, b_do, idesc_kv, (qh != 0 || kk != 0) ? 1u : 0u);
            }
#pragma unroll
            for (int kk = 0; kk < KQ16; ++kk) {
              // dK (+)= dS^T Q: A = dS^T rows (keys), K = queries (chunk kk / 4, 32 B per step inside the swizzle row)
              const uint64_t a_ds = make_sw128_kmajor_desc(ds_base + (kk >> 2) * KV_BYTES) + 2 * (kk & 3);
              const uint64_t b_q = make_sw128_mn_desc64(q_base + (qh * QH + 16 * kk) * 128);
              umma_bf16<1>(tmem_base + kColDK, a_ds, b_q, idesc_kv, (qh != 0 || kk != 0) ? 1u : 0u);
            }
            for (int kk = 0; kk < keys16; ++kk) {
              // dQ_qh (+)= dS K: A = the same bytes MN-major (M = queries: two 64-wide chunks KV_BYTES apart; K = 16 keys = two
              // 8-row groups), B = K rows (keys) MN-major
              const uint64_t a_ds = make_sw128_mn_desc_lbo(ds_base + kk * 2048, KV_BYTES);
              const uint64_t b_k = make_sw128_mn_desc64(k_base + kk * 2048);
              umma_bf16<1>(tmem_base + kColDQ + qh * kD, a_ds, b_k, idesc_q, (kt != 0 || kk != 0) ? 1u : 0u);
            }
            umma_commit<1>(&bars->acc_done);
            if (qh == n_qh - 1) umma_commit<1>(&bars->kv_empty[sk]);
            if (qh == n_qh - 1 && kt == n_kt - 1) umma_commit<1>(&bars->qdo_empty[sq]);
          }
        }
      }
    }
  } else {
    // ===================== soft-max backward (8 warps) =====================
    const int qw = warp & 3;                          // TMEM lane quarter (hardware: warp id % 4)
    const int hh = warp >> 2;                         // which 16 lanes of the quarter
    const int t4 = lane & 3;                          // position inside the row's quad
    const int r_lo = qw * 32 + hh * 16 + (lane >> 2); // tile rows of this thread: r_lo and r_lo + 8
    const uint32_t lane_addr = static_cast<uint32_t>(qw * 32 + hh * 16) << 16;
    const long row_pitch = 3L * D;
    int sub = 0;
    for (int j = 0; j < n; ++j) {
      const long w = w0 + j;
      const int h = static_cast<int>(w % H), b = static_cast<int>(w / H);
      const long stat = (static_cast<long>(b) * H + h) * S;
      float* l2s = sL2 + (j & 1) * 224;
      float* dls = sDl + (j & 1) * 224;
      if (static_cast<int>(threadIdx.x) < NQP) {
        const int qq = threadIdx.x;
        float l = INFINITY, dl = 0.f;
        if (qq < S) {
          const float v = __ldg(a.lse + stat + qq);
          l = v > -INFINITY ? v * kLog2e : INFINITY;
          dl = __ldg(a.delta + stat + qq);
        }
        l2s[qq] = l;
        dls[qq] = dl;
      }
      soft_bar_sync();
      const long boff = static_cast<long>(b) * a.bias_bstride + static_cast<long>(h) * S * a.s_pad;
      for (int kt = 0; kt < n_kt; ++kt) {
        const int key_lo = kt * kKT + r_lo, key_hi = key_lo + 8;
        const bool warp_valid = kt * kKT + qw * 32 + hh * 16 < S;      // warp-uniform: any live key in the warp's 16 rows
        bool dead_lo = key_lo >= S, dead_hi = key_hi >= S;
        if (a.key_pad != nullptr) {
          if (!dead_lo) dead_lo = a.key_pad[static_cast<long>(b) * S + key_lo] != 0;
          if (!dead_hi) dead_hi = a.key_pad[static_cast<long>(b) * S + key_hi] != 0;
        }
        const int kc_lo = min(key_lo, S - 1), kc_hi = min(key_hi, S - 1);      // clamped: loads stay inside the table
        for (int qh = 0; qh < n_qh; ++qh, ++sub) {
          // ---- bias of this sub-tile, transposed gather (prefetched under the tensor core's S^T / dP^T) ----
          uint32_t bw[2 * NBLK8];
          if (a.bias_t != nullptr && warp_valid) {
            // transposed half2 table (x log2 e, zeros past S): one 4-byte load per (row, block) at immediate offsets
            const uint32_t* bt = a.bias_t + (static_cast<long>(h) * kTKeys + key_lo) * (kTQ / 2) + (qh * QH) / 2 + t4;
#pragma unroll
            for (int blk = 0; blk < NBLK8; ++blk) {
              bw[2 * blk] = __ldg(bt + 4 * blk);
              bw[2 * blk + 1] = __ldg(bt + 8 * (kTQ / 2) + 4 * blk);
            }
          } else if (a.bias != nullptr && warp_valid) {
            const float* bp = a.bias + boff;
#pragma unroll
            for (int blk = 0; blk < NBLK8; ++blk) {
              const int q0 = min(qh * QH + 8 * blk + 2 * t4, S - 1), q1 = min(qh * QH + 8 * blk + 2 * t4 + 1, S - 1);
              const float* r0 = bp + static_cast<long>(q0) * a.s_pad;
              const float* r1 = bp + static_cast<long>(q1) * a.s_pad;
              const float b00 = __ldg(r0 + kc_lo), b01 = __ldg(r1 + kc_lo), b10 = __ldg(r0 + kc_hi), b11 = __ldg(r1 + kc_hi);
              const __half2 ha = __floats2half2_rn(b00 * kLog2e, b01 * kLog2e), hb = __floats2half2_rn(b10 * kLog2e, b11 * kLog2e);
              bw[2 * blk] = *reinterpret_cast<const uint32_t*>(&ha);
              bw[2 * blk + 1] = *reinterpret_cast<const uint32_t*>(&hb);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2 * NBLK8; ++i) bw[i] = 0u;
          }
          mbar_wait(&bars->s_full, sub & 1);
          tc_fence_after();
          if (warp_valid) {
            uint32_t st[4 * NBLK8], dp[4 * NBLK8];
            load_scores<NBLK8>(tmem_base + lane_addr + kColST, st);
            load_scores<NBLK8>(tmem_base + lane_addr + kColDPT, dp);
            tmem_ld_wait();
            const float* l2q = l2s + qh * QH + 2 * t4;
            const float* dlq = dls + qh * QH + 2 * t4;
            float* db_lo = nullptr;
            float* db_hi = nullptr;
            float* dt_lo = nullptr;
            if (a.dbias_t != nullptr) dt_lo = a.dbias_t + (static_cast<long>(h) * kTKeys + key_lo) * kTQ + qh * QH + 2 * t4;
            if (a.dbias != nullptr) {
              db_lo = a.dbias + boff + kc_lo;
              db_hi = a.dbias + boff + kc_hi;
            }
            uint8_t* ds_row = sDS + r_lo * 128 + 4 * t4;          // + chunk * KV_BYTES + ((unit ^ (row & 7)) << 4); row + 8: + 1024
            const int rx = r_lo & 7;
            auto blocks = [&](auto lo_c, auto cnt_c) {
              constexpr int LO = decltype(lo_c)::value, CNT = decltype(cnt_c)::value;
              uint32_t pw[2 * CNT];
#pragma unroll
              for (int k = 0; k < CNT; ++k) {
                const int blk = LO + k;
                const float2 l2 = *reinterpret_cast<const float2*>(l2q + 8 * blk);
                const float2 dl = *reinterpret_cast<const float2*>(dlq + 8 * blk);
                const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&bw[2 * blk]));
                const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(&bw[2 * blk + 1]));
                float p0 = ex2_fast(fmaf(__uint_as_float(st[4 * blk + 0]), kLog2e, ba.x) - l2.x);
                float p1 = ex2_fast(fmaf(__uint_as_float(st[4 * blk + 1]), kLog2e, ba.y) - l2.y);
                float p2 = ex2_fast(fmaf(__uint_as_float(st[4 * blk + 2]), kLog2e, bb.x) - l2.x);
                float p3 = ex2_fast(fmaf(__uint_as_float(st[4 * blk + 3]), kLog2e, bb.y) - l2.y);
                p0 = dead_lo ? 0.f : p0; p1 = dead_lo ? 0.f : p1;
                p2 = dead_hi ? 0.f : p2; p3 = dead_hi ? 0.f : p3;
                const float d0 = p0 * (__uint_as_float(dp[4 * blk + 0]) - dl.x), d1 = p1 * (__uint_as_float(dp[4 * blk + 1]) - dl.y);
                const float d2 = p2 * (__uint_as_float(dp[4 * blk + 2]) - dl.x), d3 = p3 * (__uint_as_float(dp[4 * blk + 3]) - dl.y);
                if (a.dbias_t != nullptr) {
                  // dead rows / columns hold ds = 0 and land in the zero padding: only whole blocks past S are skipped
                  if (qh * QH + 8 * blk < S) {
                    if (!dead_lo) red_add_v2(dt_lo + 8 * blk, d0, d1);
                    if (!dead_hi) red_add_v2(dt_lo + 8 * kTQ + 8 * blk, d2, d3);
                  }
                } else if (a.dbias != nullptr) {
                  const int q0 = qh * QH + 8 * blk + 2 * t4;
                  const long o0 = static_cast<long>(q0) * a.s_pad, o1 = o0 + a.s_pad;
                  const bool live0 = q0 < S && l2.x < INFINITY, live1 = q0 + 1 < S && l2.y < INFINITY;
                  if (live0 && !dead_lo) atomicAdd(db_lo + o0, d0);
                  if (live1 && !dead_lo) atomicAdd(db_lo + o1, d1);
                  if (live0 && !dead_hi) atomicAdd(db_hi + o0, d2);
                  if (live1 && !dead_hi) atomicAdd(db_hi + o1, d3);
                }
                pw[2 * k] = pack_bf16x2(p0, p1);
                pw[2 * k + 1] = pack_bf16x2(p2, p3);
                uint8_t* dsp = ds_row + (blk >> 3) * KV_BYTES + (((blk & 7) ^ rx) << 4);
                *reinterpret_cast<uint32_t*>(dsp) = pack_bf16x2(d0, d1);
                *reinterpret_cast<uint32_t*>(dsp + 1024) = pack_bf16x2(d2, d3);
              }
              if constexpr (CNT == 8) tmem_st_16x128b_x8(tmem_base + lane_addr + kColST + 4 * LO, pw);
              if constexpr (CNT == 4) tmem_st_16x128b_x4(tmem_base + lane_addr + kColST + 4 * LO, pw);
              if constexpr (CNT == 2) tmem_st_16x128b_x2(tmem_base + lane_addr + kColST + 4 * LO, pw);
            };
            constexpr int n8 = NBLK8 & 8, n4 = NBLK8 & 4, n2 = NBLK8 & 2;
            static_assert(NBLK8 < 16, "");
            if constexpr (n8 != 0) blocks(std::integral_constant<int, 0>{}, std::integral_constant<int, 8>{});
            if constexpr (n4 != 0) blocks(std::integral_constant<int, n8>{}, std::integral_constant<int, 4>{});
            if constexpr (n2 != 0) blocks(std::integral_constant<int, n8 + n4>{}, std::integral_constant<int, 2>{});
            tmem_st_wait_all();
            fence_proxy_async();                  // dS^T was written through the generic proxy, the MMA reads it through the async one
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_full);
          // ---- accumulators that are complete after this sub-step ----
          const bool last_qh = qh == n_qh - 1;
          if (last_qh) {
            mbar_wait(&bars->acc_done, sub & 1);
            tc_fence_after();
            if (warp_valid) {
              uint32_t gk[32], gv[32];
              tmem_ld_16x256b_x8(tmem_base + lane_addr + kColDK, gk);
              tmem_ld_16x256b_x8(tmem_base + lane_addr + kColDV, gv);
              tmem_ld_wait();
#pragma unroll
              for (int rsel = 0; rsel < 2; ++rsel) {
                const int key = rsel ? key_hi : key_lo;
                if (key < S) {
                  uint32_t* op = reinterpret_cast<uint32_t*>(a.dqkv + (static_cast<long>(b) * S + key) * row_pitch + D + h * kD + 2 * t4);
#pragma unroll
                  for (int blk = 0; blk < 8; ++blk) {
                    op[4 * blk] = pack_bf16x2(__uint_as_float(gk[4 * blk + 2 * rsel]), __uint_as_float(gk[4 * blk + 2 * rsel + 1]));
                    op[4 * blk + D / 2] = pack_bf16x2(__uint_as_float(gv[4 * blk + 2 * rsel]), __uint_as_float(gv[4 * blk + 2 * rsel + 1]));
                  }
                }
              }
            }
            if (kt == n_kt - 1) {
              // dQ of both query halves: lanes are queries now
              for (int q2 = 0; q2 < n_qh; ++q2) {
                const int qbase = q2 * QH;
                if (qw * 32 + hh * 16 < QH && qbase + qw * 32 + hh * 16 < S) {
                  uint32_t gq[32];
                  tmem_ld_16x256b_x8(tmem_base + lane_addr + kColDQ + q2 * kD, gq);
                  tmem_ld_wait();
#pragma unroll
                  for (int rsel = 0; rsel < 2; ++rsel) {
                    const int rq = r_lo + 8 * rsel;
                    const int qrow = qbase + rq;
                    if (rq < QH && qrow < S) {
                      uint32_t* op = reinterpret_cast<uint32_t*>(a.dqkv + (static_cast<long>(b) * S + qrow) * row_pitch + h * kD + 2 * t4);
#pragma unroll
                      for (int blk = 0; blk < 8; ++blk)
                        op[4 * blk] = pack_bf16x2(__uint_as_float(gq[4 * blk + 2 * rsel]) * a.q_scale,
                                                  __uint_as_float(gq[4 * blk + 2 * rsel + 1]) * a.q_scale);
                    }
                  }
                }
              }
            }
            tc_fence_before();                    // the accumulators are overwritten by MMAs issued after the next p_full
          }
        }
      }
    }
    (void)sub_per_item;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// biasT[h][key][q / 2] = half2(bias[h][q][key], bias[h][q + 1][key]) * log2 e, zero for key >= S or q >= S.  One CTA per
// (32 keys, head): a 32 x 32 shared-memory tile turns the key-contiguous reads into query-contiguous writes.
__global__ void __launch_bounds__(256)
relpos_bias_transpose_kernel(const float* __restrict__ bias, uint32_t* __restrict__ bias_t, int S, int s_pad) {
  __shared__ float tile[32][33];
  const int h = blockIdx.y, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int q0 = 0; q0 < kTQ; q0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + ty + 8 * i, key = k0 + tx;
      tile[ty + 8 * i][tx] = (q < S && key < S) ? bias[(static_cast<long>(h) * S + q) * s_pad + key] * 1.4426950408889634f : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kk = (threadIdx.x >> 4) + 16 * i, w = threadIdx.x & 15;          // key row inside the tile, query pair
      const __half2 hv = __floats2half2_rn(tile[2 * w][kk], tile[2 * w + 1][kk]);
      bias_t[(static_cast<long>(h) * kTKeys + k0 + kk) * (kTQ / 2) + q0 / 2 + w] = *reinterpret_cast<const uint32_t*>(&hv);
    }
    __syncthreads();
  }
}

// dbias[h][q][key] += dbiasT[h][key][q]   (q, key < S)
__global__ void __launch_bounds__(256)
relpos_dbias_fold_kernel(const float* __restrict__ dbias_t, float* __restrict__ dbias, int S, int s_pad) {
  __shared__ float tile[32][33];
  const int h = blockIdx.y, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int q0 = 0; q0 < S; q0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = ty + 8 * i;
      tile[kk][tx] = dbias_t[(static_cast<long>(h) * kTKeys + k0 + kk) * kTQ + min(q0 + tx, kTQ - 1)];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + ty + 8 * i, key = k0 + tx;
      if (q < S && key < S) dbias[(static_cast<long>(h) * S + q) * s_pad + key] += tile[tx][ty + 8 * i];
    }
    __syncthreads();
  }
}

// dbias[h][i][:S] -= mean_j dbias[h][i][j].  Softmax is invariant to a per-row shift of its logits, so the exact bias gradient
// has zero row sums; the accumulated one carries a small per-row offset, because delta_i = sum_d dO_id O_id is taken from the
// bf16-rounded forward output (as flash-attention does) while P is recomputed in fp32.  That offset is harmless per element
// (~1e-4 of the row's scale) but it is COHERENT along a row, and a table entry shared by a whole row — the CLS -> token bucket
// of adapter/image.py:164-171, whose exact value is a 196-term cancellation down to -dS[0][0] — inherits all of it
// (measured: 87 % error on that one entry through InfoNCE, cosine 0.93 of the table gradient; 0.99+ with the projection).
// One warp per (head, query row).
__global__ void __launch_bounds__(256)
relpos_dbias_center_kernel(float* __restrict__ dbias, int S, int s_pad, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = dbias + static_cast<long>(row) * s_pad;
  float s = 0.f;
  for (int j = lane; j < S; j += 32) s += p[j];
  s = warp_sum(s) / S;
  for (int j = lane; j < S; j += 32) p[j] -= s;
}

template <int QH>
int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tdo, const BwdArgs& a, cudaStream_t stream) {
  const size_t qdo = (static_cast<size_t>(a.n_qh) * QH * 128 + 1023) & ~static_cast<size_t>(1023);
  const size_t smem = 4 * qdo + 4ull * kKT * 128 + 2ull * kKT * 128 + 4ull * 224 * 4 + sizeof(BwdBars) + 64;
  if (smem > 227 * 1024) return OPB_ERR_UNSUPPORTED;
  static size_t configured = 0;
  auto kern = attention_bwd_tc_kernel<QH>;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return OPB_ERR_CUDA;
    configured = smem;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned grid = static_cast<unsigned>(a.n_items < sms ? a.n_items : sms);
  return launch_maybe_cluster(kern, dim3(grid), dim3(kThreads), smem, stream, tq, tkv, tdo, a) == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

}  // namespace

// bias fp32 (H, S, s_pad) -> bias_t: H x 256 x 112 half2 words (scaled by log2 e, zero-padded)
int relpos_bias_transpose(const float* bias, void* bias_t, int S, int s_pad, int H, cudaStream_t stream) {
  if (S <= 0 || S > kTQ || H <= 0 || s_pad < S) return OPB_ERR_INVALID;
  relpos_bias_transpose_kernel<<<dim3(kTKeys / 32, H), 256, 0, stream>>>(bias, reinterpret_cast<uint32_t*>(bias_t), S, s_pad);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// dbias (H, S, s_pad) += transpose of dbias_t (H x 256 x 224 fp32)
int relpos_dbias_fold(const float* dbias_t, float* dbias, int S, int s_pad, int H, cudaStream_t stream) {
  if (S <= 0 || S > kTQ || H <= 0 || s_pad < S) return OPB_ERR_INVALID;
  relpos_dbias_fold_kernel<<<dim3((S + 31) / 32, H), 256, 0, stream>>>(dbias_t, dbias, S, s_pad);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int relpos_dbias_center(float* dbias, int S, int s_pad, int H, cudaStream_t stream) {
  if (S <= 0 || H <= 0 || s_pad < S) return OPB_ERR_INVALID;
  const long rows = static_cast<long>(H) * S;
  relpos_dbias_center_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, stream>>>(dbias, S, s_pad, static_cast<int>(rows));
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// Same contract as attention_bwd (ops.h) minus the delta kernel (the caller has run attn_delta); S <= 224.
int attention_bwd_tc(const void* qkv, const void* d_out, const float* bias, const uint8_t* key_pad, const float* lse,
                     const float* delta, void* dqkv, float* dbias, int B, int S, int H, int s_pad, float q_scale,
                     long bias_bstride, const void* bias_t, float* dbias_t, cudaStream_t stream) {
  if (S > 224) return OPB_ERR_UNSUPPORTED;
  if (dbias_t != nullptr && bias_t == nullptr) return OPB_ERR_INVALID;
  const int D = H * kD;
  BwdArgs a;
  a.bias = bias; a.dbias = dbias; a.bias_bstride = bias_bstride; a.s_pad = s_pad; a.key_pad = key_pad;
  a.bias_t = reinterpret_cast<const uint32_t*>(bias_t); a.dbias_t = dbias_t;
  if (bias_t != nullptr) { a.bias = nullptr; a.dbias = nullptr; }
  a.lse = lse; a.delta = delta; a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  a.B = B; a.S = S; a.H = H; a.q_scale = q_scale;
  a.n_kt = (S + kKT - 1) / kKT;
  a.n_items = static_cast<long>(B) * H;
  int qh;
  if (S <= 48) { qh = 48; a.n_qh = 1; }
  else if (S <= 80) { qh = 80; a.n_qh = 1; }
  else if (S <= 112) { qh = 112; a.n_qh = 1; }
  else { qh = 112; a.n_qh = 2; }
  CUtensorMap tq, tkv, tdo;
  int rc = make_tmap_bf16_2d(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, a.n_qh * qh);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tkv, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, kKT);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tdo, d_out, static_cast<uint64_t>(B) * S, D, D, a.n_qh * qh);
  if (rc != OPB_OK) return rc;
  switch (qh) {
    case 48: return launch_bwd<48>(tq, tkv, tdo, a, stream);
    case 80: return launch_bwd<80>(tq, tkv, tdo, a, stream);
    default: return launch_bwd<112>(tq, tkv, tdo, a, stream);
  }
}

}  // namespace opb

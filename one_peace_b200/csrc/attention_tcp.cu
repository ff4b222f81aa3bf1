// Persistent tcgen05 self-attention for sequences of up to 224 keys (vision S = 197, text, 'vl' = 214):
//     P = softmax_fp32(q k^T + relpos_bias[h] (+ -inf on padded keys)),  o = P v            (multihead_attention.py:107-115)
//
// Round-2 redesign of attention_tc.cu, from measurements on B200 (scripts/microbench, profiles/r02_tmem_ld_bw.txt):
//   * a tcgen05.ld + wait::ld round trip costs ~250 cycles per warp; TMEM itself delivers > 350 B/clk/SM.  The one-CTA-per-
//     (b, h, q-tile) kernel paid that latency ~13 times per CTA (one 16-column chunk at a time, twice) and its 3072 short-lived
//     CTAs paid barrier init / TMEM alloc / table staging / a cold TMA each: 113 us per layer with every pipe below 35 %.
//   * here ONE CTA PER SM walks a contiguous range of (head, q-tile, batch) tiles.  Per tile a soft-max warp issues ALL loads of
//     its score rows at once (one round trip), keeps them in registers (104 per thread at S = 197) and never reads them again.
//   * the relative-position bias of a (head, q-tile) is the same for every batch element: it is kept per thread as packed half2
//     (pre-multiplied by log2 e, -inf on keys >= S) in a thread-private shared-memory strip (16-byte loads, conflict-free
//     pitch; 104 score + 52 bias registers would exceed the 168-register budget of a 10-warp CTA; for 208 < S <= 224 half of it
//     stays in registers to fit the shared memory) and rebuilt from the LUT only when the CTA crosses into another (head, q-tile).
//     Per score that leaves: unpack, fma, max, sub, ex2, add, half a cvt, half a 4-byte store.
//   * tcgen05.ld.16x256b hands each thread 2 rows x 2 adjacent columns per 8-column block — the mma.sync accumulator layout — so
//     the 4 threads of a row sit in one quad: row max / row sum / LayerNorm statistics are two shuffles, no shared-memory
//     exchange, no named barriers between the soft-max warps.
//   * everything around the soft-max is pipelined across tiles by warp 0 (TMA: Q/K/V, two stages) and warp 1 (MMA issue):
//     S_{t+1} = Q K^T is accumulated into the other TMEM score buffer while tile t is exponentiated, O_{t-1} is read out and
//     stored in the middle of tile t's exponentials, P V of tile t runs under tile t+1's loads and maxima.
//
//   * P never touches shared memory: the bf16 pairs are written back over the tile's own score columns in tensor memory
//     (tcgen05.st.16x128b lands each thread's words exactly where its score fragment came from) and P V reads its A operand
//     from there (tcgen05.mma with a TMEM A operand).  No 64 KB P buffer, no generic->async proxy fence, and the exponentials
//     of tile t no longer wait for P V of tile t-1.
//
// TMEM (512 columns): S/P_even @ 0, S/P_odd @ 224, O @ 448 (64).  Shared memory: 2 x (Q 16 KB + K + V) + bias strips (52 KB at
// S = 197) + two heads' LUT rows <= 200 KB.
#include "common.cuh"
#include "ops.h"
#include "tmem_frag.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

namespace opb {

constexpr int kPQ = 128;          // query rows per tile
constexpr int kPD = 64;           // head dim
constexpr int kPS1 = 224;         // TMEM column of the odd score buffer
constexpr int kPO = 448;          // TMEM column of the output accumulator
constexpr int kPSoftWarps = 8;    // 2 per TMEM lane quarter (16 lanes each)
constexpr int kPThreads = 32 * kPSoftWarps;   // no separate producer warps (see the kernel)

struct TcpBars {
  uint64_t qk_full[2], v_full[2], s_full[2];
  uint64_t p_full, pv_done, o_free;
  uint32_t tmem_base;
};

OPB_DEVICE void named_bar_sync_all() { asm volatile("bar.sync 1, %0;" ::"n"(32 * 8) : "memory"); }

#ifdef OPB_ATTN_TIMING
__device__ unsigned long long g_tcp_t[8];
__device__ unsigned int g_tcp_n;
#define TCP_T(i) do { if (tprobe) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define TCP_T(i) do {} while (0)
#endif

// P (bf16 pairs) of blocks [LO, LO + CNT) -> tensor memory, 2 registers per block (row A word, row B word): the 16x128b store
// shape puts thread (row = lane / 4 (+ 8), packed column = 4 blk + lane % 4) exactly where the score fragment came from
template <int CNT>
OPB_DEVICE void store_p_blocks(uint32_t taddr, const uint32_t* pw) {
  static_assert(CNT == 16 || CNT == 8 || CNT == 4 || CNT == 2, "");
  if constexpr (CNT == 16) tmem_st_16x128b_x16(taddr, pw);
  if constexpr (CNT == 8) tmem_st_16x128b_x8(taddr, pw);
  if constexpr (CNT == 4) tmem_st_16x128b_x4(taddr, pw);
  if constexpr (CNT == 2) tmem_st_16x128b_x2(taddr, pw);
}

struct TcpArgs {
  const float* lut; int lut_len;
  const int* code_row; const int* code_col;
  const uint8_t* key_pad;
  __nv_bfloat16* out; float* lse; float* ln_stats;
  int B, S, H, seg_split, q_tiles;
  long n_tiles;
};

template <int NB16, bool HAS_PAD>
__global__ void __launch_bounds__(kPThreads, 1)
attention_tcp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const TcpArgs a) {
  constexpr int NPAD = NB16 * 16;
  constexpr int NBLK8 = NB16 * 2;
  constexpr uint32_t KV_BYTES = NPAD * 128;
  extern __shared__ __align__(1024) uint8_t tcp_smem_raw[];
  uint8_t* smem = tcp_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                               // 2 stages x 16 KB
  uint8_t* sK = sQ + 2 * kPQ * 128;                 // 2 stages x NPAD rows x 128 B
  uint8_t* sV = sK + 2 * KV_BYTES;                  // 2 stages
  uint32_t* sBias = reinterpret_cast<uint32_t*>(sV + 2 * KV_BYTES);            // [256 threads][2 * NBLK8 words]: bias strips
  float* sLut = reinterpret_cast<float*>(sBias + kPThreads * 2 * NBLK8);        // [2 heads][lut_len]
  int* sCc = reinterpret_cast<int*>(sLut + 2 * a.lut_len);                      // [NPAD] column codes
  TcpBars* bars = reinterpret_cast<TcpBars*>(sCc + NPAD);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S, H = a.H, B = a.B, D = a.H * kPD;
  // contiguous tile range of this CTA; tile w = ((h * q_tiles + qt) * B + b)
  const long w0 = a.n_tiles * blockIdx.x / gridDim.x, w1 = a.n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int n = static_cast<int>(w1 - w0);
  const int h_first = static_cast<int>(w0 / B) / a.q_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qk_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->s_full[i], 1);
    }
    mbar_init(&bars->p_full, kPSoftWarps);
    mbar_init(&bars->pv_done, 1);
    mbar_init(&bars->o_free, kPSoftWarps);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<1>(&bars->tmem_base, 512);
  }
  // A CTA's <= ~21 consecutive tiles touch at most two heads (B >= 11 tiles per (head, q-tile) is checked by the host): stage
  // their LUT rows and the column codes once; the per-(head, q-tile) bias rebuild then gathers from shared memory only.
  for (int i = threadIdx.x; i < 2 * a.lut_len; i += kPThreads) {
    const int hh2 = min(h_first + i / a.lut_len, H - 1);
    sLut[i] = __ldg(a.lut + static_cast<long>(hh2) * a.lut_len + i % a.lut_len);
  }
  for (int i = threadIdx.x; i < NPAD; i += kPThreads) sCc[i] = __ldg(a.code_col + min(i, S - 1));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // ---- producer duties, carried by lane 0 of warp 0 (TMA) and lane 0 of warp 1 (MMA issue) at fixed points of the tile loop.
  // A separate producer warp would make the CTA 9-10 warps, which the register allocator treats as 12: 168 registers per
  // thread instead of 255 — not enough for 104 score registers plus working set (measured: spills, serialised code, 150 us). ----
  const bool tma_thread = threadIdx.x == 0, mma_thread = threadIdx.x == 32;
  auto tile_coords = [&](int j, int& b, int& h, int& qt) {
    const long w = w0 + j;
    b = static_cast<int>(w % B);
    const int g = static_cast<int>(w / B);
    qt = g % a.q_tiles;
    h = g / a.q_tiles;
  };
  auto load_qk = [&](int j) {                       // stage j & 1 must be free: Q K^T of tile j - 2 has completed
    int b, h, qt;
    tile_coords(j, b, h, qt);
    const int st = j & 1;
    mbar_arrive_expect_tx(&bars->qk_full[st], kPQ * 128 + KV_BYTES);
    tma_load_2d(&tm_q, &bars->qk_full[st], sQ + st * kPQ * 128, h * kPD, b * S + qt * kPQ);
    tma_load_2d(&tm_kv, &bars->qk_full[st], sK + st * KV_BYTES, D + h * kPD, b * S);
  };
  auto load_v = [&](int j) {                        // stage j & 1 must be free: P V of tile j - 2 has completed
    int b, h, qt;
    tile_coords(j, b, h, qt);
    const int st = j & 1;
    mbar_arrive_expect_tx(&bars->v_full[st], KV_BYTES);
    tma_load_2d(&tm_kv, &bars->v_full[st], sV + st * KV_BYTES, 2 * D + h * kPD, b * S);
  };
  auto issue_qk = [&](int j) {                      // score buffer j & 1 held P of tile j - 2: its P V must have completed
    constexpr uint32_t idesc_qk = make_idesc_bf16(kPQ, NPAD);
    const int st = j & 1;
    mbar_wait(&bars->qk_full[st], (j >> 1) & 1);
    if (j >= 2) mbar_wait(&bars->pv_done, (j - 2) & 1);
    tc_fence_after();
    const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ + st * kPQ * 128));
    const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + st * KV_BYTES));
#pragma unroll
    for (int kk = 0; kk < kPD / 16; ++kk) umma_bf16<1>(tmem_base + st * kPS1, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk != 0);
    umma_commit<1>(&bars->s_full[st]);
  };
  auto issue_pv = [&](int j) {
    constexpr uint32_t idesc_pv = make_idesc_bf16(kPQ, kPD) | (1u << 16);      // B (= V) MN-major
    const int st = j & 1;
    mbar_wait(&bars->p_full, j & 1);
    if (j > 0) mbar_wait(&bars->o_free, (j - 1) & 1);
    mbar_wait(&bars->v_full[st], (j >> 1) & 1);
    tc_fence_after();
    const uint32_t vbase = smem_u32(sV + st * KV_BYTES);
#pragma unroll
    for (int kk = 0; kk < NB16; ++kk) {
      // A = P from tensor memory: 16 keys = 8 packed columns of the tile's score buffer.  B = V (MN-major): 16 keys = 16 rows
      const uint64_t dv = make_sw128_mn_desc64(vbase + kk * 16 * 128);
      umma_bf16_ts(tmem_base + kPO, tmem_base + st * kPS1 + 8 * kk, dv, idesc_pv, kk != 0);
    }
    umma_commit<1>(&bars->pv_done);
  };
  if (tma_thread) {
    for (int j = 0; j < 2 && j < n; ++j) { load_qk(j); load_v(j); }
  }
  if (mma_thread) {
    for (int j = 0; j < 2 && j < n; ++j) issue_qk(j);
  }
  __syncwarp();

  // ===================== soft-max (all 8 warps) =====================
  const int qw = warp & 3;                          // TMEM lane quarter (hardware: warp id % 4)
  const int hh = warp >> 2;                         // which 16 lanes of the quarter
  const int q = lane & 3;                           // position inside the row's quad
  const int rA = qw * 32 + hh * 16 + (lane >> 2);   // tile rows of this thread: rA and rA + 8
  const uint32_t lane_addr = static_cast<uint32_t>(qw * 32 + hh * 16) << 16;
  constexpr float kLog2e = 1.4426950408889634f;

  // the thread's bias strip: per 8-column block two half2 words (row rA, row rA + 8) = (bias * log2e) of columns
  // 8 blk + 2 q + {0, 1}, -inf for keys >= S.  Pitch 2 * NBLK8 words (4 * odd for 26 blocks: conflict-free 16-byte loads).
  uint32_t* strip = sBias + threadIdx.x * (2 * NBLK8);
  // deferred epilogue state (tile it - 1)
  float lA_prev = 0.f, lB_prev = 0.f, mA_prev = 0.f, mB_prev = 0.f;
  int b_prev = 0, h_prev = 0, q0_prev = 0;

  auto epilogue_store = [&](const uint32_t (&o)[32], int b, int h, int q0, float lA, float lB, float mA, float mB) {
    const long rows_total = static_cast<long>(B) * S;
#pragma unroll
    for (int rsel = 0; rsel < 2; ++rsel) {
      const int qrow = q0 + rA + 8 * rsel;
      const float l = rsel ? lB : lA, m2 = rsel ? mB : mA;
      const float inv = l > 0.f ? 1.f / l : 0.f;
      float ssum = 0.f, ssq = 0.f;
      uint32_t pk[8];
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {
        const float y0 = __uint_as_float(o[4 * blk + 2 * rsel]) * inv, y1 = __uint_as_float(o[4 * blk + 2 * rsel + 1]) * inv;
        ssum += y0 + y1;
        ssq = fmaf(y0, y0, fmaf(y1, y1, ssq));
        pk[blk] = pack_bf16x2(y0, y1);
      }
      ssum += __shfl_xor_sync(0xffffffffu, ssum, 1); ssq += __shfl_xor_sync(0xffffffffu, ssq, 1);
      ssum += __shfl_xor_sync(0xffffffffu, ssum, 2); ssq += __shfl_xor_sync(0xffffffffu, ssq, 2);
      if (qrow < S) {
        uint32_t* op = reinterpret_cast<uint32_t*>(a.out + (static_cast<long>(b) * S + qrow) * D + h * kPD + 2 * q);
#pragma unroll
        for (int blk = 0; blk < 8; ++blk) op[4 * blk] = pk[blk];
        if (q == 0) {
          // log-sum-exp of the biased scores (natural log), kept for the backward pass like attention.cu does
          if (a.lse != nullptr) a.lse[(static_cast<long>(b) * H + h) * S + qrow] = m2 * 0.6931471805599453f + __logf(l);
          if (a.ln_stats != nullptr)
            *reinterpret_cast<float2*>(a.ln_stats + (h * rows_total + static_cast<long>(b) * S + qrow) * 2) = make_float2(ssum, ssq);
        }
      }
    }
  };

  // (head, q-tile) changed: rebuild the thread's strip from the staged LUT (bias[h][i][j] = lut[h][code_row[i] - code_col[j]],
  // block-diagonal for concatenated 'vl' / 'al' sequences, transformer_encoder.py:148-158).  Thread-private: no barrier.
  auto rebuild_bias = [&](int h, int q0) {
    const float* s_lut = sLut + (h - h_first) * a.lut_len;
#pragma unroll
    for (int rsel = 0; rsel < 2; ++rsel) {
      const int qrow = min(q0 + rA + 8 * rsel, S - 1);
      const int crow = __ldg(a.code_row + qrow);
      if (a.seg_split == 0) {
        // single-modality fast path: no segment test (sCc is clamped to the last valid column beyond S)
#pragma unroll
        for (int blk = 0; blk < NBLK8; ++blk) {
          const int2 cc = *reinterpret_cast<const int2*>(sCc + 8 * blk + 2 * q);
          float b0 = s_lut[crow - cc.x] * kLog2e, b1 = s_lut[crow - cc.y] * kLog2e;
          if (8 * blk + 8 > S) {                     // warp-uniform: blocks that reach past the sequence
            b0 = 8 * blk + 2 * q >= S ? -INFINITY : b0;
            b1 = 8 * blk + 2 * q + 1 >= S ? -INFINITY : b1;
          }
          const __half2 hb = __floats2half2_rn(b0, b1);
          strip[2 * blk + rsel] = *reinterpret_cast<const uint32_t*>(&hb);
        }
      } else {
        const int seg_lo = qrow >= a.seg_split ? a.seg_split : 0;
        const int seg_hi = qrow < a.seg_split ? a.seg_split : S;
#pragma unroll
        for (int blk = 0; blk < NBLK8; ++blk) {
          float b2[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = 8 * blk + 2 * q + j;
            const bool in_seg = c >= seg_lo && c < seg_hi;
            const float bv = s_lut[in_seg ? crow - sCc[c] : 0];        // clamped index + select, never a conditional load
            b2[j] = c >= S ? -INFINITY : (in_seg ? bv * kLog2e : 0.f);
          }
          const __half2 hb = __floats2half2_rn(b2[0], b2[1]);
          strip[2 * blk + rsel] = *reinterpret_cast<const uint32_t*>(&hb);
        }
      }
    }
  };

  // tile coordinates, advanced incrementally (no 64-bit division per tile)
  int b = static_cast<int>(w0 % B), g = static_cast<int>(w0 / B);
  int qt = g % a.q_tiles, h = g / a.q_tiles;
  bool fresh = true;
#ifdef OPB_ATTN_TIMING
  const bool tprobe = warp == 2 && lane == 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
  for (int it = 0; it < n; ++it) {
    const int q0 = qt * kPQ;
    const bool warp_valid = q0 + qw * 32 + hh * 16 < S;      // warp-uniform
    if (fresh) {
      rebuild_bias(h, q0);
      fresh = false;
    }
    uint32_t pmask[HAS_PAD ? (NPAD + 31) / 32 : 1];
    if constexpr (HAS_PAD) {
#pragma unroll
      for (int wd = 0; wd < (NPAD + 31) / 32; ++wd) {
        const int c = wd * 32 + lane;
        const bool pad = c < S && a.key_pad[static_cast<long>(b) * S + c] != 0;
        pmask[wd] = __ballot_sync(0xffffffffu, pad);
      }
    }
    // ---- scores of tile `it` -> registers (one TMEM round trip) ----
    const int sb = it & 1;
    const uint32_t s_addr = tmem_base + lane_addr + sb * kPS1;
    uint32_t v[4 * NBLK8];
    TCP_T(0);                                      // bias rebuild / pad mask / bookkeeping
    mbar_wait(&bars->s_full[sb], (it >> 1) & 1);
    tc_fence_after();
    TCP_T(1);                                      // waiting for S
    // Q K^T of tile `it` has completed, so its Q / K stage is free: fetch tile it + 2
    if (tma_thread && it + 2 < n) load_qk(it + 2);
    __syncwarp();
    if (warp_valid) {
      load_scores<NBLK8>(s_addr, v);
      tmem_ld_wait();
    }
    TCP_T(2);                                      // TMEM round trip

    // ---- t = s log2e + bias log2e, row maxima ----
    float mA = -INFINITY, mB = -INFINITY;
    if (warp_valid) {
      uint4 sb4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int blk = 0; blk < NBLK8; ++blk) {
        if ((blk & 1) == 0) sb4 = *reinterpret_cast<const uint4*>(strip + 2 * blk);      // 2 blocks per 16-byte load
        const uint32_t wa = (blk & 1) ? sb4.z : sb4.x, wb = (blk & 1) ? sb4.w : sb4.y;
        const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&wa));
        const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(&wb));
        float t0 = fmaf(__uint_as_float(v[4 * blk + 0]), kLog2e, ba.x), t1 = fmaf(__uint_as_float(v[4 * blk + 1]), kLog2e, ba.y);
        float t2 = fmaf(__uint_as_float(v[4 * blk + 2]), kLog2e, bb.x), t3 = fmaf(__uint_as_float(v[4 * blk + 3]), kLog2e, bb.y);
        if constexpr (HAS_PAD) {
          const uint32_t bits = pmask[(8 * blk) >> 5] >> (((8 * blk) & 31) + 2 * q);
          t0 = (bits & 1u) ? -INFINITY : t0; t2 = (bits & 1u) ? -INFINITY : t2;
          t1 = (bits & 2u) ? -INFINITY : t1; t3 = (bits & 2u) ? -INFINITY : t3;
        }
        v[4 * blk + 0] = __float_as_uint(t0); v[4 * blk + 1] = __float_as_uint(t1);
        v[4 * blk + 2] = __float_as_uint(t2); v[4 * blk + 3] = __float_as_uint(t3);
        mA = fmaxf(mA, fmaxf(t0, t1));
        mB = fmaxf(mB, fmaxf(t2, t3));
      }
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1)); mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2)); mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
      if (mA == -INFINITY) mA = 0.f;               // every key masked: p = 0, l = 0, output row 0
      if (mB == -INFINITY) mB = 0.f;
    }
    // the score buffer of tile it + 1 held P of tile it - 1 (its P V was issued a tile ago), Q / K of tile it + 1 were
    // requested a tile ago: accumulate S_{it+1} now
    if (mma_thread && it >= 1 && it + 1 < n) issue_qk(it + 1);
    __syncwarp();
    TCP_T(3);                                      // t values + maxima
    // ---- p = 2^(t - m), row sums, P (bf16 pairs) -> tensor memory, over the scores of this tile (they are in registers);
    //      O_{it-1} is fetched on the way ----
    float lA = 0.f, lB = 0.f;
    uint32_t o[32];
    const bool prev_valid = it > 0 && q0_prev + qw * 32 + hh * 16 < S;
    auto exp_blocks = [&](auto lo_c, auto cnt_c) {
      constexpr int LO = decltype(lo_c)::value, CNT = decltype(cnt_c)::value;
      uint32_t pw[2 * CNT];
#pragma unroll
      for (int k = 0; k < CNT; ++k) {
        const int blk = LO + k;
        const float p0 = ex2_fast(__uint_as_float(v[4 * blk + 0]) - mA), p1 = ex2_fast(__uint_as_float(v[4 * blk + 1]) - mA);
        const float p2 = ex2_fast(__uint_as_float(v[4 * blk + 2]) - mB), p3 = ex2_fast(__uint_as_float(v[4 * blk + 3]) - mB);
        lA += p0 + p1;
        lB += p2 + p3;
        pw[2 * k] = pack_bf16x2(p0, p1);
        pw[2 * k + 1] = pack_bf16x2(p2, p3);
      }
      store_p_blocks<CNT>(s_addr + 4 * LO, pw);
    };
    constexpr int n16 = NBLK8 & 16, n8 = NBLK8 & 8, n4 = NBLK8 & 4, n2 = NBLK8 & 2;
    if (warp_valid) {
      if constexpr (n16 != 0) exp_blocks(std::integral_constant<int, 0>{}, std::integral_constant<int, 16>{});
      else if constexpr (n8 != 0) exp_blocks(std::integral_constant<int, 0>{}, std::integral_constant<int, 8>{});
    }
    TCP_T(4);                                      // first part of the exponentials
    if (it > 0) {
      // O_{it-1} is complete once P V of tile it-1 has finished; its V stage is free for tile it + 1
      mbar_wait(&bars->pv_done, (it - 1) & 1);
      tc_fence_after();
      if (tma_thread && it + 1 < n) load_v(it + 1);
      __syncwarp();
      if (prev_valid) tmem_ld_16x256b_x8(tmem_base + lane_addr + kPO, o);
    }
    if (warp_valid) {
      if constexpr (n16 != 0 && n8 != 0) exp_blocks(std::integral_constant<int, n16>{}, std::integral_constant<int, 8>{});
      if constexpr (n4 != 0) exp_blocks(std::integral_constant<int, n16 + n8>{}, std::integral_constant<int, 4>{});
      if constexpr (n2 != 0) exp_blocks(std::integral_constant<int, n16 + n8 + n4>{}, std::integral_constant<int, 2>{});
      tmem_st_wait_all();
    }
    TCP_T(5);                                      // exponentials + P stores
    if (prev_valid) tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&bars->p_full);
      if (it > 0) mbar_arrive(&bars->o_free);
    }
    if (mma_thread) issue_pv(it);                  // waits for the other warps' P rows; the rest of warp 1 waits at the shuffles below
    __syncwarp();
    lA += __shfl_xor_sync(0xffffffffu, lA, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 1);
    lA += __shfl_xor_sync(0xffffffffu, lA, 2); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    if (prev_valid) epilogue_store(o, b_prev, h_prev, q0_prev, lA_prev, lB_prev, mA_prev, mB_prev);
    lA_prev = lA; lB_prev = lB; mA_prev = mA; mB_prev = mB;
    b_prev = b; h_prev = h; q0_prev = q0;
    if (++b == B) {                                // next tile: next batch element, or the next (head, q-tile)
      b = 0;
      ++g;
      qt = g % a.q_tiles;
      h = g / a.q_tiles;
      fresh = true;
    }
    TCP_T(6);                                      // O wait + arrivals + epilogue of the previous tile
  }
#ifdef OPB_ATTN_TIMING
  if (tprobe) {
    for (int i = 0; i < 7; ++i) atomicAdd(&g_tcp_t[i], static_cast<unsigned long long>(tacc[i]));
    atomicAdd(&g_tcp_n, static_cast<unsigned int>(n));
  }
#endif
  if (n > 0) {
    mbar_wait(&bars->pv_done, (n - 1) & 1);
    tc_fence_after();
    if (q0_prev + qw * 32 + hh * 16 < S) {
      uint32_t o[32];
      tmem_ld_16x256b_x8(tmem_base + lane_addr + kPO, o);
      tmem_ld_wait();
      epilogue_store(o, b_prev, h_prev, q0_prev, lA_prev, lB_prev, mA_prev, mB_prev);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

#ifdef OPB_ATTN_TIMING
extern "C" void opb_tcp_timing_dump() {
  unsigned long long t[8]; unsigned int n;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(t, g_tcp_t, sizeof(t)); cudaMemcpyFromSymbol(&n, g_tcp_n, sizeof(n));
  printf("[tcp timing] tiles=%u  avg clocks per tile (warp 2): setup=%.0f wait_S=%.0f tmem_ld=%.0f tvals_max=%.0f wait_PV=%.0f exp_store=%.0f "
         "o_wait_epilogue=%.0f\n", n, (double)t[0] / n, (double)t[1] / n, (double)t[2] / n, (double)t[3] / n, (double)t[4] / n,
         (double)t[5] / n, (double)t[6] / n);
}
#endif

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

template <int NB16>
static int launch_tcp(const CUtensorMap& tq, const CUtensorMap& tkv, const TcpArgs& a, bool has_pad, cudaStream_t stream) {
  constexpr int NPAD = NB16 * 16;
  const size_t smem = 2ull * kPQ * 128 + 4ull * NPAD * 128 + static_cast<size_t>(kPThreads) * 4 * NB16 * 4 +
                      2ull * a.lut_len * 4 + NPAD * 4 + sizeof(TcpBars);
  if (smem > 227 * 1024) return OPB_ERR_UNSUPPORTED;
  static size_t configured[2] = {0, 0};
  auto kern = has_pad ? attention_tcp_kernel<NB16, true> : attention_tcp_kernel<NB16, false>;
  if (smem > configured[has_pad]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return OPB_ERR_CUDA;
    configured[has_pad] = smem;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned grid = static_cast<unsigned>(a.n_tiles < sms ? a.n_tiles : sms);
  return launch_maybe_cluster(kern, dim3(grid), dim3(kPThreads), smem, stream, tq, tkv, a) == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// Same contract as attention_tc_fwd (ops.h); S <= 224.  Returns OPB_ERR_UNSUPPORTED for longer sequences.
int attention_tcp_fwd(const void* qkv, const float* lut, int lut_len, const int* code_row, const int* code_col,
                      const uint8_t* key_pad, void* out, float* lse, float* ln_stats, int B, int S, int H, int seg_split,
                      cudaStream_t stream) {
  if (B <= 0 || S <= 0 || H <= 0 || lut == nullptr || code_row == nullptr || code_col == nullptr) return OPB_ERR_INVALID;
  if (seg_split < 0 || seg_split >= S) return OPB_ERR_INVALID;
  if (S > 224) return OPB_ERR_UNSUPPORTED;
  const int nb16 = (S + 15) / 16;
  const int inst = nb16 <= 2 ? 2 : nb16 <= 4 ? 4 : nb16 <= 6 ? 6 : nb16 <= 9 ? 9 : nb16 <= 13 ? 13 : 14;
  // a CTA stages the LUT rows of two consecutive heads: its tile range must not span a third one, and the rows are 16-byte
  // multiples so that the strips after them stay aligned
  if (lut_len % 4 != 0) return OPB_ERR_INVALID;
  if (H > 128) return OPB_ERR_UNSUPPORTED;
  const int D = H * kPD;
  CUtensorMap tq, tkv;
  int rc = make_tmap_bf16_2d(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, kPQ);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tkv, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, inst * 16);
  if (rc != OPB_OK) return rc;
  TcpArgs a;
  a.lut = lut; a.lut_len = lut_len; a.code_row = code_row; a.code_col = code_col; a.key_pad = key_pad;
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse; a.ln_stats = ln_stats;
  a.B = B; a.S = S; a.H = H; a.seg_split = seg_split; a.q_tiles = (S + kPQ - 1) / kPQ;
  a.n_tiles = static_cast<long>(B) * H * a.q_tiles;
  const bool hp = key_pad != nullptr;
  switch (inst) {
    case 2: return launch_tcp<2>(tq, tkv, a, hp, stream);
    case 4: return launch_tcp<4>(tq, tkv, a, hp, stream);
    case 6: return launch_tcp<6>(tq, tkv, a, hp, stream);
    case 9: return launch_tcp<9>(tq, tkv, a, hp, stream);
    case 13: return launch_tcp<13>(tq, tkv, a, hp, stream);
    default: return launch_tcp<14>(tq, tkv, a, hp, stream);
  }
}

}  // namespace opb

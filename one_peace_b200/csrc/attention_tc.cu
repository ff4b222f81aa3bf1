// tcgen05 self-attention for sequences of up to 384 keys (vision S = 197 / 257, text S <= 72):
//     P = softmax_fp32(q k^T + relpos_bias[h] (+ -inf on padded keys)),  o = P v            (multihead_attention.py:107-115)
//
// One CTA = (batch, head, 128-query tile).  Warp 0 issues TMA (Q, all K blocks, all V blocks of this (b, h) straight
// from the QKV GEMM output), warp 1 issues the MMAs, warps 2-5 own one query row per thread:
//   S_kb = Q K_kb^T            tcgen05.mma  M=128 N=128 K=64, accumulators in TMEM (one 128-column slot per key block)
//   phase A  (row thread)      tcgen05.ld S, max tree -> m = max_j s_ij + max(lut[h]): an UPPER BOUND of the biased row
//                              maximum (soft-max is shift invariant; the slack is at most the spread of the head's bias
//                              table) — no LUT gather, no add, no tcgen05.st here (round 2; round 1 did all three and
//                              spent a third of the CTA's time in this phase)
//   phase B  (row thread)      tcgen05.ld S, gather the biases, p = exp2((s + bias) log2e - m log2e) (-inf on padded keys),
//                              row sum, P (bf16) -> shared memory in the K-major 128B-swizzled layout the next MMA reads
//   O += P_kb V_kb             tcgen05.mma  M=128 N=64 K=128, V consumed as an MN-major operand exactly as TMA wrote
//                              it ([key][d] rows of 128 B) — no transpose;  O aliases the first 64 columns of S_0
//   epilogue (row thread)      tcgen05.ld O, scale by 1/l, bf16, 128-byte row store (+ inner-LN partial statistics)
//
// Relative-position bias without the (H,S,S) table: every ONE-PEACE bias is table[bucket(i, j)] with the bucket a
// function of a per-position code difference (text / audio: i - j; image: 2-D offset) plus three CLS ids
// (adapter/text.py:18-29,62-68, adapter/image.py:19-34).  The host folds that into a per-head 1-D LUT with
//     bias[h][i][j] = lut[h][code_row[i] - code_col[j]]
// (CLS row / column / corner are mapped to constant LUT regions by offsetting code_row[0] / code_col[0]), so a score
// costs one shared-memory gather instead of a 4-byte read of a 1.9 MB table per (batch, head) — that L2 stream and the
// shuffle / barrier-bound online softmax made the mma.sync kernel (attention.cu) latency-bound at ~200 us per layer.
//
// Shared memory 112 KB (Q 16 + K 32 + V 32 + P 32; the LUT and code tables land in the P buffer and move into the dead
// Q tile between the phases) and
// 256 TMEM columns per CTA -> two CTAs per SM overlap each other's TMA / MMA / softmax phases.
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace opb {

constexpr int kTcQ = 128;        // query rows per CTA
constexpr int kTcK = 128;        // keys per block
constexpr int kTcD = 64;         // head dim
constexpr int kTcMaxBlocks = 3;  // S <= 384

struct TcBars {
  uint64_t qk;                   // Q + all K landed
  uint64_t v;                    // all V landed
  uint64_t s[kTcMaxBlocks];      // S_kb accumulated
  uint64_t p[kTcMaxBlocks];      // P_kb written by the 4 row warps
  uint64_t pv[kTcMaxBlocks];     // P_kb V_kb accumulated (P buffer reusable; after the last: O complete)
  uint64_t lut;                  // LUT + column codes landed (bulk copy)
  uint32_t tmem_base;
};

OPB_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));     // not volatile: the 16 exponentials of a sub-chunk must be free to issue back to back
  return y;
}
OPB_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
OPB_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
OPB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
OPB_DEVICE void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// MN-major bf16 operand stored as [k rows][64 elements = 128 B] with the 128-byte swizzle (what TMA writes for a
// {64, rows} box): 8-row groups are 1024 B apart (SBO); a single 64-wide MN chunk, so LBO is unused.
OPB_DEVICE uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int umma_m, int umma_n) {
  return make_idesc_bf16(umma_m, umma_n) | (1u << 16);     // B operand MN-major
}

#ifdef OPB_ATTN_TIMING
__device__ unsigned long long g_attn_t[8];
__device__ unsigned int g_attn_n;
#define OPB_T(i) do { if (threadIdx.x == 64) tt[i] = globaltimer_ns(); } while (0)
#else
#define OPB_T(i) do {} while (0)
#endif

constexpr int kTcTableBytes = 13 * 1024;   // LUT + codes (+ pad mask) inside the Q tile; the remaining 3 KB: row-thread exchange

template <bool HAS_PAD, int HALVES>
__global__ void __launch_bounds__(64 + 128 * HALVES, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const float* __restrict__ lut, const float* __restrict__ lut_max, int lut_len,
                    const int* __restrict__ code_row, const int* __restrict__ code_col,
                    const uint8_t* __restrict__ key_pad, __nv_bfloat16* __restrict__ out,
                    float* __restrict__ lse, float* __restrict__ ln_stats, int B, int S, int H, int nkb, uint32_t tmem_cols,
                    int seg_split, int k0, int Sk, int big_tables) {
  // k0 / Sk: the launch covers keys [k0, k0 + Sk) of every sample (Sk == S, k0 == 0 unless the sequence is split over several
  // launches whose partial results are merged by attention_merge_kernel — 384 < S <= 768, the 15 s audio sequences).  Queries
  // are always all S rows; `code_col` and `key_pad` are indexed by the LOCAL key (the host passes code_col + k0).
  // no static shared memory in this kernel: the dynamic window starts at the (1024-aligned) base of the CTA's shared
  // memory.  The 112 KB + barriers must fit twice per SM, so there is no room for alignment slack; verify instead.
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTcQ * 128;
  uint8_t* sV = sK + nkb * kTcK * 128;
  uint8_t* sP = sV + nkb * kTcK * 128;                  // 32 KB: LUT + codes (+ pad mask) in phase A, P afterwards
  TcBars* bars = reinterpret_cast<TcBars*>(sP + kTcQ * kTcK * 2);
  // big_tables (long sequences: the LUT of a 750-token audio sequence is 18 KB): the tables get their own region behind the
  // barriers instead of travelling P buffer -> dead Q tile; the Q tile then only holds the row threads' exchange area.
  uint8_t* sT = reinterpret_cast<uint8_t*>(bars) + 256;

  const int q_tiles = (S + kTcQ - 1) / kTcQ;
  const int qt = blockIdx.x % q_tiles;
  const int h = (blockIdx.x / q_tiles) % H;
  const int b = blockIdx.x / (q_tiles * H);
  const int D = H * kTcD;
  const int q0 = qt * kTcQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef OPB_ATTN_TIMING
  unsigned long long tt[8];
  OPB_T(0);
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(&bars->qk, 1);
    mbar_init(&bars->v, 1);
    mbar_init(&bars->lut, 1);
    for (int i = 0; i < kTcMaxBlocks; ++i) {
      mbar_init(&bars->s[i], 1);
      mbar_init(&bars->p[i], 4 * HALVES);
      mbar_init(&bars->pv[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<1>(&bars->tmem_base, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = b * S;
      // phase-A tables: lut [lut_len] | code_col [s4] (both padded to 16-byte multiples by the host)
      const int s4 = (Sk + 3) & ~3;
      mbar_arrive_expect_tx(&bars->lut, static_cast<uint32_t>(lut_len + s4) * 4);
      uint8_t* tdst = big_tables ? sT : sP;
      bulk_load_1d(tdst, lut + static_cast<long>(h) * lut_len, static_cast<uint32_t>(lut_len) * 4, &bars->lut);
      bulk_load_1d(tdst + static_cast<long>(lut_len) * 4, code_col, static_cast<uint32_t>(s4) * 4, &bars->lut);
      mbar_arrive_expect_tx(&bars->qk, (kTcQ + nkb * kTcK) * 128);
      tma_load_2d(&tm_qkv, &bars->qk, sQ, h * kTcD, row0 + q0);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&tm_qkv, &bars->qk, sK + kb * kTcK * 128, D + h * kTcD, row0 + k0 + kb * kTcK);
      mbar_arrive_expect_tx(&bars->v, nkb * kTcK * 128);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&tm_qkv, &bars->v, sV + kb * kTcK * 128, 2 * D + h * kTcD, row0 + k0 + kb * kTcK);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTcQ, kTcK);
      constexpr uint32_t idesc_pv = make_idesc_bf16_bmn(kTcQ, kTcD);
      mbar_wait(&bars->qk, 0);
      tc_fence_after();
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + kb * kTcK * 128));
#pragma unroll
        for (int k = 0; k < kTcD / 16; ++k) umma_bf16<1>(tmem_base + kb * kTcK, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit<1>(&bars->s[kb]);
      }
      mbar_wait(&bars->v, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bars->p[kb], 0);
        tc_fence_after();
        const uint32_t pbase = smem_u32(sP);
        const uint32_t vbase = smem_u32(sV + kb * kTcK * 128);
#pragma unroll
        for (int k = 0; k < kTcK / 16; ++k) {
          // A = P (K-major): two 64-key atoms of [128 rows][128 B]; +32 B per 16 keys inside an atom
          const uint64_t dp = make_sw128_kmajor_desc(pbase + (k >> 2) * (kTcQ * 128) + (k & 3) * 32);
          // B = V (MN-major): 16 keys = 16 rows of 128 B
          const uint64_t dv = make_sw128_mnmajor_desc(vbase + k * 16 * 128);
          umma_bf16<1>(tmem_base, dp, dv, idesc_pv, (kb | k) != 0);
        }
        umma_commit<1>(&bars->pv[kb]);
      }
    }
  } else {
    // ===================== row threads =====================
    // HALVES threads per query row (the HALVES warps that share a TMEM lane quarter); thread `half` owns the 16-key
    // sub-chunks t with t % HALVES == half.  HALVES = 2 doubles the soft-max warps per SM (16): every pipe of this kernel
    // runs below 35 % (ncu, profiles/r02_ncu_layer.summary.txt) — it is bound by the dependent ld -> gather -> ex2 -> store
    // chain with too few warps to interleave.
    constexpr int kRowThreads = 128 * HALVES;
    const int qw = warp & 3;                                 // TMEM lane quarter (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;
    const int r = qw * 32 + lane;                            // row inside the tile
    const int qrow = q0 + r;
    const bool row_valid = qrow < S;
    const bool warp_valid = (q0 + qw * 32) < S;              // warp-uniform
    const int tid4 = threadIdx.x - 64;                       // 0 .. kRowThreads - 1
    // bias tables: lut [lut_len] | code_col [S padded to 4] (| key_pad bytes).  They land in the P buffer (bulk copy at
    // kernel start) and are moved into the Q tile once the last S = Q K^T has completed (Q is dead then), because phase B
    // gathers from them while it fills the P buffer.  The last 3 KB of the Q tile are the exchange area of a row's threads.
    const int tbl_words = lut_len + ((Sk + 3) & ~3);
    const int pad_bytes = HAS_PAD ? ((Sk + 31) & ~31) : 0;
    const int tbl_bytes = (tbl_words * 4 + pad_bytes + 15) & ~15;
    const uint8_t* tbl = big_tables ? sT : sQ;
    const float* s_lut = reinterpret_cast<const float*>(tbl);
    const int* s_ccol = reinterpret_cast<const int*>(tbl) + lut_len;
    const uint8_t* s_pad_w = tbl + static_cast<long>(tbl_words) * 4;   // key-padding bytes, zero beyond S up to a 32-key boundary
    float* xch = reinterpret_cast<float*>(sQ + kTcTableBytes);         // [max | sum][half][row]
    float2* xstat = reinterpret_cast<float2*>(sQ + kTcTableBytes + 4 * kTcQ * 4);
    if constexpr (HAS_PAD) {
      uint8_t* pad_in = (big_tables ? sT : sP) + static_cast<long>(tbl_words) * 4;
      for (int i = tid4; i < pad_bytes; i += kRowThreads) pad_in[i] = i < Sk ? key_pad[static_cast<long>(b) * S + k0 + i] : 0;
    }
    const int crow = code_row[row_valid ? qrow : 0];
    // concatenated sequences ('vl' / 'al', transformer_encoder.py:148-158): the relative-position bias is block-diagonal —
    // a row only sees the bias of the keys of its own modality segment [seg_lo, seg_hi); zero across segments
    const int seg_lo = (seg_split > 0 && qrow >= seg_split) ? seg_split : 0;
    const int seg_hi = (seg_split > 0 && qrow < seg_split) ? seg_split : Sk;      // (seg_split > 0 only with k0 == 0, Sk == S)
    OPB_T(1);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(qw * 32) << 16);
    const int nsub = nkb * (kTcK / 16);

    // ---- phase A: an upper bound of the row maximum, WITHOUT touching the bias ----
    // soft-max is shift invariant, so any m >= max_j (s_ij + bias_ij) that is not absurdly loose gives the same result: bf16
    // P and the fp32 row sum / O accumulators keep full relative precision for exp(x - m) down to ~1e-38.  We use
    //     m = max_j s_ij  +  max_l lut[h][l]        (slack <= spread of the head's bias table, a few units)
    // which needs only tcgen05.ld + a max tree: no LUT gather, no add, and no tcgen05.st of biased scores back into TMEM
    // (round 1 did all three here and loaded the scores again in phase B).  Padded keys and the neighbouring sample's rows
    // inside the last block may enter the bound; they only loosen it.
    float m = -INFINITY;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&bars->s[kb], 0);
      tc_fence_after();
      if (kb == 0) OPB_T(2);
      if (warp_valid) {
        const int kvalid = min(kTcK, Sk - kb * kTcK);
        for (int c = half * 16; c < kvalid; c += 16 * HALVES) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(lane_base + kb * kTcK + c, v);
          tmem_ld_wait();
          float t4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            t4[j] = fmaxf(fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 4])), fmaxf(__uint_as_float(v[j + 8]), __uint_as_float(v[j + 12])));
          m = fmaxf(m, fmaxf(fmaxf(t4[0], t4[1]), fmaxf(t4[2], t4[3])));
        }
      }
    }
    // every S_kb has completed: Q is dead.  Exchange the partial maxima, move the tables P buffer -> Q tile (16-byte
    // pieces); after the barrier the P buffer is free.
    if constexpr (HALVES > 1) xch[half * kTcQ + r] = m;
    mbar_wait(&bars->lut, 0);
    if constexpr (HAS_PAD) named_bar_sync(1, kRowThreads);       // pad bytes written by other threads
    if (!big_tables)
      for (int i = tid4 * 16; i < tbl_bytes; i += kRowThreads * 16) sts128u(sQ + i, *reinterpret_cast<const uint4*>(sP + i));
    named_bar_sync(1, kRowThreads);
    if constexpr (HALVES > 1) m = fmaxf(m, xch[(half ^ 1) * kTcQ + r]);
    m += lut_max[h];
    OPB_T(3);

    // ---- phase B: bias gather, exp, row sum, P -> shared memory (16-key sub-chunks) ----
    const float mb = m * 1.4426950408889634f;
    float l0 = 0.f, l1 = 0.f;
    auto prefetch = [&](int t, uint32_t (&v)[16], float (&add)[16]) {
      const int key0 = t * 16;                         // TMEM column == key index (128-column slot per key block)
      __syncwarp();
      tmem_ld16(lane_base + key0, v);
      int ii[16];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int4 cc = *reinterpret_cast<const int4*>(s_ccol + key0 + j);     // keys past S: padded / stale but in-bounds
        ii[j] = crow - cc.x; ii[j + 1] = crow - cc.y; ii[j + 2] = crow - cc.z; ii[j + 3] = crow - cc.w;
      }
      // warp-uniform fast path: the whole sub-chunk lies inside the row's own modality segment (and the sequence) -> 16
      // unconditional gathers.  Otherwise the index is clamped and the value selected afterwards — never a conditional
      // LOAD: the compiler turns those into divergent branch regions (measured: 4x slower loop).
      const bool fast = __all_sync(0xffffffffu, key0 >= seg_lo && key0 + 16 <= seg_hi);
      if (fast) {
#pragma unroll
        for (int j = 0; j < 16; ++j) add[j] = fmaf(s_lut[ii[j]], 1.4426950408889634f, -mb);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool ok = (key0 + j >= seg_lo) & (key0 + j < seg_hi);
          const float bv = s_lut[ok ? ii[j] : 0];
          add[j] = fmaf(ok ? bv : 0.f, 1.4426950408889634f, -mb);
          add[j] = (key0 + j >= Sk) ? -INFINITY : add[j];
        }
      }
      if constexpr (HAS_PAD) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const uint32_t pw = *reinterpret_cast<const uint32_t*>(s_pad_w + key0 + j);   // 4 mask bytes (zero beyond S)
          add[j] = (pw & 0xffu) ? -INFINITY : add[j];
          add[j + 1] = (pw & 0xff00u) ? -INFINITY : add[j + 1];
          add[j + 2] = (pw & 0xff0000u) ? -INFINITY : add[j + 2];
          add[j + 3] = (pw & 0xff000000u) ? -INFINITY : add[j + 3];
        }
      }
    };
    // sub-chunk t is live (has keys < S) iff t * 16 < S; dead sub-chunks of the last block are stored as zeros
    auto finish = [&](int t, const uint32_t (&v)[16], const float (&add)[16], bool live) {
      uint32_t pk[8];
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), 1.4426950408889634f, add[j]));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), 1.4426950408889634f, add[j + 1]));
          l0 += p0;
          l1 += p1;
          pk[j >> 1] = pack_bf16x2(p0, p1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = 0u;
      }
      // 16 keys = 32 B = 16-byte chunks (c/8, c/8 + 1) of this row in atom c / 64, c = column inside the key block
      const int c = (t * 16) & (kTcK - 1);
      uint8_t* atom = sP + (c >> 6) * (kTcQ * 128);
      sts128u(atom + sw128_off(r, ((c & 63) >> 3)), make_uint4(pk[0], pk[1], pk[2], pk[3]));
      sts128u(atom + sw128_off(r, ((c & 63) >> 3) + 1), make_uint4(pk[4], pk[5], pk[6], pk[7]));
    };
    const bool any = warp_valid;
    if constexpr (HALVES == 1) {
      // one thread per row: software-pipelined (the loads of sub-chunk t + 1 fly while t is exponentiated and stored)
      uint32_t vA[16], vB[16];
      float aA[16], aB[16];
      if (any) prefetch(0, vA, aA);
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb > 0) mbar_wait(&bars->pv[kb - 1], 0);     // previous P consumed by the tensor core
#pragma unroll 1
        for (int u = 0; u < kTcK / 16; u += 2) {
          const int t = kb * (kTcK / 16) + u;
          const bool liveA = any && t * 16 < Sk, liveB = any && (t + 1) * 16 < Sk, liveC = any && (t + 2) * 16 < Sk && t + 2 < nsub;
          if (liveA) tmem_ld_wait();
          if (liveB) prefetch(t + 1, vB, aB);
          finish(t, vA, aA, liveA);
          if (liveB) tmem_ld_wait();
          if (liveC) prefetch(t + 2, vA, aA);
          finish(t + 1, vB, aB, liveB);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p[kb]);
      }
    } else {
      // two threads per row: alternate sub-chunks, latency hidden by the 16 soft-max warps of the SM
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb > 0) mbar_wait(&bars->pv[kb - 1], 0);
#pragma unroll 1
        for (int u = half; u < kTcK / 16; u += HALVES) {
          const int t = kb * (kTcK / 16) + u;
          const bool live = any && t * 16 < Sk;
          uint32_t v[16];
          float add[16];
          if (live) {
            prefetch(t, v, add);
            tmem_ld_wait();
          }
          finish(t, v, add, live);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p[kb]);
      }
    }
    float l = l0 + l1;

    // ---- epilogue: O / l -> bf16 rows; with two threads per row each normalises 32 of the 64 output columns ----
    OPB_T(4);
    if constexpr (HALVES > 1) xch[(2 + half) * kTcQ + r] = l;
    mbar_wait(&bars->pv[nkb - 1], 0);
    tc_fence_after();
    OPB_T(5);
    constexpr int kOutCols = kTcD / HALVES;
    uint32_t o[kOutCols];
    if (warp_valid) {
      __syncwarp();
      if constexpr (HALVES == 1) {
        uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
        uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
        tmem_ld32(lane_base, o0);
        tmem_ld32(lane_base + 32, o1);
      } else {
        uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
        tmem_ld32(lane_base + half * 32, o0);
      }
      tmem_ld_wait();
    }
    if constexpr (HALVES > 1) {
      named_bar_sync(2, kRowThreads);
      l += xch[(2 + (half ^ 1)) * kTcQ + r];
    }
    float ssum = 0.f, ssq = 0.f;
    if (warp_valid && row_valid) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* op = out + (static_cast<long>(b) * S + qrow) * D + h * kTcD + half * kOutCols;
#pragma unroll
      for (int k = 0; k < kOutCols / 8; ++k) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { y[e] = __uint_as_float(o[8 * k + e]) * inv; ssum += y[e]; ssq += y[e] * y[e]; }
        *reinterpret_cast<uint4*>(op + 8 * k) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                                           pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      }
      // log-sum-exp of the biased scores (natural log), kept for the backward pass like attention.cu does
      if (half == 0 && lse != nullptr) lse[(static_cast<long>(b) * H + h) * S + qrow] = m + __logf(l);
    }
    if (ln_stats != nullptr) {
      if constexpr (HALVES > 1) {
        if (half == 1) xstat[r] = make_float2(ssum, ssq);
        named_bar_sync(3, kRowThreads);
        if (half == 0) { const float2 q2 = xstat[r]; ssum += q2.x; ssq += q2.y; }
      }
      if (half == 0 && warp_valid && row_valid) {
        const long rows_total = static_cast<long>(B) * S;
        *reinterpret_cast<float2*>(ln_stats + (h * rows_total + static_cast<long>(b) * S + qrow) * 2) = make_float2(ssum, ssq);
      }
    }
  }

#ifdef OPB_ATTN_TIMING
  if (threadIdx.x == 64) {
    tt[6] = globaltimer_ns();
    for (int i = 1; i <= 6; ++i) atomicAdd(&g_attn_t[i], tt[i] - tt[0]);
    atomicAdd(&g_attn_n, 1u);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

#ifdef OPB_ATTN_TIMING
extern "C" void opb_attn_timing_dump() {
  unsigned long long t[8]; unsigned int n;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(t, g_attn_t, sizeof(t)); cudaMemcpyFromSymbol(&n, g_attn_n, sizeof(n));
  printf("[attn timing] ctas=%u  avg ns since CTA start: tables_ready=%.0f S0_ready=%.0f phaseA_done=%.0f phaseB_done=%.0f O_ready=%.0f end=%.0f\n",
         n, (double)t[1] / n, (double)t[2] / n, (double)t[3] / n, (double)t[4] / n, (double)t[5] / n, (double)t[6] / n);
}
#endif

// lut[h][l] = table[idx[l]][h]
__global__ void relpos_lut_kernel(const float* __restrict__ table, const int* __restrict__ idx, float* __restrict__ lut,
                                  int L, int H) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const long bk = idx[l];
  for (int h = 0; h < H; ++h) lut[static_cast<long>(h) * L + l] = table[bk * H + h];
}

int relpos_lut_build(const float* table, const int* idx, float* lut, int L, int H, cudaStream_t stream) {
  if (L <= 0 || H <= 0) return OPB_ERR_INVALID;
  relpos_lut_kernel<<<(L + 255) / 256, 256, 0, stream>>>(table, idx, lut, L, H);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int attention_tcp_fwd(const void* qkv, const float* lut, int lut_len, const int* code_row, const int* code_col,
                      const uint8_t* key_pad, void* out, float* lse, float* ln_stats, int B, int S, int H, int seg_split,
                      cudaStream_t stream);

// One launch of attention_tc_kernel over keys [k0, k0 + Sk) of every sample (all S queries).
static int launch_tc(const void* qkv, const float* lut, const float* lut_max, int lut_len, const int* code_row, const int* code_col,
                     const uint8_t* key_pad, void* out, float* lse, float* ln_stats, int B, int S, int H, int seg_split, int k0,
                     int Sk, cudaStream_t stream) {
  const int nkb = (Sk + kTcK - 1) / kTcK;
  if (nkb > kTcMaxBlocks) return OPB_ERR_UNSUPPORTED;
  if (lut_len % 4 != 0 || (reinterpret_cast<uintptr_t>(lut) & 15) != 0 || (reinterpret_cast<uintptr_t>(code_col + k0) & 15) != 0)
    return OPB_ERR_INVALID;     // bulk-copied: 16-byte granularity (code_col must hold (S + 3) & ~3 entries)
  const long table_bytes = static_cast<long>(lut_len) * 4 + static_cast<long>((Sk + 3) & ~3) * 4 + Sk + 48;
  // small tables move into the (dead) 16 KB Q tile for phase B; larger ones (long sequences) get their own region
  const int big_tables = table_bytes > kTcTableBytes ? 1 : 0;
  if (table_bytes > 40 * 1024) return OPB_ERR_UNSUPPORTED;
  const int D = H * kTcD;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, kTcQ);
  if (rc != OPB_OK) return rc;
  const size_t smem = kTcQ * 128 + 2ull * nkb * kTcK * 128 + kTcQ * kTcK * 2 + (big_tables ? 256 + ((table_bytes + 63) & ~63L) : sizeof(TcBars));
  static_assert(sizeof(TcBars) <= 256, "");
  const uint32_t tmem_cols = nkb == 1 ? 128 : (nkb == 2 ? 256 : 512);
  static const char* env_h = getenv("OPB_ATTN_ROW_THREADS");          // 1 or 2 threads per query row (A/B switch), default 2
  const int halves = (env_h != nullptr && env_h[0] == '1') ? 1 : 2;
  static size_t configured[4] = {0, 0, 0, 0};
  const int v = (key_pad != nullptr ? 1 : 0) + 2 * (halves - 1);
  auto kern = v == 0 ? attention_tc_kernel<false, 1> : (v == 1 ? attention_tc_kernel<true, 1> : (v == 2 ? attention_tc_kernel<false, 2> : attention_tc_kernel<true, 2>));
  if (smem > configured[v]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return OPB_ERR_CUDA;
    configured[v] = smem;
  }
  const int q_tiles = (S + kTcQ - 1) / kTcQ;
  const unsigned grid = static_cast<unsigned>(static_cast<long>(B) * H * q_tiles);
  kern<<<grid, 64 + 128 * halves, smem, stream>>>(tm, lut, lut_max, lut_len, code_row, code_col + k0, key_pad,
                                                 reinterpret_cast<__nv_bfloat16*>(out), lse, ln_stats, B, S, H, nkb, tmem_cols, seg_split,
                                                 k0, Sk, big_tables);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// Merge of two key-range partial results (flash-decoding style): out = w0 out0 + w1 out1 with w_c = exp(lse_c - lse), in place in
// `out` (which holds out0); writes the total log-sum-exp and the per-head inner-LayerNorm statistics.  One warp per (row, head).
__global__ void __launch_bounds__(256)
attention_merge_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ out1, const float* __restrict__ lse0,
                       const float* __restrict__ lse1, float* __restrict__ lse, float* __restrict__ ln_stats, int B, int S, int H) {
  const long w = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long rows = static_cast<long>(B) * S;
  if (w >= rows * H) return;
  const long row = w / H;
  const int h = static_cast<int>(w % H);
  const int b = static_cast<int>(row / S), sq = static_cast<int>(row % S);
  const long li = (static_cast<long>(b) * H + h) * S + sq;
  const float l0 = lse0[li], l1 = lse1[li];
  const float m = fmaxf(l0, l1);
  float w0 = 0.f, w1 = 0.f, tot = -INFINITY;
  if (m > -INFINITY) {
    const float e0 = __expf(l0 - m), e1 = __expf(l1 - m);
    const float den = e0 + e1;
    w0 = e0 / den; w1 = e1 / den;
    tot = m + __logf(den);
  }
  const long D = static_cast<long>(H) * kTcD;
  uint32_t* p0 = reinterpret_cast<uint32_t*>(out + row * D + h * kTcD) + lane;
  const uint32_t* p1 = reinterpret_cast<const uint32_t*>(out1 + row * D + h * kTcD) + lane;
  const float2 a = unpack_bf16x2(*p0), c = unpack_bf16x2(*p1);
  const float y0 = w0 * a.x + w1 * c.x, y1 = w0 * a.y + w1 * c.y;
  *p0 = pack_bf16x2(y0, y1);
  const float ssum = warp_sum(y0 + y1), ssq = warp_sum(y0 * y0 + y1 * y1);
  if (lane == 0) {
    if (lse != nullptr) lse[li] = tot;
    if (ln_stats != nullptr) *reinterpret_cast<float2*>(ln_stats + (h * rows + row) * 2) = make_float2(ssum, ssq);
  }
}

int attention_tc_fwd(const void* qkv, const float* lut, const float* lut_max, int lut_len, const int* code_row, const int* code_col,
                     const uint8_t* key_pad, void* out, float* lse, float* ln_stats, int B, int S, int H, int seg_split,
                     cudaStream_t stream) {
  if (B <= 0 || S <= 0 || H <= 0 || lut == nullptr || lut_max == nullptr || code_row == nullptr || code_col == nullptr) return OPB_ERR_INVALID;
  if (seg_split < 0 || seg_split >= S) return OPB_ERR_INVALID;
  // S <= 224: the persistent kernel (attention_tcp.cu).  OPB_ATTN_PERSIST=0 keeps the one-CTA-per-tile kernel below (A/B switch).
  const char* env_p = getenv("OPB_ATTN_PERSIST");            // read per call: tests switch it in-process
  if (S <= 224 && !(env_p != nullptr && env_p[0] == '0')) {
    const int rc = attention_tcp_fwd(qkv, lut, lut_len, code_row, code_col, key_pad, out, lse, ln_stats, B, S, H, seg_split, stream);
    if (rc != OPB_ERR_UNSUPPORTED) return rc;
  }
  if (S <= kTcMaxBlocks * kTcK) return launch_tc(qkv, lut, lut_max, lut_len, code_row, code_col, key_pad, out, lse, ln_stats, B, S, H, seg_split, 0, S, stream);
  // 384 < S <= 768 (audio: 750 tokens for 15 s): two key ranges, each a launch that holds its <= 384 scores per row in tensor
  // memory, merged afterwards.  Scratch (second partial output + the two partial log-sum-exp vectors) is stream-ordered.
  if (S > 2 * kTcMaxBlocks * kTcK || seg_split != 0) return OPB_ERR_UNSUPPORTED;
  const int Sk0 = ((S / 2 + 7) / 8) * 8, Sk1 = S - Sk0;
  const size_t out_bytes = static_cast<size_t>(B) * S * H * kTcD * 2, lse_bytes = static_cast<size_t>(B) * H * S * 4;
  uint8_t* scratch = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), out_bytes + 2 * lse_bytes, stream) != cudaSuccess) return OPB_ERR_CUDA;
  float* lse0 = reinterpret_cast<float*>(scratch + out_bytes);
  float* lse1 = lse0 + static_cast<size_t>(B) * H * S;
  int rc = launch_tc(qkv, lut, lut_max, lut_len, code_row, code_col, key_pad, out, lse0, nullptr, B, S, H, 0, 0, Sk0, stream);
  if (rc == OPB_OK) rc = launch_tc(qkv, lut, lut_max, lut_len, code_row, code_col, key_pad, scratch, lse1, nullptr, B, S, H, 0, Sk0, Sk1, stream);
  if (rc == OPB_OK) {
    const long warps = static_cast<long>(B) * S * H;
    attention_merge_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(scratch), lse0, lse1, lse, ln_stats, B, S, H);
    if (cudaGetLastError() != cudaSuccess) rc = OPB_ERR_CUDA;
  }
  cudaFreeAsync(scratch, stream);
  return rc;
}

}  // namespace opb

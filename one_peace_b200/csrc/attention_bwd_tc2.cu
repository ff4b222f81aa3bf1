// Double-buffered form of the tcgen05 attention backward (attention_bwd_tc.cu has the algebra and the operand layouts; read
// that header first).  Same transposed formulation — TMEM lanes are keys, S^T = K Q^T and dP^T = V dO^T, P^T is the TMEM A operand
// of dV, one swizzled dS^T tile in shared memory is the K-major A operand of dK and the MN-major A operand of dQ — but the
// tensor core and the soft-max warps no longer alternate:
//
//   * a sub-step is (key tile of 128) x (query QUARTER of 64 columns), so S^T + dP^T of one sub-step are 128 TMEM columns and two
//     sub-steps fit side by side: while the 8 warps turn buffer g & 1 into P^T / dS^T, the MMA issuer has already queued
//     S^T / dP^T of sub-step g + 1 in the other buffer and issues those of g + 2 right behind the dV / dK / dQ chains of g.
//       TMEM: [S^T | dP^T] x 2 @ 0..255, dV @ 256, dK @ 320, dQ (two 128-query tiles) @ 384 / 448  = 512 columns.
//   * dS^T of a quarter is exactly one 64-query chunk of the shared tile; dK consumes a chunk per sub-step, dQ both chunks after
//     the second quarter of a pair.  A chunk is rewritten two sub-steps later, after the `acc_done` of the sub-step in between.
//   * finished accumulators (dK / dV after the last quarter of a key tile, dQ after the last key tile) are read out one sub-step
//     LATE, between the soft-max work of the next sub-step and its `p_full` arrival — their MMAs have long completed by then, so the
//     warps never wait for the tensor core in steady state.
//   * only the sequence's pad16(S) query rows are staged (Q and dO, two stages each); the last quarter shrinks its MMAs to the
//     valid 16-column groups (N of S^T / dP^T, K of dV / dK) instead of reading past the staged rows.
// Bias / dbias are taken as TRANSPOSED tables only (opb_attention_bwd_t); callers with dense or per-sample tables use v1.
// Measured at B = 64, S = 197, H = 24 (profiles/r02_attention_fwd_bwd.txt): see DESIGN.md 4.6.
#include "common.cuh"
#include "ops.h"
#include "tmem_frag.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace opb {

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace {

constexpr int kD = 64;
constexpr int kKT = 128;
constexpr int kQH = 64;                 // query columns per sub-step
// soft-max warps: 8 lane groups of 16 TMEM lanes x kColSplit column halves.  With 8 warps (two per scheduler) the dependent
// ld -> fma -> ex2 -> mul -> pack -> st chains left the issue slots 70 % idle; 16 warps of half the columns each hide them.
#ifndef OPB_BWD2_COLSPLIT
#define OPB_BWD2_COLSPLIT 2
#endif
constexpr int kColSplit = OPB_BWD2_COLSPLIT;
constexpr int kNB = 8 / kColSplit;      // 8-column blocks per warp and sub-step
constexpr int kSoftWarps = 8 * kColSplit;
constexpr int kThreads = 32 * (kSoftWarps + 2);
constexpr int kColBuf = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
constexpr int kTKeys = 256, kTQ = 224;  // transposed bias tables (attention_bwd_tc.cu)
constexpr uint32_t KV_BYTES = kKT * 128;

struct Bars2 {
  uint64_t qdo_full[2], qdo_empty[2], kv_full[2], kv_empty[2];
  uint64_t s_full[2], p_full[2], acc_done;
  uint32_t tmem_base;
};

struct Args2 {
  const uint32_t* bias_t; float* dbias_t;
  const uint8_t* key_pad;
  const float* lse; const float* delta;
  __nv_bfloat16* dqkv;
  int B, S, H, n_kt, n_q, nqp;
  float q_scale;
  long n_items;
};

// no "memory" clobber: the table is only ever touched by these reductions, and a clobber would pin every block's shared-memory loads
// behind the previous block's reduction (no overlap between the 8 blocks of a sub-step)
OPB_DEVICE void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b));
}
OPB_DEVICE uint64_t mn_desc_lbo(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
OPB_DEVICE void ld_blocks(uint32_t taddr, uint32_t (&v)[4 * kNB]) {
  if constexpr (kNB == 8) tmem_ld_16x256b_x8(taddr, v); else tmem_ld_16x256b_x4(taddr, v);
}
OPB_DEVICE void st_blocks(uint32_t taddr, const uint32_t (&v)[2 * kNB]) {
  if constexpr (kNB == 8) tmem_st_16x128b_x8(taddr, v); else tmem_st_16x128b_x4(taddr, v);
}
OPB_DEVICE void soft_bar_sync2() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftWarps) : "memory"); }

// position of a sub-step inside the CTA's item range; advanced in lock-step by the three roles
struct SubIt {
  int j, kt, qh;
  OPB_DEVICE void next(int n_kt, int n_q) {
    if (++qh == n_q) { qh = 0; if (++kt == n_kt) { kt = 0; ++j; } }
  }
};

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                         const __grid_constant__ CUtensorMap tm_do, const Args2 a) {
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(1024) uint8_t bwd2_smem_raw[];
  uint8_t* smem = bwd2_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const uint32_t QDO_BYTES = static_cast<uint32_t>(a.nqp) * 128;
  const uint32_t qdo_stride = (QDO_BYTES + 1023u) & ~1023u;
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + 2 * qdo_stride;
  uint8_t* sK = sDO + 2 * qdo_stride;
  uint8_t* sV = sK + 2 * KV_BYTES;
  uint8_t* sDS = sV + 2 * KV_BYTES;                               // [2 query chunks of 64][128 keys][128 B]
  float* sL2 = reinterpret_cast<float*>(sDS + 2 * KV_BYTES);      // [2][256]
  float* sDl = sL2 + 2 * 256;
  Bars2* bars = reinterpret_cast<Bars2*>(sDl + 2 * 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S, H = a.H, D = a.H * kD;
  const long w0 = a.n_items * blockIdx.x / gridDim.x, w1 = a.n_items * (blockIdx.x + 1) / gridDim.x;
  const int n = static_cast<int>(w1 - w0);
  const int n_kt = a.n_kt, n_q = a.n_q;
  const int G = n * n_kt * n_q;                                   // sub-steps of this CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qdo_full[i], 1);
      mbar_init(&bars->qdo_empty[i], 1);
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], kSoftWarps);
    }
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
      int t = 0;
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
    if (lane == 0 && G > 0) {
      constexpr uint32_t idesc_kv = make_idesc_bf16(kKT, kD) | (1u << 16);                    // B MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(kKT, kD) | (1u << 15) | (1u << 16);        // A and B MN-major
      // 16-column groups of quarter qh that hold live queries (the last quarter of the sequence may be short)
      auto nv16 = [&](int qh) { const int v = (S - qh * kQH + 15) / 16; return v > 4 ? 4 : v; };
      // S^T / dP^T of sub-step (it, g) into buffer g & 1; t_of = running key-tile index of it
      auto issue_sp = [&](const SubIt& it, int g) {
        const int sq = it.j & 1;
        const int t = it.j * n_kt + it.kt, sk = t & 1;
        if (it.kt == 0 && it.qh == 0) mbar_wait(&bars->qdo_full[sq], (it.j >> 1) & 1);
        if (it.qh == 0) mbar_wait(&bars->kv_full[sk], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t idesc_s = make_idesc_bf16(kKT, 16 * nv16(it.qh));
        const uint64_t dk_a = make_sw128_kmajor_desc(smem_u32(sK + sk * KV_BYTES)), dv_a = make_sw128_kmajor_desc(smem_u32(sV + sk * KV_BYTES));
        const uint64_t dq_b = make_sw128_kmajor_desc(smem_u32(sQ + sq * qdo_stride) + it.qh * kQH * 128);
        const uint64_t ddo_b = make_sw128_kmajor_desc(smem_u32(sDO + sq * qdo_stride) + it.qh * kQH * 128);
        const uint32_t buf = tmem_base + (g & 1) * kColBuf;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) umma_bf16<1>(buf, dk_a + 2 * kk, dq_b + 2 * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) umma_bf16<1>(buf + kQH, dv_a + 2 * kk, ddo_b + 2 * kk, idesc_s, kk != 0);
        umma_commit<1>(&bars->s_full[g & 1]);
      };
      SubIt sp = {0, 0, 0}, it = {0, 0, 0};
      int g_sp = 0;
      for (; g_sp < 2 && g_sp < G; ++g_sp) { issue_sp(sp, g_sp); sp.next(n_kt, n_q); }
      const uint32_t ds_base = smem_u32(sDS);
      for (int g = 0; g < G; ++g) {
        const int sq = it.j & 1;
        const int t = it.j * n_kt + it.kt, sk = t & 1;
        const uint32_t q_base = smem_u32(sQ + sq * qdo_stride), do_base = smem_u32(sDO + sq * qdo_stride);
        const uint32_t k_base = smem_u32(sK + sk * KV_BYTES);
        mbar_wait(&bars->p_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const int kq = nv16(it.qh);
        const uint32_t buf = tmem_base + (g & 1) * kColBuf;
        const uint32_t chunk = (it.qh & 1) * KV_BYTES;
        for (int kk = 0; kk < kq; ++kk) {
          const uint64_t b_do = make_sw128_mn_desc64(do_base + (it.qh * kQH + 16 * kk) * 128);
          // P^T of 16 queries = 8 packed columns; with two column halves per lane group the second half's words start at column 32
          const uint32_t pcol = kColSplit == 2 ? (kk < 2 ? 8 * kk : 32 + 8 * (kk - 2)) : 8 * kk;
          umma_bf16_ts(tmem_base + kColDV, buf + pcol, b_do, idesc_kv, (it.qh != 0 || kk != 0) ? 1u : 0u);
        }
        for (int kk = 0; kk < kq; ++kk) {
          const uint64_t a_ds = make_sw128_kmajor_desc(ds_base + chunk) + 2 * kk;
          const uint64_t b_q = make_sw128_mn_desc64(q_base + (it.qh * kQH + 16 * kk) * 128);
          umma_bf16<1>(tmem_base + kColDK, a_ds, b_q, idesc_kv, (it.qh != 0 || kk != 0) ? 1u : 0u);
        }
        if ((it.qh & 1) || it.qh == n_q - 1) {
          int keys16 = (S - it.kt * kKT + 15) / 16;
          if (keys16 > 8) keys16 = 8;
          for (int kk = 0; kk < keys16; ++kk) {
            const uint64_t a_ds = mn_desc_lbo(ds_base + kk * 2048, KV_BYTES);
            const uint64_t b_k = make_sw128_mn_desc64(k_base + kk * 2048);
            umma_bf16<1>(tmem_base + kColDQ + (it.qh >> 1) * kD, a_ds, b_k, idesc_q, (it.kt != 0 || kk != 0) ? 1u : 0u);
          }
        }
        umma_commit<1>(&bars->acc_done);
        if (it.qh == n_q - 1) umma_commit<1>(&bars->kv_empty[sk]);
        if (it.qh == n_q - 1 && it.kt == n_kt - 1) umma_commit<1>(&bars->qdo_empty[sq]);
        it.next(n_kt, n_q);
        if (g_sp < G) { issue_sp(sp, g_sp); sp.next(n_kt, n_q); ++g_sp; }
      }
    }
  } else {
    // ===================== soft-max backward (8 warps) =====================
    const int qw = warp & 3, hh = (warp >> 2) & 1, ch = warp >> 3, t4 = lane & 3;     // ch: column half (kColSplit == 2)
    const int blk0 = ch * kNB;                                                          // first 8-column block of this warp
    const int r_lo = qw * 32 + hh * 16 + (lane >> 2);
    const uint32_t lane_addr = static_cast<uint32_t>(qw * 32 + hh * 16) << 16;
    const long row_pitch = 3L * D;
    uint8_t* ds_row = sDS + r_lo * 128 + 4 * t4;
    const int rx = r_lo & 7;
    // pending read-outs of the previous sub-step's finished accumulators
    bool pend_kv = false, pend_q = false;
    int pb = 0, ph = 0, pkt = 0;
    auto epilogue = [&]() {
      if (pend_kv) {
        const bool wv = pkt * kKT + qw * 32 + hh * 16 < S;
        if (wv) {
          uint32_t gk[4 * kNB], gv[4 * kNB];
          ld_blocks(tmem_base + lane_addr + kColDK + 8 * blk0, gk);
          ld_blocks(tmem_base + lane_addr + kColDV + 8 * blk0, gv);
          tmem_ld_wait();
#pragma unroll
          for (int rsel = 0; rsel < 2; ++rsel) {
            const int key = pkt * kKT + r_lo + 8 * rsel;
            if (key < S) {
              uint32_t* op = reinterpret_cast<uint32_t*>(a.dqkv + (static_cast<long>(pb) * S + key) * row_pitch + D + ph * kD + 2 * t4);
#pragma unroll
              for (int k = 0; k < kNB; ++k) {
                op[4 * (blk0 + k)] = pack_bf16x2(__uint_as_float(gk[4 * k + 2 * rsel]), __uint_as_float(gk[4 * k + 2 * rsel + 1]));
                op[4 * (blk0 + k) + D / 2] = pack_bf16x2(__uint_as_float(gv[4 * k + 2 * rsel]), __uint_as_float(gv[4 * k + 2 * rsel + 1]));
              }
            }
          }
        }
      }
      if (pend_q) {
        for (int q2 = 0; q2 < (n_q + 1) / 2; ++q2) {
          if (q2 * kKT + qw * 32 + hh * 16 < S) {
            uint32_t gq[4 * kNB];
            ld_blocks(tmem_base + lane_addr + kColDQ + q2 * kD + 8 * blk0, gq);
            tmem_ld_wait();
#pragma unroll
            for (int rsel = 0; rsel < 2; ++rsel) {
              const int qrow = q2 * kKT + r_lo + 8 * rsel;
              if (qrow < S) {
                uint32_t* op = reinterpret_cast<uint32_t*>(a.dqkv + (static_cast<long>(pb) * S + qrow) * row_pitch + ph * kD + 2 * t4);
#pragma unroll
                for (int k = 0; k < kNB; ++k)
                  op[4 * (blk0 + k)] = pack_bf16x2(__uint_as_float(gq[4 * k + 2 * rsel]) * a.q_scale, __uint_as_float(gq[4 * k + 2 * rsel + 1]) * a.q_scale);
              }
            }
          }
        }
      }
      if (pend_kv || pend_q) tc_fence_before();
      pend_kv = pend_q = false;
    };

    // bias words (transposed half2 table) of sub-step x -> dst; zero where the sub-tile has no live query / the warp no live key
    auto load_bias = [&](const SubIt& x, uint32_t (&dst)[2 * kNB]) {
#pragma unroll
      for (int i = 0; i < 2 * kNB; ++i) dst[i] = 0u;
      if (a.bias_t == nullptr || x.kt * kKT + qw * 32 + hh * 16 >= S) return;
      const int hx = static_cast<int>((w0 + x.j) % H);
      const int nb = min(8, (S - x.qh * kQH + 7) / 8);
      const uint32_t* bt = a.bias_t + (static_cast<long>(hx) * kTKeys + x.kt * kKT + r_lo) * (kTQ / 2) + (x.qh * kQH) / 2 + t4;
#pragma unroll
      for (int k = 0; k < kNB; ++k) {
        if (blk0 + k < nb) {
          dst[2 * k] = __ldg(bt + 4 * (blk0 + k));
          dst[2 * k + 1] = __ldg(bt + 8 * (kTQ / 2) + 4 * (blk0 + k));
        }
      }
    };
    SubIt it = {0, 0, 0};
    int b = 0, h = 0;
    bool dead_lo = true, dead_hi = true, warp_valid = false;
    int key_lo = 0;
    float* l2s = sL2;
    float* dls = sDl;
    for (int g = 0; g < G; ++g) {
      if (it.kt == 0 && it.qh == 0) {
        // ---- new item: per-query vectors -> shared memory (double-buffered by item parity) ----
        const long w = w0 + it.j;
        h = static_cast<int>(w % H); b = static_cast<int>(w / H);
        const long stat = (static_cast<long>(b) * H + h) * S;
        l2s = sL2 + (it.j & 1) * 256;
        dls = sDl + (it.j & 1) * 256;
        if (threadIdx.x < 256) {
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
        soft_bar_sync2();
      }
      if (it.qh == 0) {
        key_lo = it.kt * kKT + r_lo;
        const int key_hi = key_lo + 8;
        warp_valid = it.kt * kKT + qw * 32 + hh * 16 < S;
        dead_lo = key_lo >= S; dead_hi = key_hi >= S;
        if (a.key_pad != nullptr) {
          if (!dead_lo) dead_lo = a.key_pad[static_cast<long>(b) * S + key_lo] != 0;
          if (!dead_hi) dead_hi = a.key_pad[static_cast<long>(b) * S + key_hi] != 0;
        }
      }
      const int q0 = it.qh * kQH;
      const int nblk = min(8, (S - q0 + 7) / 8);                   // 8-column blocks with live queries (warp-uniform)
      // ---- bias words of this sub-tile (transposed half2 table), fetched under the wait for the tensor core ----
      uint32_t bw[2 * kNB];
      load_bias(it, bw);
      mbar_wait(&bars->s_full[g & 1], (g >> 1) & 1);
      tc_fence_after();
      uint32_t pw[2 * kNB], dsw[2 * kNB];
      if (warp_valid) {
        uint32_t st[4 * kNB], dp[4 * kNB];
        const uint32_t buf = tmem_base + lane_addr + (g & 1) * kColBuf;
        ld_blocks(buf + 8 * blk0, st);
        ld_blocks(buf + kQH + 8 * blk0, dp);
        tmem_ld_wait();
        const float* l2q = l2s + q0 + 2 * t4;
        const float* dlq = dls + q0 + 2 * t4;
        float* dt_lo = nullptr;
        if (a.dbias_t != nullptr) dt_lo = a.dbias_t + (static_cast<long>(h) * kTKeys + key_lo) * kTQ + q0 + 2 * t4;
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
          const int blk = blk0 + k;
          if (blk < nblk) {
            const float2 l2 = *reinterpret_cast<const float2*>(l2q + 8 * blk);
            const float2 dl = *reinterpret_cast<const float2*>(dlq + 8 * blk);
            const float2 ba = __half22float2(*reinterpret_cast<const __half2*>(&bw[2 * k]));
            const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(&bw[2 * k + 1]));
            float p0 = ex2_fast(fmaf(__uint_as_float(st[4 * k + 0]), kLog2e, ba.x) - l2.x);
            float p1 = ex2_fast(fmaf(__uint_as_float(st[4 * k + 1]), kLog2e, ba.y) - l2.y);
            float p2 = ex2_fast(fmaf(__uint_as_float(st[4 * k + 2]), kLog2e, bb.x) - l2.x);
            float p3 = ex2_fast(fmaf(__uint_as_float(st[4 * k + 3]), kLog2e, bb.y) - l2.y);
            p0 = dead_lo ? 0.f : p0; p1 = dead_lo ? 0.f : p1;
            p2 = dead_hi ? 0.f : p2; p3 = dead_hi ? 0.f : p3;
            const float d0 = p0 * (__uint_as_float(dp[4 * k + 0]) - dl.x), d1 = p1 * (__uint_as_float(dp[4 * k + 1]) - dl.y);
            const float d2 = p2 * (__uint_as_float(dp[4 * k + 2]) - dl.x), d3 = p3 * (__uint_as_float(dp[4 * k + 3]) - dl.y);
            if (a.dbias_t != nullptr) {
              if (!dead_lo) red_add_v2(dt_lo + 8 * blk, d0, d1);
              if (!dead_hi) red_add_v2(dt_lo + 8 * kTQ + 8 * blk, d2, d3);
            }
            pw[2 * k] = pack_bf16x2(p0, p1);
            pw[2 * k + 1] = pack_bf16x2(p2, p3);
            dsw[2 * k] = pack_bf16x2(d0, d1);
            dsw[2 * k + 1] = pack_bf16x2(d2, d3);
          } else {
            pw[2 * k] = pw[2 * k + 1] = dsw[2 * k] = dsw[2 * k + 1] = 0u;
          }
        }
        st_blocks(buf + 8 * blk0, pw);          // packed P^T of this warp's blocks over the first half of ITS OWN score columns
      }
      // the dS^T chunk of this quarter was last read by the MMAs of sub-steps g - 2 (dK) and g - 1 (dQ of the pair): wait for g - 1;
      // its finished accumulators, if any, are read out now (deferred by one sub-step: no wait in steady state)
      if (g > 0) {
        mbar_wait(&bars->acc_done, (g - 1) & 1);
        tc_fence_after();
        epilogue();
      }
      if (warp_valid) {
        uint8_t* dchunk = ds_row + (it.qh & 1) * KV_BYTES;
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
          uint8_t* dsp = dchunk + (((blk0 + k) ^ rx) << 4);
          *reinterpret_cast<uint32_t*>(dsp) = dsw[2 * k];
          *reinterpret_cast<uint32_t*>(dsp + 1024) = dsw[2 * k + 1];
        }
        tmem_st_wait_all();
        fence_proxy_async();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[g & 1]);
      // what this sub-step completes (read out during the next one)
      if (it.qh == n_q - 1) {
        pend_kv = true; pb = b; ph = h; pkt = it.kt;
        if (it.kt == n_kt - 1) pend_q = true;
      }
      it.next(n_kt, n_q);
    }
    if (G > 0) {
      mbar_wait(&bars->acc_done, (G - 1) & 1);
      tc_fence_after();
      epilogue();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace

// Transposed-table form only (bias_t / dbias_t may be null); S <= 224.  `delta` already computed.
int attention_bwd_tc2(const void* qkv, const void* d_out, const uint8_t* key_pad, const float* lse, const float* delta, void* dqkv,
                      int B, int S, int H, float q_scale, const void* bias_t, float* dbias_t, cudaStream_t stream) {
  if (S > 224) return OPB_ERR_UNSUPPORTED;
  if (dbias_t != nullptr && bias_t == nullptr) return OPB_ERR_INVALID;
  const int D = H * kD;
  Args2 a;
  a.bias_t = reinterpret_cast<const uint32_t*>(bias_t); a.dbias_t = dbias_t; a.key_pad = key_pad;
  a.lse = lse; a.delta = delta; a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  a.B = B; a.S = S; a.H = H; a.q_scale = q_scale;
  a.n_kt = (S + kKT - 1) / kKT;
  a.n_q = (S + kQH - 1) / kQH;
  a.nqp = (S + 15) / 16 * 16;
  a.n_items = static_cast<long>(B) * H;
  CUtensorMap tq, tkv, tdo;
  int rc = make_tmap_bf16_2d(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, a.nqp);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tkv, qkv, static_cast<uint64_t>(B) * S, 3ull * D, 3ull * D, kKT);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tdo, d_out, static_cast<uint64_t>(B) * S, D, D, a.nqp);
  if (rc != OPB_OK) return rc;
  const size_t qdo = (static_cast<size_t>(a.nqp) * 128 + 1023) & ~static_cast<size_t>(1023);
  const size_t smem = 4 * qdo + 6ull * KV_BYTES + 4ull * 256 * 4 + sizeof(Bars2) + 64;
  if (smem > 227 * 1024) return OPB_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return OPB_ERR_CUDA;
    configured = smem;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned grid = static_cast<unsigned>(a.n_items < sms ? a.n_items : sms);
  return launch_maybe_cluster(attention_bwd_tc2_kernel, dim3(grid), dim3(kThreads), smem, stream, tq, tkv, tdo, a) == cudaSuccess
             ? OPB_OK : OPB_ERR_CUDA;
}

}  // namespace opb

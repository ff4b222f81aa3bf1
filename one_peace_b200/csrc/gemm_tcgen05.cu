// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] . B[N,K]^T )
//
//   * operands are K-major bf16 (activations [rows, K], nn.Linear weights [out, in]) and are staged
//     into shared memory by TMA with the 128-byte swizzle, BLOCK_K = 64;
//   * tcgen05.mma (kind::f16, fp32 accumulate) issued by one thread, accumulators in TMEM,
//     two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of i+1;
//   * CG = 1: one CTA per 128 x 256 tile.  CG = 2: a CTA pair (cluster of 2) per 256 x 256 tile with
//     tcgen05.mma.cta_group::2 — each CTA stages its 128 rows of A and its 128 rows of B;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> registers
//     -> fused epilogue -> global).
//
// Fused epilogues (what the reference runs as separate ATen kernels, SURVEY.md 2.5 K1'/K3/K4):
//   EPI_STORE_BF16 : out_bf16 = (acc + bias[n]) * colscale[n]        (q/k/v projection, q pre-scaled;
//                                                                     multihead_attention.py:103-107)
//   EPI_GEGLU_BF16 : out_bf16[:, t*128+j] = gelu(acc[:, j]) * acc[:, 128+j] on a W0/W1-interleaved
//                    weight (transformer_layer.py:54-67) — the 2*ffn wide intermediate never hits HBM
//   EPI_RESID_F32  : out_f32 = resid + gamma[n] * (acc + bias[n])    (out_proj / fc2 + LayerScale +
//                                                                     residual, transformer_layer.py:70-88)
//   EPI_STORE_F32  : out_f32 = acc + bias[n]
#include "common.cuh"
#include "gemm.h"

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace opb {

constexpr int kBlockM = 128;   // rows of A per CTA
constexpr int kBlockN = 256;   // accumulator columns per tile
constexpr int kBlockK = 64;    // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kAccStages = 2;

// Epilogue staging (TMA-store path): per epilogue warp a ring of 4 KB chunk buffers ([32 rows][128 B], 128B swizzle)
//   RESID_F32 / STORE_F32 : 4 x fp32 chunk (32 cols; the residual chunk is TMA-prefetched two chunks ahead and
//                           overwritten in place; with 4 slots a slot is refilled two stores after it was stored
//                           from, so the producer never waits for the NEWEST bulk store — waiting on it cost ~1.5 us
//                           per chunk and made out_proj 3x slower than cuBLAS) + 2 x 2 KB bf16-copy chunk (32 cols,
//                           64-byte rows, 64B swizzle)                                         = 20 KB / warp
//   STORE_BF16 / GELU_BF16 / GEGLU_BF16 : 2 x bf16 chunk (64 cols)                            =  8 KB / warp
// plus 4 KB / warp holding the tile's per-column epilogue vectors (LayerNorm column sums, bias, scale / gamma):
// fetched once per tile BEFORE the accumulator is ready, so no epilogue FMA ever waits on a global load (the
// first version issued those loads just-in-time and ncu showed 30 % of all samples on the dependent FFMA).
#ifndef OPB_EPI_WARPS_BF16
#define OPB_EPI_WARPS_BF16 4   // 8 was measured no faster (profiles/r01_bench_*): the extra staging costs a pipeline stage
#endif
template <int EPI, bool TMAEPI>
struct EpiCfg {
  static constexpr bool kRegular = (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_GEGLU_BF16 ||
                                    EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32);
  static constexpr bool kF32 = (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32);
  // bf16-output TMA epilogues run 8 epilogue warps (two per TMEM lane quarter, each taking half of the tile's
  // columns): the GeGLU / QKV epilogues are instruction-bound and a single warp per SM sub-partition cannot hide the
  // TMEM-load -> math -> smem -> TMA-store chain.  The fp32 residual epilogue keeps 4 (its staging is 24 KB / warp).
  static constexpr int kWarps = (TMAEPI && !kF32 && OPB_EPI_WARPS_BF16 == 8) ? 8 : 4;
  static constexpr int kThreads = 64 + 32 * kWarps;
  static constexpr int kRing = TMAEPI ? (kF32 ? 4 : 2) : 0;
  static constexpr int kColVecOff = kRing * 4096 + (kF32 ? 2 * 2048 : 0);
  static constexpr int kWarpBytes = TMAEPI ? (kColVecOff + 4096) : (kRegular ? 4096 : 0);   // direct path: column vectors only
  static constexpr int kBytes = kWarps * kWarpBytes;
};

template <int CG, int EPI = 0, bool TMAEPI = false>
struct GemmCfg {
  static constexpr int kBRows = kBlockN / CG;                       // rows of B staged per CTA
  static constexpr int kABytes = kBlockM * kBlockK * 2;             // 16 KB
  static constexpr int kBBytes = kBRows * kBlockK * 2;              // 32 KB (CG=1) / 16 KB (CG=2)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = EpiCfg<EPI, TMAEPI>::kBytes;
  static constexpr int kBudget = 227 * 1024 - 2048 - kEpiBytes;     // barriers + alignment slack
  static constexpr int kStagesMax = (CG == 1) ? 4 : 6;
  static constexpr int kStages = (kBudget / kStageBytes) < kStagesMax ? (kBudget / kStageBytes) : kStagesMax;
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiBytes + 1024;  // + alignment slack
};

struct GemmGeom {
  int M, N, K;          // per group: rows, output columns, reduction length (= taps * kb_inner * 64 when windowed)
  int kb_inner;         // k-blocks per tap (inner A coordinate wraps every kb_inner blocks); K-blocks total = taps * kb_inner
  int num_k_blocks;
  int groups;           // independent problems sharing A rows (grouped conv); 1 otherwise
  int a_group_c0;       // inner A coordinate offset per group
  int b_group_rows;     // B row offset per group (= N)
  int n_umma;           // UMMA N (256, or N rounded up to 16 when N < 256)
  // M-tail split-K (EPI_RESID_F32 only): the partially filled last row of tiles is not scheduled as full tiles (which
  // would cost a whole extra wave when (M / tile_m) * n_tiles is a multiple of the cluster count, e.g. 49 * 6 = 294 on
  // 74 clusters) but as n_tiles * tail_pieces short work items, each covering kb_per_piece K-blocks, whose raw
  // accumulators are added (fp32 atomics) into `tail_ws` [tile rows, N]; a small kernel then applies the epilogue.
  int tail_pieces;      // 0 = off
  int kb_per_piece;
  float* tail_ws;
  // MN-major operands (C = A^T-stored x B^T-stored): the operand is given as [K rows, MN cols] row-major — a contraction over
  // the ROWS of an activation matrix (dW = dY^T X, transformer backward) or a gathered embedding matrix used as B_all^T
  // (InfoNCE gradient).  TMA then stages [64 k-rows][64 MN elements = 128 B] boxes (one per 64-wide MN chunk, 8 KB apart)
  // and the UMMA descriptors / instruction descriptor select the MN-major canonical layout: no transpose kernel.
  int a_mn;
  int b_mn;
};

// MN-major bf16 operand, 128-byte swizzle: atoms of 64 MN elements x 8 k-rows (1024 B); k-groups SBO = 1024 B apart, 64-wide MN
// chunks LBO = chunk_bytes apart (CUTLASS make_umma_desc<Major::MN>: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
OPB_DEVICE uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct SmemBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint64_t resid_full[4][4];   // per epilogue warp: residual-chunk ring (TMA-epilogue path)
  uint32_t tmem_base;
};

// Optional per-phase cycle accounting of the fp32 residual epilogue (build with EXTRA=-DOPB_GEMM_TIMING; read with
// opb_gemm_timing_dump()): lane 0 of epilogue warp 0 of every CTA accumulates clock64() deltas.
#ifdef OPB_GEMM_TIMING
__device__ unsigned long long g_gemm_t[12];
#define OPB_GT(var) const long long var = clock64()
#define OPB_GACC(i, a, b) do { gt_acc[i] += (b) - (a); } while (0)
#else
#define OPB_GT(var) do {} while (0)
#define OPB_GACC(i, a, b) do {} while (0)
#endif

// row statistics for the fused-LayerNorm epilogue: either precomputed (mu, rstd) or reduced here from partial records
OPB_DEVICE void load_ln_stats(const GemmEpilogue& ep, int row, int M, float& mu, float& rs) {
  mu = 0.f; rs = 1.f;
  const int rc = row < M ? row : M - 1;
  if (ep.ln_partial != nullptr) {
    float s1 = 0.f, s2 = 0.f;
    // independent 8-byte loads (coalesced over the warp's 32 rows); unrolled so that several are in flight — this runs
    // before the wait on the accumulator, i.e. in the epilogue warps' idle time
#pragma unroll 8
    for (int p = 0; p < ep.ln_parts; ++p) {
      const float2 v = *reinterpret_cast<const float2*>(ep.ln_partial + (static_cast<long>(p) * M + rc) * 2);
      s1 += v.x; s2 += v.y;
    }
    mu = s1 / ep.ln_dim;
    rs = rsqrtf(fmaxf(s2 / ep.ln_dim - mu * mu, 0.f) + ep.ln_eps);
  } else if (ep.ln_mu != nullptr) {
    mu = ep.ln_mu[rc];
    rs = ep.ln_rstd[rc];
  }
}

template <int CG, int EPI, bool TMAEPI>
__global__ void __launch_bounds__(EpiCfg<EPI, TMAEPI>::kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_o2,
                 const GemmEpilogue ep, const GemmGeom geo) {
  const int M = geo.M, N = geo.N;
  using Cfg = GemmCfg<CG, EPI, TMAEPI>;
  using ECfg = EpiCfg<EPI, TMAEPI>;
  // no static shared memory: the dynamic window starts at the 1024-aligned base of the CTA's shared memory (checked);
  // deriving every pointer from this array keeps the shared state space visible to the compiler (LDS / STS)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  SmemBars* bars = reinterpret_cast<SmemBars*>(smem + Cfg::kStages * Cfg::kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);

  const int tile_m_rows = kBlockM * CG;
  const int num_m_tiles = (geo.tail_pieces > 0) ? M / tile_m_rows : (M + tile_m_rows - 1) / tile_m_rows;   // full tiles only when the tail is split
  const int num_n_tiles = (N + kBlockN - 1) / kBlockN;
  const int tiles_per_group = num_m_tiles * num_n_tiles;
  const int num_main = tiles_per_group * geo.groups;
  const int num_tiles = num_main + num_n_tiles * geo.tail_pieces;
  const int num_k_blocks = geo.num_k_blocks;
  const int first_tile = blockIdx.x / CG;
  const int tile_stride = gridDim.x / CG;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], EpiCfg<EPI, TMAEPI>::kWarps * CG);
    }
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < 4; ++i) mbar_init(&bars->resid_full[w][i], 1);
    if constexpr (TMAEPI) {
      tma_prefetch_desc(&tm_o);
      tma_prefetch_desc(&tm_o2);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<CG>(&bars->tmem_base, kAccStages * kBlockN);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint64_t* full_bar0 = bars->full;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        int grp = 0, m_blk, n_blk, kb0 = 0, kb1 = num_k_blocks;
        if (tile < num_main) {
          grp = tile / tiles_per_group;
          const int tin = tile - grp * tiles_per_group;
          m_blk = tin / num_n_tiles;      // n-fastest rasterisation: the tiles running concurrently share a few A row
          n_blk = tin % num_n_tiles;      // panels (L2-resident) and sweep B; m-fastest re-streamed A from HBM per n-tile
        } else {            // M-tail piece (groups == 1, one tap)
          const int t = tile - num_main;
          n_blk = t / geo.tail_pieces;
          m_blk = num_m_tiles;
          kb0 = (t % geo.tail_pieces) * geo.kb_per_piece;
          kb1 = min(num_k_blocks, kb0 + geo.kb_per_piece);
        }
        const int a_row = m_blk * tile_m_rows + static_cast<int>(cta_rank) * kBlockM;
        const int b_row = grp * geo.b_group_rows + n_blk * kBlockN + static_cast<int>(cta_rank) * (geo.n_umma / 2);
        const int a_c0 = grp * geo.a_group_c0;
        int kin = kb0, tap = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(&full_bar0[stage], Cfg::kStageBytes);
            if (geo.a_mn) {
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i) tma_load_2d(&tm_a, &full_bar0[stage], sa + i * 8192, a_row + 64 * i, kb * kBlockK);
            } else {
              tma_load_3d(&tm_a, &full_bar0[stage], sa, a_c0 + kin * kBlockK, tap, a_row);
            }
            if (geo.b_mn) {
#pragma unroll
              for (int i = 0; i < Cfg::kBRows / 64; ++i) tma_load_2d(&tm_b, &full_bar0[stage], sb + i * 8192, b_row + 64 * i, kb * kBlockK);
            } else {
              tma_load_2d(&tm_b, &full_bar0[stage], sb, kb * kBlockK, b_row);
            }
          } else {
            if (is_leader) mbar_arrive_expect_tx(&full_bar0[stage], Cfg::kStageBytes * 2);
            if (geo.a_mn) {
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i) tma_load_2d_2sm(&tm_a, &full_bar0[stage], sa + i * 8192, a_row + 64 * i, kb * kBlockK);
            } else {
              tma_load_3d_2sm(&tm_a, &full_bar0[stage], sa, a_c0 + kin * kBlockK, tap, a_row);
            }
            if (geo.b_mn) {
#pragma unroll
              for (int i = 0; i < Cfg::kBRows / 64; ++i) tma_load_2d_2sm(&tm_b, &full_bar0[stage], sb + i * 8192, b_row + 64 * i, kb * kBlockK);
            } else {
              tma_load_2d_2sm(&tm_b, &full_bar0[stage], sb, kb * kBlockK, b_row);
            }
          }
          if (++kin == geo.kb_inner) { kin = 0; ++tap; }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (is_leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM * CG, geo.n_umma) | (geo.a_mn ? (1u << 15) : 0u) | (geo.b_mn ? (1u << 16) : 0u);
      // descriptor advance per 16-deep MMA: 32 B inside the swizzle row (K-major) or two 8-row k-groups = 2048 B (MN-major)
      const uint64_t step_a = geo.a_mn ? (2048 >> 4) : 2, step_b = geo.b_mn ? (2048 >> 4) : 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        int kb0 = 0, kb1 = num_k_blocks;
        if (tile >= num_main) {
          kb0 = ((tile - num_main) % geo.tail_pieces) * geo.kb_per_piece;
          kb1 = min(num_k_blocks, kb0 + geo.kb_per_piece);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = geo.a_mn ? make_sw128_mn_desc(sa, 8192) : make_sw128_kmajor_desc(sa);
          const uint64_t db = geo.b_mn ? make_sw128_mn_desc(sb, 8192) : make_sw128_kmajor_desc(sb);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            umma_bf16<CG>(tmem_d, da + step_a * k, db + step_b * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit<CG>(&bars->empty[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(&bars->tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
#ifdef OPB_GEMM_TIMING
    long long gt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (tile >= num_main) {
        // ---- M-tail piece: the raw partial accumulators of the valid rows go to THIS PIECE's slab of the fp32 workspace
        // ([piece][tile row][N], plain stores); the fix-up kernel adds the slabs in piece order — deterministic, no memset, no
        // atomics (round 1 added all pieces into one tile with fp32 atomics: run-to-run different low bits) ----
        const int n_blk_t = (tile - num_main) / geo.tail_pieces;
        const int piece = (tile - num_main) % geo.tail_pieces;
        const int rl = static_cast<int>(cta_rank) * kBlockM + q * 32 + lane;     // row inside the tail tile
        const bool rv = rl < (M - num_m_tiles * tile_m_rows);
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr_t = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
        float* wrow = geo.tail_ws + (static_cast<long>(piece) * tile_m_rows + rl) * N + n_blk_t * kBlockN;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(taddr_t + c, v);
          tmem_ld_wait();
          if (c + 32 == kBlockN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
          }
          if (rv) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (n_blk_t * kBlockN + c + j < N)
                *reinterpret_cast<float4*>(wrow + c + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
        }
        continue;
      }
      const int grp = tile / tiles_per_group;
      const int tin = tile - grp * tiles_per_group;
      const int m_blk = tin / num_n_tiles;
      const int n_blk = tin % num_n_tiles;
      const int row = m_blk * tile_m_rows + static_cast<int>(cta_rank) * kBlockM + q * 32 + lane;
      bool row_ok = row < M;
      if (ep.out_group > 0 && ep.out_group_valid > 0 && (row % ep.out_group) >= ep.out_group_valid) row_ok = false;
      const int gcol0 = grp * N;      // first global output column of this group
      if constexpr (TMAEPI) {
        // ---------------------------------------------------------------------------------------------
        // Coalesced epilogue: every global access is a TMA bulk copy of a [32 rows][128 B] chunk staged in this
        // warp's swizzled shared-memory ring; threads only touch TMEM, registers and their own smem row.
        // (The direct path below issues 16-byte accesses at a 3-24 KB row pitch — 32 L1 wavefronts per
        // instruction — and made the K = 1536 GEMMs epilogue-bound.)
        // ---------------------------------------------------------------------------------------------
        const int ew = warp - 2;                                        // staging slot of this warp
        const int hf = ew >> 2;                                         // column half (8-warp epilogues only)
        uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes + ew * ECfg::kWarpBytes;
        const int row0 = m_blk * tile_m_rows + static_cast<int>(cta_rank) * kBlockM + q * 32;   // warp's first row
        const int col0 = n_blk * kBlockN;
        float ln_mu, ln_rs;
        load_ln_stats(ep, row, M, ln_mu, ln_rs);
        float st_sum = 0.f, st_sq = 0.f;
        // Per-column epilogue coefficients of this tile -> this warp's smem copy ([3][256] fp32), in AFFINE form so the
        // inner loops are branch-free and every load is unconditional (the compiler can batch them):
        //     x = acc * (ln_rs * P[n]) + (nlm * Q[n] + R[n]),   nlm = -ln_rs * ln_mu
        //     P = colscale | gamma | 1,   Q = colsum * P,   R = bias * P;   columns >= N get P = Q = R = 0
        uint8_t* cv_s = stg + ECfg::kColVecOff;
        {
          const float* v2 = (EPI == EPI_RESID_F32) ? ep.gamma : ep.colscale;
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int idx = k * 128 + lane * 4;
            float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = p4, r4 = p4;
            if (col0 + idx < N) {
              p4 = v2 != nullptr ? *reinterpret_cast<const float4*>(v2 + col0 + idx) : make_float4(1.f, 1.f, 1.f, 1.f);
              if (ep.ln_colsum != nullptr) {
                const float4 c4 = *reinterpret_cast<const float4*>(ep.ln_colsum + col0 + idx);
                q4 = make_float4(c4.x * p4.x, c4.y * p4.y, c4.z * p4.z, c4.w * p4.w);
              }
              if (ep.bias != nullptr) {
                const float4 b4 = *reinterpret_cast<const float4*>(ep.bias + col0 + idx);
                r4 = make_float4(b4.x * p4.x, b4.y * p4.y, b4.z * p4.z, b4.w * p4.w);
              }
            }
            sts128(cv_s + 4 * idx, p4);
            sts128(cv_s + 4 * (256 + idx), q4);
            sts128(cv_s + 4 * (512 + idx), r4);
          }
          __syncwarp();
        }
        const float nlm = -ln_rs * ln_mu;
        if constexpr (ECfg::kF32) {
          uint8_t* bufB0 = stg + ECfg::kRing * 4096;
          uint64_t* rbar = bars->resid_full[ew];
          const bool has_res = (EPI == EPI_RESID_F32) && ep.resid != nullptr;
          // ring slot / barrier phase bookkeeping is continuous across tiles: chunk counter gc = it * 8 + c
          const int gc0 = it * 8;
          OPB_GT(g_t0);
          // the two slots refilled below were last stored from by chunks 4 and 5 of the previous tile: everything but
          // the two newest store groups (chunks 6, 7 -> the other two slots) must have finished reading shared memory
          if (lane == 0) { if (ep.dbg & 2) bulk_wait_read<0>(); else bulk_wait_read<2>(); }
          __syncwarp();
          if (has_res && lane == 0) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              const int slot = (gc0 + pc) & 3;
              mbar_arrive_expect_tx(&rbar[slot], 4096);
              tma_load_2d(&tm_o, &rbar[slot], stg + slot * 4096, col0 + 32 * pc, row0);
            }
          }
          OPB_GT(g_t1);
          mbar_wait(&bars->tmem_full[acc], acc_phase);
          tc_fence_after();
          OPB_GT(g_t2);
          OPB_GACC(0, g_t0, g_t1); OPB_GACC(1, g_t1, g_t2); OPB_GACC(10, 0, 1);
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            const int gc = gc0 + c;
            const int slot = gc & 3;
            uint8_t* buf = stg + slot * 4096;
            uint32_t v[32];
            __syncwarp();
            OPB_GT(g_c0);
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            OPB_GT(g_c1);
            if (c == 7) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
            }
            if (lane == 0) {
              // the slot of chunk c+2 was last stored from by chunk c-2: every store group except the newest one must
              // have finished reading shared memory (this also covers the bf16-copy buffer written below)
              bulk_wait_read<1>();
              if (has_res && c + 2 < 8) {
                const int ns = (gc + 2) & 3;
                mbar_arrive_expect_tx(&rbar[ns], 4096);
                tma_load_2d(&tm_o, &rbar[ns], stg + ns * 4096, col0 + 32 * (c + 2), row0);
              }
            }
            __syncwarp();
            OPB_GT(g_c2);
            float x[32];
            {
              float4 pp[8], qq[8], rr[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                pp[k] = lds128(cv_s + 4 * (c * 32 + 4 * k));
                qq[k] = lds128(cv_s + 4 * (256 + c * 32 + 4 * k));
                rr[k] = lds128(cv_s + 4 * (512 + c * 32 + 4 * k));
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                x[4 * k] = fmaf(__uint_as_float(v[4 * k]), ln_rs * pp[k].x, fmaf(nlm, qq[k].x, rr[k].x));
                x[4 * k + 1] = fmaf(__uint_as_float(v[4 * k + 1]), ln_rs * pp[k].y, fmaf(nlm, qq[k].y, rr[k].y));
                x[4 * k + 2] = fmaf(__uint_as_float(v[4 * k + 2]), ln_rs * pp[k].z, fmaf(nlm, qq[k].z, rr[k].z));
                x[4 * k + 3] = fmaf(__uint_as_float(v[4 * k + 3]), ln_rs * pp[k].w, fmaf(nlm, qq[k].w, rr[k].w));
              }
            }
            OPB_GT(g_c3);
            if (has_res) {
              mbar_wait(&rbar[slot], (gc >> 2) & 1);
              OPB_GT(g_c4);
              OPB_GACC(5, g_c3, g_c4);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 r = lds128(buf + sw128_off(lane, k));
                x[4 * k] += r.x; x[4 * k + 1] += r.y; x[4 * k + 2] += r.z; x[4 * k + 3] += r.w;
              }
            }
            if (ep.stats_out != nullptr) {      // columns >= N hold zeros (P = Q = R = 0 and zero-filled residual)
#pragma unroll
              for (int j = 0; j < 32; ++j) { st_sum += x[j]; st_sq += x[j] * x[j]; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              sts128(buf + sw128_off(lane, k), make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]));
            uint8_t* bufB = bufB0 + (c & 1) * 2048;            // [32 rows][64 B], alternating per chunk
            if (ep.out_bf16 != nullptr) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint4 o;
                o.x = pack_bf16x2(x[8 * k], x[8 * k + 1]);
                o.y = pack_bf16x2(x[8 * k + 2], x[8 * k + 3]);
                o.z = pack_bf16x2(x[8 * k + 4], x[8 * k + 5]);
                o.w = pack_bf16x2(x[8 * k + 6], x[8 * k + 7]);
                sts128u(bufB + sw64_off(lane, k), o);
              }
            }
            OPB_GT(g_c5);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tm_o, buf, col0 + c * 32, row0);
              if (ep.out_bf16 != nullptr) tma_store_2d(&tm_o2, bufB, col0 + c * 32, row0);
              bulk_commit();
            }
            OPB_GT(g_c6);
            OPB_GACC(2, g_c0, g_c1); OPB_GACC(3, g_c1, g_c2); OPB_GACC(4, g_c2, g_c3); OPB_GACC(6, g_c3, g_c5);
            OPB_GACC(7, g_c5, g_c6); OPB_GACC(8, g_c0, g_c6);
          }
          OPB_GT(g_t3);
          OPB_GACC(9, g_t0, g_t3);
          if (row_ok && ep.stats_out != nullptr)
            *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk) * M + row) * 2) = make_float2(st_sum, st_sq);
        } else {
          // ---- bf16 outputs: 64-column chunks ----
          mbar_wait(&bars->tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
          constexpr int kChunks = (EPI == EPI_GEGLU_BF16) ? 2 : 4;
          constexpr int kMine = kChunks / (ECfg::kWarps / 4);      // chunks per warp (8-warp mode: two warps share a lane quarter)
          const int gc0 = it * kMine;
#pragma unroll 1
          for (int c = hf * kMine; c < (hf + 1) * kMine; ++c) {
            uint8_t* buf = stg + ((gc0 + c) & 1) * 4096;
            if (lane == 0) bulk_wait_read<1>();     // slot last used two chunks ago
            __syncwarp();
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {        // two 32-column halves of the 64-column output chunk
              float y[32];
              if constexpr (EPI == EPI_GEGLU_BF16) {
                uint32_t g[32], l[32];
                const int ac = c * 64 + hh * 32;    // accumulator column of the gate half
                __syncwarp();
                tmem_ld32(taddr + ac, g);
                tmem_ld32(taddr + kBlockN / 2 + ac, l);
                tmem_ld_wait();
                if (c == (hf + 1) * kMine - 1 && hh == 1) {
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
                }
                {
                  float4 qg[8], rg[8], ql[8], rl[8];     // GeGLU: P == 1
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    qg[k] = lds128(cv_s + 4 * (256 + ac + 4 * k));
                    rg[k] = lds128(cv_s + 4 * (512 + ac + 4 * k));
                    ql[k] = lds128(cv_s + 4 * (256 + kBlockN / 2 + ac + 4 * k));
                    rl[k] = lds128(cv_s + 4 * (512 + kBlockN / 2 + ac + 4 * k));
                  }
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float g0 = fmaf(__uint_as_float(g[4 * k]), ln_rs, fmaf(nlm, qg[k].x, rg[k].x));
                    const float g1 = fmaf(__uint_as_float(g[4 * k + 1]), ln_rs, fmaf(nlm, qg[k].y, rg[k].y));
                    const float g2 = fmaf(__uint_as_float(g[4 * k + 2]), ln_rs, fmaf(nlm, qg[k].z, rg[k].z));
                    const float g3 = fmaf(__uint_as_float(g[4 * k + 3]), ln_rs, fmaf(nlm, qg[k].w, rg[k].w));
                    const float l0 = fmaf(__uint_as_float(l[4 * k]), ln_rs, fmaf(nlm, ql[k].x, rl[k].x));
                    const float l1 = fmaf(__uint_as_float(l[4 * k + 1]), ln_rs, fmaf(nlm, ql[k].y, rl[k].y));
                    const float l2 = fmaf(__uint_as_float(l[4 * k + 2]), ln_rs, fmaf(nlm, ql[k].z, rl[k].z));
                    const float l3 = fmaf(__uint_as_float(l[4 * k + 3]), ln_rs, fmaf(nlm, ql[k].w, rl[k].w));
                    y[4 * k] = gelu_erf(g0) * l0; y[4 * k + 1] = gelu_erf(g1) * l1;
                    y[4 * k + 2] = gelu_erf(g2) * l2; y[4 * k + 3] = gelu_erf(g3) * l3;
                  }
                }
                if (ep.stats_out != nullptr) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) { st_sum += y[j]; st_sq += y[j] * y[j]; }
                }
              } else {
                uint32_t v[32];
                const int ac = c * 64 + hh * 32;
                __syncwarp();
                tmem_ld32(taddr + ac, v);
                tmem_ld_wait();
                if (c == (hf + 1) * kMine - 1 && hh == 1) {
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
                }
                {
                  float4 pp[8], qq[8], rr[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    pp[k] = lds128(cv_s + 4 * (ac + 4 * k));
                    qq[k] = lds128(cv_s + 4 * (256 + ac + 4 * k));
                    rr[k] = lds128(cv_s + 4 * (512 + ac + 4 * k));
                  }
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    y[4 * k] = fmaf(__uint_as_float(v[4 * k]), ln_rs * pp[k].x, fmaf(nlm, qq[k].x, rr[k].x));
                    y[4 * k + 1] = fmaf(__uint_as_float(v[4 * k + 1]), ln_rs * pp[k].y, fmaf(nlm, qq[k].y, rr[k].y));
                    y[4 * k + 2] = fmaf(__uint_as_float(v[4 * k + 2]), ln_rs * pp[k].z, fmaf(nlm, qq[k].z, rr[k].z));
                    y[4 * k + 3] = fmaf(__uint_as_float(v[4 * k + 3]), ln_rs * pp[k].w, fmaf(nlm, qq[k].w, rr[k].w));
                  }
                  if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) y[j] = gelu_erf(y[j]);
                  }
                }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint4 o;
                o.x = pack_bf16x2(y[8 * k], y[8 * k + 1]);
                o.y = pack_bf16x2(y[8 * k + 2], y[8 * k + 3]);
                o.z = pack_bf16x2(y[8 * k + 4], y[8 * k + 5]);
                o.w = pack_bf16x2(y[8 * k + 6], y[8 * k + 7]);
                sts128u(buf + sw128_off(lane, hh * 4 + k), o);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int ocol = (EPI == EPI_GEGLU_BF16) ? n_blk * (kBlockN / 2) + c * 64 : col0 + c * 64;
              tma_store_2d(&tm_o, buf, ocol, row0);
              bulk_commit();
            }
          }
          if constexpr (EPI == EPI_GEGLU_BF16) {
            if (row_ok && ep.stats_out != nullptr)
            {
              *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk * 2 + hf) * M + row) * 2) = make_float2(st_sum, st_sq);
              if constexpr (ECfg::kWarps == 4)      // keep the [2 * n_tiles, M] record layout of the 8-warp mode
                *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk * 2 + 1) * M + row) * 2) = make_float2(0.f, 0.f);
            }
          }
        }
        continue;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;

      // output / residual row mapping (lets adapters scatter rows behind a CLS slot and broadcast a
      // positional table over the batch)
      long out_row = row;
      if (ep.out_group > 0) out_row = static_cast<long>(row / ep.out_group) * ep.out_group_stride + (row % ep.out_group) + ep.out_row_offset;
      long res_row = out_row;
      if (ep.resid_period > 0) res_row = (row % ep.resid_period) + ep.resid_row_offset;
      float ln_mu, ln_rs;
      load_ln_stats(ep, row, M, ln_mu, ln_rs);
      const bool has_ln = ep.ln_colsum != nullptr;
      float st_sum = 0.f, st_sq = 0.f;   // partial statistics of the stored values (next LayerNorm)
      // per-column epilogue vectors of this tile in shared memory (see the TMA path); grouped GEMMs index them by
      // global column and keep reading global memory
      bool use_cv = false;
      uint8_t* cv_s = smem;
      if constexpr (ECfg::kRegular && !TMAEPI) {
        use_cv = (geo.groups == 1);
        if (use_cv) {
          cv_s = smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes + (warp - 2) * ECfg::kWarpBytes;
          const float* v2 = (EPI == EPI_RESID_F32) ? ep.gamma : ep.colscale;
          const float* vecs[3] = {ep.ln_colsum, ep.bias, v2};
          const int tcol0 = n_blk * kBlockN;
          __syncwarp();
#pragma unroll
          for (int vv = 0; vv < 3; ++vv) {
            if (vecs[vv] != nullptr) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int idx = k * 128 + lane * 4;
                float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tcol0 + idx < N) t4 = *reinterpret_cast<const float4*>(vecs[vv] + tcol0 + idx);
                sts128(cv_s + 4 * (vv * 256 + idx), t4);
              }
            }
          }
          __syncwarp();
        }
      }
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();

      if constexpr (EPI == EPI_GEGLU_BF16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out) + out_row * ep.ldo + n_blk * (kBlockN / 2);
        const int n_out = N / 2;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 2; c += 32) {
          uint32_t g[32], l[32];
          __syncwarp();
          tmem_ld32(taddr + c, g);
          tmem_ld32(taddr + kBlockN / 2 + c, l);
          tmem_ld_wait();
          if (c + 32 == kBlockN / 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
          }
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n_blk * (kBlockN / 2) + c + j < n_out) {
                float ga[8], li[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { ga[e] = __uint_as_float(g[j + e]); li[e] = __uint_as_float(l[j + e]); }
                const int pc = n_blk * kBlockN + c + j;          // packed (interleaved) weight row of the gate half
                if (has_ln) {
                  float cg[8], cl[8];
                  if (use_cv) {
                    *reinterpret_cast<float4*>(cg) = lds128(cv_s + 4 * (c + j));
                    *reinterpret_cast<float4*>(cg + 4) = lds128(cv_s + 4 * (c + j + 4));
                    *reinterpret_cast<float4*>(cl) = lds128(cv_s + 4 * (kBlockN / 2 + c + j));
                    *reinterpret_cast<float4*>(cl + 4) = lds128(cv_s + 4 * (kBlockN / 2 + c + j + 4));
                  } else {
                  *reinterpret_cast<float4*>(cg) = *reinterpret_cast<const float4*>(ep.ln_colsum + pc);
                  *reinterpret_cast<float4*>(cg + 4) = *reinterpret_cast<const float4*>(ep.ln_colsum + pc + 4);
                  *reinterpret_cast<float4*>(cl) = *reinterpret_cast<const float4*>(ep.ln_colsum + pc + kBlockN / 2);
                  *reinterpret_cast<float4*>(cl + 4) = *reinterpret_cast<const float4*>(ep.ln_colsum + pc + kBlockN / 2 + 4);
                  }
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    ga[e] = ln_rs * (ga[e] - ln_mu * cg[e]);
                    li[e] = ln_rs * (li[e] - ln_mu * cl[e]);
                  }
                }
                if (ep.bias != nullptr) {
                  float bg[8], bl[8];
                  if (use_cv) {
                    *reinterpret_cast<float4*>(bg) = lds128(cv_s + 4 * (256 + c + j));
                    *reinterpret_cast<float4*>(bg + 4) = lds128(cv_s + 4 * (256 + c + j + 4));
                    *reinterpret_cast<float4*>(bl) = lds128(cv_s + 4 * (256 + kBlockN / 2 + c + j));
                    *reinterpret_cast<float4*>(bl + 4) = lds128(cv_s + 4 * (256 + kBlockN / 2 + c + j + 4));
                  } else {
                  *reinterpret_cast<float4*>(bg) = *reinterpret_cast<const float4*>(ep.bias + pc);
                  *reinterpret_cast<float4*>(bg + 4) = *reinterpret_cast<const float4*>(ep.bias + pc + 4);
                  *reinterpret_cast<float4*>(bl) = *reinterpret_cast<const float4*>(ep.bias + pc + kBlockN / 2);
                  *reinterpret_cast<float4*>(bl + 4) = *reinterpret_cast<const float4*>(ep.bias + pc + kBlockN / 2 + 4);
                  }
#pragma unroll
                  for (int e = 0; e < 8; ++e) { ga[e] += bg[e]; li[e] += bl[e]; }
                }
                float u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  u[e] = gelu_erf(ga[e]) * li[e];
                  st_sum += u[e];
                  st_sq += u[e] * u[e];
                }
                uint4 o;
                o.x = pack_bf16x2(u[0], u[1]);
                o.y = pack_bf16x2(u[2], u[3]);
                o.z = pack_bf16x2(u[4], u[5]);
                o.w = pack_bf16x2(u[6], u[7]);
                *reinterpret_cast<uint4*>(out + c + j) = o;
              }
            }
          }
        }
        if (row_ok && ep.stats_out != nullptr) {   // same [2 * n_tiles, M] layout as the 8-warp TMA epilogue
          *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk * 2) * M + row) * 2) = make_float2(st_sum, st_sq);
          *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk * 2 + 1) * M + row) * 2) = make_float2(0.f, 0.f);
        }
      } else if constexpr (EPI == EPI_LSE_PARTIAL) {
        // z = scale * acc.  Each thread owns one row of the tile: all reductions are thread-local.
        const float scale = *ep.scale_ptr;
        const int col0 = n_blk * kBlockN;
        const int tgt = row + ep.target_offset;
        const int nv = ep.n_valid > 0 ? ep.n_valid : N;
        float m = -INFINITY, ssum = 0.f, zsum = 0.f, best = -INFINITY, ztgt = 0.f;
        int best_idx = 0;
        bool has_tgt = false;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (c + 32 == kBlockN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
          }
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + c + j;
            const float z = (col < nv) ? scale * __uint_as_float(v[j]) : -INFINITY;
            v[j] = __float_as_uint(z);
            cm = fmaxf(cm, z);
            if (z > best) { best = z; best_idx = col; }
            if (col == tgt) { ztgt = z; has_tgt = true; }
          }
          if (cm > -INFINITY) {
            const float mn = fmaxf(m, cm);
            float add = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float z = __uint_as_float(v[j]);
              if (z > -INFINITY) { add += __expf(z - mn); zsum += z; }
            }
            ssum = ssum * __expf(m - mn) + add;
            m = mn;
          }
        }
        if (row_ok) {
          float* w = ep.ws + (static_cast<long>(n_blk) * M + row) * 8;
          *reinterpret_cast<float4*>(w) = make_float4(m, ssum, zsum, best);
          *reinterpret_cast<float4*>(w + 4) = make_float4(__int_as_float(best_idx), ztgt, has_tgt ? 1.f : 0.f, 0.f);
        }
      } else if constexpr (EPI == EPI_SOFTMAX_GRAD) {
        const float scale = *ep.scale_ptr;
        const float gscale = scale * ep.coef;
        const int col0 = n_blk * kBlockN;
        const int tgt = row + ep.target_offset;
        const float lse = row_ok ? ep.row_lse[row] : 0.f;
        const float hit = 1.f - ep.eps - ep.eps_i;
        const int nv = ep.n_valid > 0 ? ep.n_valid : N;
        float gz = 0.f;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (c + 32 == kBlockN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = col0 + c + j;
            if (!row_ok || col >= N) continue;
            float gq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float z = scale * __uint_as_float(v[j + e]);
              float g = __expf(z - lse) - ep.eps_i;
              if (col + e == tgt) g -= hit;
              if (col + e >= nv) g = 0.f;
              gz += g * z;
              gq[e] = g * gscale;
            }
            uint4 o;
            o.x = pack_bf16x2(gq[0], gq[1]);
            o.y = pack_bf16x2(gq[2], gq[3]);
            o.z = pack_bf16x2(gq[4], gq[5]);
            o.w = pack_bf16x2(gq[6], gq[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + out_row * ep.ldo + col) = o;
          }
        }
        if (row_ok) ep.ws[static_cast<long>(n_blk) * M + row] = gz;
      } else {
        const int col0 = n_blk * kBlockN;
#pragma unroll 1
        for (int c = 0; c < kBlockN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (c + 32 == kBlockN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (ep.dbg & 1) mbar_arrive_cluster(&bars->tmem_empty[acc], 0); else mbar_arrive_cluster_relaxed(&bars->tmem_empty[acc], 0); }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (!row_ok || col0 + c + j >= N) continue;
            const int col = gcol0 + col0 + c + j;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[j + e]);
            const int tcc = c + j;   // column inside the tile
            if (has_ln) {
              const float4 c0 = use_cv ? lds128(cv_s + 4 * tcc) : *reinterpret_cast<const float4*>(ep.ln_colsum + col);
              const float4 c1 = use_cv ? lds128(cv_s + 4 * (tcc + 4)) : *reinterpret_cast<const float4*>(ep.ln_colsum + col + 4);
              x[0] = ln_rs * (x[0] - ln_mu * c0.x); x[1] = ln_rs * (x[1] - ln_mu * c0.y);
              x[2] = ln_rs * (x[2] - ln_mu * c0.z); x[3] = ln_rs * (x[3] - ln_mu * c0.w);
              x[4] = ln_rs * (x[4] - ln_mu * c1.x); x[5] = ln_rs * (x[5] - ln_mu * c1.y);
              x[6] = ln_rs * (x[6] - ln_mu * c1.z); x[7] = ln_rs * (x[7] - ln_mu * c1.w);
            }
            if (ep.bias != nullptr) {
              const float4 b0 = use_cv ? lds128(cv_s + 4 * (256 + tcc)) : *reinterpret_cast<const float4*>(ep.bias + col);
              const float4 b1 = use_cv ? lds128(cv_s + 4 * (256 + tcc + 4)) : *reinterpret_cast<const float4*>(ep.bias + col + 4);
              x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
              x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
            if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16) {
              if (ep.colscale != nullptr) {
                const float4 s0 = use_cv ? lds128(cv_s + 4 * (512 + tcc)) : *reinterpret_cast<const float4*>(ep.colscale + col);
                const float4 s1 = use_cv ? lds128(cv_s + 4 * (512 + tcc + 4)) : *reinterpret_cast<const float4*>(ep.colscale + col + 4);
                x[0] *= s0.x; x[1] *= s0.y; x[2] *= s0.z; x[3] *= s0.w;
                x[4] *= s1.x; x[5] *= s1.y; x[6] *= s1.z; x[7] *= s1.w;
              }
              if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = gelu_erf(x[e]);
              }
              uint4 o;
              o.x = pack_bf16x2(x[0], x[1]);
              o.y = pack_bf16x2(x[2], x[3]);
              o.z = pack_bf16x2(x[4], x[5]);
              o.w = pack_bf16x2(x[6], x[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + out_row * ep.ldo + col) = o;
            } else {
              if constexpr (EPI == EPI_RESID_F32) {
                if (ep.gamma != nullptr) {
                  const float4 g0 = use_cv ? lds128(cv_s + 4 * (512 + tcc)) : *reinterpret_cast<const float4*>(ep.gamma + col);
                  const float4 g1 = use_cv ? lds128(cv_s + 4 * (512 + tcc + 4)) : *reinterpret_cast<const float4*>(ep.gamma + col + 4);
                  x[0] *= g0.x; x[1] *= g0.y; x[2] *= g0.z; x[3] *= g0.w;
                  x[4] *= g1.x; x[5] *= g1.y; x[6] *= g1.z; x[7] *= g1.w;
                }
                if (ep.resid != nullptr) {
                  const float* r = ep.resid + res_row * ep.ldr + col;
                  const float4 r0 = *reinterpret_cast<const float4*>(r);
                  const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
                  x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
                  x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
                }
              }
              float* o = reinterpret_cast<float*>(ep.out) + out_row * ep.ldo + col;
              *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
              if constexpr (EPI == EPI_RESID_F32) {
                if (ep.stats_out != nullptr) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) { st_sum += x[e]; st_sq += x[e] * x[e]; }
                }
                if (ep.out_bf16 != nullptr) {
                  uint4 ob;
                  ob.x = pack_bf16x2(x[0], x[1]);
                  ob.y = pack_bf16x2(x[2], x[3]);
                  ob.z = pack_bf16x2(x[4], x[5]);
                  ob.w = pack_bf16x2(x[6], x[7]);
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_bf16) + out_row * ep.ldo_bf16 + col) = ob;
                }
              }
            }
          }
        }
        if constexpr (EPI == EPI_RESID_F32) {
          if (row_ok && ep.stats_out != nullptr)
            *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(n_blk) * M + row) * 2) = make_float2(st_sum, st_sq);
        }
      }
    }
#ifdef OPB_GEMM_TIMING
    if (warp == 2 && lane == 0) {
      for (int i = 0; i < 12; ++i) atomicAdd(&g_gemm_t[i], static_cast<unsigned long long>(gt_acc[i]));
    }
#endif
  }

  // ===================== teardown =====================
  if constexpr (TMAEPI) {
    if (warp >= 2 && lane == 0) bulk_wait_all<0>();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kAccStages * kBlockN);
  }
}

#ifdef OPB_GEMM_TIMING
extern "C" void opb_gemm_timing_dump(int reset) {
  unsigned long long t[12];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(t, g_gemm_t, sizeof(t));
  const double n = t[10] > 0 ? static_cast<double>(t[10]) : 1.0;
  printf("[gemm timing] fp32-residual epilogue, warp 0 of every CTA, %.0f tiles; avg cycles / tile: pre(store drain+prefetch)=%.0f "
         "wait_acc=%.0f | per tile over 8 chunks: tmem_ld=%.0f drain+prefetch=%.0f coef_math=%.0f resid_wait=%.0f "
         "resid_add+stats+sts=%.0f fence+store=%.0f chunks_total=%.0f | tile_total=%.0f\n",
         n, t[0] / n, t[1] / n, t[2] / n, t[3] / n, t[4] / n, t[5] / n, (double)(t[6] - t[5]) / n, t[7] / n, t[8] / n, t[9] / n);
  if (reset) {
    unsigned long long z[12] = {0};
    cudaMemcpyToSymbol(g_gemm_t, z, sizeof(z));
  }
}
#endif

// Epilogue of the split-K M-tail rows: x = resid + gamma * (rstd * (acc - mu * colsum) + bias) from the summed raw
// accumulators; writes the fp32 row, its bf16 copy and the per-256-column (sum, sum of squares) statistics records.
// One CTA (256 threads) per (tail row, 256-column tile) — grid (rows, n_tiles): the round-1 kernel walked the tiles of a row
// serially (6 dependent ~1 us iterations; 18 for the q/k/v projection of a small batch).
__global__ void __launch_bounds__(256)
gemm_tail_epilogue_kernel(const float* __restrict__ ws, const GemmEpilogue ep, int M, int N, int row0, int n_tiles, int pieces,
                          int tile_rows, int epi) {
  __shared__ float red[2][8];
  const int row = row0 + blockIdx.x;
  float mu, rs;
  if (ep.ln_partial != nullptr) {
    // the whole CTA works on ONE row: reduce its partial records cooperatively (a per-thread loop over up to 96
    // records made this kernel 3x slower than the separate finalize launch it replaced)
    float s1 = 0.f, s2 = 0.f;
    for (int p = threadIdx.x; p < ep.ln_parts; p += blockDim.x) {
      const float2 v = *reinterpret_cast<const float2*>(ep.ln_partial + (static_cast<long>(p) * M + row) * 2);
      s1 += v.x; s2 += v.y;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    s1 = s2 = 0.f;
    for (int i = 0; i < 8; ++i) { s1 += red[0][i]; s2 += red[1][i]; }
    mu = s1 / ep.ln_dim;
    rs = rsqrtf(fmaxf(s2 / ep.ln_dim - mu * mu, 0.f) + ep.ln_eps);
  } else {
    load_ln_stats(ep, row, M, mu, rs);
  }
  const float* w = ws + static_cast<long>(blockIdx.x) * N;
  const long slab = static_cast<long>(tile_rows) * N;
  {
    const int t = blockIdx.y;
    const int col = t * kBlockN + threadIdx.x;
    float x = 0.f;
    if (col < N) {
      for (int p = 0; p < pieces; ++p) x += w[p * slab + col];          // fixed piece order
      if (ep.ln_colsum) x = rs * (x - mu * ep.ln_colsum[col]);
      if (ep.bias) x += ep.bias[col];
      if (epi == EPI_STORE_BF16) {          // q/k/v projection of a small batch: (LN-folded acc + bias) * colscale -> bf16
        if (ep.colscale) x *= ep.colscale[col];
        reinterpret_cast<__nv_bfloat16*>(ep.out)[static_cast<long>(row) * ep.ldo + col] = __float2bfloat16(x);
      } else {
        if (ep.gamma) x *= ep.gamma[col];
        if (ep.resid) x += ep.resid[static_cast<long>(row) * ep.ldr + col];
        reinterpret_cast<float*>(ep.out)[static_cast<long>(row) * ep.ldo + col] = x;
        if (ep.out_bf16) reinterpret_cast<__nv_bfloat16*>(ep.out_bf16)[static_cast<long>(row) * ep.ldo_bf16 + col] = __float2bfloat16(x);
      }
    }
    if (ep.stats_out != nullptr) {
      float s1 = warp_sum(x), s2 = warp_sum(x * x);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
      __syncthreads();
      if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
        *reinterpret_cast<float2*>(ep.stats_out + (static_cast<long>(t) * M + row) * 2) = make_float2(a, b);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

// 2D bf16 row-major [rows, cols] with row pitch ld (elements); box = 64 cols x box_rows, 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return OPB_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) return OPB_ERR_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OPB_OK : OPB_ERR_CUDA;
}

// generic 2D row-major tensor map for the epilogue's TMA loads / stores: box = box_cols x box_rows, 128B swizzle
static int make_tmap_2d_out(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_cols, uint32_t box_rows, bool swizzle64 = false) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return OPB_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * elem_bytes) % 16 != 0) return OPB_ERR_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OPB_OK : OPB_ERR_CUDA;
}

// 3D view of the A operand: dims {k_inner, taps, rows}; element (c, j, r) lives at ptr + r*row_stride + j*tap_stride + c.
// A plain [rows, K] matrix is the taps == 1 case.  box = 64 x 1 x box_rows, 128B swizzle (same smem image as 2D).
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t k_inner, uint64_t taps, uint64_t tap_stride,
                      uint64_t rows, uint64_t row_stride, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return OPB_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (tap_stride * 2) % 16 != 0 || (row_stride * 2) % 16 != 0)
    return OPB_ERR_INVALID;
  cuuint64_t gdim[3] = {k_inner, taps, rows};
  cuuint64_t gstride[2] = {tap_stride * 2, row_stride * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kBlockK), 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OPB_OK : OPB_ERR_CUDA;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int CG, int EPI, bool TMAEPI>
static int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                         const GemmEpilogue& ep, const GemmGeom& geo, cudaStream_t stream) {
  using Cfg = GemmCfg<CG, EPI, TMAEPI>;
  auto kern = gemm_bf16_kernel<CG, EPI, TMAEPI>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return OPB_ERR_CUDA;
    configured = true;
  }
  const int tile_m_rows = kBlockM * CG;
  const int n_tiles_n = (geo.N + kBlockN - 1) / kBlockN;
  const int num_tiles = geo.tail_pieces > 0 ? (geo.M / tile_m_rows) * n_tiles_n + n_tiles_n * geo.tail_pieces
                                            : ((geo.M + tile_m_rows - 1) / tile_m_rows) * n_tiles_n * geo.groups;
  int clusters = sm_count() / CG;
  if (clusters > num_tiles) clusters = num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(EpiCfg<EPI, TMAEPI>::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, to, to2, ep, geo);
  return e == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// Chooses the coalesced TMA-store epilogue whenever the output is a plain [M, N] matrix (no row remapping,
// residual updated in place); the direct-store epilogue handles the adapter scatter cases and the InfoNCE epilogues.
template <int CG, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, const GemmGeom& geo,
                       cudaStream_t stream) {
  constexpr bool kCanTma = (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_GEGLU_BF16 ||
                            EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32);
  if constexpr (kCanTma) {
    static const char* env = getenv("OPB_GEMM_TMA_EPILOGUE");
    const bool f32 = (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32);
    const int eb = f32 ? 4 : 2;
    const long n_out = (EPI == EPI_GEGLU_BF16) ? geo.N / 2 : geo.N;
    // OPB_GEMM_TMA_EPILOGUE: 0 = direct stores everywhere, 1 = TMA-store epilogue everywhere, 2 = TMA only for the fp32
    // residual epilogues
    const char mode = env != nullptr ? env[0] : '1';
    bool ok = mode != '0' && !(mode == '2' && !f32) && geo.groups == 1 && ep.out_group == 0 && ep.resid_period == 0 &&
              (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && (ep.ldo * eb) % 16 == 0 && (n_out * eb) % 16 == 0;
    if (EPI == EPI_RESID_F32 && ep.resid != nullptr && (ep.resid != ep.out || ep.ldr != ep.ldo)) ok = false;
    if (ep.out_bf16 != nullptr && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) != 0 || (ep.ldo_bf16 * 2) % 16 != 0))
      ok = false;
    if (ok) {
      CUtensorMap to, to2;
      int rc = make_tmap_2d_out(&to, ep.out, eb, geo.M, n_out, ep.ldo, f32 ? 32 : 64, 32);
      if (rc != OPB_OK) return rc;
      if (ep.out_bf16 != nullptr) {
        rc = make_tmap_2d_out(&to2, ep.out_bf16, 2, geo.M, geo.N, ep.ldo_bf16, 32, 32, true);
        if (rc != OPB_OK) return rc;
      } else {
        to2 = to;
      }
      return launch_gemm_t<CG, EPI, true>(ta, tb, to, to2, ep, geo, stream);
    }
  }
  return launch_gemm_t<CG, EPI, false>(ta, tb, ta, ta, ep, geo, stream);
}

static int dispatch_gemm(int cta_group, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep,
                         const GemmGeom& geo, cudaStream_t stream) {
#define OPB_DISPATCH(CGV)                                                                       \
  switch (epi) {                                                                                \
    case EPI_STORE_BF16: return launch_gemm<CGV, EPI_STORE_BF16>(ta, tb, ep, geo, stream);      \
    case EPI_GELU_BF16: return launch_gemm<CGV, EPI_GELU_BF16>(ta, tb, ep, geo, stream);        \
    case EPI_GEGLU_BF16: return launch_gemm<CGV, EPI_GEGLU_BF16>(ta, tb, ep, geo, stream);      \
    case EPI_RESID_F32: return launch_gemm<CGV, EPI_RESID_F32>(ta, tb, ep, geo, stream);        \
    case EPI_STORE_F32: return launch_gemm<CGV, EPI_STORE_F32>(ta, tb, ep, geo, stream);        \
    case EPI_LSE_PARTIAL: return launch_gemm<CGV, EPI_LSE_PARTIAL>(ta, tb, ep, geo, stream);    \
    case EPI_SOFTMAX_GRAD: return launch_gemm<CGV, EPI_SOFTMAX_GRAD>(ta, tb, ep, geo, stream);  \
    default: return OPB_ERR_INVALID;                                                            \
  }
  if (cta_group == 1) { OPB_DISPATCH(1) } else { OPB_DISPATCH(2) }
#undef OPB_DISPATCH
}

int gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int epi, const GemmEpilogue& ep_in,
              int cta_group, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return OPB_ERR_INVALID;
  GemmEpilogue ep = ep_in;
  static const char* env_dbg = getenv("OPB_GEMM_DBG");
  if (env_dbg != nullptr) ep.dbg = atoi(env_dbg);
  if (K % 8 != 0 || N % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0) return OPB_ERR_INVALID;
  if (epi == EPI_GEGLU_BF16 && N % kBlockN != 0) return OPB_ERR_INVALID;
  if (cta_group == 0) cta_group = (M > 2 * kBlockM) ? 2 : 1;
  if (cta_group != 1 && cta_group != 2) return OPB_ERR_INVALID;
  GemmGeom geo;
  geo.M = M; geo.N = N; geo.K = K;
  geo.num_k_blocks = (K + kBlockK - 1) / kBlockK;
  geo.kb_inner = geo.num_k_blocks;      // one "tap": the inner coordinate never wraps
  geo.groups = 1;
  geo.a_group_c0 = 0;
  geo.b_group_rows = 0;
  geo.n_umma = (N >= kBlockN) ? kBlockN : ((N + 15) / 16) * 16;
  geo.tail_pieces = 0;
  geo.kb_per_piece = 0;
  geo.tail_ws = nullptr;
  geo.a_mn = geo.b_mn = 0;
  if (cta_group == 2 && geo.n_umma % 32 != 0) cta_group = 1;   // each CTA of a pair stages n_umma / 2 rows of B
  // M-tail split-K: worth it when dropping the partial row of tiles saves a whole wave
  static const char* env_tail = getenv("OPB_GEMM_TAIL_SPLITK");
  const bool tail_ok = epi == EPI_RESID_F32 && ep.workspace != nullptr && ep.out_group == 0 && ep.resid_period == 0 &&
                       !(env_tail != nullptr && env_tail[0] == '0');
  // Small-M mode (M < 256: a handful of texts through the 4B stack): N / 256 = 6 output tiles would leave 142 SMs idle while six
  // CTAs stream the whole weight (fc2: 18.9 MB at ~40 GB/s per CTA — 8 texts took 150 us per layer, launch-free or not).  The
  // GEMM becomes ONE 256-row "tail" (CTA pairs) whose K range is split over all clusters; the fix-up kernel applies the epilogue.
  const bool small_ok = (epi == EPI_RESID_F32 || (epi == EPI_STORE_BF16 && ep.stats_out == nullptr)) && ep.workspace != nullptr &&
                        ep.out_group == 0 && ep.resid_period == 0 && !(env_tail != nullptr && env_tail[0] == '0');
  const bool small_m = small_ok && M < 2 * kBlockM && N >= kBlockN && geo.num_k_blocks >= 8 &&
                       ep.workspace_bytes >= 2L * (2 * kBlockM) * N * 4;
  if (small_m) cta_group = 2;
  const int tile_m = kBlockM * cta_group;
  const int m_full = M / tile_m, tail_rows = M % tile_m;
  const int n_t = (N + kBlockN - 1) / kBlockN;
  // (measured: pays for K >= 3072 — fc2 232 -> 212 us; for K = 1536 the memset + atomics + tail kernel cost more than
  //  the saved wave — out_proj 78 -> 89 us — so short-K GEMMs keep the plain schedule unless the whole GEMM is one small tile)
  if ((tail_ok || small_m) && tail_rows > 0 && (small_m || (m_full > 0 && geo.num_k_blocks >= 48)) &&
      ep.workspace_bytes >= 2L * tile_m * N * 4) {
    const int clusters = sm_count() / cta_group;
    const int waves_all = ((m_full + 1) * n_t + clusters - 1) / clusters;
    const int waves_main = (m_full * n_t + clusters - 1) / clusters;
    if (small_m || waves_main < waves_all) {
      int pieces = clusters / n_t;
      if (pieces > geo.num_k_blocks) pieces = geo.num_k_blocks;
      const long slabs = ep.workspace_bytes / (static_cast<long>(tile_m) * N * 4);      // one fp32 [tile_m, N] slab per piece
      if (pieces > slabs) pieces = static_cast<int>(slabs);
      if (pieces >= 2) {
        geo.kb_per_piece = (geo.num_k_blocks + pieces - 1) / pieces;
        geo.tail_pieces = (geo.num_k_blocks + geo.kb_per_piece - 1) / geo.kb_per_piece;
        geo.tail_ws = reinterpret_cast<float*>(ep.workspace);
      }
    }
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_3d(&ta, A, K, 1, lda, M, lda, kBlockM);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, kBlockN / cta_group);
  if (rc != OPB_OK) return rc;
  rc = dispatch_gemm(cta_group, epi, ta, tb, ep, geo, stream);
  if (rc != OPB_OK || geo.tail_pieces == 0) return rc;
  gemm_tail_epilogue_kernel<<<dim3(tail_rows, n_t), 256, 0, stream>>>(geo.tail_ws, ep, M, N, m_full * tile_m, n_t, geo.tail_pieces, tile_m, epi);
  return cudaGetLastError() == cudaSuccess ? OPB_OK : OPB_ERR_CUDA;
}

// C[M, N] = epilogue(A . B^T) with either operand given MN-major (see GemmGeom::a_mn): A as [K, M] row-major (pitch lda), B as
// [K, N] row-major (pitch ldb).  K may be any length (TMA zero-fills past the last row); STORE epilogues only.
int gemm_bf16_t(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K, int epi,
                const GemmEpilogue& ep, int cta_group, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || lda % 8 != 0 || ldb % 8 != 0 || N % 8 != 0) return OPB_ERR_INVALID;
  if (epi != EPI_STORE_BF16 && epi != EPI_STORE_F32) return OPB_ERR_INVALID;
  if ((!a_mn && K % 8 != 0) || (!b_mn && K % 8 != 0) || (a_mn && M % 8 != 0)) return OPB_ERR_INVALID;
  if (cta_group == 0) cta_group = (M > 2 * kBlockM) ? 2 : 1;
  GemmGeom geo;
  geo.M = M; geo.N = N; geo.K = K;
  geo.num_k_blocks = (K + kBlockK - 1) / kBlockK;
  geo.kb_inner = geo.num_k_blocks;
  geo.groups = 1;
  geo.a_group_c0 = 0;
  geo.b_group_rows = 0;
  geo.n_umma = (N >= kBlockN) ? kBlockN : ((N + 15) / 16) * 16;
  if (b_mn && geo.n_umma % 64 != 0) geo.n_umma = (geo.n_umma + 63) / 64 * 64;     // whole 64-wide MN chunks
  geo.tail_pieces = 0;
  geo.kb_per_piece = 0;
  geo.tail_ws = nullptr;
  geo.a_mn = a_mn ? 1 : 0;
  geo.b_mn = b_mn ? 1 : 0;
  if (cta_group == 2 && (geo.n_umma % 32 != 0 || (b_mn && geo.n_umma != kBlockN))) cta_group = 1;
  CUtensorMap ta, tb;
  int rc = a_mn ? make_tmap_bf16_2d(&ta, A, K, M, lda, 64) : make_tmap_bf16_3d(&ta, A, K, 1, lda, M, lda, kBlockM);
  if (rc != OPB_OK) return rc;
  rc = b_mn ? make_tmap_bf16_2d(&tb, B, K, N, ldb, 64) : make_tmap_bf16_2d(&tb, B, N, K, ldb, kBlockN / cta_group);
  if (rc != OPB_OK) return rc;
  return dispatch_gemm(cta_group, epi, ta, tb, ep, geo, stream);
}

int gemm_bf16_grouped_window(const void* X, const void* W, int rows, int groups, int c_pad, int taps, int n_per_group,
                             int epi, const GemmEpilogue& ep, cudaStream_t stream) {
  if (rows <= 0 || groups <= 0 || taps <= 0 || n_per_group <= 0) return OPB_ERR_INVALID;
  if (c_pad % kBlockK != 0 || n_per_group % 8 != 0 || n_per_group > kBlockN) return OPB_ERR_INVALID;
  if (epi == EPI_GEGLU_BF16 || epi == EPI_LSE_PARTIAL || epi == EPI_SOFTMAX_GRAD) return OPB_ERR_INVALID;
  GemmGeom geo;
  geo.M = rows; geo.N = n_per_group; geo.K = taps * c_pad;
  geo.kb_inner = c_pad / kBlockK;
  geo.num_k_blocks = taps * geo.kb_inner;
  geo.groups = groups;
  geo.a_group_c0 = c_pad;
  geo.b_group_rows = n_per_group;
  geo.n_umma = ((n_per_group + 15) / 16) * 16;
  geo.tail_pieces = 0;
  geo.kb_per_piece = 0;
  geo.tail_ws = nullptr;
  geo.a_mn = geo.b_mn = 0;
  const long row_stride = static_cast<long>(groups) * c_pad;
  CUtensorMap ta, tb;
  // dims {groups*c_pad, taps, rows}: tap j of output row r reads input row r + j (tap stride == row stride)
  int rc = make_tmap_bf16_3d(&ta, X, row_stride, taps, row_stride, rows, row_stride, kBlockM);
  if (rc != OPB_OK) return rc;
  rc = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(groups) * n_per_group, geo.K, geo.K, kBlockN);
  if (rc != OPB_OK) return rc;
  return dispatch_gemm(1, epi, ta, tb, ep, geo, stream);
}

}  // namespace opb

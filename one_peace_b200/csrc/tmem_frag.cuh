// tcgen05.ld / tcgen05.st fragment helpers shared by the attention kernels (attention_tcp.cu, attention_bwd_tc.cu).
//
// tcgen05.ld.16x256b hands each thread of a warp 2 rows x 2 adjacent fp32 columns per 8-column block (the mma.sync accumulator
// layout): thread `lane` of a warp that addresses 16 TMEM lanes holds rows lane / 4 and lane / 4 + 8, columns 8 blk + 2 (lane % 4)
// + {0, 1}; registers v[4 blk + {0, 1}] = first row, v[4 blk + {2, 3}] = second row.  tcgen05.st.16x128b with two registers per
// block writes packed bf16 pairs back over the same positions (packed column 4 blk + lane % 4), which is the layout a TMEM A
// operand of tcgen05.mma expects.
#pragma once
#include "common.cuh"

namespace opb {

OPB_DEVICE float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
OPB_DEVICE void tmem_ld_16x256b_x16(uint32_t taddr, uint32_t* p) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7]), "=r"(p[8]), "=r"(p[9]), "=r"(p[10]), "=r"(p[11]), "=r"(p[12]), "=r"(p[13]), "=r"(p[14]), "=r"(p[15]), "=r"(p[16]), "=r"(p[17]), "=r"(p[18]), "=r"(p[19]), "=r"(p[20]), "=r"(p[21]), "=r"(p[22]), "=r"(p[23]), "=r"(p[24]), "=r"(p[25]), "=r"(p[26]), "=r"(p[27]), "=r"(p[28]), "=r"(p[29]), "=r"(p[30]), "=r"(p[31]), "=r"(p[32]), "=r"(p[33]), "=r"(p[34]), "=r"(p[35]), "=r"(p[36]), "=r"(p[37]), "=r"(p[38]), "=r"(p[39]), "=r"(p[40]), "=r"(p[41]), "=r"(p[42]), "=r"(p[43]), "=r"(p[44]), "=r"(p[45]), "=r"(p[46]), "=r"(p[47]), "=r"(p[48]), "=r"(p[49]), "=r"(p[50]), "=r"(p[51]), "=r"(p[52]), "=r"(p[53]), "=r"(p[54]), "=r"(p[55]), "=r"(p[56]), "=r"(p[57]), "=r"(p[58]), "=r"(p[59]), "=r"(p[60]), "=r"(p[61]), "=r"(p[62]), "=r"(p[63])
               : "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* p) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7]), "=r"(p[8]), "=r"(p[9]), "=r"(p[10]), "=r"(p[11]), "=r"(p[12]), "=r"(p[13]), "=r"(p[14]), "=r"(p[15]), "=r"(p[16]), "=r"(p[17]), "=r"(p[18]), "=r"(p[19]), "=r"(p[20]), "=r"(p[21]), "=r"(p[22]), "=r"(p[23]), "=r"(p[24]), "=r"(p[25]), "=r"(p[26]), "=r"(p[27]), "=r"(p[28]), "=r"(p[29]), "=r"(p[30]), "=r"(p[31])
               : "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* p) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7]), "=r"(p[8]), "=r"(p[9]), "=r"(p[10]), "=r"(p[11]), "=r"(p[12]), "=r"(p[13]), "=r"(p[14]), "=r"(p[15])
               : "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* p) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7])
               : "r"(taddr)
               : "memory");
}

OPB_DEVICE void tmem_st_16x128b_x16(uint32_t taddr, const uint32_t* p) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x16.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
               :
               : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(p[8]), "r"(p[9]), "r"(p[10]), "r"(p[11]), "r"(p[12]), "r"(p[13]), "r"(p[14]), "r"(p[15]), "r"(p[16]), "r"(p[17]), "r"(p[18]), "r"(p[19]), "r"(p[20]), "r"(p[21]), "r"(p[22]), "r"(p[23]), "r"(p[24]), "r"(p[25]), "r"(p[26]), "r"(p[27]), "r"(p[28]), "r"(p[29]), "r"(p[30]), "r"(p[31]), "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t* p) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
               :
               : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(p[8]), "r"(p[9]), "r"(p[10]), "r"(p[11]), "r"(p[12]), "r"(p[13]), "r"(p[14]), "r"(p[15]), "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_st_16x128b_x4(uint32_t taddr, const uint32_t* p) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               :
               : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_st_16x128b_x2(uint32_t taddr, const uint32_t* p) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%4], {%0, %1, %2, %3};"
               :
               : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(taddr)
               : "memory");
}
OPB_DEVICE void tmem_st_wait_all() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (M = 128 rows = lanes, bf16 pairs packed in 32-bit columns) is read from
// tensor memory (FlashAttention-4's P V form) — P never touches shared memory
OPB_DEVICE void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
OPB_DEVICE uint64_t make_sw128_mn_desc64(uint32_t smem_addr) {     // MN-major, one 64-element MN chunk, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// all score columns of the warp's 16 lanes: NBLK8 blocks of 8 columns, 4 registers per block (2 rows x 2 columns)
template <int NBLK8>
OPB_DEVICE void load_scores(uint32_t taddr, uint32_t (&v)[4 * NBLK8]) {
  static_assert(NBLK8 % 2 == 0 && NBLK8 <= 28, "");
  constexpr int n16 = NBLK8 & 16, n8 = NBLK8 & 8, n4 = NBLK8 & 4, n2 = NBLK8 & 2;
  if constexpr (n16 != 0) tmem_ld_16x256b_x16(taddr, &v[0]);
  if constexpr (n8 != 0) tmem_ld_16x256b_x8(taddr + n16 * 8, &v[4 * n16]);
  if constexpr (n4 != 0) tmem_ld_16x256b_x4(taddr + (n16 + n8) * 8, &v[4 * (n16 + n8)]);
  if constexpr (n2 != 0) tmem_ld_16x256b_x2(taddr + (n16 + n8 + n4) * 8, &v[4 * (n16 + n8 + n4)]);
}

}  // namespace opb

// Micro-benchmark: cost of a tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16) as a function of N, the accumulator pattern
// and the A-operand source, measured as cycles per instruction over a long back-to-back chain issued by one thread and closed by one
// tcgen05.commit.  Why: the attention kernels issue many SMALL MMAs (P V: N = 64, 13 per tile; backward: N = 64, ~20 per sub-step) and
// their tensor pipe shows ~11 % activity — is a small-N MMA bound by a fixed issue cost rather than by its MACs (128 x N x 16)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_cost mma_issue_cost.cu ; run on a B200 (one CTA, one SM).
// Operands are whatever the (zero-initialised) shared memory holds: only the timing matters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: A from shared memory, ONE accumulator (dependent chain, the P V / dV pattern)
// mode 1: A from shared memory, accumulators rotate over 4 column slots (independent MMAs)
// mode 2: A from tensor memory (ts form), one accumulator
__global__ void bench(int n, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t da = kmajor_desc(smem_u32(smem)), db = kmajor_desc(smem_u32(smem + 32 * 1024));
    const uint32_t id = idesc(128, n);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {        // rep 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tbase + 256 + (mode == 1 ? (i & 3) * 64 : 0);     // accumulators in columns 256.. (A for mode 2 in 0..)
        if (mode == 2) mma_ts(d, tbase + 8 * (i & 7), db + 2 * (i & 3), id, i != 0);
        else mma_ss(d, da + 2 * (i & 3), db + 2 * (i & 3), id, i != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      const long long t1 = clock64();
      while (!try_wait(&bar, parity)) {}
      parity ^= 1;
      const long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 512;
  const char* names[3] = {"A smem, one accumulator", "A smem, 4 accumulators", "A tmem, one accumulator"};
  printf("tcgen05.mma kind::f16 cta_group::1 M=128 K=16: cycles per instruction over a chain of %d (issue / until commit arrives); "
         "MAC-bound floor = 128*N*16 / 4096 MAC/clk\n", iters);
  for (int mode = 0; mode < 3; ++mode)
    for (int n : {16, 32, 64, 112, 128, 208, 256}) {
      if (mode == 1 && n > 64) continue;
      bench<<<1, 128, 96 * 1024>>>(n, mode, iters, d);
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (cudaGetLastError() != cudaSuccess) { printf("error\n"); return 1; }
      printf("  %-26s N=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA (floor %5.1f)\n", names[mode], n, (double)h[0] / iters,
             (double)h[1] / iters, 128.0 * n * 16 / 4096);
    }
  return 0;
}

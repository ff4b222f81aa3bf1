// Companion of mma_issue_cost.cu: the same measurement for cta_group::2 (a CTA pair, M = 256: 128 rows of A from each CTA's shared
// memory, N / 2 rows of B from each) and for cta_group::1 with M = 64.  Question: does one M = 256 instruction cost the same ~120-160
// cycles as an M = 128 one (then a CTA pair per (batch, head) halves the attention kernels' issue-bound floors)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_cost_2cta mma_issue_cost_2cta.cu ; run on a B200.
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
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__global__ void bench(int m, int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const bool leader = CG == 1 || cluster_rank() == 0;
  if (threadIdx.x == 0 && leader) {
    const uint64_t da = kmajor_desc(smem_u32(smem)), db = kmajor_desc(smem_u32(smem + 32 * 1024));
    const uint32_t id = idesc(m, n);
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tbase + 256;
        const uint32_t acc = i != 0;
        if (CG == 1)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(d), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(id), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(d), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(id), "r"(acc) : "memory");
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"(static_cast<uint16_t>(1)) : "memory");
      const long long t1 = clock64();
      for (long spin = 0; !try_wait(&bar, parity); ++spin) if (spin > 50000000L) { out[0] = out[1] = -1; return; }   // never hang the GPU
      parity ^= 1;
      const long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 512;
  printf("tcgen05.mma kind::f16 K=16, cycles per instruction over a chain of %d (issue / until the commit arrives)\n", iters);
  for (int cg = 1; cg <= 2; ++cg)
    for (int m : {64, 128, 256}) {
      if ((cg == 1 && m == 256) || (cg == 2 && m == 64)) continue;
      for (int n : {64, 128, 208, 256}) {
        if (cg == 2 && n % 32) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cg); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 96 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cg == 1 ? cudaLaunchKernelEx(&cfg, bench<1>, m, n, iters, d) : cudaLaunchKernelEx(&cfg, bench<2>, m, n, iters, d);
        long long h[2] = {0, 0};
        cudaError_t e2 = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("  cta_group::%d M=%3d N=%3d: error %s\n", cg, m, n, cudaGetErrorString(e != cudaSuccess ? e : e2)); return 1; }
        printf("  cta_group::%d M=%3d N=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA (MAC floor per SM %5.1f)\n", cg, m, n,
               (double)h[0] / iters, (double)h[1] / iters, (m / cg) * (double)n * 16 / 4096);
      }
    }
  return 0;
}

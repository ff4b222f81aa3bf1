// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps and the .xN width.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= v[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= v[i];
  return s;
}

// each warp issues `iters` loads of 32 lanes x X columns x 4 B; a wait::ld every `batch` loads
template <int X>
__global__ void bench(int iters, int batch, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += batch) {
    for (int j = 0; j < batch; ++j) acc ^= ld<X>(base + (((i + j) * X + (warp >> 2) * 64) & (511 - X + 1) & ~(X - 1)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

int main() {
  long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int x : {16, 32}) {
    for (int warps : {1, 4, 8, 16}) {
      for (int batch : {1, 4}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (x == 16) bench<16><<<148, warps * 32>>>(iters, batch, d_cyc, d_sink);
          else bench<32><<<148, warps * 32>>>(iters, batch, d_cyc, d_sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long cyc[148]; cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        double bytes = double(warps) * iters * 32.0 * x * 4.0;
        printf("x%-2d warps %2d wait-every %d : %8lld cycles  %.1f B/clk/SM  (%.1f B/clk/warp)\n", x, warps, batch, cyc[0],
               bytes / cyc[0], bytes / cyc[0] / warps);
      }
    }
  }
  return 0;
}

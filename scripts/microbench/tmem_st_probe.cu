// Probe: register <-> (lane, column) mapping of tcgen05.st.16x128b, read back with tcgen05.ld.32x32b.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(&base_s));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s + (static_cast<uint32_t>(warp * 32) << 16);
  for (int half = 0; half < 2; ++half) {
    uint32_t r[4];
    for (int i = 0; i < 4; ++i) r[i] = (half << 12) | (lane << 4) | i;      // value = half, lane, register index
    const uint32_t addr = base + (static_cast<uint32_t>(half * 16) << 16);
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 8; ++c) out[(warp * 32 + lane) * 8 + c] = v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "r"(32) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 8 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[128 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int L = 0; L < 128; ++L)
    for (int c = 0; c < 8; ++c) {
      // hypothesis: TMEM (lane L, col c) was written by half = (L % 32) / 16, thread lane = 4 * (L % 8) + c % 4, register 2 * (c / 4) + ((L % 16) / 8)
      const uint32_t want = (((L % 32) / 16) << 12) | ((4 * (L % 8) + (c % 4)) << 4) | (2 * (c / 4) + ((L % 16) / 8));
      if (h[L * 8 + c] != want) ++bad;
    }
  printf("hypothesis mismatches: %d of 1024\n", bad);
  for (int L = 32; L < 48; L += 1) {
    printf("tmem lane %2d:", L);
    for (int c = 0; c < 8; ++c) { uint32_t g = h[L * 8 + c]; printf("  (h%u,t%u,r%u)", g >> 12, (g >> 4) & 31, g & 15); }
    printf("\n");
  }
  return 0;
}

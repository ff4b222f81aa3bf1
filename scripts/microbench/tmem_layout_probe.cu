// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b, against data written with tcgen05.st.32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_layout_probe tmem_layout_probe.cu
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
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = (warp * 32 + lane) * 256 + c;        // value = tmem_lane * 256 + column
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t addr = base + (static_cast<uint32_t>(half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "r"(32) : "memory");
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int warp = 0; warp < 4; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 8; ++i) {
          const uint32_t got = h[((warp * 2 + half) * 32 + lane) * 8 + i];
          // hypothesis: reg i -> block i / 4 (8 columns each), row = lane / 4 + 8 * ((i % 4) / 2), col = 8 * (i / 4) + 2 * (lane % 4) + (i % 2)
          const int row = warp * 32 + half * 16 + lane / 4 + 8 * ((i % 4) / 2);
          const int col = 8 * (i / 4) + 2 * (lane % 4) + (i % 2);
          if (got != static_cast<uint32_t>(row * 256 + col)) ++bad;
        }
  printf("hypothesis mismatches: %d of %d\n", bad, 4 * 2 * 32 * 8);
  for (int lane = 0; lane < 32; lane += 1) {
    if (lane > 5 && lane < 28) continue;
    printf("warp1 half0 lane %2d:", lane);
    for (int i = 0; i < 8; ++i) { uint32_t g = h[((1 * 2 + 0) * 32 + lane) * 8 + i]; printf("  (L%u,c%u)", g >> 8, g & 255); }
    printf("\n");
  }
  return 0;
}

// Microbenchmark: TMEM -> register read-back rate (tcgen05.ld.32x32b.x16) on B200.
// Why: a tcgen05 version of the pair scan would leave S[64][64] of every pileup entry in TMEM (16 KB) and
// must read it back to multiply it into the running products; this rate bounds that design.
// One CTA per SM allocates all 512 TMEM columns; warp w reads the 32 lanes of its SM sub-partition (w % 4), 16 columns
// per instruction, 4 instructions in flight before tcgen05.wait::ld.  Output: bytes per clock and SM for 4 / 8 / 16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__global__ void __launch_bounds__(512) k(uint32_t* out, int iters) {
  __shared__ uint32_t base_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&base_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = base_slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (uint32_t c = 0; c < 512; c += 64) {
      uint32_t r[4][16];
#pragma unroll
      for (int u = 0; u < 4; ++u) tmem_ld16(base + c + 16 * u, r[u]);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[u][j];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base_slot));
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  uint32_t* out;
  cudaMalloc(&out, sizeof(uint32_t) * sms * 512);
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<<<sms, warps * 32>>>(out, 10);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * khz * 1e3;
    const double bytes = (double)warps * 32 * 512 * 4 * iters;   // per SM
    printf("{\"bench\": \"tcgen05.ld.32x32b.x16\", \"warps_per_sm\": %d, \"ms\": %.3f, \"bytes_per_clk_sm\": %.1f, "
           "\"clk_per_64x64_f32_tile\": %.1f, \"err\": \"%s\"}\n",
           warps, ms, bytes / clk, 16384.0 / (bytes / clk), cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
  return 0;
}

// Microbenchmark for the round-2 lever "3xTF32 pair tiles" (DESIGN.md 6): issue rate of the legacy tensor path
// mma.sync.m16n8k8.tf32 on B200 and of the consumer loop it would replace.
//   S[j][k] = sum_l a_j[l] w_k[l]  (l < 3) per pileup entry and hypothesis.  With a = ah + al, w = wh + wl (tf32 pieces)
//   S ~= ah.wh + ah.wl + al.wh : a K = 9 product, zero-padded to K = 16 = two m16n8k8 per 16 x 8 tile of hypotheses, then
//   4 FMUL per thread and tile into the running products.  Today's fp32 path: FMUL + 2 FFMA + FMUL per hypothesis.
// Modes: 0 FFMA only (4 per tile slot: the FP32 ceiling), 1 mma k8 accumulate in place, 2 the candidate consumer step
// (2 mma with C = 0 then 4 FMUL; one fragment word changes per entry so that nothing is loop invariant), 3 the same
// with the A / B fragments re-loaded from shared memory every entry.  Today's consumer arithmetic (FMUL + 2 FFMA + FMUL
// per hypothesis on 8 x 8 register tiles) is tools/f32_bench.cu: 97 lane-ops / clk / SM = 24 hypothesis-entries.
// Output: per mode and warps per SM, hypothesis-entries per clock and SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_bench tf32_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b, const float* c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

constexpr int NT = 5;   // 16 x 8 tiles per warp (20 tiles of the 64 x 64 upper triangle over 4 warps)

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  __shared__ uint32_t sm[16][NT + 1][8][16];   // (two lanes share a word: bandwidth, not content, matters)   // per warp: fragments of one entry (mode 3)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[NT][4];
  for (int i = 0; i < NT; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 1.0f + 1e-3f * (i + j + lane);
  uint32_t a[2][4], b[NT][2][2];
  for (int s = 0; s < 2; ++s) for (int j = 0; j < 4; ++j) a[s][j] = __float_as_uint(seed * (1.0f + 0.01f * (s + j)));
  for (int i = 0; i < NT; ++i) for (int s = 0; s < 2; ++s) for (int j = 0; j < 2; ++j) b[i][s][j] = __float_as_uint(seed * (1.0f + 0.001f * (i + s + j)));
  for (int i = 0; i <= NT; ++i) for (int j = 0; j < 8; ++j) sm[warp][i][j][lane & 15] = __float_as_uint(seed + 0.001f * (i + j));
  __syncwarp();
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) { a[0][0] += 8192u; a[1][1] += 8192u; }   // next entry: other operands
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[0][j] = sm[warp][NT][j][lane & 15]; a[1][j] = sm[warp][NT][4 + j][lane & 15]; }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (MODE == 0) {
        acc[i][0] = fmaf(acc[i][0], seed, 1e-9f); acc[i][1] = fmaf(acc[i][1], seed, 1e-9f);
        acc[i][2] = fmaf(acc[i][2], seed, 1e-9f); acc[i][3] = fmaf(acc[i][3], seed, 1e-9f);
      } else if (MODE == 1) {
        mma_tf32(acc[i], a[0], b[i][0], acc[i]);
      } else {
        if (MODE == 3) {
#pragma unroll
          for (int j = 0; j < 2; ++j) { b[i][0][j] = sm[warp][i][j][lane & 15]; b[i][1][j] = sm[warp][i][2 + j][lane & 15]; }
        }
        float s[4];
        mma_tf32(s, a[0], b[i][0], zero);
        mma_tf32(s, a[1], b[i][1], s);
        acc[i][0] *= s[0]; acc[i][1] *= s[1]; acc[i][2] *= s[2]; acc[i][3] *= s[3];
      }
    }
  }
  float r = 0;
  for (int i = 0; i < NT; ++i) for (int j = 0; j < 4; ++j) r += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double hyp_per_iter_warp) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 512);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<MODE><<<sms, warps * 32>>>(out, 100, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32>>>(out, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("{\"mode\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"hyp_entries_per_clk_sm\": %.1f, \"clk_per_iter_warp\": %.1f, \"err\": \"%s\"}\n",
           name, warps, ms, hyp_per_iter_warp * warps * iters / clk, clk / iters, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
}

int main() {
  run<0>("ffma x4 per tile slot (fp32 ceiling, 128 lane-ops per instr quartet)", NT * 128.0);
  run<1>("mma.m16n8k8.tf32 accumulate in place", NT * 128.0);
  run<2>("2 mma (C=0) + 4 fmul per 16x8 tile, fragments in registers", NT * 128.0);
  run<3>("2 mma (C=0) + 4 fmul per 16x8 tile, fragments from shared memory", NT * 128.0);
  return 0;
}

// Microbenchmark: FP64 issue rates on B200 (DFMA, DMMA m8n8k4 / m16n8k4 / m16n8k8, and DMMA + DMUL mixed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ void mma1684(double* d, double a0, double a1, double b, const double* c) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
               : "d"(a0), "d"(a1), "d"(b), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b, const double* c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}

template <int MODE>
__global__ void k(double* out, int iters, double seed) {
  double acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = seed + i + j + threadIdx.x;
  double a0 = seed * 1.0001, a1 = seed * 0.9999, b = seed * 1.00001;
  double zero[4] = {0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {   // DFMA: 4 per i
        acc[i][0] = fma(acc[i][0], a0, b); acc[i][1] = fma(acc[i][1], a0, b);
        acc[i][2] = fma(acc[i][2], a0, b); acc[i][3] = fma(acc[i][3], a0, b);
      } else if (MODE == 1) {   // m8n8k4 accumulate in place
        mma884(acc[i][0], acc[i][1], a0, b, acc[i][0], acc[i][1]);
        mma884(acc[i][2], acc[i][3], a1, b, acc[i][2], acc[i][3]);
      } else if (MODE == 2) {   // m16n8k4 accumulate in place
        mma1684(acc[i], a0, a1, b, acc[i]);
      } else if (MODE == 3) {   // m16n8k4 with C = 0, then 4 DMUL into the product accumulators (our use)
        double s[4];
        mma1684(s, a0 + i, a1, b, zero);
        acc[i][0] *= s[0]; acc[i][1] *= s[1]; acc[i][2] *= s[2]; acc[i][3] *= s[3];
      } else if (MODE == 4) {   // m16n8k8
        double a[4] = {a0, a1, a0, a1}, bb[2] = {b, b};
        mma1688(acc[i], a, bb, acc[i]);
      } else if (MODE == 5) {   // DMUL only: 4 per i
        acc[i][0] *= a0; acc[i][1] *= a0; acc[i][2] *= a0; acc[i][3] *= a0;
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flop_per_inst_warp, int inst_per_i) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 512);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps * 32>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const double insts = (double)iters * 8 * inst_per_i * warps;   // per SM
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s warps/SM=%2d  %.3f ms  inst/clk/SM=%.3f  (cycles per warp-inst per SM %.2f)  TFLOP/s=%.1f\n", name, warps, ms,
           insts / cyc, cyc / insts, insts * sms * flop_per_inst_warp / (ms * 1e-3) / 1e12);
  }
  cudaFree(out);
}

int main() {
  run<0>("DFMA", 64, 4);
  run<5>("DMUL", 32, 4);
  run<1>("DMMA m8n8k4", 512, 2);
  run<2>("DMMA m16n8k4", 1024, 1);
  run<4>("DMMA m16n8k8", 2048, 1);
  run<3>("DMMA m16n8k4(C=0)+4 DMUL", 1024 + 4 * 32, 1);
  return 0;
}

// Microbenchmark for the fp32 pair-scan consumer design on B200:
//   (1) the 8 x 8 register tile inner loop (3 FFMA + 1 FMUL per hypothesis and entry, operands from shared
//       memory, exponent peel every 4 entries), scalar FFMA vs packed fma.rn.f32x2 (SASS FFMA2)
//   (2) cost of LDS.128 under several broadcast patterns (how many wavefronts a warp request needs)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32_bench f32_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// MODE 0: scalar, 1: packed along k (a duplicated with MOVs), 2: packed, no peel, 3: scalar, no peel
// LANE_STRIDE: byte distance between the operand blocks of consecutive lanes (0 = broadcast)
template <int MODE>
__global__ void __launch_bounds__(512) tile_kernel(float* out, int iters, int lane_stride, int* scratch) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = make_float4(0.99f, 0.999f, 0.9999f, 1.0f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float4* pa = sm + (lane * lane_stride) / 16;
  const float4* pw = pa + 6;
  float acc[8][8];
  int* es = scratch + (blockIdx.x * blockDim.x + threadIdx.x);
  const int estride = gridDim.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j] = 1.0f; }
  for (int it = 0; it < iters; ++it) {
    const int off = (it & 7) * 12;   // 8 "entries" of 192 B per lane block
    float a[8][3], w[8][3];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const float4 x = pa[off + q];
      const float4 y = pw[off + q];
      const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        a[(q * 4 + c) / 3][(q * 4 + c) % 3] = xs[c];
        w[(q * 4 + c) / 3][(q * 4 + c) % 3] = ys[c];
      }
    }
    if (MODE == 0 || MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = fmaf(a[i][2], w[j][2], fmaf(a[i][1], w[j][1], a[i][0] * w[j][0]));
          acc[i][j] *= s;
        }
    } else {
      // operands as they would sit in a component-major pair layout: here built with MOV/PRMT-free packs
      uint64_t w2[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int l = 0; l < 3; ++l) w2[j][l] = pack2(w[2 * j][l], w[2 * j + 1][l]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t a0 = pack2(a[i][0], a[i][0]), a1 = pack2(a[i][1], a[i][1]), a2 = pack2(a[i][2], a[i][2]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint64_t s = mul2(a0, w2[j][0]);
          s = fma2(a1, w2[j][1], s);
          s = fma2(a2, w2[j][2], s);
          uint64_t p = pack2(acc[i][2 * j], acc[i][2 * j + 1]);
          p = mul2(p, s);
          acc[i][2 * j] = __uint_as_float((uint32_t)p);
          acc[i][2 * j + 1] = __uint_as_float((uint32_t)(p >> 32));
        }
      }
    }
    if ((MODE == 0 || MODE == 1) && (it & 3) == 3) {
      float mn = 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; j += 2) mn = fminf(mn, fminf(acc[i][j], acc[i][j + 1]));
      if (__any_sync(0xffffffffu, mn < 9.0e-13f)) {   // 2^-40: rare path, exponents live in global scratch
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t b = __float_as_uint(acc[i][j]);
            es[(i * 8 + j) * estride] += (int)(b >> 23);
            acc[i][j] = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
          }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j] + (float)es[(i * 8 + j) * estride];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 patterns: PAT 0 all lanes same address; 1: (lane % 8) * 16; 2: (lane / 4) * 16; 3: lane * 16;
// 4: (lane % 8) * 48 (stride-48 quarter warps, all quarters identical); 5: lane * 48
template <int PAT>
__global__ void lds_kernel(float* out, int iters) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = PAT == 0 ? 0 : PAT == 1 ? (lane % 8) : PAT == 2 ? (lane / 4) : PAT == 3 ? lane : PAT == 4 ? (lane % 8) * 3 : lane * 3;
  float4 acc = make_float4(0, 0, 0, 0);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + idx * 16;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(base + (uint32_t)(((it & 3) * 16 + u) * 1536)));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

static double cycles(float ms) {
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  return ms * 1e-3 * clk_khz * 1e3;
}

template <int MODE>
void run_tile(const char* name, int lane_stride) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  int* scratch;
  cudaMalloc(&out, sizeof(float) * sms * 512);
  cudaMalloc(&scratch, sizeof(int) * sms * 512 * 64);
  cudaMemset(scratch, 0, sizeof(int) * sms * 512 * 64);
  cudaFuncSetAttribute(tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    tile_kernel<MODE><<<sms, warps * 32, 196608>>>(out, 100, lane_stride, scratch);
    cudaEventRecord(e0);
    tile_kernel<MODE><<<sms, warps * 32, 196608>>>(out, iters, lane_stride, scratch);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = cycles(ms);
    // per SM: warps * iters tile-entries of 64 hypotheses x 32 lanes
    const double hyp = (double)warps * iters * 64 * 32;
    printf("%-34s stride=%4d warps/SM=%2d  %.3f ms  hyp-entries/clk/SM=%.2f  (fp32 lane-ops/clk/SM %.1f)  err=%s\n", name,
           lane_stride, warps, ms, hyp / cyc, 4 * hyp / cyc, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
}

template <int PAT>
void run_lds(const char* name) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 512);
  cudaFuncSetAttribute(lds_kernel<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  const int warps = 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  lds_kernel<PAT><<<sms, warps * 32, 131072>>>(out, 100);
  cudaEventRecord(e0);
  lds_kernel<PAT><<<sms, warps * 32, 131072>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = cycles(ms);
  printf("LDS.128 %-40s cycles per warp request per SM = %.2f  err=%s\n", name, cyc / ((double)iters * 16 * warps),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run_lds<0>("all lanes one address");
  run_lds<1>("(lane%8)*16: 8 addrs, quarters identical");
  run_lds<2>("(lane/4)*16: 8 addrs, 2 per quarter");
  run_lds<3>("lane*16: contiguous 512 B");
  run_lds<4>("(lane%8)*48: 8 addrs stride 48");
  run_lds<5>("lane*48: 32 addrs stride 48");
  run_tile<3>("scalar FFMA, no peel", 0);
  run_tile<0>("scalar FFMA + peel/4", 0);
  run_tile<2>("packed FFMA2, no peel", 0);
  run_tile<1>("packed FFMA2 + peel/4", 0);
  run_tile<0>("scalar FFMA + peel/4", 768);
  run_tile<1>("packed FFMA2 + peel/4", 768);
  return 0;
}

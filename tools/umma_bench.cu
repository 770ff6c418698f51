// Microbenchmark + correctness probe: one tcgen05.mma (kind::tf32, M = 64, N = 64, K = 8; operands in shared memory in
// the canonical K-major no-swizzle layout, accumulator in TMEM) -- the building block of the round-2 pair scan
// (DESIGN.md 6): S[j][k] = sum_l a_j[l] w_k[l] of one pileup entry per instruction.
//   1. numerics: D = A x B^T against the host, through tcgen05.ld (which TMEM lanes hold which rows is reported)
//   2. rate: clocks per MMA for back-to-back dependent (same accumulator) and independent (rotating TMEM columns) issues
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 64, N = 64, K = 8;
constexpr uint32_t LBO = 128, SBO = 256;   // bytes: next core matrix along K / next 8-row group

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) | ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) |
         (1ull << 46);   // version 1 (sm_100), layout type 0 = no swizzle
}

// canonical K-major no-swizzle layout: core matrix = 8 rows x 16 bytes (4 tf32), rows 16 bytes apart
__device__ __forceinline__ uint32_t canon_off(int row, int k) {
  return (uint32_t)((row >> 3) * SBO + (k >> 2) * LBO + (row & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

template <int F16>
__global__ void __launch_bounds__(128) k(const float* a_in, const float* b_in, float* d_out, long long* clk_out, int nmma,
                                         int* status, int m_rate, int n_rate) {
  constexpr int f16 = F16;
  __shared__ __align__(1024) unsigned char sA[4096];
  __shared__ __align__(1024) unsigned char sB[8192];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) *(float*)(sA + canon_off(i / K, i % K)) = a_in[i];
  for (int i = tid; i < N * K; i += 128) *(float*)(sB + canon_off(i / K, i % K)) = b_in[i];
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
  uint32_t phase = 0;
  bool ok = true;
  // 1. one MMA, read back (and its latency: issue -> commit -> mbarrier phase seen by the waiting threads)
  const long long t_one = clock64();
  if (tid == 0) {
    mma_tf32(tmem, da, db, idesc, 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
  }
  ok = mbar_wait_bounded(smem_u32(&bar), phase);
  phase ^= 1;
  if (tid == 0) clk_out[blockIdx.x * 2 + 1] = clock64() - t_one;   // (overwritten by the rate part unless nmma == 0)
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(ta));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 16; ++j) d_out[((size_t)(32 * warp + lane)) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  __syncthreads();
  // 2. rates: mode 0 = dependent (accumulate into the same tile), mode 1 = independent (8 tiles of 64 columns)
  // (the rate part takes any legal shape and kind: the operand bytes beyond the initialised 2 KB are whatever shared
  // memory holds)
  const uint32_t idesc_r = (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(n_rate >> 3) << 17) |
                           ((uint32_t)(m_rate >> 4) << 24);
  const uint32_t ntile = 512u / (uint32_t)n_rate;
  for (int mode = 0; mode < 2 && ok && nmma > 0; ++mode) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      const uint32_t tmask = mode ? ntile - 1 : 0u, acc = mode ? 0u : 1u;
#pragma unroll 8
      for (int i = 0; i < nmma; ++i) {
        const uint32_t td = tmem + ((uint32_t)i & tmask) * (uint32_t)n_rate;
        if (f16) mma_f16(td, da, db, idesc_r, acc);
        else mma_tf32(td, da, db, idesc_r, acc);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
    }
    ok = mbar_wait_bounded(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0) clk_out[blockIdx.x * 2 + mode] = clock64() - t0;
    __syncthreads();
  }
  if (tid == 0 && !ok) atomicExch(status, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

static float to_tf32(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  memcpy(&x, &b, 4);
  return x;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::vector<float> a(M * K), b(N * K), d(128 * N, -1.f);
  for (int i = 0; i < M * K; ++i) a[i] = to_tf32(0.001f + 0.37f * ((i * 7919) % 101) / 101.f);
  for (int i = 0; i < N * K; ++i) b[i] = to_tf32(0.002f + 0.61f * ((i * 104729) % 97) / 97.f);
  float *da, *db, *dd;
  long long* dclk;
  int* dst;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMalloc(&dclk, sizeof(long long) * sms * 2); cudaMalloc(&dst, 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dd, d.data(), d.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0, 4);
  const int nmma = 4096;
  k<0><<<1, 128>>>(da, db, dd, dclk, nmma, dst, 64, 64);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  int st = 0;
  long long clk[2];
  cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
  // which TMEM lane holds row i?  candidates: lane = i (rows in lanes 0..63) / lane = 32 (i / 16) + i % 16
  double err[2] = {0, 0};
  for (int cand = 0; cand < 2; ++cand)
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double s = 0;
        for (int l = 0; l < K; ++l) s += (double)a[i * K + l] * b[j * K + l];
        const int lane = cand == 0 ? i : 32 * (i / 16) + i % 16;
        err[cand] = fmax(err[cand], fabs(d[(size_t)lane * N + j] - s) / s);
      }
  printf("{\"bench\": \"tcgen05.mma kind::tf32 M64 N64 K8 (SS, no swizzle)\", \"timeout\": %d, \"max_rel_err_rows_in_lanes_0_63\": %.3g, "
         "\"max_rel_err_rows_16_per_subpartition\": %.3g, \"clk_per_mma_dependent\": %.1f, \"clk_per_mma_independent\": %.1f}\n",
         st, err[0], err[1], (double)clk[0] / nmma, (double)clk[1] / nmma);
  {
    k<0><<<1, 128>>>(da, db, dd, dclk, 0, dst, 64, 64);
    cudaDeviceSynchronize();
    cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
    printf("{\"bench\": \"latency of one tcgen05.mma M64 N64 K8: issue -> commit -> mbarrier observed\", \"clk\": %lld}\n", clk[1]);
  }
  // all SMs at once, shapes and kinds: does the rate hold chip-wide, what do M = 128, wider N and kind::f16 cost?
  const int shapes[][3] = {{64, 64, 0}, {128, 64, 0}, {64, 128, 0}, {128, 128, 0}, {128, 256, 0},
                           {64, 64, 1}, {128, 64, 1}, {128, 128, 1}, {128, 256, 1}};
  for (auto& sh : shapes) {
    if (sh[2]) k<1><<<sms, 128>>>(da, db, dd, dclk, nmma, dst, sh[0], sh[1]);
    else k<0><<<sms, 128>>>(da, db, dd, dclk, nmma, dst, sh[0], sh[1]);
    e = cudaDeviceSynchronize();
    std::vector<long long> all(sms * 2);
    cudaMemcpy(all.data(), dclk, sizeof(long long) * sms * 2, cudaMemcpyDeviceToHost);
    long long mx0 = 0, mx1 = 0;
    for (int i = 0; i < sms; ++i) { mx0 = std::max(mx0, all[i * 2]); mx1 = std::max(mx1, all[i * 2 + 1]); }
    printf("{\"bench\": \"all %d SMs\", \"kind\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"err\": \"%s\", \"clk_per_mma_dependent\": %.1f, "
           "\"clk_per_mma_independent\": %.1f, \"mac_per_clk_sm\": %.0f}\n", sms, sh[2] ? "f16" : "tf32", sh[0], sh[1], sh[2] ? 16 : 8,
           cudaGetErrorString(e), (double)mx0 / nmma, (double)mx1 / nmma, (double)sh[0] * sh[1] * (sh[2] ? 16 : 8) * nmma / mx1);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}

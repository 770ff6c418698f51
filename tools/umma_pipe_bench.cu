// Pipeline microbenchmark for the round-2 pair scan: per pileup entry one thread issues NK
// tcgen05.mma (kind::tf32, N = 64, K = 8 each; M = 64 or 128) into one of 8 TMEM tiles, four consumer warps (one per SM
// sub-partition, or CW per sub-partition taking the entries in turn) read the tile back with tcgen05.ld and multiply it
// into 64 running products per thread; mbarriers hand
// the tiles back and forth (tcgen05.commit -> full, consumer arrive -> empty).  Operands stay fixed in shared memory
// (the transform side is not part of this measurement).  Output: clocks per entry-tile for NK = 1, 2 and M = 64, 128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_pipe_bench umma_pipe_bench.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

constexpr int NT = 8;                       // TMEM tiles of 64 columns
constexpr uint32_t LBO = 128, SBO = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) | ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
               "r"(idesc), "r"(accumulate), "r"(0u));
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 24); ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ bool mbar_poll(uint32_t bar, uint32_t parity) {   // non-blocking test_wait spin (issuer thread)
  for (int i = 0; i < (1 << 24); ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

template <int NK, int CW, int NOWAIT, int CONSUME>
__global__ void __launch_bounds__(32 * (4 * CW + 1)) k(float* out, long long* clk_out, int nent, int m, int* status) {
  __shared__ __align__(1024) unsigned char sA[2][4096];
  __shared__ __align__(1024) unsigned char sB[2][2048];
  __shared__ __align__(8) unsigned long long full[NT], empty[NT];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 4096 / 4; i += blockDim.x) ((float*)sA)[i] = 1.0f + 1e-4f * (i % 97);
  for (int i = tid; i < 2 * 2048 / 4; i += blockDim.x) ((float*)sB)[i] = 0.125f + 1e-4f * (i % 89);
  if (tid == 0)
    for (int t = 0; t < NT; ++t) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full[t])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;\n" ::"r"(smem_u32(&empty[t])));
    }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  bool ok = true;
  const long long t0 = clock64();
  if (warp == 4 * CW) {                  // MMA issuer
    if (lane == 0) {
      uint64_t da[2] = {make_desc(smem_u32(sA[0])), make_desc(smem_u32(sA[1]))};
      uint64_t db[2] = {make_desc(smem_u32(sB[0])), make_desc(smem_u32(sB[1]))};
      for (int e = 0; e < nent && ok; ++e) {
        const int t = e % NT;
        const uint32_t ph = (uint32_t)(e / NT) & 1u;
        if (e >= NT && !NOWAIT) ok = mbar_wait(smem_u32(&empty[t]), ph ^ 1u);   // NOWAIT: tiles are overwritten unseen
        asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) mma_tf32(tmem + (uint32_t)t * 64u, da[kk], db[kk], idesc, kk ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&full[t])));
      }
    }
  } else {                               // consumers: warp w owns TMEM lanes 32 w .. 32 w + 31
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 1.0f;
    for (int e = warp >> 2; e < nent && ok; e += CW) {   // CW warps per sub-partition take the entries in turn
      const int t = e % NT;
      const uint32_t ph = (uint32_t)(e / NT) & 1u;
      ok = mbar_wait(smem_u32(&full[t]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      if (!CONSUME) {                    // barrier traffic only
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[t]));
        continue;
      }
      const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)t * 64u;
      uint32_t r[64];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[16 * u + 0]), "=r"(r[16 * u + 1]), "=r"(r[16 * u + 2]), "=r"(r[16 * u + 3]), "=r"(r[16 * u + 4]),
                       "=r"(r[16 * u + 5]), "=r"(r[16 * u + 6]), "=r"(r[16 * u + 7]), "=r"(r[16 * u + 8]), "=r"(r[16 * u + 9]),
                       "=r"(r[16 * u + 10]), "=r"(r[16 * u + 11]), "=r"(r[16 * u + 12]), "=r"(r[16 * u + 13]),
                       "=r"(r[16 * u + 14]), "=r"(r[16 * u + 15])
                     : "r"(ta + 16u * u));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[t]));     // the tile may be overwritten while the products are formed
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] *= __uint_as_float(r[j]) * 0.25f;   // (scaled so that the products stay finite)
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) s += acc[j];
    out[blockIdx.x * 512 + tid] = s;
  }
  if (!ok) atomicExch(status, 1);
  __syncthreads();
  if (tid == 0) clk_out[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

// Staged variant: the issuer hands over STAGES of 4 entries (4 TMEM tiles = 256 columns, two stages in TMEM): one wait
// for a free stage, 4 x NK MMAs, ONE commit; consumer warp group cw (4 warps, one per sub-partition) reads tile cw of
// every stage.  The barrier round trips are paid per stage, not per entry.
template <int NK>
__global__ void __launch_bounds__(32 * 17) k2(float* out, long long* clk_out, int nstage, int m, int* status) {
  __shared__ __align__(1024) unsigned char sA[2][4096];
  __shared__ __align__(1024) unsigned char sB[2][2048];
  __shared__ __align__(8) unsigned long long full[2], empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 4096 / 4; i += blockDim.x) ((float*)sA)[i] = 1.0f + 1e-4f * (i % 97);
  for (int i = tid; i < 2 * 2048 / 4; i += blockDim.x) ((float*)sB)[i] = 0.125f + 1e-4f * (i % 89);
  if (tid == 0)
    for (int g = 0; g < 2; ++g) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full[g])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;\n" ::"r"(smem_u32(&empty[g])));
    }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  bool ok = true;
  const long long t0 = clock64();
  if (warp == 16) {
    if (lane == 0) {
      uint64_t da[2] = {make_desc(smem_u32(sA[0])), make_desc(smem_u32(sA[1]))};
      uint64_t db[2] = {make_desc(smem_u32(sB[0])), make_desc(smem_u32(sB[1]))};
      for (int s = 0; s < nstage && ok; ++s) {
        const int g = s & 1;
        const uint32_t ph = (uint32_t)(s >> 1) & 1u;
        if (s >= 2) ok = mbar_wait(smem_u32(&empty[g]), ph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int kk = 0; kk < NK; ++kk) mma_tf32(tmem + (uint32_t)(g * 4 + i) * 64u, da[kk], db[kk], idesc, kk ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&full[g])));
      }
    }
  } else {
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 1.0f;
    for (int s = 0; s < nstage && ok; ++s) {
      const int g = s & 1;
      const uint32_t ph = (uint32_t)(s >> 1) & 1u;
      ok = mbar_wait(smem_u32(&full[g]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(g * 4 + (warp >> 2)) * 64u;
      uint32_t r[64];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[16 * u + 0]), "=r"(r[16 * u + 1]), "=r"(r[16 * u + 2]), "=r"(r[16 * u + 3]), "=r"(r[16 * u + 4]),
                       "=r"(r[16 * u + 5]), "=r"(r[16 * u + 6]), "=r"(r[16 * u + 7]), "=r"(r[16 * u + 8]), "=r"(r[16 * u + 9]),
                       "=r"(r[16 * u + 10]), "=r"(r[16 * u + 11]), "=r"(r[16 * u + 12]), "=r"(r[16 * u + 13]),
                       "=r"(r[16 * u + 14]), "=r"(r[16 * u + 15])
                     : "r"(ta + 16u * u));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[g]));
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] *= __uint_as_float(r[j]) * 0.25f;
    }
    float sum = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) sum += acc[j];
    out[blockIdx.x * 512 + tid] = sum;
  }
  if (!ok) atomicExch(status, 1);
  __syncthreads();
  if (tid == 0) clk_out[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

// (CHUNKED = 1: the consumer overlaps its loads with its multiplies, 16 columns at a time; CHUNKED = 2: in addition the
// issuer polls with test_wait instead of try_wait.  Neither changes the result.)
// Staged variant with both lane halves: a stage is 8 entries -- 4 column tiles x 2 M = 64 accumulators that share the
// columns (TMEM lanes 0-15 and 16-31 of every sub-partition), so one tcgen05.ld delivers the rows of two entries and all
// 32 lanes of a consumer warp work.  Otherwise as k2: one wait
// for a free stage, 4 x NK MMAs, ONE commit; consumer warp group cw (4 warps, one per sub-partition) reads tile cw of
// every stage.  The barrier round trips are paid per stage, not per entry.
template <int NK, int CHUNKED>
__global__ void __launch_bounds__(32 * 17) k3(float* out, long long* clk_out, int nstage, int m, int* status) {
  __shared__ __align__(1024) unsigned char sA[2][4096];
  __shared__ __align__(1024) unsigned char sB[2][2048];
  __shared__ __align__(8) unsigned long long full[2], empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 4096 / 4; i += blockDim.x) ((float*)sA)[i] = 1.0f + 1e-4f * (i % 97);
  for (int i = tid; i < 2 * 2048 / 4; i += blockDim.x) ((float*)sB)[i] = 0.125f + 1e-4f * (i % 89);
  if (tid == 0)
    for (int g = 0; g < 2; ++g) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full[g])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;\n" ::"r"(smem_u32(&empty[g])));
    }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  bool ok = true;
  const long long t0 = clock64();
  if (warp == 16) {
    if (lane == 0) {
      uint64_t da[2] = {make_desc(smem_u32(sA[0])), make_desc(smem_u32(sA[1]))};
      uint64_t db[2] = {make_desc(smem_u32(sB[0])), make_desc(smem_u32(sB[1]))};
      for (int s = 0; s < nstage && ok; ++s) {
        const int g = s & 1;
        const uint32_t ph = (uint32_t)(s >> 1) & 1u;
        if (s >= 2) ok = CHUNKED == 2 ? mbar_poll(smem_u32(&empty[g]), ph ^ 1u) : mbar_wait(smem_u32(&empty[g]), ph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n");
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int kk = 0; kk < NK; ++kk)
            mma_tf32(tmem + ((uint32_t)(i & 1) << 20) + (uint32_t)(g * 4 + (i >> 1)) * 64u, da[kk], db[kk], idesc, kk ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&full[g])));
      }
    }
  } else {
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 1.0f;
    for (int s = 0; s < nstage && ok; ++s) {
      const int g = s & 1;
      const uint32_t ph = (uint32_t)(s >> 1) & 1u;
      ok = mbar_wait(smem_u32(&full[g]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(g * 4 + (warp >> 2)) * 64u;
      uint32_t r[64];
      if (CHUNKED) {
        // 16 columns at a time: the load of chunk u + 1 is in flight while chunk u is multiplied in
#define DMX_LD16(u)                                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" \
               : "=r"(r[16 * u + 0]), "=r"(r[16 * u + 1]), "=r"(r[16 * u + 2]), "=r"(r[16 * u + 3]), "=r"(r[16 * u + 4]),  \
                 "=r"(r[16 * u + 5]), "=r"(r[16 * u + 6]), "=r"(r[16 * u + 7]), "=r"(r[16 * u + 8]), "=r"(r[16 * u + 9]),  \
                 "=r"(r[16 * u + 10]), "=r"(r[16 * u + 11]), "=r"(r[16 * u + 12]), "=r"(r[16 * u + 13]),                  \
                 "=r"(r[16 * u + 14]), "=r"(r[16 * u + 15])                                                               \
               : "r"(ta + 16u * u))
        DMX_LD16(0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (u == 0) DMX_LD16(1);
          if (u == 1) DMX_LD16(2);
          if (u == 2) DMX_LD16(3);
          if (u == 3) {
            asm volatile("tcgen05.fence::before_thread_sync;\n");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty[g]));
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[16 * u + j] *= __uint_as_float(r[16 * u + j]) * 0.25f;
        }
#undef DMX_LD16
        continue;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[16 * u + 0]), "=r"(r[16 * u + 1]), "=r"(r[16 * u + 2]), "=r"(r[16 * u + 3]), "=r"(r[16 * u + 4]),
                       "=r"(r[16 * u + 5]), "=r"(r[16 * u + 6]), "=r"(r[16 * u + 7]), "=r"(r[16 * u + 8]), "=r"(r[16 * u + 9]),
                       "=r"(r[16 * u + 10]), "=r"(r[16 * u + 11]), "=r"(r[16 * u + 12]), "=r"(r[16 * u + 13]),
                       "=r"(r[16 * u + 14]), "=r"(r[16 * u + 15])
                     : "r"(ta + 16u * u));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[g]));
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] *= __uint_as_float(r[j]) * 0.25f;
    }
    float sum = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) sum += acc[j];
    out[blockIdx.x * 512 + tid] = sum;
  }
  if (!ok) atomicExch(status, 1);
  __syncthreads();
  if (tid == 0) clk_out[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

template <int NK>
void run2(int sms, int m) {
  float* out; long long* clk; int* st;
  cudaMalloc(&out, sizeof(float) * sms * 512); cudaMalloc(&clk, sizeof(long long) * sms); cudaMalloc(&st, 4);
  cudaMemset(st, 0, 4);
  const int nstage = 5000;
  k2<NK><<<sms, 32 * 17>>>(out, clk, nstage, m, st);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> c(sms);
  int status = 0;
  cudaMemcpy(c.data(), clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : c) mx = std::max(mx, v);
  printf("{\"bench\": \"staged pipeline: 4 entries per commit, 16 consumer warps\", \"mma_per_entry\": %d, \"M\": %d, \"N\": 64, \"err\": \"%s\", "
         "\"timeout\": %d, \"clk_per_entry_tile\": %.1f}\n", NK, m, cudaGetErrorString(e), status, (double)mx / (nstage * 4.0));
  cudaFree(out); cudaFree(clk); cudaFree(st);
}

template <int NK, int CHUNKED = 0>
void run3(int sms, int m) {
  float* out; long long* clk; int* st;
  cudaMalloc(&out, sizeof(float) * sms * 512); cudaMalloc(&clk, sizeof(long long) * sms); cudaMalloc(&st, 4);
  cudaMemset(st, 0, 4);
  const int nstage = 5000;
  k3<NK, CHUNKED><<<sms, 32 * 17>>>(out, clk, nstage, m, st);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> c(sms);
  int status = 0;
  cudaMemcpy(c.data(), clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : c) mx = std::max(mx, v);
  printf("{\"bench\": \"staged pipeline: 8 entries per commit (both lane halves), 16 consumer warps\", \"consumer_overlaps_ld_and_fmul\": %d, \"mma_per_entry\": %d, \"M\": %d, \"N\": 64, \"err\": \"%s\", "
         "\"timeout\": %d, \"clk_per_entry_tile\": %.1f}\n", CHUNKED, NK, m, cudaGetErrorString(e), status, (double)mx / (nstage * 8.0));
  cudaFree(out); cudaFree(clk); cudaFree(st);
}

template <int NK, int CW, int NOWAIT = 0, int CONSUME = 1>
void run(int sms, int m) {
  float* out; long long* clk; int* st;
  cudaMalloc(&out, sizeof(float) * sms * 512); cudaMalloc(&clk, sizeof(long long) * sms); cudaMalloc(&st, 4);
  cudaMemset(st, 0, 4);
  const int nent = 20000;
  k<NK, CW, NOWAIT, CONSUME><<<sms, 32 * (4 * CW + 1)>>>(out, clk, nent, m, st);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> c(sms);
  int status = 0;
  cudaMemcpy(c.data(), clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : c) mx = std::max(mx, v);
  printf("{\"bench\": \"mma -> tmem -> ld -> fmul pipeline\", \"mma_per_entry\": %d, \"consumer_warps_per_subpartition\": %d, \"issuer_waits_for_free_tile\": %d, \"consumers_read_and_multiply\": %d, \"M\": %d, \"N\": 64, \"err\": \"%s\", \"timeout\": %d, "
         "\"clk_per_entry_tile\": %.1f}\n", NK, CW, !NOWAIT, CONSUME, m, cudaGetErrorString(e), status, (double)mx / nent);
  cudaFree(out); cudaFree(clk); cudaFree(st);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run3<1, 2>(sms, 64);
  run3<2, 2>(sms, 64);
  run3<1, 1>(sms, 64);
  run3<2, 1>(sms, 64);
  run3<1>(sms, 64);
  run3<2>(sms, 64);
  run2<1>(sms, 64);
  run2<2>(sms, 64);
  run2<2>(sms, 128);
  run<1, 1>(sms, 64);
  run<1, 2>(sms, 64);
  run<1, 3>(sms, 64);
  run<1, 4>(sms, 64);
  run<2, 2>(sms, 64);
  run<2, 3>(sms, 64);
  run<2, 4>(sms, 64);
  run<2, 4>(sms, 128);
  run<1, 4, 1>(sms, 64);
  run<2, 4, 1>(sms, 64);
  run<1, 4, 1, 0>(sms, 64);
  run<2, 4, 1, 0>(sms, 64);
  run<1, 4, 0, 0>(sms, 64);
  run<2, 4, 0, 0>(sms, 64);
  return 0;
}

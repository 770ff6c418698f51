// dmx_field.cu -- VCF FORMAT field -> genotype-probability matrix on the device.
//
// What the reference does once per VCF record on one host thread (BCFFilteredReader::parse_posteriors,
// /root/reference/bcf_filtered_reader.cpp:360-454; parse_genotypes :186-242; parse_likelihoods :244-320) and then
// widens into scl.snps[i].gps (cmd_cram_demuxlet.cpp:227-231), here as one warp per SNP:
//
//   GT  allele counts acs / an over the non-missing alleles (:231-240; integer warp reduction), then per sample the
//       called-genotype row (1 - gt_error, gt_error / 2; :398-402) or, for a missing call, the Hardy-Weinberg row from
//       (acs + 1/2) / (an + 1) (:381-388)
//   GP  per-sample sum-normalisation in float (:424-435), the cohort sums -- a running FLOAT sum in sample order
//       seeded with (1/4, 1/2, 1/4) (:417-421), which one lane per genotype adds up sequentially because float
//       addition is not associative --, mean over nv + 1 (:437-438), blend with weight gt_error in double (:441-446)
//   PL  ten EM iterations (:271-310): per sample gp = {1,2,1} * acs[j] * acs[k] * 10^(-PL/10), normalised; the new
//       allele counts are running DOUBLE sums in sample order (newacs[j] += gp; newacs[k] += gp), again one lane per
//       allele; the phred table (PhredHelper.cpp:32) is the host libm's pow(), uploaded by the caller
//
// Every operation is the reference's, separately rounded (__fadd_rn / __dmul_rn ...: no contraction), so the matrix
// is bit-identical to the reference's floats (tests/test_field.py).  Rows are staged in shared memory: global loads
// and stores are fully coalesced 128-byte warp accesses.  Bound: HBM (GT 2 + 12, GP / PL 12 + 12 bytes per sample and
// SNP); the sequential sums are nv (GP) or 3 nv x 10 (PL) dependent adds per SNP and hide behind the other resident
// warps.
#include <cuda_runtime.h>
#include <stdint.h>
#include "dmx_internal.h"

namespace dmx {

namespace {

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

// coalesced store of a staged row (nv3 floats) -- plain 4-byte lanes: rows of nv * 12 bytes are not 16-byte aligned
// for every nv
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float* __restrict__ sh, int nv3, int lane) {
  for (int i = lane; i < nv3; i += 32) dst[i] = sh[i];
}

__global__ void __launch_bounds__(256) field_gt_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = p.nv, nv3 = nv * 3;
  float* sh = reinterpret_cast<float*>(smem) + (size_t)warp * nv3;
  const double ge = p.gt_error;
  const float f_hit = (float)__dsub_rn(1.0, ge), f_miss = (float)__ddiv_rn(ge, 2.0);   // :400
  for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < p.nsnps; s += (int64_t)gridDim.x * wpb) {
    const int8_t* al = static_cast<const int8_t*>(p.in) + s * nv * 2;
    int c0 = 0, c1 = 0, bad = 0;
    for (int i = lane; i < nv; i += 32) {
      const char2 a = reinterpret_cast<const char2*>(al)[i];
      c0 += (a.x == 0) + (a.y == 0);
      c1 += (a.x == 1) + (a.y == 1);
      bad |= (a.x > 1) | (a.y > 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (bad && lane == 0) atomicOr(p.err_flag, 1);
    // :385: (j == k ? 1.0 : 2.0) * (acs[j] + 1.0 / nalleles) / (an + 1.0) * (acs[k] + 1.0 / nalleles) / (an + 1.0)
    const double a0 = __dadd_rn((double)c0, 0.5), a1 = __dadd_rn((double)c1, 0.5), den = __dadd_rn((double)(c0 + c1), 1.0);
    const float m0 = (float)__ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(1.0, a0), den), a0), den);
    const float m1 = (float)__ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(2.0, a1), den), a0), den);
    const float m2 = (float)__ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(1.0, a1), den), a1), den);
    for (int i = lane; i < nv; i += 32) {
      const char2 a = reinterpret_cast<const char2*>(al)[i];
      float o0, o1, o2;
      if (a.x < 0 || a.y < 0) {
        o0 = m0; o1 = m1; o2 = m2;
      } else {
        const int g = a.x + a.y;   // bcf_alleles2gt on {0, 1} x {0, 1}
        o0 = g == 0 ? f_hit : f_miss;
        o1 = g == 1 ? f_hit : f_miss;
        o2 = g == 2 ? f_hit : f_miss;
      }
      sh[3 * i] = o0; sh[3 * i + 1] = o1; sh[3 * i + 2] = o2;
    }
    __syncwarp();
    store_row(p.out + s * p.out_stride, sh, nv3, lane);
    if (p.af && lane == 0) p.af[s] = __ddiv_rn(a1, den);   // get_af(1), bcf_filtered_reader.h:177-179
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) field_gp_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = p.nv, nv3 = nv * 3;
  float* sh = reinterpret_cast<float*>(smem) + (size_t)warp * nv3;
  const double ge = p.gt_error, w = __dsub_rn(1.0, ge);
  const float fden = (float)(int32_t)(nv + 1.0);   // :438
  for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < p.nsnps; s += (int64_t)gridDim.x * wpb) {
    const float* row = static_cast<const float*>(p.in) + s * nv3;
    for (int i = lane; i < nv3; i += 32) sh[i] = __ldcs(row + i);
    __syncwarp();
    for (int i = lane; i < nv; i += 32) {   // :424-431 (stride of 3 floats: no bank conflicts)
      const float a = sh[3 * i], b = sh[3 * i + 1], c = sh[3 * i + 2];
      const float sum = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, a), b), c);
      sh[3 * i] = __fdiv_rn(a, sum);
      sh[3 * i + 1] = __fdiv_rn(b, sum);
      sh[3 * i + 2] = __fdiv_rn(c, sum);
    }
    __syncwarp();
    float acc = lane == 1 ? 0.5f : 0.25f;    // :417-421: the pseudo sample
    if (lane < 3) {
      const float* col = sh + lane;
      int i = 0;
      for (; i + 8 <= nv; i += 8) {          // loads batched ahead of the dependent add chain
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[3 * (i + u)];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
      }
      for (; i < nv; ++i) acc = __fadd_rn(acc, col[3 * i]);
      acc = __fdiv_rn(acc, fden);
    }
    const double s0 = __dmul_rn(ge, (double)__shfl_sync(0xffffffffu, acc, 0));
    const double s1 = __dmul_rn(ge, (double)__shfl_sync(0xffffffffu, acc, 1));
    const double s2 = __dmul_rn(ge, (double)__shfl_sync(0xffffffffu, acc, 2));
    float* dst = p.out + s * p.out_stride;
    for (int i = lane; i < nv3; i += 32) {   // :441-446
      const int j = i % 3;
      const double sj = j == 0 ? s0 : (j == 1 ? s1 : s2);
      dst[i] = (float)__dadd_rn(__dmul_rn(w, (double)sh[i]), sj);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) field_pl_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = p.nv, nv3 = nv * 3;
  // per warp: normalised gp[nv][3] (double) | phred index [nv][3] (uint8, padded to 8 bytes)
  const size_t per_warp = (size_t)nv3 * 8 + (((size_t)nv3 + 7) & ~(size_t)7);
  double* gpn = reinterpret_cast<double*>(smem + warp * per_warp);
  uint8_t* ph = reinterpret_cast<uint8_t*>(gpn + nv3);
  const double dan = (double)(2 * nv);       // an (:289), as newacs[i] / an converts it (:308)
  for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < p.nsnps; s += (int64_t)gridDim.x * wpb) {
    const int32_t* row = static_cast<const int32_t*>(p.in) + s * nv3;
    for (int i = lane; i < nv3; i += 32) {
      const uint32_t v = (uint32_t)__ldcs(row + i);   // PhredHelper.h:40: uint32_t phred, > 255 reads entry 255
      ph[i] = (uint8_t)(v > 255u ? 255u : v);
    }
    __syncwarp();
    double acs0 = 0.5, acs1 = 0.5;                    // :257-258
    for (int it = 0; it < 10; ++it) {
      const double q0 = __dmul_rn(acs0, acs0);        // (1 * acs[0]) * acs[0]
      const double q1 = __dmul_rn(__dmul_rn(2.0, acs1), acs0);
      const double q2 = __dmul_rn(acs1, acs1);
      for (int i = lane; i < nv; i += 32) {           // :276-287
        const double g0 = __dmul_rn(q0, __ldg(p.prob + ph[3 * i]));
        const double g1 = __dmul_rn(q1, __ldg(p.prob + ph[3 * i + 1]));
        const double g2 = __dmul_rn(q2, __ldg(p.prob + ph[3 * i + 2]));
        const double sum = __dadd_rn(__dadd_rn(__dadd_rn(0.0, g0), g1), g2);
        gpn[3 * i] = __ddiv_rn(g0, sum);
        gpn[3 * i + 1] = __ddiv_rn(g1, sum);
        gpn[3 * i + 2] = __ddiv_rn(g2, sum);
      }
      __syncwarp();
      if (it == 9) {                                  // :301-304
        float* dst = p.out + s * p.out_stride;
        for (int i = lane; i < nv3; i += 32) dst[i] = (float)gpn[i];
        if (!p.af) break;
      }
      // newacs[0] takes gp0, gp0, gp1 of every sample, newacs[1] takes gp1, gp2, gp2, in sample order (:283-286)
      double acc = 0.0;
      if (lane < 2) {
        const double* col = gpn + lane;
        int i = 0;
        for (; i + 4 <= nv; i += 4) {
          double a[4], b[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = col[3 * (i + u)]; b[u] = col[3 * (i + u) + 1]; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc = __dadd_rn(acc, a[u]);
            acc = __dadd_rn(acc, lane == 0 ? a[u] : b[u]);
            acc = __dadd_rn(acc, b[u]);
          }
        }
        for (; i < nv; ++i) {
          const double a = col[3 * i], b = col[3 * i + 1];
          acc = __dadd_rn(acc, a);
          acc = __dadd_rn(acc, lane == 0 ? a : b);
          acc = __dadd_rn(acc, b);
        }
        acc = __ddiv_rn(acc, dan);                    // :308
      }
      acs0 = shfl_d(acc, 0);
      acs1 = shfl_d(acc, 1);
      __syncwarp();
    }
    if (p.af && lane == 0)                            // :312-313, then get_af(1)
      p.af[s] = __ddiv_rn(__dadd_rn(__dmul_rn(acs1, dan), 0.5), __dadd_rn(dan, 1.0));
    __syncwarp();
  }
}

}  // namespace

size_t field_smem_per_warp(int field, int nv) {
  const size_t nv3 = (size_t)nv * 3;
  return field == DMX_FIELD_PL ? nv3 * 8 + ((nv3 + 7) & ~(size_t)7) : nv3 * 4;
}

int launch_field(const FieldParams& p, int sm_count, void* stream) {
  if (p.nsnps == 0) return 0;
  const size_t per_warp = field_smem_per_warp(p.field, p.nv);
  int wpb = 8;
  while (wpb > 1 && per_warp * wpb > 48 * 1024) wpb >>= 1;   // keep several CTAs per SM resident
  const size_t smem = per_warp * wpb;
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  void (*k)(const FieldParams) = p.field == DMX_FIELD_GT ? field_gt_kernel
                                 : (p.field == DMX_FIELD_GP ? field_gp_kernel : field_pl_kernel);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, wpb * 32, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  // whole waves of resident CTAs over the SMs; warps stride over the SNPs
  int64_t grid = (int64_t)sm_count * per_sm;
  const int64_t need = (p.nsnps + wpb - 1) / wpb;
  if (grid > need) grid = need;
  k<<<(unsigned)grid, wpb * 32, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace dmx

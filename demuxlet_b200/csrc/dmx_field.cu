// dmx_field.cu -- VCF FORMAT field -> genotype-probability matrix on the device.
//
// What the reference does once per VCF record on one host thread (BCFFilteredReader::parse_posteriors,
// /root/reference/bcf_filtered_reader.cpp:360-454; parse_genotypes :186-242; parse_likelihoods :244-320) and then
// widens into scl.snps[i].gps (cmd_cram_demuxlet.cpp:227-231), here as one warp per SNP (GT) or per group of
// consecutive SNPs (GP, PL):
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
// SNP) for GT and GP; PL is bound by its 30 IEEE double divisions per sample and SNP (ten normalisations of three
// genotypes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "dmx_internal.h"

namespace dmx {

__host__ __device__ size_t field_smem_per_warp(int field, int nv, int group, int pitch);

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
#pragma unroll
  for (int i = lane; i < nv3; i += 32) dst[i] = sh[i];
}

template <int NV>
__global__ void __launch_bounds__(256) field_gt_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = NV ? NV : p.nv, nv3 = nv * 3;
  float* sh = reinterpret_cast<float*>(smem) + (size_t)warp * nv3;
  const double ge = p.gt_error;
  const float f_hit = (float)__dsub_rn(1.0, ge), f_miss = (float)__ddiv_rn(ge, 2.0);   // :400
  for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < p.nsnps; s += (int64_t)gridDim.x * wpb) {
    const int8_t* al = static_cast<const int8_t*>(p.in) + s * nv * 2;
    int cnt = 0;                                     // c0 | c1 << 13 | bad << 26 (nv <= 4095: counts <= 8190)
#pragma unroll
    for (int i = lane; i < nv; i += 32) {
      const char2 a = reinterpret_cast<const char2*>(al)[i];
      cnt += ((a.x == 0) + (a.y == 0)) + (((a.x == 1) + (a.y == 1)) << 13) + (((a.x > 1) | (a.y > 1)) << 26);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int c0 = cnt & 0x1fff, c1 = (cnt >> 13) & 0x1fff, bad = cnt >> 26;
    if (bad && lane == 0) atomicOr(p.err_flag, 1);
    if (p.pass && lane == 0) {
      // passed_vfilter with the thresholds demuxlet exposes (bcf_filtered_reader.cpp:553-561; --min-callrate, --min-mac)
      const int an = c0 + c1, ac = c1;
      const bool lowcall = p.min_callrate > __ddiv_rn((double)an, __dmul_rn(2.0, (double)nv));
      p.pass[s] = (uint8_t)(!lowcall && !(ac < p.min_mac || an - ac < p.min_mac));
    }
    if (!p.out) continue;                            // filter only
    // :385: (j == k ? 1.0 : 2.0) * (acs[j] + 1.0 / nalleles) / (an + 1.0) * (acs[k] + 1.0 / nalleles) / (an + 1.0)
    const double a0 = __dadd_rn((double)c0, 0.5), a1 = __dadd_rn((double)c1, 0.5), den = __dadd_rn((double)(c0 + c1), 1.0);
    // lanes 0..2 take one genotype each (two divisions per warp instead of six)
    const double fj = lane == 1 ? 2.0 : 1.0, aj = lane == 0 ? a0 : a1, ak = lane == 2 ? a1 : a0;
    const float mj = (float)__ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(fj, aj), den), ak), den);
    const float m0 = __shfl_sync(0xffffffffu, mj, 0), m1 = __shfl_sync(0xffffffffu, mj, 1),
                m2 = __shfl_sync(0xffffffffu, mj, 2);
#pragma unroll
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

// GP and PL process a GROUP of G consecutive SNPs per warp pass: the order-dependent running sums then occupy 3 G
// (GP) or 2 G (PL) lanes instead of 3 or 2, and their dependent add chains -- nv float adds (GP), 3 nv x 10 double adds
// (PL) per SNP -- cost 1 / G of the issue slots.  Rows sit in shared memory at a pitch that puts the chain lanes of
// different SNPs on different banks.
template <int NV>
__global__ void __launch_bounds__(256) field_gp_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = NV ? NV : p.nv, nv3 = nv * 3, G = p.group, pitch = p.pitch;
  float* sh = reinterpret_cast<float*>(smem) + (size_t)warp * G * pitch;
  const double ge = p.gt_error, w = __dsub_rn(1.0, ge);
  const float fden = (float)(int32_t)(nv + 1.0);   // :438
  const int cg = lane / 3, cj = lane - cg * 3;     // chain lane: SNP cg of the group, genotype cj
  const int64_t ngroups = (p.nsnps + G - 1) / G;
  for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups; grp += (int64_t)gridDim.x * wpb) {
    const int64_t s0 = grp * G;
    const int ng = (int)min((int64_t)G, p.nsnps - s0);
    const float* src = static_cast<const float*>(p.in) + s0 * nv3;
    for (int g = 0; g < ng; ++g) {
#pragma unroll
      for (int i = lane; i < nv3; i += 32) sh[g * pitch + i] = __ldcs(src + g * nv3 + i);
    }
    __syncwarp();
    for (int g = 0; g < ng; ++g)
#pragma unroll
      for (int i = lane; i < nv; i += 32) {          // :424-431 (stride of 3 floats: no bank conflicts)
        float* q = sh + g * pitch + 3 * i;
        const float a = q[0], b = q[1], c = q[2];
        const float sum = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, a), b), c);
        q[0] = __fdiv_rn(a, sum);
        q[1] = __fdiv_rn(b, sum);
        q[2] = __fdiv_rn(c, sum);
      }
    __syncwarp();
    float acc = cj == 1 ? 0.5f : 0.25f;              // :417-421: the pseudo sample
    if (cg < ng) {
      const float* col = sh + cg * pitch + cj;
      int i = 0;
      for (; i + 8 <= nv; i += 8) {                  // loads batched ahead of the dependent add chain
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[3 * (i + u)];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
      }
      for (; i < nv; ++i) acc = __fadd_rn(acc, col[3 * i]);
      acc = __fdiv_rn(acc, fden);                    // :437-438
    }
    const double gacc = __dmul_rn(ge, (double)acc);  // gt_error * gpSums[j]
    for (int g = 0; g < ng; ++g) {
      const double t0 = shfl_d(gacc, 3 * g), t1 = shfl_d(gacc, 3 * g + 1), t2 = shfl_d(gacc, 3 * g + 2);
      float* dst = p.out + (s0 + g) * p.out_stride;
      const float* row = sh + g * pitch;
      int j = lane % 3;
#pragma unroll
      for (int i = lane; i < nv3; i += 32) {         // :441-446
        const double sj = j == 0 ? t0 : (j == 1 ? t1 : t2);
        dst[i] = (float)__dadd_rn(__dmul_rn(w, (double)row[i]), sj);
        j += 2;                                      // (i + 32) % 3
        if (j >= 3) j -= 3;
      }
    }
    __syncwarp();
  }
}

// RN(a / s) given r = RN(1 / s), for normal operands far from the exponent limits (here: 1e-60 < a <= s <= 3)
__device__ __forceinline__ double div_shared(double a, double s, double r) {
  const double q = __dmul_rn(a, r);
  const double rem = __fma_rn(-q, s, a);
  return __fma_rn(rem, r, q);
}

template <int NV>
__global__ void __launch_bounds__(256) field_pl_kernel(const FieldParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = NV ? NV : p.nv, nv3 = nv * 3, G = p.group, pitch = p.pitch;
  // per warp: normalised gp[G][pitch] (double) | phred index [G][nv3] (uint8)
  const size_t per_warp = field_smem_per_warp(DMX_FIELD_PL, nv, G, pitch);
  double* gpn = reinterpret_cast<double*>(smem + warp * per_warp);
  uint8_t* ph = reinterpret_cast<uint8_t*>(gpn + (size_t)G * pitch);
  const double dan = (double)(2 * nv);       // an (:289), as newacs[i] / an converts it (:308)
  const int cg = lane >> 1, cl = lane & 1;   // chain lane: SNP cg of the group, allele cl
  const int64_t ngroups = (p.nsnps + G - 1) / G;
  for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups; grp += (int64_t)gridDim.x * wpb) {
    const int64_t s0 = grp * G;
    const int ng = (int)min((int64_t)G, p.nsnps - s0);
    const int32_t* src = static_cast<const int32_t*>(p.in) + s0 * nv3;
    for (int i = lane; i < ng * nv3; i += 32) {       // the group's rows are contiguous
      const uint32_t v = (uint32_t)__ldcs(src + i);   // PhredHelper.h:40: uint32_t phred, > 255 reads entry 255
      ph[i] = (uint8_t)(v > 255u ? 255u : v);
    }
    __syncwarp();
    double acs = 0.5;                                 // :257-258; lane (cg, cl) holds acs[cl] of SNP cg
    for (int it = 0; it < 10; ++it) {
      for (int g = 0; g < ng; ++g) {
        const double acs0 = shfl_d(acs, 2 * g), acs1 = shfl_d(acs, 2 * g + 1);
        const double q0 = __dmul_rn(acs0, acs0);      // (1 * acs[0]) * acs[0]
        const double q1 = __dmul_rn(__dmul_rn(2.0, acs1), acs0);
        const double q2 = __dmul_rn(acs1, acs1);
        const uint8_t* pg = ph + g * nv3;
        double* og = gpn + (size_t)g * pitch;
#pragma unroll
        for (int i = lane; i < nv; i += 32) {         // :276-287
          const double g0 = __dmul_rn(q0, __ldg(p.prob + pg[3 * i]));
          const double g1 = __dmul_rn(q1, __ldg(p.prob + pg[3 * i + 1]));
          const double g2 = __dmul_rn(q2, __ldg(p.prob + pg[3 * i + 2]));
          const double sum = __dadd_rn(__dadd_rn(__dadd_rn(0.0, g0), g1), g2);
          // the three IEEE quotients by the same sum (:276-287) from ONE correctly rounded reciprocal: q = a * r,
          // residual a - q * sum exact in an FMA, q' = q + residual * r -- Markstein's correction step, the tail of the
          // hardware division sequence itself; checked bit for bit against __ddiv_rn over all 256^3 phred triples and
          // 2^30 random operands by tools/pl_div_check.cu (profiles/)
          const double rs = __drcp_rn(sum);
          og[3 * i] = div_shared(g0, sum, rs);
          og[3 * i + 1] = div_shared(g1, sum, rs);
          og[3 * i + 2] = div_shared(g2, sum, rs);
        }
      }
      __syncwarp();
      if (it == 9) {                                  // :301-304
        for (int g = 0; g < ng; ++g) {
          float* dst = p.out + (s0 + g) * p.out_stride;
          const double* og = gpn + (size_t)g * pitch;
#pragma unroll
          for (int i = lane; i < nv3; i += 32) dst[i] = (float)og[i];
        }
        if (!p.af) break;
      }
      // newacs[0] takes gp0, gp0, gp1 of every sample, newacs[1] takes gp1, gp2, gp2, in sample order (:283-286)
      double acc = 0.0;
      if (cg < ng) {
        const double* col = gpn + (size_t)cg * pitch + cl;
        int i = 0;
        for (; i + 4 <= nv; i += 4) {
          double a[4], b[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = col[3 * (i + u)]; b[u] = col[3 * (i + u) + 1]; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc = __dadd_rn(acc, a[u]);
            acc = __dadd_rn(acc, cl == 0 ? a[u] : b[u]);
            acc = __dadd_rn(acc, b[u]);
          }
        }
        for (; i < nv; ++i) {
          const double a = col[3 * i], b = col[3 * i + 1];
          acc = __dadd_rn(acc, a);
          acc = __dadd_rn(acc, cl == 0 ? a : b);
          acc = __dadd_rn(acc, b);
        }
        acc = __ddiv_rn(acc, dan);                    // :308
      }
      acs = acc;
      __syncwarp();
    }
    if (p.af && cl == 1 && cg < ng)                   // :312-313, then get_af(1)
      p.af[s0 + cg] = __ddiv_rn(__dadd_rn(__dmul_rn(acs, dan), 0.5), __dadd_rn(dan, 1.0));
    __syncwarp();
  }
}

}  // namespace

__host__ __device__ size_t field_smem_per_warp(int field, int nv, int group, int pitch) {
  const size_t nv3 = (size_t)nv * 3;
  if (field == DMX_FIELD_GT) return nv3 * 4;
  if (field == DMX_FIELD_GP) return (size_t)group * pitch * 4;
  return (size_t)group * pitch * 8 + (((size_t)group * nv3 + 7) & ~(size_t)7);
}

// SNPs per warp pass and the shared-memory row pitch (elements) of the grouped kernels: chain lane (g, j) starts at
// element g * pitch + j, so pitch = 3 (mod 32) floats / 2 (mod 16) doubles spreads the chain lanes over all banks
void field_geometry(int field, int nv, int* group, int* pitch) {
  const int nv3 = nv * 3;
  if (field == DMX_FIELD_GP) {
    *pitch = nv3 + ((3 - nv3 % 32) + 32) % 32;
    *group = std::max(1, std::min(10, 4096 / (*pitch * 4)));
  } else if (field == DMX_FIELD_PL) {
    *pitch = nv3 + ((2 - nv3 % 16) + 16) % 16;
    *group = std::max(1, std::min(16, 7168 / (*pitch * 8 + nv3)));
  } else {
    *pitch = nv3;
    *group = 1;
  }
}

int launch_field(const FieldParams& p0, int sm_count, void* stream) {
  if (p0.nsnps == 0) return 0;
  FieldParams p = p0;
  field_geometry(p.field, p.nv, &p.group, &p.pitch);
  const size_t per_warp = field_smem_per_warp(p.field, p.nv, p.group, p.pitch);
  int wpb = 8;
  while (wpb > 1 && per_warp * wpb > 56 * 1024) wpb >>= 1;   // keep several CTAs per SM resident
  const size_t smem = per_warp * wpb;
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  // sample counts of the benchmark configurations are compile-time constants (unrolled row loops)
  void (*k)(const FieldParams) = nullptr;
#define DMX_FIELD_PICK(NVC)                                                                       \
  k = p.field == DMX_FIELD_GT ? field_gt_kernel<NVC>                                              \
                              : (p.field == DMX_FIELD_GP ? field_gp_kernel<NVC> : field_pl_kernel<NVC>)
  if (p.nv == 64) DMX_FIELD_PICK(64);
  else if (p.nv == 8) DMX_FIELD_PICK(8);
  else if (p.nv == 256) DMX_FIELD_PICK(256);
  else DMX_FIELD_PICK(0);
#undef DMX_FIELD_PICK
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, wpb * 32, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  // whole waves of resident CTAs over the SMs; warps stride over the SNP groups
  int64_t grid = (int64_t)sm_count * per_sm;
  const int64_t need = ((p.nsnps + p.group - 1) / p.group + wpb - 1) / wpb;
  if (grid > need) grid = need;
  k<<<(unsigned)grid, wpb * 32, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace dmx

// dmx_kernels.cu -- sm_100a kernels of the demuxlet likelihood hot path.
//
// What is computed (reference: /root/reference/cmd_cram_demuxlet.cpp):
//   entq_kernel       :424-452 + :597-663  per (cell, snp) entry: read likelihood table Q'(p) on the
//                     distinct mixing proportions p(n,l,m) (:613) and the pass-1 GLs
//   gprow_kernel      :391-401  cohort mean genotype probability per SNP (gp0) in front of every GP row
//   pair_scan_kernel  :455-459 + :665-710  per cell: sum over its entries of log(sum_{l,m} gp_j[l] gp_k[m]
//                     Q'[n][l][m]) for every sample pair (j,k) of every dense alpha plane, the
//                     separable alpha = 0 plane factors, llks00, and the pass-1 singlet llks;
//                     then the per-task reductions of :712-734 and :799-814
//   merge_kernel      combines the tasks of a cell (:712-734, :799-814)
//
// Design: the bilinear form is factored as S[j][k] = sum_l a_j[l] * w_k[l] with
// w_k[l] = sum_m Q'[l][m] gp_k[m] (3 FMA per hypothesis instead of 9 terms over a precomputed table); entries with at
// most one informative read collapse further to S' = X_j + Y_k (one add, one multiply); sum-of-logs is replaced by
// the log of a running product with the exponent kept in an integer (one log per hypothesis and cell).
// Three persistent, warp-specialised scan kernels, all mbarrier pipelines fed by TMA (tile::gather4 of table rows by
// SNP id + bulk copies of the entry records):
//   pair_scan_v2_kernel   (dmx_scan2.cuh) default fp32 scan: the gathers deliver ready-to-use operands (dosage rows,
//                         row sums, raw GP rows): 3 loader, 3 "singles" (per-sample products) and 9 consumer warps,
//                         a thread owns an 8 x 8 tile of hypotheses for one of the 8 entries of a stage
//   pair_scan_f32_kernel  first-generation fp32 scan (DMX_SCAN_GEN=1): 2 loader, 4 transform, 9 consumer warps and an
//                         accumulator warp; operands pass through a processed ring
//   pair_scan_kernel      fp64 scan on the FP64 tensor path (DMMA): two pipelines of 1 loader, 2 transform and 5
//                         consumer warps; dense llksAB blocks (--write-pair) and the DMX_FP32_PAIRS-off mode
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <type_traits>

#include "dmx_internal.h"

namespace dmx {

__constant__ double c_err[256];   // phred2Err, PhredHelper.cpp:30-32 (computed on the host with libm pow)
__constant__ double c_err3[256];  // phred2Err / 3.0, the quotient the reference forms at :606-607 (host-computed)
__constant__ double c_logtab[18 + 256];   // the host libm's __log_data (read from the loaded libm.so.6 at run time)
__device__ double g_logtab[18 + 256];     // same table in global memory: the per-interval entries are indexed
                                          // differently by every lane (a constant-bank access would serialise)

// Bit-for-bit transcription of the x86-64 FMA build of glibc's log() (sysdeps/ieee754/dbl-64/e_log.c as
// compiled into libm.so.6 2.3x: instruction sequence of __log_fma, table __log_data).  IEEE fma/mul/add are
// exact on both machines, so with the host's own table this returns the very double the reference's
// `log(sumPs)` (cmd_cram_demuxlet.cpp:683) yields on an FMA-capable x86 host.  dmx_create verifies that
// claim against the host libm on a large sample before the device re-scorer is trusted.
__device__ __forceinline__ double glibc_log_fma(double x) {
  const double* T = c_logtab;
  unsigned long long ix = (unsigned long long)__double_as_longlong(x);
  if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {          // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = __dadd_rn(x, -1.0);
    double p1 = __fma_rn(r, T[9], T[8]);
    double p2 = __fma_rn(r, T[12], T[11]);
    const double r2 = __dmul_rn(r, r);
    double p3 = __fma_rn(r, T[15], T[14]);
    p1 = __fma_rn(r2, T[10], p1);
    p2 = __fma_rn(r2, T[13], p2);
    const double r3 = __dmul_rn(r, r2);
    p3 = __fma_rn(r2, T[16], p3);
    p3 = __fma_rn(r3, T[17], p3);
    const double q = __fma_rn(p3, r3, p2);
    const double poly = __fma_rn(q, r3, p1);
    const double t = __fma_rn(r, 134217728.0, r);
    const double rhi = __fma_rn(-134217728.0, r, t);
    const double rhi2 = __dmul_rn(rhi, rhi);
    const double rlo = __dsub_rn(r, rhi);
    const double hi = __fma_rn(rhi2, T[7], r);
    double lo = __fma_rn(rhi2, T[7], __dsub_rn(r, hi));
    const double s = __dadd_rn(r, rhi);
    const double b = __dmul_rn(T[7], rlo);
    lo = __fma_rn(b, s, lo);
    const double y = __fma_rn(poly, r3, lo);
    return __dadd_rn(hi, y);
  }
  const unsigned top = (unsigned)(ix >> 48);
  if (top - 0x10u > 0x7fdfu) {
    if (ix * 2 == 0) return -INFINITY;                              // log(+-0)
    if (ix == 0x7ff0000000000000ull) return x;                      // log(inf)
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return __longlong_as_double(0x7ff8000000000000ll);
    ix = (unsigned long long)__double_as_longlong(__dmul_rn(x, 4503599627370496.0));   // subnormal
    ix -= 52ull << 52;
  }
  const unsigned long long tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((long long)tmp >> 52);
  const unsigned long long iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = __ldg(&g_logtab[18 + 2 * i]), logc = __ldg(&g_logtab[19 + 2 * i]);
  const double z = __longlong_as_double((long long)iz);
  const double kd = (double)k;
  const double w = __fma_rn(kd, T[0], logc);
  const double r = __fma_rn(z, invc, -1.0);
  const double t1 = __fma_rn(r, T[4], T[3]);
  const double hi = __dadd_rn(r, w);
  const double r2 = __dmul_rn(r, r);
  double lo = __dadd_rn(__dsub_rn(w, hi), r);
  lo = __fma_rn(kd, T[1], lo);
  const double r3 = __dmul_rn(r, r2);
  const double t2 = __fma_rn(r, T[6], T[5]);
  lo = __fma_rn(r2, T[2], lo);
  const double t3 = __fma_rn(t2, r2, t1);
  const double y = __fma_rn(r3, t3, lo);
  return __dadd_rn(y, hi);
}

__global__ void logtest_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = glibc_log_fma(x[i]);
}
int launch_logtest(const double* x, double* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  logtest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return (int)cudaGetLastError();
}
int set_log_table(const double* tab274) {
  cudaError_t e = cudaMemcpyToSymbol(c_logtab, tab274, (18 + 256) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_logtab, tab274, (18 + 256) * sizeof(double));
  return (int)e;
}

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA 1-D)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
#ifndef DMX_WAIT_HINT_NS
#define DMX_WAIT_HINT_NS 2000
#endif
#ifndef DMX_WAIT_SLEEP_NS
#define DMX_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if DMX_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // suspend-time hint (ns): fewer retry issues
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "n"(DMX_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
#if DMX_WAIT_SLEEP_NS > 0
    __nanosleep(DMX_WAIT_SLEEP_NS);   // a waiting warp must not eat the issue slots of the warps it waits for
#endif
  }
}
// wait of a role that runs far ahead of its consumers (loaders): sleep between polls instead of competing for the
// issue slots of the warps that do the work
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#ifndef DMX_RELAX_NS
#define DMX_RELAX_NS 256
#endif
  while (!mbar_try_wait(bar, parity)) {
#if DMX_RELAX_NS > 0
    __nanosleep(DMX_RELAX_NS);
#endif
  }
}
// DMX_PROF (experiment builds only, tools/build_variant.sh): cycles every warp role of the fp32 scan spends inside its
// mbarrier waits, summed over warps and CTAs; launch_pair_scan prints and clears the table after every launch.
// role 0 loader, 1 transform, 2 accumulator, 3 consumer; counter 0 = role total, 7 = warps, 1..6 = waits (see roles)
#ifndef DMX_PROF
#define DMX_PROF 0
#endif
#if DMX_PROF
__device__ unsigned long long g_prof[4][8];
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0_ = clock64();
#define PROF_WAIT(i, stmt) { const long long t0_ = clock64(); stmt; prof_[i] += clock64() - t0_; }
#define PROF_END(role, lane) if ((lane) == 0) { prof_[0] = clock64() - prof_t0_; prof_[7] = 1; \
    for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_prof[role][i_], (unsigned long long)prof_[i_]); }
#else
#define PROF_DECL
#define PROF_WAIT(i, stmt) { stmt; }
#define PROF_END(role, lane)
#endif
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// TMA tensor copies (descriptors built on the host, dmx_api.cu:encode_maps).  The *_elect forms are executed by a
// converged warp with warp-uniform operands; elect.sync picks the one lane that issues.
__device__ __forceinline__ void tma_gather4_elect(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2,
                                                  int r3, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4, %5, %6}], [%7];\n\t}" ::"r"(dst),
      "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_tile2d_elect(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar),
      "r"(bytes)
      : "memory");
}

// RN(a / s) given r = RN(1 / s) (__drcp_rn): q = a * r, the residual a - q * s exact in an FMA, q' = q + residual * r --
// Markstein's correction step, the tail of the hardware's own division sequence.  For normal operands far from the
// exponent limits; checked bit for bit against __ddiv_rn by tools/pl_div_check.cu (3.5 G + 2^30 quotients, 0 mismatches).
__device__ __forceinline__ double div_shared(double a, double s, double r) {
  const double q = __dmul_rn(a, r);
  const double rem = __fma_rn(-q, s, a);
  return __fma_rn(rem, r, q);
}

// peel the binary exponent of a positive double into an integer accumulator; mantissa stays in [1,2)
__device__ __forceinline__ void peel(double& x, int& esum) {
  int hi = __double2hiint(x);
  esum += hi >> 20;
  x = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));
}
// fp32 flavour; a zero (or denormal) product is left alone: its log is -inf whatever the exponent sum says
__device__ __forceinline__ void peelf(float& x, int& esum) {
  const uint32_t b = __float_as_uint(x);
  if (b >= 0x00800000u) {
    esum += (int)(b >> 23);
    x = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
  } else {
    esum += 127;
  }
}
__device__ __forceinline__ double unpeelf_log(float x, int esum, int nflush) {
  return (double)(esum - 127 * nflush) * 0.693147180559945309417232121458 + log((double)x);
}
// log of a positive double to ~1e-7 absolute: binary exponent exactly, mantissa through the fp32 logf (the fp32 scan's
// values carry ~1e-6 of product rounding anyway; a double-precision log() costs ~100 dependent FP64 instructions)
__device__ __forceinline__ double log_pos_fast(double x) {
  const long long b = __double_as_longlong(x);
  const int k = (int)(b >> 52) - 1023;
  if (k <= -1023 || k >= 1024) return log(x);   // zero, denormal, inf / nan: the library's answer
  const float f = (float)__longlong_as_double((b & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
  return (double)k * 0.693147180559945309417232121458 + (double)logf(f);
}
__device__ __forceinline__ double unpeel_log(double x, int esum, int nflush) {
  return (double)(esum - 1023 * nflush) * 0.693147180559945309417232121458 + log(x);
}

// ---------------------------------------------------------------------------------------------
// gprow: builds the GP rows the TMA gathers read, [gp0: 3 doubles, 8 B pad][nvp x 3 floats], from the uploaded
// [nsnps][nvp*3] matrix.  gp0 = cohort mean GP per SNP, cmd_cram_demuxlet.cpp:391-401 (ascending-j double sums, then
// / nv): lanes 0..2 of the row's warp sum one genotype each while the other lanes copy the row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gprow_kernel(const float* __restrict__ gp, float* __restrict__ rows, int64_t nsnps,
                                                    int nv, int nvp) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= nsnps) return;
  const int n3 = nvp * 3;
  const float* g = gp + s * (int64_t)n3;
  float* out = rows + s * (int64_t)(n3 + 8);
  for (int i = lane; i < n3; i += 32) out[8 + i] = g[i];
  if (lane < 3) {
    double a = 0;
    for (int j = 0; j < nv; ++j) a = __dadd_rn(a, (double)g[3 * j + lane]);
    reinterpret_cast<double*>(out)[lane] = __ddiv_rn(a, (double)nv);
  } else if (lane == 3) {
    reinterpret_cast<double*>(out)[3] = 0.0;
  }
}

int launch_gprow(const float* gp_flat, float* gp_rows, int64_t nsnps, int nv, int nvp, void* stream) {
  if (nsnps == 0) return 0;
  gprow_kernel<<<(unsigned)((nsnps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gp_flat, gp_rows, nsnps, nv, nvp);
  return (int)cudaGetLastError();
}

// dtab: the tables the second-generation fp32 scan gathers for affine entries (dmx_scan2.cuh), from the GP rows:
//   drow[2 snp + a][4 + j]  expected dosage of allele a under sample j's genotype vector normalised to sum 1
//                           (a = 0: 2 g0 + g1, a = 1: g1 + 2 g2, over s = (g0 + g1) + g2)
//   srow[snp][4 + j]        s (its logs come back as the lb term of the scan's epilogue)
// Rows are nvp + 8 floats: four floats of padding on either side of the samples, so that a gather4 box may start four
// floats to the left of the block it fetches.  Padding samples (all-zero GP) get D = 0, s = 0.
__global__ void __launch_bounds__(256) dtab_kernel(const float* __restrict__ rows, float* __restrict__ drow,
                                                   float* __restrict__ srow, int64_t nsnps, int nvp) {
  const int pitch = nvp + 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsnps * pitch) return;
  const int64_t snp = i / pitch;
  const int c = (int)(i - snp * pitch), j = c - 4;
  float d0 = 0.f, d1 = 0.f, s = 0.f;
  if (j >= 0 && j < nvp) {
    const float* g = rows + snp * (int64_t)(nvp * 3 + 8) + 8 + 3 * j;
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    s = (g0 + g1) + g2;
    if (s > 0.f) {
      d0 = __fdiv_rn(fmaf(2.0f, g0, g1), s);
      d1 = __fdiv_rn(fmaf(2.0f, g2, g1), s);
    }
  }
  drow[(2 * snp) * pitch + c] = d0;
  drow[(2 * snp + 1) * pitch + c] = d1;
  srow[i] = s;
}

int launch_dtab(const float* gp_rows, float* drow, float* srow, int64_t nsnps, int nvp, void* stream) {
  if (nsnps == 0) return 0;
  const int64_t n = nsnps * (nvp + 8);
  dtab_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gp_rows, drow, srow, nsnps, nvp);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// entq: per-entry read likelihood on the distinct p values (:597-663) and pass-1 GLs (:424-452), in the
// reference's arithmetic (separately rounded mul/add/div, per-read max normalisation).  Outputs per entry:
//   entq[e][UQ]        fp64 record: GL[3] | f[3] (separable plane) | llk0s factor | llks00 factor of a separable
//                      plane 0 | per dense plane Q'[3][3] + its llks00 factor
//   entf_head/_plane   the same values rounded to fp32 (what the fp32 pair scan streams)
//   ents[e][2 + nd]    the three kinds of sample-independent factors (cell_scalars_kernel sums their logs per cell)
// Two kernels: entq1 (one thread per entry, all distinct values in registers; U <= 8, the common grids) and
// entq (LPE lanes per entry; lane `sub` owns p values sub, sub+LPE, ...) for long alpha grids.
// ---------------------------------------------------------------------------------------------
// GETV(slot): value of record slot `slot` (GL, f and Q' are already final when this runs)
template <class GETV>
__device__ __forceinline__ void entq_emit(const EntqParams& p, int64_t e, double* __restrict__ out, int nvr, uint32_t al1,
                                          GETV getv) {
  // per-entry factors of the sample-independent likelihoods (they need only this record and the cohort mean gp0 of
  // the entry's SNP): llk0s (:459) -> head[6], llks00 of a separable plane 0 -> head[7], llks00[n] (:700-709) -> slot 9
  // of each dense plane; then the fp32 copies
  const double* z = reinterpret_cast<const double*>(p.gprow + (int64_t)__ldg(p.ent_snp + e) * p.gp_stride);
  const double z0 = __ldg(z), z1 = __ldg(z + 1), z2 = __ldg(z + 2);
  const double g0 = getv(0), g1 = getv(1), g2 = getv(2);
  double f0 = 0.0, f1 = 0.0, f2 = 0.0;
  const double s0 = fma(g2, z2, fma(g1, z1, g0 * z0));
  double s1 = 1.0;
  if (p.sep0) {
    f0 = getv(3); f1 = getv(4); f2 = getv(5);
    s1 = fma(f2, z2, fma(f1, z1, f0 * z0)) * ((z0 + z1) + z2);
  }
  out[6] = s0;
  out[7] = s1;
  double* sc = p.ents + e * (2 + 2 * p.nd);
  sc[0] = s0;
  sc[1] = s1;
  // fp32 tables of the fp32 scan, class-sorted inside the cell (affine entries first)
  const bool affine = nvr <= 1;
  const bool refread = al1 == 0;           // the informative read shows the reference allele: D = ref dosage
  const int64_t pos = p.perm ? e + (int64_t)__ldg(p.perm + e) : e;
  if (p.entf_head) {
    if (p.snp_perm) p.snp_perm[pos] = affine ? 2 * __ldg(p.ent_snp + e) + (refread ? 0 : 1) : __ldg(p.ent_snp + e);
    float4* fh = reinterpret_cast<float4*>(p.entf_head + pos * FHEAD);
    if (affine) {
      // value at the dosage-0 corner and half the difference to the dosage-2 corner (>= 0 by construction)
      const double h1 = refread ? g2 : g0, c1 = 0.5 * (refread ? g0 - g2 : g2 - g0);
      const double hs = refread ? f2 : f0, cs = 0.5 * (refread ? f0 - f2 : f2 - f0);
      fh[0] = make_float4((float)h1, (float)c1, (float)hs, (float)cs);
      fh[1] = make_float4(refread ? 1.f : 0.f, 0.f, 0.f, 0.f);
    } else {
      fh[0] = make_float4((float)g0, (float)g1, (float)g2, (float)f0);
      fh[1] = make_float4((float)f1, (float)f2, 0.f, 0.f);
    }
  }
  for (int d = 0; d < p.nd; ++d) {
    double* q = out + QHEAD + QPLANE * d;
    double Q[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] = getv(QHEAD + QPLANE * d + i);
    const double t0 = fma(Q[2], z2, fma(Q[1], z1, Q[0] * z0));
    const double t1 = fma(Q[5], z2, fma(Q[4], z1, Q[3] * z0));
    const double t2 = fma(Q[8], z2, fma(Q[7], z1, Q[6] * z0));
    const double s2 = fma(z2, t2, fma(z1, t1, z0 * t0));
    q[9] = s2;
    sc[2 + d] = s2;
    sc[2 + p.nd + d] = 1.0;
    if (p.entf_plane) {
      float4* fp = reinterpret_cast<float4*>(p.entf_plane + d * p.entf_pstride + pos * FPLANE);
      if (affine) {
        // Q'[l][m] = q + cl * (dosage of l) + cm * (dosage of m) around the corner where the read's allele is absent
        const double q = refread ? Q[8] : Q[0];
        const double cl = 0.5 * (refread ? Q[2] - Q[8] : Q[6] - Q[0]);   // along l (j side)
        const double cm = 0.5 * (refread ? Q[6] - Q[8] : Q[2] - Q[0]);   // along m (k side)
        fp[0] = make_float4((float)(0.5 * q), (float)cl, (float)cm, 0.f);
        // second-generation scan: Q' / M with M = the largest corner value of Q' (both genotypes with two copies of the
        // read's allele, unless the base error is above 3/4 and a slope negative): alpha + bl * (dosage of l) +
        // bm * (dosage of m) in (0, 1]
        const double M = q + 2.0 * (fmax(cl, 0.0) + fmax(cm, 0.0));
        fp[1] = make_float4((float)(q / M), (float)(cl / M), (float)(cm / M), 0.f);
        if (affine) sc[2 + p.nd + d] = M;
      } else {
        fp[0] = make_float4((float)Q[0], (float)Q[1], (float)Q[2], (float)Q[3]);
        fp[1] = make_float4((float)Q[4], (float)Q[5], (float)Q[6], (float)Q[7]);
        fp[2] = make_float4((float)Q[8], 0.f, 0.f, 0.f);
      }
    }
  }
}

template <int LPE, int VPL>
__global__ void __launch_bounds__(256) entq_kernel(const __grid_constant__ EntqParams p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPE;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t e = p.E0 + gw * (32 / LPE) + lane / LPE;
  if (e >= p.E) return;   // whole groups leave together
  if (p.glist) {            // fp32 scan: by position, general entries only
    const int32_t v = __ldg(p.glist + e);
    if (v >= 0) return;
    e = (int64_t)(v & 0x7fffffff);
  }
  const unsigned gmask = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << (lane - sub));
  double P[VPL], pu[VPL];
  bool ok[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    int u = sub + v * LPE;
    ok[v] = u < p.U;
    pu[v] = ok[v] ? p.pu[u] : 0.0;
    P[v] = 1.0;
  }
  double G0 = 1.0, G1 = 1.0, G2 = 1.0;
  int nvr = 0;
  uint32_t al1 = 1;
  const int64_t r0 = p.ent_ptr[e], r1 = p.ent_ptr[e + 1];
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t rd = p.reads[r];
    const uint32_t al = rd >> 8, bq = rd & 0xff;
    if (al == 2) continue;                                                  // :435, :604
    ++nvr;
    al1 = al;
    const double err = c_err[bq];
    const double e3 = __ddiv_rn(err, 3.0);
    const double mat = __dsub_rn(1.0, err);
    const double pR = (al == 0) ? mat : e3;                                 // :606
    const double pA = (al == 1) ? mat : e3;                                 // :607
    double mx = 0.0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const double t = __dadd_rn(__dmul_rn(pR, __dsub_rn(1.0, pu[v])), __dmul_rn(pA, pu[v]));   // :625
      P[v] = __dmul_rn(P[v], t);
      if (ok[v]) mx = fmax(mx, P[v]);
    }
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(gmask, mx, o));
#pragma unroll
    for (int v = 0; v < VPL; ++v) P[v] = __ddiv_rn(P[v], mx);               // :632-639
    if (sub == 0) {   // pass-1 GLs, :437-443
      G0 = __dmul_rn(G0, pR);
      G1 = __dmul_rn(G1, __dsub_rn(0.5, e3));
      G2 = __dmul_rn(G2, pA);
      const double sm = __dadd_rn(__dadd_rn(G0, G1), G2);
      G0 = __ddiv_rn(G0, sm);
      G1 = __ddiv_rn(G1, sm);
      G2 = __ddiv_rn(G2, sm);
    }
  }
  double mx = 0.0;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    P[v] = __dadd_rn(P[v], 1e-6);                                           // :643-654
    if (ok[v]) mx = fmax(mx, P[v]);
  }
#pragma unroll
  for (int o = LPE / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(gmask, mx, o));
  double* out = p.entq + e * p.UQ;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
    if (ok[v]) {
      const int u = sub + v * LPE;
      const double val = __ddiv_rn(P[v], mx);                                // :656-663
      for (int o = p.oslot_ptr[u]; o < p.oslot_ptr[u + 1]; ++o) out[p.oslot[o]] = val;
    }
  if (sub == 0) {
    G0 = __dadd_rn(G0, 1e-6);                                                // :446-452
    G1 = __dadd_rn(G1, 1e-6);
    G2 = __dadd_rn(G2, 1e-6);
    const double sm = __dadd_rn(__dadd_rn(G0, G1), G2);
    out[0] = __ddiv_rn(G0, sm);
    out[1] = __ddiv_rn(G1, sm);
    out[2] = __ddiv_rn(G2, sm);
  }
  __syncwarp(gmask);
  if (sub == 0) entq_emit(p, e, out, nvr, al1, [&](int slot) { return __ldcg(out + slot); });
}

// one thread per entry: the U <= UMAX distinct values live in registers (independent chains: the divisions of a read
// overlap), the phred tables in shared memory (lanes index them with different qualities)
template <int UMAX>
__global__ void __launch_bounds__(128) entq1_kernel(const __grid_constant__ EntqParams p) {
  __shared__ double s_err[256], s_err3[256];
  __shared__ double s_val[UMAX][128];
  for (int i = threadIdx.x; i < 256; i += 128) { s_err[i] = c_err[i]; s_err3[i] = c_err3[i]; }
  __syncthreads();
  int64_t e = p.E0 + (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= p.E) return;
  if (p.glist) {            // fp32 scan: by position, general entries only
    const int32_t v = __ldg(p.glist + e);
    if (v >= 0) return;
    e = (int64_t)(v & 0x7fffffff);
  }
  const int nu = p.U;
  double P[UMAX], pu[UMAX], qu[UMAX];
#pragma unroll
  for (int v = 0; v < UMAX; ++v) {
    pu[v] = v < nu ? p.pu[v] : 0.0;
    qu[v] = __dsub_rn(1.0, pu[v]);
    P[v] = 1.0;
  }
  double G0 = 1.0, G1 = 1.0, G2 = 1.0;
  int nvr = 0;
  uint32_t al1 = 1;
  const int64_t r0 = p.ent_ptr[e], r1 = p.ent_ptr[e + 1];
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t rd = p.reads[r];
    const uint32_t al = rd >> 8, bq = rd & 0xff;
    if (al == 2) continue;                                                  // :435, :604
    ++nvr;
    al1 = al;
    const double err = s_err[bq];
    const double e3 = s_err3[bq];                                           // err / 3 (same rounded quotient)
    const double mat = __dsub_rn(1.0, err);
    const double pR = (al == 0) ? mat : e3;                                 // :606
    const double pA = (al == 1) ? mat : e3;                                 // :607
    double mx = 0.0;
#pragma unroll
    for (int v = 0; v < UMAX; ++v) {
      const double t = __dadd_rn(__dmul_rn(pR, qu[v]), __dmul_rn(pA, pu[v]));   // :625
      P[v] = __dmul_rn(P[v], t);
      if (v < nu) mx = fmax(mx, P[v]);
    }
    {   // :632-639: the quotients by the same maximum from one correctly rounded reciprocal (bit-identical)
      const double rmx = __drcp_rn(mx);
#pragma unroll
      for (int v = 0; v < UMAX; ++v) P[v] = div_shared(P[v], mx, rmx);
    }
    // pass-1 GLs, :437-443.  Unlike Q' (which the exact re-scorer consumes bit for bit) the GLs only feed sums of
    // logs that are compared at 1e-4, so the three quotients by the same sum are one reciprocal and three products
    G0 = __dmul_rn(G0, pR);
    G1 = __dmul_rn(G1, __dsub_rn(0.5, e3));
    G2 = __dmul_rn(G2, pA);
    const double rs = __drcp_rn(__dadd_rn(__dadd_rn(G0, G1), G2));
    G0 *= rs; G1 *= rs; G2 *= rs;
  }
  double mx = 0.0;
#pragma unroll
  for (int v = 0; v < UMAX; ++v) {
    P[v] = __dadd_rn(P[v], 1e-6);                                           // :643-654
    if (v < nu) mx = fmax(mx, P[v]);
  }
  {
    const double rmx = __drcp_rn(mx);
#pragma unroll
    for (int v = 0; v < UMAX; ++v) s_val[v][threadIdx.x] = div_shared(P[v], mx, rmx);   // :656-663
  }
  double* out = p.entq + e * p.UQ;
  G0 = __dadd_rn(G0, 1e-6);                                                  // :446-452
  G1 = __dadd_rn(G1, 1e-6);
  G2 = __dadd_rn(G2, 1e-6);
  const double rs = __drcp_rn(__dadd_rn(__dadd_rn(G0, G1), G2));
  G0 *= rs; G1 *= rs; G2 *= rs;
  out[0] = G0;
  out[1] = G1;
  out[2] = G2;
  const int tid = threadIdx.x;
  if (p.sep0)
    for (int s = 3; s < 6; ++s) out[s] = s_val[p.slot_u[s]][tid];
  for (int d = 0; d < p.nd; ++d)
    for (int i = 0; i < 9; ++i) {
      const int s = QHEAD + QPLANE * d + i;
      out[s] = s_val[p.slot_u[s]][tid];
    }
  entq_emit(p, e, out, nvr, al1, [&](int slot) { return slot == 0 ? G0 : (slot == 1 ? G1 : (slot == 2 ? G2 : s_val[p.slot_u[slot]][tid])); });
}

// ent_ptr from read counts: one block per cell, block-wide exclusive scan of the counts in rounds of 128 entries
__global__ void __launch_bounds__(128) entptr_kernel(const int64_t* __restrict__ cell_ptr, const int64_t* __restrict__ cell_r,
                                                     const uint8_t* __restrict__ cnt, int64_t c0, int64_t c1,
                                                     int64_t* __restrict__ ent_ptr) {
  __shared__ int ws[4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t cell = c0 + (int64_t)blockIdx.x;
  if (cell >= c1) return;
  const int64_t e0 = cell_ptr[cell], e1 = cell_ptr[cell + 1];
  int64_t run = cell_r[cell];
  for (int64_t b = e0; b < e1; b += 128) {
    const int64_t e = b + threadIdx.x;
    const int v = e < e1 ? (int)cnt[e] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < w) pre += ws[q];
      tot += ws[q];
    }
    if (e < e1) ent_ptr[e] = run + pre + inc - v;
    run += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) ent_ptr[e1] = run;   // == cell_r[cell + 1]: the next cell's first offset, written by both
}

int launch_entptr(const int64_t* cell_ptr, const int64_t* cell_r, const uint8_t* cnt, int64_t c0, int64_t c1,
                  int64_t* ent_ptr, void* stream) {
  if (c1 <= c0) return 0;
  entptr_kernel<<<(unsigned)(c1 - c0), 128, 0, (cudaStream_t)stream>>>(cell_ptr, cell_r, cnt, c0, c1, ent_ptr);
  return (int)cudaGetLastError();
}

// entry classes: one warp per cell, one pass over its entries.  The affine entries take the front positions of the
// cell's block in entry order, the general ones the back positions in reverse entry order (any fixed order serves: the
// products commute and the exact re-scorer walks the original order).  glist[pos] = entry at that position, sign bit
// set for a general entry: the general entry-table kernels run over positions and skip the others.
__global__ void __launch_bounds__(128) classify_kernel(const int64_t* __restrict__ cell_ptr,
                                                       const int64_t* __restrict__ ent_ptr,
                                                       const uint16_t* __restrict__ reads, int64_t c0, int64_t c1,
                                                       int32_t* __restrict__ perm, int32_t* __restrict__ cell_na,
                                                       int32_t* __restrict__ glist) {
  // one block per cell (a warp per cell left the kernel waiting for the few largest cells): 128 entries per round,
  // positions from the warps' ballots and a prefix over the four warps
  __shared__ int wa[4], wg[4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t cell = c0 + (int64_t)blockIdx.x;
  if (cell >= c1) return;
  const int64_t e0 = cell_ptr[cell], e1 = cell_ptr[cell + 1];
  int ra = 0, rg = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t b = e0; b < e1; b += 128) {
    const int64_t e = b + threadIdx.x;
    const bool in = e < e1;
    bool aff = false;
    if (in) {
      int nvr = 0;
      const int64_t r1 = ent_ptr[e + 1];
      for (int64_t r = ent_ptr[e]; r < r1; ++r) nvr += (reads[r] >> 8) != 2;
      aff = nvr <= 1;
    }
    const unsigned ma = __ballot_sync(0xffffffffu, aff), mg = __ballot_sync(0xffffffffu, in && !aff);
    if (lane == 0) { wa[w] = __popc(ma); wg[w] = __popc(mg); }
    __syncthreads();
    int pa = 0, pg = 0, ta = 0, tg = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < w) { pa += wa[q]; pg += wg[q]; }
      ta += wa[q]; tg += wg[q];
    }
    if (in) {
      const int64_t pos = aff ? e0 + ra + pa + __popc(ma & lt) : e1 - 1 - (rg + pg + __popc(mg & lt));
      perm[e] = (int32_t)(pos - e);
      glist[pos] = (int32_t)e | (aff ? 0 : (int32_t)0x80000000);
    }
    ra += ta;
    rg += tg;
    __syncthreads();
  }
  if (threadIdx.x == 0) cell_na[cell] = ra;
}

int launch_classify(const int64_t* cell_ptr, const int64_t* ent_ptr, const uint16_t* reads, int64_t c0, int64_t c1,
                    int32_t* perm, int32_t* cell_na, int32_t* glist, void* stream) {
  if (c1 <= c0) return 0;
  classify_kernel<<<(unsigned)(c1 - c0), 128, 0, (cudaStream_t)stream>>>(cell_ptr, ent_ptr, reads, c0, c1, perm, cell_na,
                                                                         glist);
  return (int)cudaGetLastError();
}

// affine entries of the fp32 scan: records from the memo tables, the SNP-dependent factors computed here
__global__ void __launch_bounds__(256) entq_aff_kernel(const __grid_constant__ EntqParams p) {
  const int64_t e = p.E0 + (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= p.E) return;
  int nvr = 0;
  uint32_t rd1 = 0;
  for (int64_t r = p.ent_ptr[e]; r < p.ent_ptr[e + 1]; ++r) {
    const uint32_t rd = p.reads[r];
    if ((rd >> 8) != 2) { ++nvr; rd1 = rd; }
  }
  if (nvr > 1) {
    p.ecode[e] = 0xffffu;
    return;
  }
  const int idx = nvr == 0 ? MEMO_ROWS - 1 : (int)rd1;      // allele * 256 + quality
  p.ecode[e] = (uint16_t)idx;
  const double* mq = p.memo_q + (int64_t)idx * p.UQ;
  const int snp = __ldg(p.ent_snp + e);
  const double* z = reinterpret_cast<const double*>(p.gprow + (int64_t)snp * p.gp_stride);
  const double z0 = __ldg(z), z1 = __ldg(z + 1), z2 = __ldg(z + 2);
  // the same expressions as entq_emit
  const double g0 = __ldg(mq), g1 = __ldg(mq + 1), g2 = __ldg(mq + 2);
  const double s0 = fma(g2, z2, fma(g1, z1, g0 * z0));
  double s1 = 1.0;
  if (p.sep0) {
    const double f0 = __ldg(mq + 3), f1 = __ldg(mq + 4), f2 = __ldg(mq + 5);
    s1 = fma(f2, z2, fma(f1, z1, f0 * z0)) * ((z0 + z1) + z2);
  }
  double* sc = p.ents + e * (2 + 2 * p.nd);
  sc[0] = s0;
  sc[1] = s1;
  const int64_t pos = e + (int64_t)__ldg(p.perm + e);
  const bool refread = nvr == 1 && (rd1 >> 8) == 0;
  p.snp_perm[pos] = 2 * snp + (refread ? 0 : 1);          // row of the dosage table
  const float4* mh = reinterpret_cast<const float4*>(p.memo_fh + (int64_t)idx * FHEAD);
  float4* fh = reinterpret_cast<float4*>(p.entf_head + pos * FHEAD);
  fh[0] = __ldg(mh);
  fh[1] = __ldg(mh + 1);
  for (int d = 0; d < p.nd; ++d) {
    const double* Q = mq + QHEAD + QPLANE * d;
    const double t0 = fma(__ldg(Q + 2), z2, fma(__ldg(Q + 1), z1, __ldg(Q + 0) * z0));
    const double t1 = fma(__ldg(Q + 5), z2, fma(__ldg(Q + 4), z1, __ldg(Q + 3) * z0));
    const double t2 = fma(__ldg(Q + 8), z2, fma(__ldg(Q + 7), z1, __ldg(Q + 6) * z0));
    sc[2 + d] = fma(z2, t2, fma(z1, t1, z0 * t0));
    {   // the scale of the affine form, the same expression as entq_emit
      const double q = refread ? __ldg(Q + 8) : __ldg(Q + 0);
      const double cl = 0.5 * (refread ? __ldg(Q + 2) - __ldg(Q + 8) : __ldg(Q + 6) - __ldg(Q + 0));
      const double cm = 0.5 * (refread ? __ldg(Q + 6) - __ldg(Q + 8) : __ldg(Q + 2) - __ldg(Q + 0));
      sc[2 + p.nd + d] = q + 2.0 * (fmax(cl, 0.0) + fmax(cm, 0.0));
    }
    const float4* mp = reinterpret_cast<const float4*>(p.memo_fp + ((int64_t)d * MEMO_ROWS + idx) * FPLANE);
    float4* fp = reinterpret_cast<float4*>(p.entf_plane + d * p.entf_pstride + pos * FPLANE);
    fp[0] = __ldg(mp);
    fp[1] = __ldg(mp + 1);
  }
}

int launch_entq_aff(const EntqParams& p, void* stream) {
  const int64_t n = p.E - p.E0;
  if (n <= 0) return 0;
  entq_aff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

int launch_entq(const EntqParams& p, void* stream) {
  const int64_t n = p.E - p.E0;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  auto grid = [&](int lpe) { return (unsigned)((n * lpe + 255) / 256); };
  const unsigned g1 = (unsigned)((n + 127) / 128);
  if (p.U <= 5) entq1_kernel<5><<<g1, 128, 0, st>>>(p);
  else if (p.U <= 8) entq1_kernel<8><<<g1, 128, 0, st>>>(p);
  else if (p.U <= 16) entq_kernel<16, 1><<<grid(16), 256, 0, st>>>(p);
  else if (p.U <= 32) entq_kernel<32, 1><<<grid(32), 256, 0, st>>>(p);
  else if (p.U <= 64) entq_kernel<32, 2><<<grid(32), 256, 0, st>>>(p);
  else entq_kernel<32, 4><<<grid(32), 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}

int set_err_table(const double* err256, const double* err3_256) {
  cudaError_t e = cudaMemcpyToSymbol(c_err, err256, 256 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_err3, err3_256, 256 * sizeof(double));
  return (int)e;
}

// ---------------------------------------------------------------------------------------------
// cell_scalars: llk0s (:459) and llks00[n] (:700-709) of a cell = sum over its entries of the log of the per-entry
// factors entq wrote.  One warp per cell; running fp64 products with the exponent peeled every 16 factors, one log
// per lane and scalar.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cell_scalars_kernel(const __grid_constant__ CellScalarParams p, int64_t nactive) {
  const int lane = threadIdx.x & 31;
  const int64_t aidx = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (aidx >= nactive) return;
  const int cell = p.cell_order[aidx];
  const int64_t e0 = p.cell_ptr[cell], e1 = p.cell_ptr[cell + 1];
  const int ns = 2 + 2 * p.nd;             // even: the rows are read as double2 pairs, four columns (32 B) per pass
  for (int s0 = 0; s0 < ns; s0 += 4) {
    const bool two = s0 + 2 < ns;          // the last pass of a row with ns % 4 == 2 has one pair only
    double prod[4] = {1.0, 1.0, 1.0, 1.0};
    int es[4] = {0, 0, 0, 0}, nfl = 0, cnt = 0;
#pragma unroll 4
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const double2* row = reinterpret_cast<const double2*>(p.ents + e * ns + s0);
      const double2 a = __ldg(row);
      const double2 b = two ? __ldg(row + 1) : make_double2(1.0, 1.0);
      prod[0] *= a.x; prod[1] *= a.y; prod[2] *= b.x; prod[3] *= b.y;
      if (++cnt == 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) peel(prod[i], es[i]);
        ++nfl;
        cnt = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = s0 + i;
      double v = unpeel_log(prod[i], es[i], nfl);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane != 0 || s >= ns) continue;
      if (s == 0) { if (p.want_single) p.p1_0[cell] = v; }
      else if (s == 1) { if (p.sep0) p.l00[(int64_t)cell * p.nA] = v; }
      else if (s < 2 + p.nd) p.l00[(int64_t)cell * p.nA + p.plane_n[s - 2]] = v;
      else if (p.lm) p.lm[(int64_t)cell * p.nA + p.plane_n[s - 2 - p.nd]] = v;
    }
  }
}

int launch_cell_scalars(const CellScalarParams& p, int64_t nactive, void* stream) {
  if (nactive == 0) return 0;
  cell_scalars_kernel<<<(unsigned)((nactive + 3) / 4), 128, 0, (cudaStream_t)stream>>>(p, nactive);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// pair scan: persistent kernels.  Two variants share the loader and transform roles:
//   pair_scan_kernel      fp64 hypotheses on the FP64 tensor path (DMMA), two task pipelines per CTA
//   pair_scan_f32_kernel  fp32 hypotheses (3 FFMA + 1 FMUL per hypothesis and entry), one pipeline per CTA
// ---------------------------------------------------------------------------------------------
namespace {
// raw ring stage (what the TMA unit writes).  A GP row in HBM is [gp0: 3 doubles + pad = 8 floats][nvp x 3 floats]
// (gprow_kernel), so one tile::gather4 of 4 SNP rows x 200 floats brings gp0 and 64 samples of each row: the j side
// is [entry][800 B], samples from byte 32 (gp0 in front is valid for tasks with j0 = 0).  The k side of a
// rectangular task is [entry][416 B] (8 + 96 floats).  The records of the NB consecutive entries arrive as one tiled
// 2-D copy [entry][128 B] when there is a single dense plane, else as heads [entry][48 B] + planes [entry][80 B]
// (ScanParams::q_hp / q_po / q_pp give the pitches).
constexpr uint32_t GPFX = 8;                        // floats in front of the samples of a GP row
constexpr uint32_t GROWJ = (GPFX + BLK * 3) * 4;    // 800 B: smem pitch of a j-side row
constexpr uint32_t GROWK = (GPFX + BLK_RECT_K * 3) * 4;   // 416 B: k-side row of a rectangular task
constexpr uint32_t OFF_GPJ = 0;
constexpr uint32_t OFF_GPK = OFF_GPJ + NB * GROWJ;
constexpr uint32_t OFF_Q = OFF_GPK + NB * GROWK;
constexpr uint32_t RAW_SLOT = OFF_Q + NB * (QHEAD + QPLANE) * 8;     // 10752 B
static_assert((4 * GROWJ) % 128 == 0 && (4 * GROWK) % 128 == 0 && OFF_GPK % 128 == 0 && OFF_Q % 128 == 0,
              "TMA alignment");
static_assert(BLK_RECT_K == 32 && BLK == 64 && NB == 8, "gather geometry");
// fp64 processed ring: per entry and side a 64 x 4 matrix (sample x genotype, 4th column zero) stored as 8 blocks
// of 8 samples; inside a block the layout is column-major with the column index XOR-swizzled by the block parity,
// so that both the transform's stores (one column, 32 consecutive samples) and the consumers' DMMA fragment
// loads (lane -> row lane/4, column lane%4 of one block) are bank-conflict free.
constexpr uint32_t SIDE_DBL = BLK * 4;                               // doubles per entry and side (256)
constexpr uint32_t PROC_SLOT = NB * 2 * SIDE_DBL * 8;                // 32768 B
constexpr uint32_t SLOT_CTRL = 1024;                                 // barriers + epilogue scratch
constexpr uint32_t SLOT_BYTES = SLOT_CTRL + S_RAW * RAW_SLOT + S_PROC * PROC_SLOT;
static_assert(RAW_SLOT % 128 == 0 && PROC_SLOT % 128 == 0 && SLOT_BYTES % 128 == 0, "alignment");
// fp32 processed ring: per side 16 blocks of 4 samples; a block holds, for each of the NB entries of the stage,
// the 12 floats (4 samples x 3 genotypes) a consumer thread fetches with three LDS.128.  Entry stride 48 B makes
// the 8 entry lanes of a quarter warp conflict free; block stride 432 B (= 12 words mod 32) makes the transform's
// stores (lane = sample, 3 words apart) conflict free as well.
constexpr uint32_t F_BLK4_FLT = NB * 12 + 12;                        // 108 floats
constexpr uint32_t F_SIDE_FLT = (BLK / 4) * F_BLK4_FLT;              // 1728 floats
constexpr uint32_t F_PROC_SLOT = 2 * F_SIDE_FLT * 4;                 // 13824 B
static_assert(F_PROC_SLOT % 128 == 0 && NB == 8, "fp32 scan geometry");
static_assert(2 * 12 + 4 * F_S_PROC <= 64 && 2 * S_RAW + 2 * S_PROC <= 64, "barriers must fit below the scratch area");

constexpr uint32_t SCRATCH_OFF = 512;   // barriers live below
struct Scratch {                     // at slot offset SCRATCH_OFF
  double red[2][16];
  double red2[16];
  int ncand;
  uint32_t cand[MAX_CAND];
};
static_assert(sizeof(Scratch) + SCRATCH_OFF <= SLOT_CTRL, "scratch");

struct TaskInfo {
  int64_t aidx;
  int tti, cell, nst;
  int64_t e0, e1;
};

__device__ __forceinline__ bool next_task(const ScanParams& p, int r, int u, int NS, TaskInfo* t) {
  const int64_t g = (int64_t)r * NS + ((r & 1) ? (NS - 1 - u) : u);   // snake order over size-sorted tasks
  if (g >= p.ntasks) return false;
  t->aidx = g / p.T;
  t->tti = (int)(g - t->aidx * p.T);
  t->cell = __ldg(p.cell_order + t->aidx);
  t->e0 = __ldg(p.cell_ptr + t->cell);
  t->e1 = __ldg(p.cell_ptr + t->cell + 1);
  t->nst = (int)((t->e1 - t->e0 + NB - 1) / NB);
  return true;
}

template <int NTHR>
__device__ __forceinline__ void consumer_bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "n"(NTHR) : "memory");
}

// one stage of the fp64 transform for lane `tl` (sample slot on each side).  No lane-dependent branches: lanes beyond
// the task's sample count work on padding (their ring slots are never read), so the compiler can batch the
// shared-memory loads of several entries.
struct QLay { uint32_t hp, po, pp; };   // byte pitches of the record heads / planes in a raw stage

template <bool SAME, bool FULL, bool EXTRAS>
__device__ __forceinline__ void transform_stage(const unsigned char* __restrict__ rs, void* __restrict__ pav, QLay ql,
                                                int tl, int nvalid, double& PA, double& PB, double& P1) {
  const int blk = tl >> 3, sw = blk & 1;
  double* ad = reinterpret_cast<double*>(pav) + blk * 32 + (tl & 7);
  double* wd = ad + NB * SIDE_DBL;
  const int c0 = (0 ^ sw) * 8, c1 = (1 ^ sw) * 8, c2 = (2 ^ sw) * 8;
  const float* gj = reinterpret_cast<const float*>(rs + OFF_GPJ) + GPFX + tl * 3;
  const float* gk = SAME ? gj : reinterpret_cast<const float*>(rs + OFF_GPK) + GPFX + (tl & 31) * 3;
  constexpr int GK_STRIDE = SAME ? GROWJ / 4 : GROWK / 4, GJ_STRIDE = GROWJ / 4;
  const unsigned char* qhb = rs + OFF_Q;                 // head of entry e at qhb + e * ql.hp
  const unsigned char* qpb = rs + OFF_Q + ql.po;         // plane of entry e at qpb + e * ql.pp
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    if (FULL || e < nvalid) {
      const float f0 = gk[e * GK_STRIDE], f1 = gk[e * GK_STRIDE + 1], f2 = gk[e * GK_STRIDE + 2];
      const double g0 = (double)f0, g1 = (double)f1, g2 = (double)f2;
      const double2* qp = reinterpret_cast<const double2*>(qpb + e * ql.pp);
      const double2 qa = qp[0], qb = qp[1], qc = qp[2], qd = qp[3];
      const double q22 = reinterpret_cast<const double*>(qp + 4)[0];
      // Q'[l][m] row-major: qa = (00, 01), qb = (02, 10), qc = (11, 12), qd = (20, 21), q22
      const double w0 = fma(qb.x, g2, fma(qa.y, g1, qa.x * g0));
      const double w1 = fma(qc.y, g2, fma(qc.x, g1, qb.y * g0));
      const double w2 = fma(q22, g2, fma(qd.y, g1, qd.x * g0));
      if (EXTRAS) {
        const double2* qh = reinterpret_cast<const double2*>(qhb + e * ql.hp);
        const double2 h0 = qh[0], h1 = qh[1], h2 = qh[2];   // GL0 GL1 | GL2 f0 | f1 f2
        PA *= fma(h2.y, g2, fma(h2.x, g1, h1.y * g0));
        PB *= (g0 + g1) + g2;
        P1 *= fma(h1.x, g2, fma(h0.y, g1, h0.x * g0));
      }
      wd[e * SIDE_DBL + c0] = w0;
      wd[e * SIDE_DBL + c1] = w1;
      wd[e * SIDE_DBL + c2] = w2;
      if (SAME) {
        ad[e * SIDE_DBL + c0] = g0; ad[e * SIDE_DBL + c1] = g1; ad[e * SIDE_DBL + c2] = g2;
      } else {
        ad[e * SIDE_DBL + c0] = (double)gj[e * (GJ_STRIDE)];
        ad[e * SIDE_DBL + c1] = (double)gj[e * (GJ_STRIDE) + 1];
        ad[e * SIDE_DBL + c2] = (double)gj[e * (GJ_STRIDE) + 2];
      }
    } else {   // identity entry: S = 1 for every hypothesis
      wd[e * SIDE_DBL + c0] = 1.0; wd[e * SIDE_DBL + c1] = 0.0; wd[e * SIDE_DBL + c2] = 0.0;
      ad[e * SIDE_DBL + c0] = 1.0; ad[e * SIDE_DBL + c1] = 0.0; ad[e * SIDE_DBL + c2] = 0.0;
    }
  }
}

template <bool SAME, bool EXTRAS>
__device__ __forceinline__ void transform_dispatch(const unsigned char* rs, void* pa, QLay ql, int tl, int nvalid,
                                                   double& PA, double& PB, double& P1) {
  if (nvalid == NB) transform_stage<SAME, true, EXTRAS>(rs, pa, ql, tl, nvalid, PA, PB, P1);
  else transform_stage<SAME, false, EXTRAS>(rs, pa, ql, tl, nvalid, PA, PB, P1);
}

// ring pointers of one task pipeline
struct Pipe {
  uint64_t *raw_full, *raw_empty, *proc_full, *proc_empty;
  unsigned char *raw_base, *proc_base;
};

// =============================== loader warp =====================================================
// Per stage of NB entries the warp issues (one elected lane, straight-line code): two tile::gather4 copies for the j
// side (4 SNP rows x [gp0 | 64 samples] each), two more for the k side of a rectangular task, and one tiled 2-D copy
// of the NB consecutive entry records (two when there are several dense planes).  Every stage moves the same number
// of bytes: entry slots beyond the cell's last entry gather row 0 / the next cell's records (rows beyond the table are
// zero filled by the TMA unit) and the transform ignores them.  SNP ids are fetched one stage ahead.
template <int SRAW>
__device__ __forceinline__ void loader_role(const ScanParams& p, const ScanMaps& tm, const Pipe& pp, int lane, int u,
                                            int NS) {
  const uint32_t raw0 = smem_u32(pp.raw_base), full0 = smem_u32(pp.raw_full);
  const bool one_q = p.q_hp == (QHEAD + QPLANE) * 8;
  uint32_t git = 0;
  TaskInfo ti;
  for (int r = 0; next_task(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const bool need_head = tt.flags & (TF_EXTRAS | TF_SCALARS);
    const int colj = tt.j0 * 3, colk = tt.k0 * 3, qoff = tt.qoff;
    const uint32_t bytes = (uint32_t)NB * (GROWJ + (same ? 0u : GROWK)) +
                           (one_q ? (uint32_t)NB * (QHEAD + QPLANE) * 8u
                                  : (uint32_t)NB * ((need_head ? (uint32_t)QHEAD * 8u : 0u) + (uint32_t)QPLANE * 8u));
    // SNP ids are fetched 32 entries (4 stages) at a time, one block ahead: the global-load latency of the ids must
    // not sit on the per-stage critical path
    int snp_cur = (ti.e0 + lane < ti.e1) ? __ldg(p.ent_snp + ti.e0 + lane) : 0;
    int snp_nxt = (ti.e0 + 32 + lane < ti.e1) ? __ldg(p.ent_snp + ti.e0 + 32 + lane) : 0;
    for (int it = 0; it < ti.nst; ++it, ++git) {
      const uint32_t rslot = git % SRAW;
      const uint32_t ph = (git / SRAW) & 1u;
      const int64_t eb = ti.e0 + (int64_t)it * NB;
      if (it && (it & 3) == 0) {
        snp_cur = snp_nxt;
        snp_nxt = (eb + 32 + lane < ti.e1) ? __ldg(p.ent_snp + eb + 32 + lane) : 0;
      }
      int s[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) s[i] = __shfl_sync(0xffffffffu, snp_cur, (it & 3) * NB + i);
      mbar_wait(&pp.raw_empty[rslot], ph ^ 1u);
      const uint32_t rs = raw0 + rslot * RAW_SLOT, bar = full0 + rslot * 8u;
      mbar_expect_tx_elect(bar, bytes);
      tma_gather4_elect(rs + OFF_GPJ, &tm.gpj, colj, s[0], s[1], s[2], s[3], bar);
      tma_gather4_elect(rs + OFF_GPJ + 4 * GROWJ, &tm.gpj, colj, s[4], s[5], s[6], s[7], bar);
      if (!same) {
        tma_gather4_elect(rs + OFF_GPK, &tm.gpk, colk, s[0], s[1], s[2], s[3], bar);
        tma_gather4_elect(rs + OFF_GPK + 4 * GROWK, &tm.gpk, colk, s[4], s[5], s[6], s[7], bar);
      }
      if (one_q) {
        tma_tile2d_elect(rs + OFF_Q, &tm.q, 0, (int)eb, bar);
      } else {
        if (need_head) tma_tile2d_elect(rs + OFF_Q, &tm.qh, 0, (int)eb, bar);
        tma_tile2d_elect(rs + OFF_Q + NB * QHEAD * 8, &tm.qp, qoff, (int)eb, bar);
      }
    }
  }
}

// =============================== transform warps =================================================
// one lane per sample: GP row -> (a, w) vectors in the processed ring; the lanes also carry the per-sample
// "extras" (separable-plane factor, row-sum factor, pass-1 llks) and, in NB lanes, the scalars llk0s / llks00
template <int SRAW, int SPROC, uint32_t PSLOT_BYTES>
__device__ __forceinline__ void transform_role(const ScanParams& p, const Pipe& pp, int rw, int lane, int u, int NS) {
  const int tl = rw * 32 + lane;        // sample slot on each side
  const QLay ql = {(uint32_t)p.q_hp, (uint32_t)p.q_po, (uint32_t)p.q_pp};
  uint32_t git = 0;
  TaskInfo ti;
  for (int r = 0; next_task(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const bool do_k = tl < tt.nk;
    const bool k_valid = (tt.k0 + tl) < p.nv;
    const bool extras = (tt.flags & TF_EXTRAS) && do_k && k_valid;
    const bool scal = (tt.flags & (TF_SCALARS | TF_L00)) && rw == 0 && lane < NB;
    double PA = 1.0, PB = 1.0, P1 = 1.0;            // running products (extras)
    int eA = 0, eB = 0, e1x = 0;
    double PS0 = 1.0, PS1 = 1.0, PS2 = 1.0;         // scalars: llk0s, llks00[sep], llks00[n]; lane = entry slot
    int eS0 = 0, eS1 = 0, eS2 = 0;
    int nflush = 0;
    for (int it = 0; it < ti.nst; ++it, ++git) {
      const int rslot = git % SRAW, pslot = git % SPROC;
      const uint32_t rph = (git / SRAW) & 1u, pph = (git / SPROC) & 1u;
      const int nvalid = (int)min((int64_t)NB, ti.e1 - (ti.e0 + (int64_t)it * NB));
      mbar_wait(&pp.raw_full[rslot], rph);
      mbar_wait(&pp.proc_empty[pslot], pph ^ 1u);
      const unsigned char* rs = pp.raw_base + (size_t)rslot * RAW_SLOT;
      void* pa = pp.proc_base + (size_t)pslot * PSLOT_BYTES;
      if (same) {
        if (tt.flags & TF_EXTRAS) transform_dispatch<true, true>(rs, pa, ql, tl, nvalid, PA, PB, P1);
        else transform_dispatch<true, false>(rs, pa, ql, tl, nvalid, PA, PB, P1);
      } else {
        if (tt.flags & TF_EXTRAS) transform_dispatch<false, true>(rs, pa, ql, tl, nvalid, PA, PB, P1);
        else transform_dispatch<false, false>(rs, pa, ql, tl, nvalid, PA, PB, P1);
      }
      if (scal && lane < nvalid) {
        // per-entry factors precomputed by entq_kernel
        const double* hq = reinterpret_cast<const double*>(rs + OFF_Q + lane * ql.hp);
        const double* q = reinterpret_cast<const double*>(rs + OFF_Q + ql.po + lane * ql.pp);
        if (tt.flags & TF_SCALARS) {
          if (p.want_single) PS0 *= hq[6];                                                          // :459
          if (p.sep0) PS1 *= hq[7];
        }
        if (tt.flags & TF_L00) PS2 *= q[9];                                                         // :700-709
      }
      if ((it % FLUSH_STAGES) == FLUSH_STAGES - 1) {
        peel(PA, eA); peel(PB, eB); peel(P1, e1x);
        peel(PS0, eS0); peel(PS1, eS1); peel(PS2, eS2);
        ++nflush;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&pp.proc_full[pslot]);
        mbar_arrive(&pp.raw_empty[rslot]);
      }
    }
    if (extras) {
      const int k = tt.k0 + tl;
      if (p.sep0) p.singA[(int64_t)ti.cell * p.nv + k] = unpeel_log(PA, eA, nflush);
      p.lb[(int64_t)ti.cell * p.nv + k] = unpeel_log(PB, eB, nflush);
      if (p.want_single) p.p1[(int64_t)ti.cell * p.nv + k] = unpeel_log(P1, e1x, nflush);
    }
    if ((tt.flags & (TF_SCALARS | TF_L00)) && rw == 0) {
      // each of the first NB lanes holds the product over "its" entries of every stage
      double s0 = scal ? unpeel_log(PS0, eS0, nflush) : 0.0;
      double s1 = scal ? unpeel_log(PS1, eS1, nflush) : 0.0;
      double s2 = scal ? unpeel_log(PS2, eS2, nflush) : 0.0;
#pragma unroll
      for (int o = NB / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) {
        if (tt.flags & TF_SCALARS) {
          if (p.want_single) p.p1_0[ti.cell] = s0;
          if (p.sep0) p.l00[(int64_t)ti.cell * p.nA] = s1;
        }
        if (tt.flags & TF_L00) p.l00[(int64_t)ti.cell * p.nA + tt.n] = s2;
      }
    }
  }
}

// =============================== fp32 pair scan: loader and transform ==============================
// Raw stage of the fp32 scan: the gathered GP rows as in the fp64 scan, then the fp32 record heads and the task's
// fp32 plane of the NB consecutive entries (two contiguous 1-D bulk copies: the tables are entry-major).
// The k-side rows exist only when the task list has rectangular tasks (nv > 64 or a non-symmetric plane); without
// them a stage is 7040 B instead of 10368 B and the same shared memory holds 12 raw stages instead of 8
// (ScanParams::f_sraw / f_rawslot / f_off_fh, set by launch_pair_scan).
//
// The fp32 tables are class-sorted inside a cell (classify_kernel / entq): the na affine entries first, then the
// general ones; a task walks ceil(na / NB) affine stages and then ceil((n - na) / NB) general stages, every stage
// holding entries of one class only.
constexpr uint32_t F_OFF_GPJ = 0;
constexpr uint32_t F_OFF_GPK = F_OFF_GPJ + NB * GROWJ;
constexpr uint32_t F_REC_BYTES = NB * (FHEAD + FPLANE) * 4;                    // 640
constexpr uint32_t F_RAW_SLOT_RECT = F_OFF_GPK + NB * GROWK + F_REC_BYTES;     // 10368 B
constexpr uint32_t F_RAW_SLOT_TRI = F_OFF_GPK + F_REC_BYTES;                   // 7040 B
#ifndef DMX_S_RAW_TRI
#define DMX_S_RAW_TRI 12
#endif
constexpr uint32_t F_S_RAW_RECT = 8, F_S_RAW_TRI = DMX_S_RAW_TRI;
constexpr uint32_t F_RAW_BYTES = 12 * F_RAW_SLOT_TRI;                          // 84480 B
static_assert(F_OFF_GPK % 128 == 0 && F_RAW_SLOT_RECT % 128 == 0 && F_RAW_SLOT_TRI % 128 == 0 &&
              F_S_RAW_RECT * F_RAW_SLOT_RECT <= F_RAW_BYTES && F_S_RAW_TRI <= 12, "raw ring layout");
constexpr uint32_t F_SMEM_NOTS = SLOT_CTRL + F_RAW_BYTES + F_S_PROC * F_PROC_SLOT;   // + TScratch
constexpr uint32_t A_SIDE_FLT = (BLK / 4) * A_BLK4_FLT;                        // floats per side of an affine stage (576)
static_assert(2 * A_SIDE_FLT * 4 <= F_PROC_SLOT, "an affine stage fits a processed-ring slot");

struct TaskInfoF {
  int64_t aidx;
  int tti, cell, nst;
  int64_t e0, e1;
  int na, nstA;      // affine entries / stages of the cell
};

__device__ __forceinline__ bool next_task_f32(const ScanParams& p, int r, int u, int NS, TaskInfoF* t) {
  const int64_t g = (int64_t)r * NS + ((r & 1) ? (NS - 1 - u) : u);   // snake order over size-sorted tasks
  if (g >= p.ntasks) return false;
  t->aidx = g / p.T;
  t->tti = (int)(g - t->aidx * p.T);
  t->cell = __ldg(p.cell_order + t->aidx);
  t->e0 = __ldg(p.cell_ptr + t->cell);
  t->e1 = __ldg(p.cell_ptr + t->cell + 1);
  t->na = __ldg(p.cell_na + t->cell);
  t->nstA = (t->na + NB - 1) / NB;
  t->nst = t->nstA + ((int)(t->e1 - t->e0) - t->na + NB - 1) / NB;
  return true;
}

// entries of stage `it`: first position in the class-sorted tables, number of real entries, class
__device__ __forceinline__ void stage_span(const TaskInfoF& t, int it, int64_t* eb, int* nvalid, bool* affine) {
  if (it < t.nstA) {
    *affine = true;
    *eb = t.e0 + (int64_t)it * NB;
    *nvalid = min(NB, t.na - it * NB);
  } else {
    const int g = it - t.nstA;
    *affine = false;
    *eb = t.e0 + t.na + (int64_t)g * NB;
    *nvalid = min(NB, (int)(t.e1 - t.e0) - t.na - g * NB);
  }
}

__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// F_NLOAD loader warps share the stages round-robin (stage g belongs to loader g % F_NLOAD): one warp's issue path
// (barrier wait, eight shuffles, expect_tx, two to four gather4 and two bulk copies, all from one elected lane) takes
// about 900 cycles per stage even on an otherwise idle SM, more than the consumers need for the stage's arithmetic.
__device__ __forceinline__ void loader_role_f32(const ScanParams& p, const ScanMaps& tm, const Pipe& pp, int lane, int u,
                                                int NS, int lid) {
  const uint32_t raw0 = smem_u32(pp.raw_base), full0 = smem_u32(pp.raw_full);
  const uint32_t sraw = (uint32_t)p.f_sraw, rawslot = (uint32_t)p.f_rawslot, off_fh = (uint32_t)p.f_off_fh;
  uint32_t gbase = 0;   // stages of the tasks before this one
  TaskInfoF ti;
  PROF_DECL
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const int colj = tt.j0 * 3, colk = tt.k0 * 3;
    const float* fplane = p.entf_plane + (int64_t)((tt.qoff - QHEAD) / QPLANE) * p.entf_pstride;
    const bool nogp = (p.dbg & 32) != 0;   // timing experiment: no GP gathers
    const uint32_t bytes = (nogp ? 0u : (uint32_t)NB * (GROWJ + (same ? 0u : GROWK))) + F_REC_BYTES;
    int it = (int)((F_NLOAD + lid - gbase % F_NLOAD) % F_NLOAD);   // first stage of this task that is ours
    // SNP ids of a stage are fetched by lanes 0..7 one own stage ahead
    auto fetch = [&](int st) -> int {
      int64_t eb;
      int nvalid;
      bool aff;
      stage_span(ti, st, &eb, &nvalid, &aff);
      return lane < nvalid ? __ldg(p.snp_perm + eb + lane) : 0;
    };
    int snp_nxt = it < ti.nst ? fetch(it) : 0;
    for (; it < ti.nst; it += F_NLOAD) {
      const uint32_t git = gbase + (uint32_t)it;
      const uint32_t rslot = git % sraw;
      const uint32_t ph = (git / sraw) & 1u;
      int64_t eb;
      int nvalid;
      bool aff;
      stage_span(ti, it, &eb, &nvalid, &aff);
      const int snp_cur = aff ? snp_nxt >> 1 : snp_nxt;   // affine entries carry the row of the dosage table
      if (it + F_NLOAD < ti.nst) snp_nxt = fetch(it + F_NLOAD);
      int s[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) s[i] = __shfl_sync(0xffffffffu, snp_cur, i);
      PROF_WAIT(1, mbar_wait_relaxed(&pp.raw_empty[rslot], ph ^ 1u));
      const uint32_t rs = raw0 + rslot * rawslot, bar = full0 + rslot * 8u;
      mbar_expect_tx_elect(bar, bytes);
      if (!nogp) {
        tma_gather4_elect(rs + F_OFF_GPJ, &tm.gpj, colj, s[0], s[1], s[2], s[3], bar);
        tma_gather4_elect(rs + F_OFF_GPJ + 4 * GROWJ, &tm.gpj, colj, s[4], s[5], s[6], s[7], bar);
        if (!same) {
          tma_gather4_elect(rs + F_OFF_GPK, &tm.gpk, colk, s[0], s[1], s[2], s[3], bar);
          tma_gather4_elect(rs + F_OFF_GPK + 4 * GROWK, &tm.gpk, colk, s[4], s[5], s[6], s[7], bar);
        }
      }
      // entry slots beyond the stage's last entry read the records that follow (the tables have NB records of slack)
      bulk_g2s_elect(rs + off_fh, p.entf_head + eb * FHEAD, NB * FHEAD * 4, bar);
      bulk_g2s_elect(rs + off_fh + NB * FHEAD * 4, fplane + eb * FPLANE, NB * FPLANE * 4, bar);
    }
    gbase += (uint32_t)ti.nst;
  }
  PROF_END(0, lane)
}

// F_NTW = 4 transform warps take whole stages.  A lane serves one 4-sample block q = lane % 16 of each side for the
// four entries e = lane / 16 + {0, 2, 4, 6} of its stage, so everything moves as 16-byte vectors: the 12 floats of its
// samples' GP rows (3 LDS.128, 48-byte lane stride: conflict free), the entry's fp32 record, and the operand blocks
// of the processed ring.  Every genotype vector is normalised to sum 1 here (a_hat = g / s, s = g0 + g1 + g2 ~ 1 up to
// float rounding), and the product of the sums s is carried as a per-sample quantity of its own (lb); the consumers'
// S'[j][k] is the likelihood over s_j s_k and the epilogues add lb[j] + lb[k] back.
//   general stage  (a_hat, w) blocks: w_k[l] = sum_m Q'[l][m] a_hat_k[m]; consumers form S' = a_hat_j . w_k
//   affine stage   X_j = h + cj D_j, Y_k = h + ck D_k, D = dosage of the read's allele under a_hat; consumers form
//                  S' = X_j + Y_k
// The per-sample extras (singlet factor of the separable plane, pass 1, pairs with sample 0) and the sums s leave the
// transform warps as per-stage products (four entries per lane) in a products ring; the accumulator warp multiplies
// them up in stage order, so the result does not depend on which transform warp happened to take which stage.
constexpr int NQ = 5;   // per-sample quantities: PA (separable plane), PB (sums s, k side), P1 (pass 1), P0 (pairs with sample 0),
                        // PJ (sums s, j side of a rectangular task)
struct TScratch {
  float prod[F_S_PROC][2][NQ][BLK];  // products ring: [stage slot][entry half][quantity][sample], the stage's
                                     // factors of every per-sample running product (transform -> accumulator warp)
  double lb[2][2][BLK];              // [task parity][j side, k side][sample]: sum log s of the finished task
                                     // (accumulator -> consumer epilogue)
  unsigned next_stage;               // stage counter of the CTA: transform warps take stages as they become free
};
constexpr uint32_t F_TS_BYTES = (sizeof(TScratch) + 127) / 128 * 128;

__device__ __forceinline__ void store_neutral_quad(float* dst) {   // four (1, 0, 0) vectors
  float4* d = reinterpret_cast<float4*>(dst);
  d[0] = make_float4(1.f, 0.f, 0.f, 1.f);
  d[1] = make_float4(0.f, 0.f, 1.f, 0.f);
  d[2] = make_float4(0.f, 1.f, 0.f, 0.f);
}

// 1 / s for a row sum (0 for the all-zero padding samples)
// (one MUFU.RCP, ~1 ulp; an all-zero padding sample gives 0 * 1e30 = 0 downstream, never inf or NaN)
__device__ __forceinline__ float rcp_sum(float s) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s, 1e-30f)));
  return r;
}

// normalise the four genotype vectors of a block in place; sums in s[4]
__device__ __forceinline__ void normalise_quad(float (&g)[12], float (&s)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[i] = (g[3 * i] + g[3 * i + 1]) + g[3 * i + 2];
    const float rs = rcp_sum(s[i]);
    g[3 * i] *= rs; g[3 * i + 1] *= rs; g[3 * i + 2] *= rs;
  }
}
__device__ __forceinline__ void load_quad(const unsigned char* src, float (&g)[12]) {
  const float4* p4 = reinterpret_cast<const float4*>(src);
  const float4 x = p4[0], y = p4[1], z = p4[2];
  g[0] = x.x; g[1] = x.y; g[2] = x.z; g[3] = x.w; g[4] = y.x; g[5] = y.y; g[6] = y.z; g[7] = y.w;
  g[8] = z.x; g[9] = z.y; g[10] = z.z; g[11] = z.w;
}
__device__ __forceinline__ void store_quad(float* dst, const float (&g)[12]) {
  float4* d = reinterpret_cast<float4*>(dst);
  d[0] = make_float4(g[0], g[1], g[2], g[3]);
  d[1] = make_float4(g[4], g[5], g[6], g[7]);
  d[2] = make_float4(g[8], g[9], g[10], g[11]);
}

// per-stage factors of the per-sample running products, four samples per lane
struct Fac { float PA[4], PB[4], P1[4], P0[4], PJ[4]; };

// ---- general stage ---------------------------------------------------------------------------------------------------
// ROW0: 0 none, 1 symmetric plane (P0 of k-sample x *= a_0 . w_x), 2 full plane (P0 of j-sample x *= a_x . w_0)
template <bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void transform_quad_f32(const unsigned char* __restrict__ rs, uint32_t off_fh,
                                                   float* __restrict__ pf, int e, int q, bool jq, bool kq, Fac& F) {
  const uint32_t off_fp = off_fh + NB * FHEAD * 4;
  float* af = pf + q * F_BLK4_FLT + e * 12;
  float* wf = af + F_SIDE_FLT;
  float g[12], sk[4];
  const float4* qp = reinterpret_cast<const float4*>(rs + off_fp + e * (FPLANE * 4));
  if (SAME || q < BLK_RECT_K / 4) {
    if (kq) {
      load_quad(rs + (SAME ? F_OFF_GPJ + e * GROWJ : F_OFF_GPK + e * GROWK) + GPFX * 4 + q * 48, g);
      normalise_quad(g, sk);
      const float4 qa = qp[0], qb = qp[1];       // Q'[l][m] row-major: (00 01 02 10) (11 12 20 21) (22 . . .)
      const float q22 = reinterpret_cast<const float*>(qp + 2)[0];
      float w[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f0 = g[3 * i], f1 = g[3 * i + 1], f2 = g[3 * i + 2];
        w[3 * i + 0] = fmaf(qa.z, f2, fmaf(qa.y, f1, qa.x * f0));
        w[3 * i + 1] = fmaf(qb.y, f2, fmaf(qb.x, f1, qa.w * f0));
        w[3 * i + 2] = fmaf(q22, f2, fmaf(qb.w, f1, qb.z * f0));
        F.PB[i] *= sk[i];
      }
      if (EXTRAS) {
        const float4* hp = reinterpret_cast<const float4*>(rs + off_fh + e * (FHEAD * 4));
        const float4 h0 = hp[0];                  // GL0 GL1 GL2 f0
        const float2 h1 = reinterpret_cast<const float2*>(hp + 1)[0];   // f1 f2
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f0 = g[3 * i], f1 = g[3 * i + 1], f2 = g[3 * i + 2];
          F.PA[i] *= fmaf(h1.y, f2, fmaf(h1.x, f1, h0.w * f0));
          F.P1[i] *= fmaf(h0.z, f2, fmaf(h0.y, f1, h0.x * f0));
        }
      }
      if (ROW0 == 1) {   // pairs (0, x): sample 0 is the first sample of the j-side row (j0 == 0)
        const float* a0 = reinterpret_cast<const float*>(rs + F_OFF_GPJ + e * GROWJ + GPFX * 4);
        const float r0 = rcp_sum((a0[0] + a0[1]) + a0[2]);
        const float a00 = a0[0] * r0, a01 = a0[1] * r0, a02 = a0[2] * r0;
#pragma unroll
        for (int i = 0; i < 4; ++i) F.P0[i] *= fmaf(a02, w[3 * i + 2], fmaf(a01, w[3 * i + 1], a00 * w[3 * i]));
      }
      store_quad(wf, w);
    } else {
      store_neutral_quad(wf);
    }
  }
  if (jq) {
    if (!SAME) {
      float a[12], sj[4];
      load_quad(rs + F_OFF_GPJ + e * GROWJ + GPFX * 4 + q * 48, a);
      normalise_quad(a, sj);
#pragma unroll
      for (int i = 0; i < 4; ++i) F.PJ[i] *= sj[i];
      store_quad(af, a);
      if (ROW0 == 2) {   // pairs (x, 0): sample 0 is the first sample of the k-side row (k0 == 0)
        const float* g0p = reinterpret_cast<const float*>(rs + F_OFF_GPK + e * GROWK + GPFX * 4);
        const float r0 = rcp_sum((g0p[0] + g0p[1]) + g0p[2]);
        const float f0 = g0p[0] * r0, f1 = g0p[1] * r0, f2 = g0p[2] * r0;
        const float4 qa = qp[0], qb = qp[1];
        const float q22 = reinterpret_cast<const float*>(qp + 2)[0];
        const float w0 = fmaf(qa.z, f2, fmaf(qa.y, f1, qa.x * f0));
        const float w1 = fmaf(qb.y, f2, fmaf(qb.x, f1, qa.w * f0));
        const float w2 = fmaf(q22, f2, fmaf(qb.w, f1, qb.z * f0));
#pragma unroll
        for (int i = 0; i < 4; ++i) F.P0[i] *= fmaf(a[3 * i + 2], w2, fmaf(a[3 * i + 1], w1, a[3 * i] * w0));
      }
    } else {
      store_quad(af, g);
    }
  } else {
    store_neutral_quad(af);
  }
}

// ---- affine stage ----------------------------------------------------------------------------------------------------
// dosage of the read's allele under the normalised genotype vector: ref read 2 a0 + a1, alt read a1 + 2 a2
__device__ __forceinline__ float dosage(float g0, float g1, float g2, bool refread, float* s_out) {
  const float s = (g0 + g1) + g2;
  *s_out = s;
  return (refread ? fmaf(2.0f, g0, g1) : fmaf(2.0f, g2, g1)) * rcp_sum(s);
}

template <bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void transform_quad_aff(const unsigned char* __restrict__ rs, uint32_t off_fh,
                                                   float* __restrict__ pf, int e, int q, bool jq, bool kq, Fac& F) {
  const uint32_t off_fp = off_fh + NB * FHEAD * 4;
  const float4* hp = reinterpret_cast<const float4*>(rs + off_fh + e * (FHEAD * 4));
  const float4 hd = hp[0];                                   // h1 c1 hs cs
  const bool refread = reinterpret_cast<const float*>(hp + 1)[0] != 0.f;
  const float4 pl = *reinterpret_cast<const float4*>(rs + off_fp + e * (FPLANE * 4));   // h cj ck
  float* xf = pf + q * A_BLK4_FLT + e * 4;                   // j side (the only side of a triangular task)
  float* yf = SAME ? xf : xf + A_SIDE_FLT;
  if (SAME || q < BLK_RECT_K / 4) {
    float y[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (kq) {
      float g[12];
      load_quad(rs + (SAME ? F_OFF_GPJ + e * GROWJ : F_OFF_GPK + e * GROWK) + GPFX * 4 + q * 48, g);
      float x0 = 0.f;
      if (ROW0 == 1) {   // pairs (0, x): sample 0 is the first sample of the j-side row (j0 == 0)
        const float* a0 = reinterpret_cast<const float*>(rs + F_OFF_GPJ + e * GROWJ + GPFX * 4);
        float s0;
        x0 = fmaf(pl.y, dosage(a0[0], a0[1], a0[2], refread, &s0), pl.x);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s;
        const float D = dosage(g[3 * i], g[3 * i + 1], g[3 * i + 2], refread, &s);
        y[i] = fmaf(pl.z, D, pl.x);
        F.PB[i] *= s;
        if (EXTRAS) {
          F.P1[i] *= fmaf(hd.y, D, hd.x);
          F.PA[i] *= fmaf(hd.w, D, hd.z);
        }
        if (ROW0 == 1) F.P0[i] *= x0 + y[i];
      }
    }
    *reinterpret_cast<float4*>(yf) = make_float4(y[0], y[1], y[2], y[3]);
  }
  if (!SAME) {
    float x[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (jq) {
      float a[12];
      load_quad(rs + F_OFF_GPJ + e * GROWJ + GPFX * 4 + q * 48, a);
      float y0 = 0.f;
      if (ROW0 == 2) {   // pairs (x, 0): sample 0 is the first sample of the k-side row (k0 == 0)
        const float* g0p = reinterpret_cast<const float*>(rs + F_OFF_GPK + e * GROWK + GPFX * 4);
        float s0;
        y0 = fmaf(pl.z, dosage(g0p[0], g0p[1], g0p[2], refread, &s0), pl.x);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s;
        x[i] = fmaf(pl.y, dosage(a[3 * i], a[3 * i + 1], a[3 * i + 2], refread, &s), pl.x);
        F.PJ[i] *= s;
        if (ROW0 == 2) F.P0[i] *= x[i] + y0;
      }
    }
    *reinterpret_cast<float4*>(xf) = make_float4(x[0], x[1], x[2], x[3]);
  }
}

// conditional peel of a running fp32 product: exponent into `es` (unbiased), mantissa back to [1, 2)
__device__ __forceinline__ void peelf_unbiased(float& x, int& es) {
  const uint32_t b = __float_as_uint(x);
  if (b >= 0x00800000u) {
    es += (int)(b >> 23) - 127;
    x = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
  }
}

// all four entries of a lane in one stage: e = ebase + 2 i.  Rolled on purpose: the 16 warps of the CTA run five
// different roles out of a 32 KB instruction cache, and an unrolled transform alone would take a third of it.
template <bool AFF, bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void transform_stage_f32(const unsigned char* __restrict__ rs, uint32_t off_fh,
                                                    float* __restrict__ pf, int ebase, int q, int nvalid, bool jq,
                                                    bool kq, Fac& F) {
#ifndef DMX_T_UNROLL
#define DMX_T_UNROLL 1
#endif
  constexpr int TU = DMX_T_UNROLL;   // entries in flight per lane (instruction-cache footprint against latency hiding)
#pragma unroll TU
  for (int i = 0; i < NB / 2; ++i) {
    const int e = ebase + 2 * i;
    if (AFF) {
      if (e < nvalid) {
        transform_quad_aff<SAME, EXTRAS, ROW0>(rs, off_fh, pf, e, q, jq, kq, F);
      } else {   // identity entry: S' = 0.5 + 0.5 for every hypothesis
        const float4 half4 = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
        *reinterpret_cast<float4*>(pf + q * A_BLK4_FLT + e * 4) = half4;
        if (!SAME) *reinterpret_cast<float4*>(pf + A_SIDE_FLT + q * A_BLK4_FLT + e * 4) = half4;
      }
    } else {
      if (e < nvalid) {
        transform_quad_f32<SAME, EXTRAS, ROW0>(rs, off_fh, pf, e, q, jq, kq, F);
      } else {   // identity entry: S' = 1 for every hypothesis
        store_neutral_quad(pf + q * F_BLK4_FLT + e * 12);
        store_neutral_quad(pf + F_SIDE_FLT + q * F_BLK4_FLT + e * 12);
      }
    }
  }
}

template <bool AFF>
__device__ __forceinline__ void transform_dispatch_f32(bool same, bool extras_t, int row0, const unsigned char* rs,
                                                       uint32_t off_fh, float* pf, int ebase, int q, int nvalid, bool jq,
                                                       bool kq, Fac& F) {
  if (same) {
    if (extras_t) {
      if (row0) transform_stage_f32<AFF, true, true, 1>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else transform_stage_f32<AFF, true, true, 0>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
    } else {
      if (row0) transform_stage_f32<AFF, true, false, 1>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else transform_stage_f32<AFF, true, false, 0>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
    }
  } else {
    if (extras_t) {
      if (row0 == 2) transform_stage_f32<AFF, false, true, 2>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else if (row0 == 1) transform_stage_f32<AFF, false, true, 1>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else transform_stage_f32<AFF, false, true, 0>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
    } else {
      if (row0 == 2) transform_stage_f32<AFF, false, false, 2>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else if (row0 == 1) transform_stage_f32<AFF, false, false, 1>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
      else transform_stage_f32<AFF, false, false, 0>(rs, off_fh, pf, ebase, q, nvalid, jq, kq, F);
    }
  }
}

// A transform warp takes whole stages from the CTA's stage counter, so the F_NTW warps work on different stages at a
// time (which hides the latency of a stage's dependent load -> FMA -> store chain) and a warp that shares its
// sub-partition with busier consumers simply takes fewer.  In-flight stages span at most F_NTW consecutive
// indices, fewer than either ring has slots, so the parity waits below cannot alias.
__device__ __forceinline__ void transform_role_f32(const ScanParams& p, const Pipe& pp, TScratch* ts, uint64_t* prod_full,
                                                   uint64_t* prod_empty, int lane, int u, int NS) {
  const int ebase = lane >> 4, q = lane & 15;
  const uint32_t sraw = (uint32_t)p.f_sraw, rawslot = (uint32_t)p.f_rawslot, off_fh = (uint32_t)p.f_off_fh;
  uint32_t gbase = 0;   // stages of the tasks before this one
  auto grab = [&]() -> uint32_t {
    uint32_t g = 0;
    if (lane == 0) g = atomicAdd(&ts->next_stage, 1u);
    return __shfl_sync(0xffffffffu, g, 0);
  };
  PROF_DECL
  uint32_t git = grab();
  TaskInfoF ti;
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const bool extras_t = (tt.flags & TF_EXTRAS) != 0 && !(p.dbg & 4);
    const bool jq = 4 * q < tt.nj;                                  // nj, nk are multiples of 4
    const bool kq = same ? jq : (4 * q < tt.nk);
    // pairs with sample 0: on the k side of symmetric tasks, on the j side of the (rectangular) tasks of full planes
    const int row0 = (!(tt.flags & TF_ROW0) || (p.dbg & 16)) ? 0 : ((tt.flags & TF_SYM) ? 1 : (same ? 0 : 2));
    for (; git < gbase + (uint32_t)ti.nst;) {
      const int it = (int)(git - gbase);
      const uint32_t rslot = git % sraw, pslot = git % F_S_PROC;
      const uint32_t rph = (git / sraw) & 1u, pph = (git / F_S_PROC) & 1u;
      int64_t eb;
      int nvalid;
      bool aff;
      stage_span(ti, it, &eb, &nvalid, &aff);
      Fac F;                                          // this stage's factors of the running products (4 entries each)
#pragma unroll
      for (int i = 0; i < 4; ++i) F.PA[i] = F.PB[i] = F.P1[i] = F.P0[i] = F.PJ[i] = 1.0f;
      PROF_WAIT(1, mbar_wait(&pp.raw_full[rslot], rph));
      PROF_WAIT(2, mbar_wait(&pp.proc_empty[pslot], pph ^ 1u));
      const unsigned char* rs = pp.raw_base + (size_t)rslot * rawslot;
      float* pf = reinterpret_cast<float*>(pp.proc_base + (size_t)pslot * F_PROC_SLOT);
      if (p.dbg & 2) {
      } else if (aff) {
        PROF_WAIT(5, transform_dispatch_f32<true>(same, extras_t, row0, rs, off_fh, pf, ebase, q, nvalid, jq, kq, F));
      } else {
        PROF_WAIT(6, transform_dispatch_f32<false>(same, extras_t, row0, rs, off_fh, pf, ebase, q, nvalid, jq, kq, F));
      }
      // the stage's factors go to the products ring (every factor is >= ~2^-20, four of them cannot underflow)
      PROF_WAIT(3, mbar_wait(&prod_empty[pslot], pph ^ 1u));
      {
        float4* dst = reinterpret_cast<float4*>(&ts->prod[pslot][ebase][0][4 * q]);
        dst[0 * (BLK / 4)] = make_float4(F.PA[0], F.PA[1], F.PA[2], F.PA[3]);
        dst[1 * (BLK / 4)] = make_float4(F.PB[0], F.PB[1], F.PB[2], F.PB[3]);
        dst[2 * (BLK / 4)] = make_float4(F.P1[0], F.P1[1], F.P1[2], F.P1[3]);
        dst[3 * (BLK / 4)] = make_float4(F.P0[0], F.P0[1], F.P0[2], F.P0[3]);
        if (!same) dst[4 * (BLK / 4)] = make_float4(F.PJ[0], F.PJ[1], F.PJ[2], F.PJ[3]);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&pp.proc_full[pslot]);
        mbar_arrive(&prod_full[pslot]);
        mbar_arrive(&pp.raw_empty[rslot]);
      }
      PROF_WAIT(4, git = grab());
    }
    gbase += (uint32_t)ti.nst;
  }
  PROF_END(1, lane)
}

// =============================== accumulator warp =================================================
// Multiplies the per-stage factors of the products ring into the per-sample running products in stage order (the
// order the stages were numbered in, not the order the transform warps finished them: the result is reproducible).
// Lane l owns quantity l / 8 of the samples 8 (l % 8) .. + 7 (lanes 0..7 also the j-side sums of a rectangular task);
// a product is renormalised into an integer exponent after every factor.  At the end of a task it hands lb (sum of
// log s per sample, both sides) to the consumers' epilogue through shared memory and writes the per-sample outputs:
// singlet factor of the separable plane, row-sum term, pass-1 likelihoods, pairs with sample 0 -- each with the lb
// terms of its samples added back.
__device__ __forceinline__ void accumulator_role_f32(const ScanParams& p, TScratch* ts, uint64_t* prod_full,
                                                     uint64_t* prod_empty, uint64_t* lb_full, uint64_t* lb_empty, int lane,
                                                     int u, int NS) {
  const int c = lane >> 3, s0 = (lane & 7) * 8;
  const double LN2 = 0.693147180559945309417232121458;
  uint32_t git = 0;
  TaskInfoF ti;
  PROF_DECL
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const bool extras_t = (tt.flags & TF_EXTRAS) != 0 && !(p.dbg & 4);
    const int row0 = (!(tt.flags & TF_ROW0) || (p.dbg & 16)) ? 0 : ((tt.flags & TF_SYM) ? 1 : (same ? 0 : 2));
    const bool do_j = !same && c == 0;      // lanes 0..7: second set of products for the j-side sums
    float acc[8], accj[8];
    int ex[8], exj[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = accj[i] = 1.0f; ex[i] = exj[i] = 0; }
    for (int it = 0; it < ti.nst; ++it, ++git) {
      const uint32_t slot = git % F_S_PROC, ph = (git / F_S_PROC) & 1u;
      PROF_WAIT(1, mbar_wait(&prod_full[slot], ph));
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float4* src = reinterpret_cast<const float4*>(&ts->prod[slot][half][c][s0]);
        const float4 x = src[0], y = src[1];
        const float f[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        if (p.dbg & 8) continue;             // timing experiment
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] *= f[i];
          peelf_unbiased(acc[i], ex[i]);     // exact zeros (padding samples) stay zero
        }
        if (do_j) {
          const float4* sj = reinterpret_cast<const float4*>(&ts->prod[slot][half][4][s0]);
          const float4 xj = sj[0], yj = sj[1];
          const float fj[8] = {xj.x, xj.y, xj.z, xj.w, yj.x, yj.y, yj.z, yj.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            accj[i] *= fj[i];
            peelf_unbiased(accj[i], exj[i]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&prod_empty[slot]);
    }
    // lb of this task -> shared memory (double buffered by task parity; the consumers release a buffer after reading)
    const int par = r & 1;
    PROF_WAIT(2, mbar_wait(&lb_empty[par], ((uint32_t)(r >> 1) & 1u) ^ 1u));
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (double)ex[i] * LN2 + (double)logf(acc[i]);
    if (c == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ts->lb[par][1][s0 + i] = v[i];
        if (same) ts->lb[par][0][s0 + i] = v[i];
      }
    }
    if (do_j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) ts->lb[par][0][s0 + i] = (double)exj[i] * LN2 + (double)logf(accj[i]);
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (generic-proxy writes; the fence is cheap insurance)
    if (lane == 0) mbar_arrive(&lb_full[par]);
    // results of this task: quantities 0..2 belong to the k-side samples of tasks with TF_EXTRAS, quantity 3 to the
    // pairs with sample 0
    const bool want = c == 3 ? row0 != 0 : (extras_t && !(c == 0 && !p.sep0) && !(c == 2 && !p.want_single));
    if (!want) continue;
    const bool on_j = c == 3 && row0 == 2;
    const double lb_other = c == 3 ? (row0 == 2 ? ts->lb[par][1][0] : ts->lb[par][0][0]) : 0.0;   // sample 0's share
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int smp = s0 + i;
      const int gs = (on_j ? tt.j0 : tt.k0) + smp;
      if (smp >= (on_j ? tt.nj : tt.nk) || gs >= p.nv) continue;
      double out = v[i];
      if (c != 1) out += ts->lb[par][on_j ? 0 : 1][smp] + lb_other;
      if (c == 3) p.row0[((int64_t)ti.cell * p.nA + tt.n) * p.nv + gs] = out;
      else (c == 0 ? p.singA : (c == 1 ? p.lb : p.p1))[(int64_t)ti.cell * p.nv + gs] = out;
    }
  }
  PROF_END(2, lane)
}

// =============================== per-task reductions (:712-734, :799-814) ==========================
// Every consumer thread brings NV llk values with their (j, k); the block of NW consumer warps reduces them to
// the task's Part (max, sum of exp, best pair value, near-tie candidates) and stores row 0 / the dense block.
template <int NV, int NW, bool ROW0, class TI>
__device__ __forceinline__ void task_epilogue(const ScanParams& p, const TI& ti, const TaskType& tt, Scratch* scr,
                                              int bar_id, int rw, int lane, const double (&v)[NV], const int (&vj)[NV],
                                              const int (&vk)[NV], uint32_t vmask) {
  const bool sym = tt.flags & TF_SYM;
  const bool noarg = tt.flags & TF_NOARG;
  double m_all = -INFINITY, best = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double x = (vmask >> i) & 1u ? v[i] : -INFINITY;
    m_all = fmax(m_all, x);
    if (vj[i] != vk[i]) best = fmax(best, x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m_all = fmax(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  }
  if (lane == 0) { scr->red[0][rw] = m_all; scr->red[1][rw] = best; }
  consumer_bar_sync<NW * 32>(bar_id);
  m_all = scr->red[0][0]; best = scr->red[1][0];
#pragma unroll
  for (int wq = 1; wq < NW; ++wq) {
    m_all = fmax(m_all, scr->red[0][wq]);
    best = fmax(best, scr->red[1][wq]);
  }
  const double thr = best - (p.eps_rel * fmax(1.0, fabs(best)) + p.eps_ent * (double)(ti.e1 - ti.e0));
  double sd = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (!((vmask >> i) & 1u)) continue;
    const int j = vj[i], k = vk[i];
    const double x = v[i];
    if (j != k) {
      // relative to the task's best pair (not to m_all, which may be a singlet-like diagonal value hundreds of nats
      // above every pair: the sum must keep its leading digits even then)
      sd += ROW0 ? exp(x - best) : (double)expf((float)(x - best));   // fp32 scan: 1e-7 relative is plenty here
      if (!noarg && x >= thr) {
        const int cs = atomicAdd(&scr->ncand, 1);
        if (cs < MAX_CAND) scr->cand[cs] = (uint32_t)j | ((uint32_t)k << 12);
      }
    }
    // row k = 0 of the plane (LLK10 / LLK20, :824-825; singlets when plane 0 is dense)
    if (ROW0) {
      if (sym) {
        if (j == 0) p.row0[((int64_t)ti.cell * p.nA + tt.n) * p.nv + k] = x;
      } else if (k == 0) {
        p.row0[((int64_t)ti.cell * p.nA + tt.n) * p.nv + j] = x;
      }
    }
    if (p.dense) {
      double* d = p.dense + ti.aidx * p.nv * p.nv * p.nA;
      d[((int64_t)j * p.nv + k) * p.nA + tt.n] = x;
      if (sym) d[((int64_t)k * p.nv + j) * p.nA + tt.n] = x;
    }
  }
  if (sym) sd *= 2.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
  if (lane == 0) scr->red2[rw] = sd;
  consumer_bar_sync<NW * 32>(bar_id);
  if (rw == 0 && lane == 0) {
    double tot = 0.0;
    for (int wq = 0; wq < NW; ++wq) tot += scr->red2[wq];
    Part* out = p.parts + ti.aidx * p.T + ti.tti;
    out->m_all = m_all;
    out->sd = tot;
    out->best_v = best;
    out->ncand = scr->ncand;
    out->pad = 0;
    for (int c = 0; c < MAX_CAND; ++c) out->cand[c] = scr->cand[c];
    scr->ncand = 0;
  }
  consumer_bar_sync<NW * 32>(bar_id);
}

// D (8 x 8) = A (8 x 4, row) * B (4 x 8, col) in FP64 on the tensor cores (SASS: DMMA.8x8x4).  Fragments: lane
// holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// Consumer tile shapes: a warp owns 8 tiles of 8 x 8 hypotheses built from <= 4 row blocks and <= 4 column blocks.
//   shape 0: 2 row blocks x 4 col blocks      tile t -> (A[t >> 2], B[t & 3])
//   shape 1: 4 row blocks x 2 col blocks      tile t -> (A[t >> 1], B[t & 1])
//   shape 2: two groups of 2 x 2              tile t -> (A[2g + ((t >> 1) & 1)], B[2g + (t & 1)]), g = t >> 2
template <int SHAPE> __device__ __forceinline__ constexpr int tile_ia(int t) {
  return SHAPE == 0 ? (t >> 2) : SHAPE == 1 ? (t >> 1) : (2 * (t >> 2) + ((t >> 1) & 1));
}
template <int SHAPE> __device__ __forceinline__ constexpr int tile_ib(int t) {
  return SHAPE == 0 ? (t & 3) : SHAPE == 1 ? (t & 1) : (2 * (t >> 2) + (t & 1));
}
constexpr int NTILE = 8;

template <int SHAPE>
__device__ __forceinline__ void consume_stage(const double* __restrict__ pa, const int (&aoff)[4], const int (&boff)[4],
                                              uint32_t tmask, double (&acc)[NTILE][2]) {
  constexpr int NA = SHAPE == 0 ? 2 : 4, NBF = SHAPE == 1 ? 2 : 4;
  const double* pw = pa + NB * SIDE_DBL;
#pragma unroll 2
  for (int e = 0; e < NB; ++e) {
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < NA; ++i) af[i] = pa[e * SIDE_DBL + aoff[i]];
#pragma unroll
    for (int i = 0; i < NBF; ++i) bf[i] = pw[e * SIDE_DBL + boff[i]];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
      if (tmask & (1u << t)) {
        double s0, s1;
        dmma884(s0, s1, af[tile_ia<SHAPE>(t)], bf[tile_ib<SHAPE>(t)]);
        acc[t][0] *= s0;
        acc[t][1] *= s1;
      }
    }
  }
}
}  // namespace

// Warp roles of the fp64 kernel.  SM sub-partition = warp id % 4; the FP64 pipe is the bound, so the consumer
// warps (C) are spread 3/3/2/2 over the four sub-partitions and the lighter transform (T) and loader (L) warps
// fill the rest:
//   wid  0  1  2  3 |  4  5  6  7 |  8  9 10 11 | 12 13 14 15
//        CA CA CA CA| CA CB CB CB | CB CB TA TA | LA LB TB TB        (A, B = slot)
__global__ void __launch_bounds__(NTHREADS, 1) pair_scan_kernel(const __grid_constant__ ScanParams p,
                                                                const __grid_constant__ ScanMaps tm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // role decode (see table above)
  int role, slot, rw;   // role 0 consumer, 1 transform, 2 loader; rw = warp index inside the role
  if (warp < 5) { role = 0; slot = 0; rw = warp; }
  else if (warp < 10) { role = 0; slot = 1; rw = warp - 5; }
  else if (warp < 12) { role = 1; slot = 0; rw = warp - 10; }
  else if (warp < 14) { role = 2; slot = warp - 12; rw = 0; }
  else { role = 1; slot = 1; rw = warp - 14; }

  unsigned char* sb = smem + (size_t)slot * SLOT_BYTES;
  Pipe pp;
  pp.raw_full = reinterpret_cast<uint64_t*>(sb);
  pp.raw_empty = pp.raw_full + S_RAW;
  pp.proc_full = pp.raw_empty + S_RAW;
  pp.proc_empty = pp.proc_full + S_PROC;
  Scratch* scr = reinterpret_cast<Scratch*>(sb + SCRATCH_OFF);
  pp.raw_base = sb + SLOT_CTRL;
  pp.proc_base = pp.raw_base + S_RAW * RAW_SLOT;

  if (threadIdx.x < NSLOT) {
    unsigned char* b = smem + (size_t)threadIdx.x * SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b);
    for (int s = 0; s < S_RAW; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[S_RAW + s], NTW);
    }
    for (int s = 0; s < S_PROC; ++s) {
      mbar_init(&bars[2 * S_RAW + s], NTW);
      mbar_init(&bars[2 * S_RAW + S_PROC + s], NCW);
    }
    reinterpret_cast<Scratch*>(b + SCRATCH_OFF)->ncand = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // the 4th (padding) column of every A / B matrix must be 0: it is never written again
  for (int sl = 0; sl < NSLOT; ++sl) {
    double* pr = reinterpret_cast<double*>(smem + (size_t)sl * SLOT_BYTES + SLOT_CTRL + S_RAW * RAW_SLOT);
    for (uint32_t i = threadIdx.x; i < S_PROC * PROC_SLOT / 8; i += NTHREADS) pr[i] = 0.0;
  }
  __syncthreads();

  const int NS = (int)gridDim.x * NSLOT;
  const int u = (int)blockIdx.x * NSLOT + slot;

  if (role == 2) {
    loader_role<S_RAW>(p, tm, pp, lane, u, NS);
  } else if (role == 1) {
    transform_role<S_RAW, S_PROC, PROC_SLOT>(p, pp, rw, lane, u, NS);
  } else {
    // =============================== consumer warps ============================================
    uint32_t git = 0;   // stages consumed so far in this slot (same sequence in every role)
    TaskInfo ti;
    const int lr = lane >> 2, lc = lane & 3;
    for (int r = 0; next_task(p, r, u, NS, &ti); ++r) {
      const TaskType tt = p.tt[ti.tti];
      const bool tri = tt.flags & TF_TRI;
      const int nbj = (tt.nj + 7) >> 3, nbk = (tt.nk + 7) >> 3;   // 8-sample blocks on each side
      // tile cover of the task (see the shape comment above)
      int shape = 0, ar[4] = {0, 0, 0, 0}, bc[4] = {0, 0, 0, 0};
      if (tri) {
        if (rw == 0) { ar[0] = 0; ar[1] = 1; bc[0] = 0; bc[1] = 1; bc[2] = 2; bc[3] = 3; }
        else if (rw == 1) { ar[0] = 0; ar[1] = 1; bc[0] = 4; bc[1] = 5; bc[2] = 6; bc[3] = 7; }
        else if (rw == 2) { ar[0] = 2; ar[1] = 3; bc[0] = 2; bc[1] = 3; bc[2] = 4; bc[3] = 5; }
        else if (rw == 3) { shape = 1; ar[0] = 2; ar[1] = 3; ar[2] = 4; ar[3] = 5; bc[0] = 6; bc[1] = 7; }
        else { shape = 2; ar[0] = 4; ar[1] = 5; ar[2] = 6; ar[3] = 7; bc[0] = 4; bc[1] = 5; bc[2] = 6; bc[3] = 7; }
      } else {
        ar[0] = 2 * rw; ar[1] = 2 * rw + 1; bc[0] = 0; bc[1] = 1; bc[2] = 2; bc[3] = 3;
      }
      int trow[NTILE], tcol[NTILE];
      uint32_t tmask = 0;
#pragma unroll
      for (int t = 0; t < NTILE; ++t) {
        const int ia = shape == 0 ? (t >> 2) : shape == 1 ? (t >> 1) : (2 * (t >> 2) + ((t >> 1) & 1));
        const int ib = shape == 0 ? (t & 3) : shape == 1 ? (t & 1) : (2 * (t >> 2) + (t & 1));
        trow[t] = ar[ia];
        tcol[t] = bc[ib];
        const bool ok = trow[t] < nbj && tcol[t] < nbk && (!tri || tcol[t] >= trow[t]) && (tri || rw < 4);
        if (ok) tmask |= 1u << t;
      }
      // fragment offsets (doubles) inside one entry's side matrix: block * 32 + swizzled column * 8 + row
      int aoff[4], boff[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        aoff[i] = ar[i] * 32 + ((lc ^ (ar[i] & 1)) << 3) + lr;
        boff[i] = bc[i] * 32 + ((lc ^ (bc[i] & 1)) << 3) + lr;
      }
      double acc[NTILE][2];
      int es[NTILE][2];
#pragma unroll
      for (int t = 0; t < NTILE; ++t) { acc[t][0] = acc[t][1] = 1.0; es[t][0] = es[t][1] = 0; }
      int nflush = 0;
      for (int it = 0; it < ti.nst; ++it, ++git) {
        const int pslot = git % S_PROC;
        const uint32_t pph = (git / S_PROC) & 1u;
        mbar_wait(&pp.proc_full[pslot], pph);
        if (tmask) {
          const double* pa = reinterpret_cast<const double*>(pp.proc_base + (size_t)pslot * PROC_SLOT);
          if (shape == 0) consume_stage<0>(pa, aoff, boff, tmask, acc);
          else if (shape == 1) consume_stage<1>(pa, aoff, boff, tmask, acc);
          else consume_stage<2>(pa, aoff, boff, tmask, acc);
          if ((it % FLUSH_STAGES) == FLUSH_STAGES - 1) {
#pragma unroll
            for (int t = 0; t < NTILE; ++t) { peel(acc[t][0], es[t][0]); peel(acc[t][1], es[t][1]); }
            ++nflush;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pp.proc_empty[pslot]);
      }
      double v[NTILE * 2];
      int vj[NTILE * 2], vk[NTILE * 2];
      uint32_t vmask = 0;
#pragma unroll
      for (int t = 0; t < NTILE; ++t)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int j = tt.j0 + trow[t] * 8 + lr, k = tt.k0 + tcol[t] * 8 + 2 * lc + i;
          const bool valid = ((tmask >> t) & 1u) && j < p.nv && k < p.nv && (!tri || j <= k);
          v[t * 2 + i] = valid ? unpeel_log(acc[t][i], es[t][i], nflush) : -INFINITY;
          vj[t * 2 + i] = j;
          vk[t * 2 + i] = k;
          if (valid) vmask |= 1u << (t * 2 + i);
        }
      task_epilogue<NTILE * 2, NCW, true>(p, ti, tt, scr, slot, rw, lane, v, vj, vk, vmask);
    }
  }
}

// S'[i][c] = a_i . w_c and acc *= S' for one 8 x 8 tile and one general entry, a few tile rows at a time (independent
// chains that cover the 4-cycle FMA latency).  A diagonal tile of a triangular task skips the hypotheses below the
// diagonal.
__device__ __forceinline__ void tile_math(const float (&a)[8][3], const float (&w)[8][3], float (&acc)[64]) {
  constexpr bool DIAG = false;   // diagonal tiles of a triangular task compute all 64 products too: one code path for
                                 // every consumer warp (the instruction cache is the scarcer resource)
#ifndef DMX_TILE_ROWS
#define DMX_TILE_ROWS 2
#endif
  constexpr int R = DMX_TILE_ROWS;   // tile rows per pass: R x 8 independent chains
#pragma unroll
  for (int i0 = 0; i0 < 8; i0 += R) {
    float s[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = a[i0 + r][0] * w[c][0];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = fmaf(a[i0 + r][1], w[c][1], s[r][c]);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = fmaf(a[i0 + r][2], w[c][2], s[r][c]);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) acc[(i0 + r) * 8 + c] *= s[r][c];
  }
}

// the same for an affine entry: S'[i][c] = X_i + Y_c, one add and one multiply per hypothesis
__device__ __forceinline__ void tile_math_aff(const float (&x)[8], const float (&y)[8], float (&acc)[64]) {
  constexpr bool DIAG = false;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (!DIAG || c >= i) acc[i * 8 + c] *= x[i] + y[c];
}

// Range of the running products: a thread's 64 products share one integer exponent (block floating point).  A factor
// is >= ~2^-20 (the 1e-6 floors), so a tile whose largest product is above 2^-46 now keeps it inside the fp32 range
// for F_PEEL_STAGES more entries; below that every thread of the warp rescales its tile by a power of two.  A diagonal
// tile looks at (and rescales) only the hypotheses it computes: the products below the diagonal stay 1 and must not
// pass for the tile's largest.
__device__ __forceinline__ void range_check(float (&acc)[64], int& ex, bool active) {
  constexpr bool DIAG = false;
  float m8[8];   // eight short chains instead of one long one
#pragma unroll
  for (int c = 0; c < 8; ++c) m8[c] = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (!DIAG || c >= i) m8[c] = fmaxf(m8[c], acc[i * 8 + c]);
  const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
  if (__any_sync(0xffffffffu, active && mx < 0x1p-46f)) {
    uint32_t eb = __float_as_uint(mx) >> 23;                  // biased exponent of the largest product
    if (eb == 0u || eb >= 254u) eb = 127u;                    // all-zero tile (padding samples only): leave it alone
    const float sc = __uint_as_float((254u - eb) << 23);      // 2^(127 - eb): brings it back to [1, 2)
    ex += (int)eb - 127;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i) acc[i * 8 + c] *= sc;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 pair scan: one task pipeline per CTA, 16 warps in four warpgroups.  SM sub-partition = warp id % 4; roles are
// placed so that the issue load of the four sub-partitions is even (C = consumer, D = consumer of diagonal tiles:
// 36 / 64 of the work, T = transform, L = loader, A = accumulator of the per-sample products):
//   wid   0  1  2  3 |  4  5  6  7 |  8  9 10 11 | 12 13 14 15
//         C  C  C  C |  C  C  C  C |  C  T  T  T |  L  T  L  A
// The last warpgroup gives registers back (setmaxnreg) so that the consumers' 64 accumulators + 48 operands fit.
// Every role is written for a small instruction footprint (one consumer code path, rolled transform): the warps of a
// CTA run five different programs out of one 32 KB instruction cache.
//
// A consumer thread owns the 8 x 8 tile (bj, bk) of hypotheses for the entries e = lane & 7 of every stage.  Per
// general entry: 12 LDS.128 (two 4-sample blocks of a_hat, two of w), 64 x (FMUL + 2 FFMA) for S'[j][k] and 64 FMUL
// into the running products; per affine entry: 4 LDS.128 (X of its 8 j samples, Y of its 8 k samples), 64 x (FADD +
// FMUL).  Range: a thread's 64 products share one integer exponent (block floating point): every F_PEEL_STAGES entries
// the warp checks whether some thread's LARGEST product is about to leave the fp32 range and, if so, every thread
// rescales its tile by a power of two.  A hypothesis more than ~2^-80 below the best of its own tile underflows to
// zero (llk = -inf): it cannot be the best pair, a near tie or a visible term of sumDouble.  The pairs with sample 0
// that .best prints (LLK10 / LLK20) do not come from here but from the transform warps (TF_ROW0), and dense llksAB
// blocks (--write-pair) from the fp64 scan.  At the end of a task the eight entry lanes of a tile multiply their
// products together (fp64 reduce-scatter: lane e ends up with tile row e) and add the lb terms of their samples.
// ---------------------------------------------------------------------------------------------
constexpr int F_LAUNCH_THREADS = 512;
// Consumer hand-over: with DMX_CONS_BAR one consumer warp polls the stage's mbarrier and a named hardware barrier
// releases the other eight (a warp parked in bar.sync costs no issue slots, nine polling warps cost hundreds per
// stage); the barrier of stage i + 1 also tells that every warp is done with stage i, so the polling warp alone
// hands the slot back.
#ifndef DMX_CONS_BAR
#define DMX_CONS_BAR 0
#endif
// Role placement (SM sub-partition = warp id % 4).  DMX_ROLEMAP 1: the four transform warps own sub-partition 3, the
// nine consumers sit three to a sub-partition on 0..2 with a loader or the accumulator as fourth warp.
#ifndef DMX_ROLEMAP
#define DMX_ROLEMAP 0
#endif
#define DMX_CONS_POLLERS (DMX_CONS_BAR ? 1 : F_NCW)
static_assert(F_S_RAW_RECT % F_NLOAD == 0 && F_S_RAW_TRI % F_NLOAD == 0 && F_NLOAD == 2 && F_NTW <= 4 &&
              F_NTW < (int)F_S_RAW_RECT && F_NTW < F_S_PROC, "ring ownership");
static_assert(2 * 12 + 4 * F_S_PROC + 4 <= 64, "barriers must fit below the scratch area");
constexpr int F_REGS_HI = 136, F_REGS_LO = 104;   // (3 * 136 + 104) * 32 = 16384 registers per sub-partition

__global__ void __launch_bounds__(F_LAUNCH_THREADS, 1) pair_scan_f32_kernel(const __grid_constant__ ScanParams p,
                                                                          const __grid_constant__ ScanMaps tm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Pipe pp;
  pp.raw_full = reinterpret_cast<uint64_t*>(smem);
  pp.raw_empty = pp.raw_full + F_S_RAW_TRI;
  pp.proc_full = pp.raw_empty + F_S_RAW_TRI;
  pp.proc_empty = pp.proc_full + F_S_PROC;
  uint64_t* prod_full = pp.proc_empty + F_S_PROC;      // products ring: transform -> accumulator warp
  uint64_t* prod_empty = prod_full + F_S_PROC;
  uint64_t* lb_full = prod_empty + F_S_PROC;           // lb of a finished task: accumulator -> consumers
  uint64_t* lb_empty = lb_full + 2;
  Scratch* scr = reinterpret_cast<Scratch*>(smem + SCRATCH_OFF);
  pp.raw_base = smem + SLOT_CTRL;
  pp.proc_base = pp.raw_base + F_RAW_BYTES;
  TScratch* ts = reinterpret_cast<TScratch*>(smem + F_SMEM_NOTS);

  if (threadIdx.x == 0) {
    for (int s = 0; s < (int)F_S_RAW_TRI; ++s) {
      mbar_init(&pp.raw_full[s], 1);
      mbar_init(&pp.raw_empty[s], 1);
    }
    for (int s = 0; s < F_S_PROC; ++s) {
      mbar_init(&pp.proc_full[s], 1);
      mbar_init(&pp.proc_empty[s], DMX_CONS_POLLERS);
      mbar_init(&prod_full[s], 1);
      mbar_init(&prod_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&lb_full[s], 1);
      mbar_init(&lb_empty[s], F_NCW);
    }
    scr->ncand = 0;
    ts->next_stage = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int NS = (int)gridDim.x, u = (int)blockIdx.x;
#if DMX_ROLEMAP == 1
  const bool is_cons = warp < 12 && (warp & 3) != 3;
  const bool is_tr = (warp & 3) == 3;
  const bool is_acc = warp == 14;
  const int lid = warp - 12;                       // loaders: warps 12, 13
  const int rw_map = (warp >> 2) * 3 + (warp & 3); // consumer index 0..8
#else
  const bool is_cons = warp < F_NCW;
  const bool is_tr = (warp >= F_NCW && warp < 12) || warp == 13;
  const bool is_acc = warp == 15;
  const int lid = (warp - 12) >> 1;                // loaders: warps 12, 14
  const int rw_map = warp;
#endif
  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F_REGS_LO));
    if (is_tr) { if (F_NTW > 3) transform_role_f32(p, pp, ts, prod_full, prod_empty, lane, u, NS); }
    else if (is_acc) accumulator_role_f32(p, ts, prod_full, prod_empty, lb_full, lb_empty, lane, u, NS);
    else loader_role_f32(p, tm, pp, lane, u, NS, lid);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F_REGS_HI));
  if (is_tr) {
    transform_role_f32(p, pp, ts, prod_full, prod_empty, lane, u, NS);
  } else if (is_cons) {
    // =============================== consumer warps ============================================
    // consumer warp rw owns tile positions 4 rw .. 4 rw + 3: 0..27 the off-diagonal tiles of a triangular task, 28..35
    // the diagonal ones
    const int rw = rw_map;
    const int el = lane & 7;
    const int pos = rw * 4 + (lane >> 3);
    uint32_t git = 0;
    TaskInfoF ti;
    PROF_DECL
    for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
      const TaskType tt = p.tt[ti.tti];
      const bool tri = tt.flags & TF_TRI;
      const int nbj = (tt.nj + 7) >> 3, nbk = (tt.nk + 7) >> 3;   // 8-sample blocks on each side
      // tile positions of a triangular task: 0..27 the off-diagonal tiles (bj < bk, row-major), 28..35 the diagonal
      // tiles, which evaluate only the 36 hypotheses with j <= k of their tile
      int bj, bk;
      const bool diag = tri && pos >= 28;
      if (tri) {
        if (diag) {
          bj = bk = pos - 28;
        } else {
          int q = pos;
          bj = 0;
          while (q >= 7 - bj) { q -= 7 - bj; ++bj; }
          bk = bj + 1 + q;
        }
      } else {
        bj = pos >> 2;
        bk = pos & 3;
      }
      const bool active = (tri || pos < 32) && bj < nbj && bk < nbk;
      const uint32_t aoff = (uint32_t)(2 * bj) * F_BLK4_FLT + (uint32_t)el * 12u;                 // floats
      const uint32_t woff = F_SIDE_FLT + (uint32_t)(2 * bk) * F_BLK4_FLT + (uint32_t)el * 12u;
      // affine stages: X blocks of the j side, Y blocks of the k side (one shared side in a triangular task)
      const uint32_t xoff = (uint32_t)(2 * bj) * A_BLK4_FLT + (uint32_t)el * 4u;
      const uint32_t yoff = (tri ? 0u : A_SIDE_FLT) + (uint32_t)(2 * bk) * A_BLK4_FLT + (uint32_t)el * 4u;
      float acc[64];
#pragma unroll
      for (int h = 0; h < 64; ++h) acc[h] = 1.0f;
      int ex = 0;                               // true products = acc * 2^ex
      for (int it = 0; it < ti.nst; ++it, ++git) {
        const int pslot = git % F_S_PROC;
        const uint32_t pph = (git / F_S_PROC) & 1u;
#if DMX_CONS_BAR
        if (rw == 0) mbar_wait(&pp.proc_full[pslot], pph);
        consumer_bar_sync<F_NCW * 32>(1);
        if (rw == 0 && lane == 0 && it > 0) mbar_arrive(&pp.proc_empty[(git - 1) % F_S_PROC]);
#else
        PROF_WAIT(1, mbar_wait(&pp.proc_full[pslot], pph));
#endif
        const float* pf = reinterpret_cast<const float*>(pp.proc_base + (size_t)pslot * F_PROC_SLOT);
        if (!active || (p.dbg & 1)) {
        } else if (it < ti.nstA) {
          const float4 x0 = *reinterpret_cast<const float4*>(pf + xoff);
          const float4 x1 = *reinterpret_cast<const float4*>(pf + xoff + A_BLK4_FLT);
          const float4 y0 = *reinterpret_cast<const float4*>(pf + yoff);
          const float4 y1 = *reinterpret_cast<const float4*>(pf + yoff + A_BLK4_FLT);
          const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
          tile_math_aff(x, y, acc);
        } else {
          float a[8][3], w[8][3];
          float4 ld[12];   // issue order: first w block, the two a blocks, second w block
#pragma unroll
          for (int q = 0; q < 3; ++q) ld[q] = *reinterpret_cast<const float4*>(pf + woff + q * 4);
#pragma unroll
          for (int q = 0; q < 6; ++q)
            ld[3 + q] = *reinterpret_cast<const float4*>(pf + aoff + (q / 3) * F_BLK4_FLT + (q % 3) * 4);
#pragma unroll
          for (int q = 0; q < 3; ++q) ld[9 + q] = *reinterpret_cast<const float4*>(pf + woff + F_BLK4_FLT + q * 4);
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const float4 x = ld[3 + q], y = q < 3 ? ld[q] : ld[9 + q - 3];
            const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int f = (q % 3) * 4 + c;   // float index inside the 4-sample block
              a[(q / 3) * 4 + f / 3][f % 3] = xs[c];
              w[(q / 3) * 4 + f / 3][f % 3] = ys[c];
            }
          }
          tile_math(a, w, acc);
        }
        if ((it % F_PEEL_STAGES) == F_PEEL_STAGES - 1 && !(p.dbg & 64)) {
          range_check(acc, ex, active);
        }
#if !DMX_CONS_BAR
        __syncwarp();
        if (lane == 0) mbar_arrive(&pp.proc_empty[pslot]);
#endif
      }
#if DMX_CONS_BAR
      consumer_bar_sync<F_NCW * 32>(1);          // every warp is done with the task's last stage
      if (rw == 0 && lane == 0 && ti.nst > 0) mbar_arrive(&pp.proc_empty[(git - 1) % F_S_PROC]);
#endif
      // ---- fold the eight entry lanes of each tile (reduce-scatter over lane bits 2, 1, 0): lane e ends up with the
      // products of tile row e.  Values of different lanes are multiplied in fp64 (no range problem).
      double d3[8];
      const bool b2 = el & 4, b1 = el & 2, b0 = el & 1;
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {          // columns 4 cg .. 4 cg + 3
        double d1[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float lo = acc[i * 8 + cg * 4 + c], hi = acc[(i + 4) * 8 + cg * 4 + c];
            const double keep = (double)(b2 ? hi : lo), send = (double)(b2 ? lo : hi);
            d1[i * 4 + c] = keep * __shfl_xor_sync(0xffffffffu, send, 4);
          }
        double d2[8];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double lo = d1[i * 4 + c], hi = d1[(i + 2) * 4 + c];
            d2[i * 4 + c] = (b1 ? hi : lo) * __shfl_xor_sync(0xffffffffu, b1 ? lo : hi, 2);
          }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double lo = d2[c], hi = d2[4 + c];
          d3[cg * 4 + c] = (b0 ? hi : lo) * __shfl_xor_sync(0xffffffffu, b0 ? lo : hi, 1);
        }
      }
      int exs = ex;
      exs += __shfl_xor_sync(0xffffffffu, exs, 4);
      exs += __shfl_xor_sync(0xffffffffu, exs, 2);
      exs += __shfl_xor_sync(0xffffffffu, exs, 1);
      // the sums s the transform divided out come back as lb[j] + lb[k] (from the accumulator warp)
      const int par = r & 1;
      PROF_WAIT(2, mbar_wait(&lb_full[par], (uint32_t)(r >> 1) & 1u));
      const double lbj = ts->lb[par][0][bj * 8 + el];
      double v[8];
      int vj[8], vk[8];
      uint32_t vmask = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int j = tt.j0 + bj * 8 + el, k = tt.k0 + bk * 8 + c;
        const bool valid = active && j < p.nv && k < p.nv && (!tri || j <= k);
        v[c] = valid ? (double)exs * 0.693147180559945309417232121458 + log_pos_fast(d3[c]) + (lbj + ts->lb[par][1][bk * 8 + c])
                     : -INFINITY;
        vj[c] = j;
        vk[c] = k;
        if (valid) vmask |= 1u << c;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&lb_empty[par]);
      PROF_WAIT(3, (task_epilogue<8, F_NCW, false>(p, ti, tt, scr, 0, rw, lane, v, vj, vk, vmask)));
    }
    PROF_END(3, lane)
  }
}

#include "dmx_scan2.cuh"

size_t pair_scan_smem_bytes(const ScanParams& p, int fp32) {
  if (fp32 == 2) return p.has_rect ? (size_t)V_SMEM_RECT : (size_t)V_SMEM_TRI;
  return fp32 ? (size_t)(F_SMEM_NOTS + F_TS_BYTES) : (size_t)NSLOT * SLOT_BYTES;
}

static int g_num_sms = 0;
int pair_scan_init() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(pair_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NSLOT * SLOT_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(pair_scan_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_NOTS + F_TS_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(pair_scan_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             V_SMEM_TRI > V_SMEM_RECT ? V_SMEM_TRI : V_SMEM_RECT);
  return (int)e;
}

int launch_pair_scan(const ScanParams& p, const ScanMaps& tm, int fp32, void* stream) {
  if (p.ntasks == 0) return 0;
  if (g_num_sms == 0) {
    int rc = pair_scan_init();
    if (rc) return rc;
  }
  const size_t smem = pair_scan_smem_bytes(p, fp32);
  if (fp32 == 2) {
    int64_t grid = p.ntasks;
    if (grid > g_num_sms) grid = g_num_sms;
    ScanParams q = p;
    q.v_nslot = p.has_rect ? V_NSLOT_RECT : V_NSLOT_TRI;
    q.v_slot = p.has_rect ? V_SLOT_RECT : V_SLOT_TRI;
    pair_scan_v2_kernel<<<(unsigned)grid, F_LAUNCH_THREADS, smem, (cudaStream_t)stream>>>(q, tm);
#if DMX_PROF
    {
      unsigned long long h[4][8];
      cudaStreamSynchronize((cudaStream_t)stream);
      cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
      static const char* role[4] = {"loader", "singles", "-", "consumer"};
      for (int r = 0; r < 4; ++r) {
        const double w = h[r][7] ? (double)h[r][7] : 1.0;
        fprintf(stderr, "prof2 %-9s warps %4llu  total %.0f kclk/warp  waits(kclk/warp):", role[r], h[r][7], h[r][0] / w / 1e3);
        for (int i = 1; i < 7; ++i) fprintf(stderr, " %.0f", h[r][i] / w / 1e3);
        fprintf(stderr, "\n");
      }
      memset(h, 0, sizeof(h));
      cudaMemcpyToSymbol(g_prof, h, sizeof(h));
    }
#endif
  } else if (fp32) {
    int64_t grid = p.ntasks;
    if (grid > g_num_sms) grid = g_num_sms;
    ScanParams q = p;
    q.f_sraw = p.has_rect ? F_S_RAW_RECT : F_S_RAW_TRI;
    q.f_rawslot = p.has_rect ? F_RAW_SLOT_RECT : F_RAW_SLOT_TRI;
    q.f_off_fh = q.f_rawslot - F_REC_BYTES;
    pair_scan_f32_kernel<<<(unsigned)grid, F_LAUNCH_THREADS, smem, (cudaStream_t)stream>>>(q, tm);
#if DMX_PROF
    {
      unsigned long long h[4][8];
      cudaStreamSynchronize((cudaStream_t)stream);
      cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
      static const char* role[4] = {"loader", "transform", "accumulator", "consumer"};
      for (int r = 0; r < 4; ++r) {
        const double w = h[r][7] ? (double)h[r][7] : 1.0;
        fprintf(stderr, "prof %-11s warps %4llu  total %.0f kclk/warp  waits(kclk/warp):", role[r], h[r][7], h[r][0] / w / 1e3);
        for (int i = 1; i < 7; ++i) fprintf(stderr, " %.0f", h[r][i] / w / 1e3);
        fprintf(stderr, "\n");
      }
      memset(h, 0, sizeof(h));
      cudaMemcpyToSymbol(g_prof, h, sizeof(h));
    }
#endif
  } else {
    int64_t grid = (p.ntasks + NSLOT - 1) / NSLOT;
    if (grid > g_num_sms) grid = g_num_sms;
    pair_scan_kernel<<<(unsigned)grid, NTHREADS, smem, (cudaStream_t)stream>>>(p, tm);
  }
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// merge: one thread per active cell combines its tasks (:712-734, :799-814)
// ---------------------------------------------------------------------------------------------
__global__ void merge_kernel(const __grid_constant__ MergeParams p, int64_t nactive) {
  const int64_t aidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (aidx >= nactive) return;
  const int cell = p.cell_order[aidx];
  const int nv = p.nv;
  double* sing = p.sing + (int64_t)cell * nv;
  double mx = -INFINITY;
  if (p.sep0) {
    double ma = -INFINITY, mb = -INFINITY;
    const double lb0 = p.lb[(int64_t)cell * nv];
    for (int j = 0; j < nv; ++j) {
      const double a = p.singA[(int64_t)cell * nv + j];
      ma = fmax(ma, a);
      mb = fmax(mb, p.lb[(int64_t)cell * nv + j]);
      sing[j] = a + lb0;
    }
    mx = ma + mb;
  } else {
    for (int j = 0; j < nv; ++j) sing[j] = p.row0[((int64_t)cell * p.nA) * nv + j];
  }
  const Part* parts = p.parts + aidx * p.T;
  double best = -INFINITY;
  for (int t = 0; t < p.T; ++t) {
    mx = fmax(mx, parts[t].m_all);
    if (!(p.tt[t].flags & TF_NOARG)) best = fmax(best, parts[t].best_v);
  }
  double ss = 0.0, sdb = 0.0;
  for (int j = 0; j < nv; ++j) ss += exp(sing[j] - mx) * p.w_single;
  const double eps_abs = p.eps_ent * (double)(p.cell_ptr[cell + 1] - p.cell_ptr[cell]);
  const double thr = best - (p.eps_rel * fmax(1.0, fabs(best)) + eps_abs);
  CellOut o;
  o.ncand = 0;
  o.pad = 0;
  for (int c = 0; c < MAX_CAND; ++c) o.cand[c] = 0;
  for (int t = 0; t < p.T; ++t) {
    const TaskType& tt = p.tt[t];
    if (tt.flags & TF_NOARG) continue;
    if (parts[t].best_v > -INFINITY) sdb += parts[t].sd * exp(parts[t].best_v - mx) * p.w_double[tt.n];
    if (parts[t].best_v >= thr) {
      const int nc = parts[t].ncand;
      for (int c = 0; c < nc && c < MAX_CAND; ++c)
        if (o.ncand + c < MAX_CAND) o.cand[o.ncand + c] = parts[t].cand[c] | ((uint32_t)tt.n << 24);
      o.ncand += nc;   // > MAX_CAND marks overflow
    }
  }
  o.max_llk = mx;
  o.sum_single = ss;
  o.sum_double = sdb;
  o.best_v = best;
  p.out[cell] = o;
  if (p.lists) {
    // hypotheses whose reference-exact value decides an output: every ordered candidate pair, and the
    // singlets if the top two are too close to call (same windows as the host finaliser)
    RescList* L = p.lists + cell;
    int n = 0;
    bool over = o.ncand > MAX_CAND;
    if (!over && o.ncand > 1 - 1) {
      int pairs = 0;
      for (int c = 0; c < o.ncand; ++c) pairs += (p.mode[o.cand[c] >> 24] == PM_SYM) ? 2 : 1;
      if (pairs > 1)
        for (int c = 0; c < o.ncand; ++c) {
          const uint32_t h = o.cand[c];
          L->hyp[n++] = h;
          if (p.mode[h >> 24] == PM_SYM) L->hyp[n++] = ((h >> 12) & 0xfffu) | ((h & 0xfffu) << 12) | (h & 0xff000000u);
        }
    }
    int i1 = 0;
    for (int j = 1; j < nv; ++j)
      if (sing[j] > sing[i1]) i1 = j;
    double m2 = -INFINITY;
    for (int j = 0; j < nv; ++j)
      if (j != i1) m2 = fmax(m2, sing[j]);
    const double w = 2.0 * (p.eps_rel * fmax(1.0, fabs(m2)) + eps_abs);
    int close2 = 0;
    for (int j = 0; j < nv; ++j)
      if (j != i1 && sing[j] >= m2 - w) ++close2;
    if (sing[i1] - m2 <= w || close2 >= 2)
      for (int j = 0; j < nv; ++j)
        if (sing[j] >= m2 - w) {
          if (n < RMAX) L->hyp[n] = (uint32_t)j;   // (j, 0, 0)
          ++n;
        }
    // decision thresholds (:837, :844) that fall inside their window want their operands exact as well: listing them
    // here lets finalise_kernel decide the cell on the device (same tests as the host finaliser, on the scanned values)
    if (!over && n <= RMAX && o.ncand >= 1) {
      auto add = [&](uint32_t h) {
        for (int q = 0; q < n && q < RMAX; ++q)
          if (L->hyp[q] == h) return;
        if (n < RMAX) L->hyp[n] = h;
        ++n;
      };
      auto nearw = [&](double a, double b) {
        return fabs(a - b) <= p.eps_rel * (fmax(1.0, fabs(a)) + fmax(1.0, fabs(b))) + 2.0 * eps_abs;
      };
      int i2 = -1;
      for (int j = 0; j < nv; ++j)
        if (j != i1 && (i2 < 0 || sing[j] > sing[i2])) i2 = j;
      const double s1 = sing[i1], s2 = i2 >= 0 ? sing[i2] : -INFINITY;
      bool want_pairs = false;
      if (nearw(best, s1 + 2.0)) { add((uint32_t)i1); want_pairs = true; }
      if (i2 >= 0 && nearw(s1, s2 + 2.0)) { add((uint32_t)i1); add((uint32_t)i2); }
      for (int c = 0; c < o.ncand; ++c) {
        const int j = o.cand[c] & 0xfff, k = (o.cand[c] >> 12) & 0xfff;
        if (nearw(best, sing[j])) { add((uint32_t)j); want_pairs = true; }
        if (nearw(best, sing[k])) { add((uint32_t)k); want_pairs = true; }
      }
      if (want_pairs)
        for (int c = 0; c < o.ncand; ++c) {
          const uint32_t h = o.cand[c];
          add(h);
          if (p.mode[h >> 24] == PM_SYM) add(((h >> 12) & 0xfffu) | ((h & 0xfffu) << 12) | (h & 0xff000000u));
        }
    }
    L->n = (over || n > RMAX) ? -1 : n;
    if (n > RESC_FAST && n <= RMAX && !over && p.any_long) *p.any_long = 1;   // (benign race: everybody writes 1)
  }
}

int launch_merge(const MergeParams& p, int64_t nactive, void* stream) {
  if (nactive == 0) return 0;
  merge_kernel<<<(unsigned)((nactive + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, nactive);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// finalise: one thread per active cell applies the reference's comparisons to the values of merge_kernel, with the
// re-scored (reference-exact) values of the cell's RescList in place of the scanned ones -- the logic of
// dmx_finalise.cpp:finalise_cell for the case that every value it wants exact is on the list.  Anything else
// (candidate overflow, a list that did not fit, a decision threshold inside its window that needs one more exact
// value) is flagged for the host finaliser, which re-derives the cell from the same arrays.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool fin_near(double a, double b, double eps_rel, double eps_abs) {
  return fabs(a - b) <= eps_rel * (fmax(1.0, fabs(a)) + fmax(1.0, fabs(b))) + 2.0 * eps_abs;
}

__global__ void __launch_bounds__(128) finalise_kernel(const __grid_constant__ FinaliseParams p, int64_t nactive) {
  const int64_t aidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (aidx >= nactive) return;
  const int cell = p.cell_order[aidx];
  const int nv = p.nv;
  const CellOut co = p.out[cell];
  const RescList* L = p.lists + cell;
  const int ln = L->n;
  if (co.ncand > MAX_CAND || co.ncand < 1 || ln < 0) { p.flag[cell] = 1; return; }
  const double* sing = p.sing + (int64_t)cell * nv;
  const double eps_abs = p.eps_ent * (double)(p.cell_ptr[cell + 1] - p.cell_ptr[cell]);
  // the listed singlets (j, 0, 0) carry exact values
  auto sval = [&](int j, bool* exact) -> double {
    for (int q = 0; q < ln; ++q)
      if (L->hyp[q] == (uint32_t)j) { *exact = true; return L->val[q]; }
    *exact = false;
    return sing[j];
  };
  bool any_sing = false;
  for (int q = 0; q < ln; ++q) any_sing |= L->hyp[q] < 0x1000u;
  // ---- were the singlets too close to call?  then merge_kernel listed every one inside the window (same test)
  {
    int i1 = 0;
    for (int j = 1; j < nv; ++j)
      if (sing[j] > sing[i1]) i1 = j;
    double m2 = -INFINITY;
    for (int j = 0; j < nv; ++j)
      if (j != i1) m2 = fmax(m2, sing[j]);
    const double w = 2.0 * (p.eps_rel * fmax(1.0, fabs(m2)) + eps_abs);
    int close2 = 0;
    for (int j = 0; j < nv; ++j)
      if (j != i1 && sing[j] >= m2 - w) ++close2;
    if (sing[i1] - m2 <= w || close2 >= 2) {
      for (int j = 0; j < nv; ++j)
        if (sing[j] >= m2 - w) {
          bool ex;
          sval(j, &ex);
          if (!ex) { p.flag[cell] = 1; return; }
        }
    }
  }
  // ---- top-2 singlets, :746-758
  int iS1 = -1, iS2 = -1;
  double max1 = -1e300, max2 = -1e300;
  bool x1 = false, x2 = false;
  for (int j = 0; j < nv; ++j) {
    bool ex = false;
    const double v = any_sing ? sval(j, &ex) : sing[j];
    if (max1 < v) { max2 = max1; iS2 = iS1; x2 = x1; iS1 = j; max1 = v; x1 = ex; }
    else if (max2 < v) { iS2 = j; max2 = v; x2 = ex; }
  }
  // ---- best pair, :799-814: the candidates in (j, k, n) lexicographic order, both orders of a symmetric plane
  uint32_t ch[2 * MAX_CAND];
  int nc = 0;
  for (int c = 0; c < co.ncand; ++c) {
    const uint32_t h = co.cand[c];
    ch[nc++] = h;
    if (p.mode[h >> 24] == PM_SYM) ch[nc++] = ((h >> 12) & 0xfffu) | ((h & 0xfffu) << 12) | (h & 0xff000000u);
  }
  auto key = [](uint32_t h) { return ((uint64_t)(h & 0xfffu) << 24) | ((uint64_t)((h >> 12) & 0xfffu) << 12) | (uint64_t)(h >> 24); };
  for (int a = 1; a < nc; ++a) {
    const uint32_t h = ch[a];
    int b = a - 1;
    while (b >= 0 && key(ch[b]) > key(h)) { ch[b + 1] = ch[b]; --b; }
    ch[b + 1] = h;
  }
  int jB = -1, kB = -1, nB = -1;
  double llk12 = co.best_v;
  bool llk12_exact = false;
  if (nc == 1) {
    jB = ch[0] & 0xfff; kB = (ch[0] >> 12) & 0xfff; nB = ch[0] >> 24;
    for (int q = 0; q < ln; ++q)   // listed by merge_kernel when a decision threshold is close
      if (L->hyp[q] == ch[0]) { llk12 = L->val[q]; llk12_exact = true; break; }
  } else {
    double mx = -1e300;
    for (int a = 0; a < nc; ++a) {
      double v = 0.0;
      bool have = false;
      for (int q = 0; q < ln; ++q)
        if (L->hyp[q] == ch[a]) { v = L->val[q]; have = true; break; }
      if (!have) { p.flag[cell] = 1; return; }
      if (mx < v) { mx = v; jB = ch[a] & 0xfff; kB = (ch[a] >> 12) & 0xfff; nB = ch[a] >> 24; llk12_exact = true; }
    }
    llk12 = mx;
  }
  if (iS1 < 0 || iS2 < 0 || jB < 0) { p.flag[cell] = 1; return; }   // NaN inputs: the host keeps the indices valid
  // ---- decision thresholds :837, :844: every comparison clear of its window, or between exact values
  bool xj = false, xk = false;
  const double l1 = any_sing ? sval(jB, &xj) : sing[jB], l2 = any_sing ? sval(kB, &xk) : sing[kB];
  {
    bool bad = false;
    if (fin_near(llk12, l1, p.eps_rel, eps_abs)) bad |= !(llk12_exact && xj);
    if (fin_near(llk12, l2, p.eps_rel, eps_abs)) bad |= !(llk12_exact && xk);
    if (fin_near(llk12, max1 + 2, p.eps_rel, eps_abs)) bad |= !(llk12_exact && x1);
    if (fin_near(max1, max2 + 2, p.eps_rel, eps_abs)) bad |= !(x1 && x2);
    if (bad) { p.flag[cell] = 1; return; }
  }
  dmx_best_rec r;
  r.status = 0; r.n_snp = (int32_t)(p.cell_ptr[cell + 1] - p.cell_ptr[cell]);
  r.sng1 = iS1; r.sng2 = iS2; r.dbl1 = jB; r.dbl2 = kB; r.alpha_idx = nB;
  r.sng_llk1 = max1; r.sng_llk2 = max2;
  r.sng_llk0 = p.l00[(int64_t)cell * p.nA];
  r.llk12 = llk12; r.llk1 = l1; r.llk2 = l2;
  r.llk10 = p.row0[((int64_t)cell * p.nA + nB) * nv + jB];   // :824 (nB >= 1: always a dense plane)
  r.llk20 = p.row0[((int64_t)cell * p.nA + nB) * nv + kB];   // :825
  r.llk00 = p.l00[(int64_t)cell * p.nA + nB];
  r.max_llk = co.max_llk; r.sum_single = co.sum_single; r.sum_double = co.sum_double;
  r.prb_dbl = __ddiv_rn(co.sum_double, __dadd_rn(co.sum_single, co.sum_double));   // :827
  r.prb_sng1 = exp(max1 - co.max_llk) * p.w_single / co.sum_single;                 // :828
  r.n_rescored = ln;
  r.exact_mask = 0;
  if (x1) r.exact_mask |= DMX_X_SNG1 | (jB == iS1 ? DMX_X_LLK1 : 0) | (kB == iS1 ? DMX_X_LLK2 : 0);
  if (x2) r.exact_mask |= DMX_X_SNG2;
  if (xj) r.exact_mask |= DMX_X_LLK1;
  if (xk) r.exact_mask |= DMX_X_LLK2;
  if (llk12_exact) r.exact_mask |= DMX_X_LLK12;
  if ((r.llk12 > r.llk1) && (r.llk12 > r.llk2) && (r.llk12 > r.sng_llk1 + 2)) r.call = 0;   // :837
  else if (r.sng_llk1 > r.sng_llk2 + 2) r.call = 1;                                           // :844
  else r.call = 2;
  p.best[cell] = r;
  p.flag[cell] = 0;
}

int launch_finalise(const FinaliseParams& p, int64_t nactive, void* stream) {
  if (nactive == 0) return 0;
  finalise_kernel<<<(unsigned)((nactive + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, nactive);
  return (int)cudaGetLastError();
}

// dense [aidx][j][k][0] = A_j + B_k for a separable plane 0 (what --write-pair prints for n = 0 is only
// the (j, 0, 0) column, but dmx_get_pair returns the whole llksAB block)
__global__ void fill_sep_dense_kernel(const __grid_constant__ ScanParams p, int64_t nactive) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)p.nv * p.nv;
  if (i >= nactive * per) return;
  const int64_t aidx = i / per;
  const int j = (int)((i - aidx * per) / p.nv), k = (int)(i % p.nv);
  const int cell = p.cell_order[aidx];
  p.dense[i * p.nA] = p.singA[(int64_t)cell * p.nv + j] + p.lb[(int64_t)cell * p.nv + k];
}

int launch_fill_sep_dense(const ScanParams& p, int64_t nactive, void* stream) {
  if (nactive == 0 || !p.sep0 || !p.dense) return 0;
  const int64_t n = nactive * p.nv * p.nv;
  fill_sep_dense_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, nactive);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// rescore: the listed hypotheses of every cell in the reference's exact arithmetic
// (cmd_cram_demuxlet.cpp:671-684: gpAB product, (l,m) row-major sum, log, SNP-ascending accumulation).
// Two phases, RESC_FAST hypotheses of every cell per pass (the common list is the two orders of the best pair):
//   rescore_terms_kernel  one block per cell, threads over its entries: log(S) of the listed hypotheses -> terms
//   rescore_sum_kernel    one warp per (cell, hypothesis): the sequential sum in entry order (the chain of dependent
//                         adds the reference's accumulation order demands), coalesced loads a batch ahead of it
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_term(const float* __restrict__ g, const double* __restrict__ q, uint32_t hy,
                                             const int8_t* dense_idx) {
  const int j = hy & 0xfff, k = (hy >> 12) & 0xfff, n = hy >> 24;
  const double gj0 = (double)__ldg(g + 3 * j), gj1 = (double)__ldg(g + 3 * j + 1), gj2 = (double)__ldg(g + 3 * j + 2);
  const double gk0 = (double)__ldg(g + 3 * k), gk1 = (double)__ldg(g + 3 * k + 1), gk2 = (double)__ldg(g + 3 * k + 2);
  double Q[9];
  if (dense_idx[n] < 0) {   // separable plane 0: Q'[0][l][m] = f[l]
#pragma unroll
    for (int l = 0; l < 3; ++l) Q[l * 3] = Q[l * 3 + 1] = Q[l * 3 + 2] = q[3 + l];
  } else {
    const double* qp = q + QHEAD + QPLANE * dense_idx[n];
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] = qp[i];
  }
  double s = 0.0;
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj0, gk0), Q[0]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj0, gk1), Q[1]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj0, gk2), Q[2]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj1, gk0), Q[3]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj1, gk1), Q[4]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj1, gk2), Q[5]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj2, gk0), Q[6]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj2, gk1), Q[7]));
  s = __dadd_rn(s, __dmul_rn(__dmul_rn(gj2, gk2), Q[8]));
  return glibc_log_fma(s);
}

template <int HM>
__device__ __forceinline__ void rescore_terms_cell(const RescoreParams& p, const uint32_t (&hyp)[RESC_FAST], int H, int64_t e0,
                                                   int64_t n, double* __restrict__ out) {
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t e = e0 + i;
    const float* g = p.gp + (int64_t)__ldg(p.ent_snp + e) * p.gp_stride;
    const double* q = p.entq + e * p.UQ;
    if (p.ecode) {
      const uint32_t code = __ldg(p.ecode + e);
      if (code != 0xffffu) q = p.memo_q + (int64_t)code * p.UQ;
    }
#pragma unroll
    for (int h = 0; h < HM; ++h)
      if (h < H) out[h * n + i] = exact_term(g, q, hyp[h], p.dense_idx);
  }
}

__global__ void __launch_bounds__(256) rescore_terms_kernel(const __grid_constant__ RescoreParams p, int64_t nactive,
                                                            int pass) {
  if (pass > 0 && p.any_long && !*p.any_long) return;        // no list of this launch set needs a further pass
  for (int64_t aidx = blockIdx.x; aidx < nactive; aidx += gridDim.x) {
    const int cell = p.cell_order[aidx];
    const RescList* L = p.lists + cell;
    const int H = min(L->n - pass * RESC_FAST, RESC_FAST);   // hypotheses of this pass
    if (H <= 0) continue;
    uint32_t hyp[RESC_FAST];
#pragma unroll
    for (int h = 0; h < RESC_FAST; ++h) hyp[h] = h < H ? L->hyp[pass * RESC_FAST + h] : 0u;
    const int64_t e0 = p.cell_ptr[cell], n = p.cell_ptr[cell + 1] - e0;
    double* out = p.terms + e0 * RESC_FAST;                  // this cell's block: [h][entry]
    // the common list is the two orders of the best pair: a body with two hypotheses inlined instead of eight
    if (H <= 2) rescore_terms_cell<2>(p, hyp, H, e0, n, out);
    else rescore_terms_cell<RESC_FAST>(p, hyp, H, e0, n, out);
  }
}

// one warp per cell: the lanes fetch 32 consecutive terms of each listed hypothesis at a time (coalesced, a batch
// ahead) and stage them in shared memory; lane h owns the chain of dependent additions of hypothesis h, in entry order
// as the reference's accumulation demands.  (Handing the terms to one lane by shuffle cost 148 M warp instructions for
// 20 k chains; a thread per chain with its own loads was latency bound at 1 ms.)
__global__ void __launch_bounds__(128) rescore_sum_kernel(const __grid_constant__ RescoreParams p, int64_t nactive,
                                                          int pass) {
  __shared__ double buf[4][RESC_FAST][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (pass > 0 && p.any_long && !*p.any_long) return;
  const int64_t aidx = (int64_t)blockIdx.x * 4 + w;
  if (aidx >= nactive) return;
  const int cell = p.cell_order[aidx];
  RescList* L = p.lists + cell;
  const int H = min(L->n - pass * RESC_FAST, RESC_FAST);   // hypotheses of this pass
  if (H <= 0) return;
  const int64_t e0 = p.cell_ptr[cell], n = p.cell_ptr[cell + 1] - e0;
  const double* x = p.terms + e0 * RESC_FAST;               // this cell's block: [h][entry]
  double cur[RESC_FAST];
#pragma unroll
  for (int h = 0; h < RESC_FAST; ++h) cur[h] = (h < H && lane < n) ? __ldcg(x + h * n + lane) : 0.0;
  double acc = 0.0;
  for (int64_t i = 0; i < n; i += 32) {
#pragma unroll
    for (int h = 0; h < RESC_FAST; ++h)
      if (h < H) buf[w][h][lane] = cur[h];
#pragma unroll
    for (int h = 0; h < RESC_FAST; ++h)
      if (h < H) cur[h] = (i + 32 + lane < n) ? __ldcg(x + h * n + i + 32 + lane) : 0.0;
    __syncwarp();
    const int cnt = (int)min((int64_t)32, n - i);
    if (lane < H) {
      if (cnt == 32) {
#pragma unroll
        for (int u2 = 0; u2 < 32; ++u2) acc = __dadd_rn(acc, buf[w][lane][u2]);
      } else {
        for (int u2 = 0; u2 < cnt; ++u2) acc = __dadd_rn(acc, buf[w][lane][u2]);
      }
    }
    __syncwarp();
  }
  if (lane < H) L->val[pass * RESC_FAST + lane] = acc;
}

int launch_rescore(const RescoreParams& p, int64_t nactive, void* stream) {
  if (nactive == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // lists longer than RESC_FAST take further passes over the same terms buffer (their blocks exit at once otherwise)
  for (int pass = 0; pass * RESC_FAST < RMAX; ++pass) {
    const int64_t grid = pass == 0 ? nactive : (nactive < 1184 ? nactive : 1184);
    rescore_terms_kernel<<<(unsigned)grid, 256, 0, st>>>(p, nactive, pass);
    rescore_sum_kernel<<<(unsigned)((nactive + 3) / 4), 128, 0, st>>>(p, nactive, pass);
  }
  return (int)cudaGetLastError();
}

}  // namespace dmx

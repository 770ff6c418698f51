// dmx_kernels.cu -- sm_100a kernels of the demuxlet likelihood hot path.
//
// What is computed (reference: /root/reference/cmd_cram_demuxlet.cpp):
//   entq_kernel       :424-452 + :597-663  per (cell, snp) entry: read likelihood table Q'(p) on the
//                     distinct mixing proportions p(n,l,m) (:613) and the pass-1 GLs
//   gp0_kernel        :391-401  cohort mean genotype probability per SNP
//   pair_scan_kernel  :455-459 + :665-710  per cell: sum over its entries of log(sum_{l,m} gp_j[l] gp_k[m]
//                     Q'[n][l][m]) for every sample pair (j,k) of every dense alpha plane, the
//                     separable alpha = 0 plane factors, llks00, and the pass-1 singlet llks;
//                     then the per-task reductions of :712-734 and :799-814
//   merge_kernel      combines the tasks of a cell (:712-734, :799-814)
//
// Design: the bilinear form is factored as S[j][k] = sum_l a_j[l] * w_k[l] with
// w_k[l] = sum_m Q'[l][m] gp_k[m] (3 FMA per hypothesis instead of 9 terms over a precomputed table)
// and sum-of-logs is replaced by the log of a running product whose exponent is peeled off into an
// integer every FLUSH_STAGES stages (1 MUL per hypothesis and entry, one log per hypothesis and cell).
// A CTA is warp specialised: 1 loader warp issues cp.async.bulk (TMA) copies of the gathered GP rows
// and entry tables into a raw ring; 2 transform warps turn them into (a, w) vectors in a processed ring;
// 5 consumer warps each own 4 x 4 tiles of hypotheses in registers.  Rings are mbarrier pipelines.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "dmx_internal.h"

namespace dmx {

__constant__ double c_err[256];   // phred2Err, PhredHelper.cpp:30-32 (computed on the host with libm pow)

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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
}

// peel the binary exponent of a positive double into an integer accumulator; mantissa stays in [1,2)
__device__ __forceinline__ void peel(double& x, int& esum) {
  int hi = __double2hiint(x);
  esum += hi >> 20;
  x = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));
}
__device__ __forceinline__ double unpeel_log(double x, int esum, int nflush) {
  return (double)(esum - 1023 * nflush) * 0.693147180559945309417232121458 + log(x);
}

// ---------------------------------------------------------------------------------------------
// gp0: cohort mean GP per SNP, cmd_cram_demuxlet.cpp:391-401 (ascending-j double sums, then / nv)
// ---------------------------------------------------------------------------------------------
__global__ void gp0_kernel(const float* __restrict__ gp, double* __restrict__ gp0, int64_t nsnps, int nv,
                           int nvp) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsnps) return;
  const float* g = gp + s * (int64_t)nvp * 3;
  double a = 0, b = 0, c = 0;
  for (int j = 0; j < nv; ++j) {
    a = __dadd_rn(a, (double)g[3 * j]);
    b = __dadd_rn(b, (double)g[3 * j + 1]);
    c = __dadd_rn(c, (double)g[3 * j + 2]);
  }
  gp0[s * 4 + 0] = __ddiv_rn(a, (double)nv);
  gp0[s * 4 + 1] = __ddiv_rn(b, (double)nv);
  gp0[s * 4 + 2] = __ddiv_rn(c, (double)nv);
  gp0[s * 4 + 3] = 0.0;
}

int launch_gp0(const float* gp, double* gp0, int64_t nsnps, int nv, int nvp, void* stream) {
  if (nsnps == 0) return 0;
  gp0_kernel<<<(unsigned)((nsnps + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gp, gp0, nsnps, nv, nvp);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// entq: per-entry read likelihood on the distinct p values (:597-663) and pass-1 GLs (:424-452).
// LPE lanes cooperate on one entry; lane `sub` owns p values sub, sub+LPE, ...  The arithmetic is the
// reference's (separately rounded mul/add/div, per-read max normalisation).
// ---------------------------------------------------------------------------------------------
template <int LPE, int VPL>
__global__ void __launch_bounds__(256) entq_kernel(const __grid_constant__ EntqParams p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPE;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t e = gw * (32 / LPE) + lane / LPE;
  if (e >= p.E) return;   // whole groups leave together
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
  const int64_t r0 = p.ent_ptr[e], r1 = p.ent_ptr[e + 1];
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t rd = p.reads[r];
    const uint32_t al = rd >> 8, bq = rd & 0xff;
    if (al == 2) continue;                                                  // :435, :604
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
    // pass-1 GLs, :437-443
    G0 = __dmul_rn(G0, pR);
    G1 = __dmul_rn(G1, __dsub_rn(0.5, e3));
    G2 = __dmul_rn(G2, pA);
    const double sm = __dadd_rn(__dadd_rn(G0, G1), G2);
    G0 = __ddiv_rn(G0, sm);
    G1 = __ddiv_rn(G1, sm);
    G2 = __ddiv_rn(G2, sm);
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
    if (ok[v]) out[sub + v * LPE] = __ddiv_rn(P[v], mx);                     // :656-663
  if (sub == 0) {
    G0 = __dadd_rn(G0, 1e-6);                                                // :446-452
    G1 = __dadd_rn(G1, 1e-6);
    G2 = __dadd_rn(G2, 1e-6);
    const double sm = __dadd_rn(__dadd_rn(G0, G1), G2);
    out[p.U] = __ddiv_rn(G0, sm);
    out[p.U + 1] = __ddiv_rn(G1, sm);
    out[p.U + 2] = __ddiv_rn(G2, sm);
    if (p.U + 3 < p.UQ) out[p.U + 3] = 0.0;
  }
}

int launch_entq(const EntqParams& p, void* stream) {
  if (p.E == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  auto grid = [&](int lpe) { return (unsigned)((p.E * lpe + 255) / 256); };
  if (p.U <= 8) entq_kernel<8, 1><<<grid(8), 256, 0, st>>>(p);
  else if (p.U <= 16) entq_kernel<16, 1><<<grid(16), 256, 0, st>>>(p);
  else if (p.U <= 32) entq_kernel<32, 1><<<grid(32), 256, 0, st>>>(p);
  else if (p.U <= 64) entq_kernel<32, 2><<<grid(32), 256, 0, st>>>(p);
  else entq_kernel<32, 4><<<grid(32), 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}

int set_err_table(const double* err256) {
  return (int)cudaMemcpyToSymbol(c_err, err256, 256 * sizeof(double));
}

// ---------------------------------------------------------------------------------------------
// pair scan
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
  uint32_t rowj_bytes, rowk_bytes;      // GP row slices per entry
  uint32_t off_gpj, off_gpk, off_gp0, off_q, raw_slot_bytes;
  uint32_t ent_a_dbl, ent_w_dbl;        // doubles per entry in the processed ring
  uint32_t off_raw, off_proc, proc_slot_bytes, total;
};

__host__ __device__ inline SmemLayout make_layout(int nj, int nk, bool same, int UQ) {
  SmemLayout L;
  L.rowj_bytes = (uint32_t)nj * 12u;
  L.rowk_bytes = same ? 0u : (uint32_t)nk * 12u;
  L.off_gpj = 0;
  L.off_gpk = L.off_gpj + NB * L.rowj_bytes;
  L.off_gp0 = L.off_gpk + NB * L.rowk_bytes;
  L.off_q = L.off_gp0 + NB * 32u;
  L.raw_slot_bytes = L.off_q + NB * (uint32_t)UQ * 8u;
  L.ent_a_dbl = (uint32_t)(nj / TILE) * GRP_DBL;
  L.ent_w_dbl = (uint32_t)(nk / TILE) * GRP_DBL;
  L.off_raw = 512;   // barriers + epilogue scratch
  L.off_proc = L.off_raw + S_RAW * L.raw_slot_bytes;
  L.proc_slot_bytes = NB * (L.ent_a_dbl + L.ent_w_dbl) * 8u;
  L.total = L.off_proc + S_PROC * L.proc_slot_bytes;
  return L;
}

struct Scratch {                     // lives at smem offset 128
  double red[2][NCW];
  double bcast[2];
  int ncand;
  uint32_t cand[MAX_CAND];
};

__global__ void __launch_bounds__(NTHREADS, 2) pair_scan_kernel(const __grid_constant__ ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* raw_empty = raw_full + S_RAW;
  uint64_t* proc_full = raw_empty + S_RAW;
  uint64_t* proc_empty = proc_full + S_PROC;
  Scratch* scr = reinterpret_cast<Scratch*>(smem + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t aidx = blockIdx.x / p.T;
  const int tti = blockIdx.x - (int)(aidx * p.T);
  const TaskType tt = p.tt[tti];
  const int cell = p.cell_order[aidx];
  const int64_t e0 = p.cell_ptr[cell], e1 = p.cell_ptr[cell + 1];
  const int nst = (int)((e1 - e0 + NB - 1) / NB);
  const bool tri = tt.flags & TF_TRI;
  const bool same = tri;
  const int nj = tt.nj, nk = tt.nk;
  const SmemLayout L = make_layout(nj, nk, same, p.UQ);
  const int ntj = nj / TILE, ntk = nk / TILE;
  const int ntiles = tri ? ntj * (ntj + 1) / 2 : ntj * ntk;
  const int ncw_active = (ntiles + 31) / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S_RAW; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], NTW);
    }
    for (int s = 0; s < S_PROC; ++s) {
      mbar_init(&proc_full[s], NTW);
      mbar_init(&proc_empty[s], ncw_active);
    }
    scr->ncand = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  unsigned char* raw_base = smem + L.off_raw;
  unsigned char* proc_base = smem + L.off_proc;

  if (warp == NCW + NTW) {
    // =============================== loader warp ===============================================
    const uint32_t per_entry = L.rowj_bytes + L.rowk_bytes + 32u + (uint32_t)p.UQ * 8u;
    const int64_t rowstride = (int64_t)p.nvp * 3;   // floats
    for (int it = 0; it < nst; ++it) {
      const int slot = it % S_RAW;
      const uint32_t ph = (uint32_t)(it / S_RAW) & 1u;
      const int64_t eb = e0 + (int64_t)it * NB;
      const int nvalid = (int)min((int64_t)NB, e1 - eb);
      int snp = 0;
      if (lane < nvalid) snp = __ldg(p.ent_snp + eb + lane);
      mbar_wait(&raw_empty[slot], ph ^ 1u);
      unsigned char* rs = raw_base + (size_t)slot * L.raw_slot_bytes;
      if (lane == 0) mbar_expect_tx(&raw_full[slot], (uint32_t)nvalid * per_entry);
      __syncwarp();
      if (lane < nvalid) {
        const float* row = p.gp + (int64_t)snp * rowstride;
        bulk_g2s(rs + L.off_gpj + lane * L.rowj_bytes, row + tt.j0 * 3, L.rowj_bytes, &raw_full[slot]);
        if (!same)
          bulk_g2s(rs + L.off_gpk + lane * L.rowk_bytes, row + tt.k0 * 3, L.rowk_bytes, &raw_full[slot]);
        bulk_g2s(rs + L.off_gp0 + lane * 32, p.gp0 + (int64_t)snp * 4, 32u, &raw_full[slot]);
      }
      if (lane == 0)
        bulk_g2s(rs + L.off_q, p.entq + eb * p.UQ, (uint32_t)nvalid * (uint32_t)p.UQ * 8u, &raw_full[slot]);
    }
  } else if (warp >= NCW) {
    // =============================== transform warps ===========================================
    const int tl = (warp - NCW) * 32 + lane;        // sample slot on each side
    const bool do_j = tl < nj, do_k = tl < nk;
    const bool extras = (tt.flags & TF_EXTRAS) && do_k;
    const bool scal = (tt.flags & (TF_SCALARS | TF_L00)) && tl == 0;
    const bool k_valid = (tt.k0 + tl) < p.nv;
    double PA = 1.0, PB = 1.0, P1 = 1.0;            // running products (extras)
    int eA = 0, eB = 0, e1x = 0;
    double PS0 = 1.0, PS1 = 1.0, PS2 = 1.0;         // scalars: llk0s, llks00[sep], llks00[n]
    int eS0 = 0, eS1 = 0, eS2 = 0;
    int nflush = 0;
    const int gj = tl >> 2, pj = tl & 3;
    for (int it = 0; it < nst; ++it) {
      const int rslot = it % S_RAW, pslot = it % S_PROC;
      const uint32_t rph = (uint32_t)(it / S_RAW) & 1u, pph = (uint32_t)(it / S_PROC) & 1u;
      const int nvalid = (int)min((int64_t)NB, e1 - (e0 + (int64_t)it * NB));
      mbar_wait(&raw_full[rslot], rph);
      mbar_wait(&proc_empty[pslot], pph ^ 1u);
      const unsigned char* rs = raw_base + (size_t)rslot * L.raw_slot_bytes;
      double* pa = reinterpret_cast<double*>(proc_base + (size_t)pslot * L.proc_slot_bytes);
      double* pw = pa + NB * L.ent_a_dbl;
#pragma unroll 2
      for (int e = 0; e < NB; ++e) {
        double* ad = pa + e * L.ent_a_dbl + gj * GRP_DBL + pj * 3;
        double* wd = pw + e * L.ent_w_dbl + gj * GRP_DBL + pj * 3;
        if (e < nvalid) {
          const double* q = reinterpret_cast<const double*>(rs + L.off_q) + e * p.UQ;
          const float* gkf = reinterpret_cast<const float*>(rs + (same ? L.off_gpj : L.off_gpk) +
                                                            e * (same ? L.rowj_bytes : L.rowk_bytes));
          double g0 = 0, g1 = 0, g2 = 0;
          if (do_k) {
            g0 = (double)gkf[tl * 3];
            g1 = (double)gkf[tl * 3 + 1];
            g2 = (double)gkf[tl * 3 + 2];
            const double q00 = q[tt.qidx[0]], q01 = q[tt.qidx[1]], q02 = q[tt.qidx[2]];
            const double q10 = q[tt.qidx[3]], q11 = q[tt.qidx[4]], q12 = q[tt.qidx[5]];
            const double q20 = q[tt.qidx[6]], q21 = q[tt.qidx[7]], q22 = q[tt.qidx[8]];
            wd[0] = fma(q02, g2, fma(q01, g1, q00 * g0));
            wd[1] = fma(q12, g2, fma(q11, g1, q10 * g0));
            wd[2] = fma(q22, g2, fma(q21, g1, q20 * g0));
          }
          if (do_j) {
            if (same) {
              ad[0] = g0; ad[1] = g1; ad[2] = g2;
            } else {
              const float* gjf = reinterpret_cast<const float*>(rs + L.off_gpj + e * L.rowj_bytes);
              ad[0] = (double)gjf[tl * 3];
              ad[1] = (double)gjf[tl * 3 + 1];
              ad[2] = (double)gjf[tl * 3 + 2];
            }
          }
          if (extras && k_valid) {
            if (p.sep0) PA *= fma(q[p.fidx[2]], g2, fma(q[p.fidx[1]], g1, q[p.fidx[0]] * g0));
            PB *= (g0 + g1) + g2;
            if (p.want_single) P1 *= fma(q[p.U + 2], g2, fma(q[p.U + 1], g1, q[p.U] * g0));
          }
          if (scal) {
            const double* z = reinterpret_cast<const double*>(rs + L.off_gp0) + e * 4;
            const double z0 = z[0], z1 = z[1], z2 = z[2];
            if (tt.flags & TF_SCALARS) {
              if (p.want_single) PS0 *= fma(q[p.U + 2], z2, fma(q[p.U + 1], z1, q[p.U] * z0));   // :459
              if (p.sep0)
                PS1 *= fma(q[p.fidx[2]], z2, fma(q[p.fidx[1]], z1, q[p.fidx[0]] * z0)) * ((z0 + z1) + z2);
            }
            if (tt.flags & TF_L00) {                                                            // :700-709
              const double t0 = fma(q[tt.qidx[2]], z2, fma(q[tt.qidx[1]], z1, q[tt.qidx[0]] * z0));
              const double t1 = fma(q[tt.qidx[5]], z2, fma(q[tt.qidx[4]], z1, q[tt.qidx[3]] * z0));
              const double t2 = fma(q[tt.qidx[8]], z2, fma(q[tt.qidx[7]], z1, q[tt.qidx[6]] * z0));
              PS2 *= fma(z2, t2, fma(z1, t1, z0 * t0));
            }
          }
        } else {   // identity entry: S = 1 for every hypothesis
          if (do_k) { wd[0] = 1.0; wd[1] = 0.0; wd[2] = 0.0; }
          if (do_j) { ad[0] = 1.0; ad[1] = 0.0; ad[2] = 0.0; }
        }
      }
      if ((it % FLUSH_STAGES) == FLUSH_STAGES - 1) {
        peel(PA, eA); peel(PB, eB); peel(P1, e1x);
        peel(PS0, eS0); peel(PS1, eS1); peel(PS2, eS2);
        ++nflush;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&proc_full[pslot]);
        mbar_arrive(&raw_empty[rslot]);
      }
    }
    if (extras && k_valid) {
      const int k = tt.k0 + tl;
      if (p.sep0) p.singA[(int64_t)cell * p.nv + k] = unpeel_log(PA, eA, nflush);
      p.lb[(int64_t)cell * p.nv + k] = unpeel_log(PB, eB, nflush);
      if (p.want_single) p.p1[(int64_t)cell * p.nv + k] = unpeel_log(P1, e1x, nflush);
    }
    if (scal) {
      if (tt.flags & TF_SCALARS) {
        if (p.want_single) p.p1_0[cell] = unpeel_log(PS0, eS0, nflush);
        if (p.sep0) p.l00[(int64_t)cell * p.nA] = unpeel_log(PS1, eS1, nflush);
      }
      if (tt.flags & TF_L00) p.l00[(int64_t)cell * p.nA + tt.n] = unpeel_log(PS2, eS2, nflush);
    }
  } else {
    // =============================== consumer warps ============================================
    const int t = warp * 32 + lane;
    const bool lane_active = t < ntiles;
    int tj = 0, tk = 0;
    if (lane_active) {
      if (tri) {
        int rem = t, row = 0;
        while (rem >= ntj - row) { rem -= ntj - row; ++row; }
        tj = row; tk = row + rem;
      } else {
        tj = t / ntk; tk = t - tj * ntk;
      }
    }
    double acc[TILE][TILE];
    int es[TILE][TILE];
#pragma unroll
    for (int a = 0; a < TILE; ++a)
#pragma unroll
      for (int b = 0; b < TILE; ++b) { acc[a][b] = 1.0; es[a][b] = 0; }
    int nflush = 0;
    if (warp < ncw_active) {
      for (int it = 0; it < nst; ++it) {
        const int pslot = it % S_PROC;
        const uint32_t pph = (uint32_t)(it / S_PROC) & 1u;
        mbar_wait(&proc_full[pslot], pph);
        const double* pa = reinterpret_cast<const double*>(proc_base + (size_t)pslot * L.proc_slot_bytes);
        const double* pw = pa + NB * L.ent_a_dbl;
        const double* ap = pa + tj * GRP_DBL;
        const double* wp = pw + tk * GRP_DBL;
#pragma unroll 2
        for (int e = 0; e < NB; ++e) {
          double a[12], w[12];
          const double2* a2 = reinterpret_cast<const double2*>(ap + e * L.ent_a_dbl);
          const double2* w2 = reinterpret_cast<const double2*>(wp + e * L.ent_w_dbl);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double2 x = a2[i], y = w2[i];
            a[2 * i] = x.x; a[2 * i + 1] = x.y;
            w[2 * i] = y.x; w[2 * i + 1] = y.y;
          }
#pragma unroll
          for (int jj = 0; jj < TILE; ++jj)
#pragma unroll
            for (int kk = 0; kk < TILE; ++kk) {
              const double s = fma(a[jj * 3 + 2], w[kk * 3 + 2],
                                   fma(a[jj * 3 + 1], w[kk * 3 + 1], a[jj * 3] * w[kk * 3]));
              acc[jj][kk] *= s;
            }
        }
        if ((it % FLUSH_STAGES) == FLUSH_STAGES - 1) {
#pragma unroll
          for (int jj = 0; jj < TILE; ++jj)
#pragma unroll
            for (int kk = 0; kk < TILE; ++kk) peel(acc[jj][kk], es[jj][kk]);
          ++nflush;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&proc_empty[pslot]);
      }
    }
    // ---- per-task reductions (:712-734, :799-814) ---------------------------------------------
    const bool sym = tt.flags & TF_SYM;
    const bool noarg = tt.flags & TF_NOARG;
    double v[TILE][TILE];
    uint32_t vmask = 0;
    double m_all = -INFINITY, best = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < TILE; ++jj)
#pragma unroll
      for (int kk = 0; kk < TILE; ++kk) {
        const int j = tt.j0 + tj * TILE + jj, k = tt.k0 + tk * TILE + kk;
        const bool valid = lane_active && j < p.nv && k < p.nv && (!tri || tj < tk || jj <= kk);
        const double x = valid ? unpeel_log(acc[jj][kk], es[jj][kk], nflush) : -INFINITY;
        v[jj][kk] = x;
        if (valid) vmask |= 1u << (jj * TILE + kk);
        m_all = fmax(m_all, x);
        if (j != k) best = fmax(best, x);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m_all = fmax(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
      best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    }
    if (lane == 0) { scr->red[0][warp] = m_all; scr->red[1][warp] = best; }
    consumer_bar_sync();
    m_all = scr->red[0][0]; best = scr->red[1][0];
#pragma unroll
    for (int wq = 1; wq < NCW; ++wq) {
      m_all = fmax(m_all, scr->red[0][wq]);
      best = fmax(best, scr->red[1][wq]);
    }
    const double thr = best - p.eps_rel * fmax(1.0, fabs(best));
    double sd = 0.0;
#pragma unroll
    for (int jj = 0; jj < TILE; ++jj)
#pragma unroll
      for (int kk = 0; kk < TILE; ++kk) {
        if (!(vmask & (1u << (jj * TILE + kk)))) continue;
        const int j = tt.j0 + tj * TILE + jj, k = tt.k0 + tk * TILE + kk;
        const double x = v[jj][kk];
        if (j != k) {
          sd += exp(x - m_all);
          if (!noarg && x >= thr) {
            const int slot = atomicAdd(&scr->ncand, 1);
            if (slot < MAX_CAND) scr->cand[slot] = (uint32_t)j | ((uint32_t)k << 12);
          }
        }
        // row k = 0 of the plane (LLK10 / LLK20, :824-825; singlets when plane 0 is dense)
        if (sym) {
          if (j == 0) p.row0[((int64_t)cell * p.nA + tt.n) * p.nv + k] = x;
        } else if (k == 0) {
          p.row0[((int64_t)cell * p.nA + tt.n) * p.nv + j] = x;
        }
        if (p.dense) {
          double* d = p.dense + (int64_t)aidx * p.nv * p.nv * p.nA;
          d[((int64_t)j * p.nv + k) * p.nA + tt.n] = x;
          if (sym) d[((int64_t)k * p.nv + j) * p.nA + tt.n] = x;
        }
      }
    if (sym) sd *= 2.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
    consumer_bar_sync();   // everyone has read red[] before it is reused
    if (lane == 0) scr->red[0][warp] = sd;
    consumer_bar_sync();
    if (t == 0) {
      double tot = 0.0;
      for (int wq = 0; wq < NCW; ++wq) tot += scr->red[0][wq];
      Part* out = p.parts + aidx * p.T + tti;
      out->m_all = m_all;
      out->sd = tot;
      out->best_v = best;
      out->ncand = scr->ncand;
      out->pad = 0;
      for (int c = 0; c < MAX_CAND; ++c) out->cand[c] = scr->cand[c];
    }
  }
}

size_t pair_scan_smem_bytes(const ScanParams& p, bool /*fp32*/) {
  // worst case over the two task shapes (tri 64 x 64 and rect 64 x 32), clipped to nvp
  const int nj = p.nvp < BLK ? p.nvp : BLK;
  const int nk = p.nvp < BLK_RECT_K ? p.nvp : BLK_RECT_K;
  const SmemLayout a = make_layout(nj, nj, true, p.UQ);
  const SmemLayout b = make_layout(nj, nk, false, p.UQ);
  return a.total > b.total ? a.total : b.total;
}

int pair_scan_init() {
  return (int)cudaFuncSetAttribute(pair_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int launch_pair_scan(const ScanParams& p, int64_t nactive, bool fp32, void* stream) {
  if (nactive == 0) return 0;
  const size_t smem = pair_scan_smem_bytes(p, fp32);
  const int64_t grid = nactive * p.T;
  if (grid > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  pair_scan_kernel<<<(unsigned)grid, NTHREADS, smem, (cudaStream_t)stream>>>(p);
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
  const double thr = best - p.eps_rel * fmax(1.0, fabs(best));
  CellOut o;
  o.ncand = 0;
  o.pad = 0;
  for (int c = 0; c < MAX_CAND; ++c) o.cand[c] = 0;
  for (int t = 0; t < p.T; ++t) {
    const TaskType& tt = p.tt[t];
    if (tt.flags & TF_NOARG) continue;
    sdb += parts[t].sd * exp(parts[t].m_all - mx) * p.w_double[tt.n];
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
}

int launch_merge(const MergeParams& p, int64_t nactive, void* stream) {
  if (nactive == 0) return 0;
  merge_kernel<<<(unsigned)((nactive + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, nactive);
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

}  // namespace dmx

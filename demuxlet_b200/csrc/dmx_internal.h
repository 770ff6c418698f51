// dmx_internal.h -- structures shared by the host side (dmx_api.cu, dmx_finalise.cpp) and the
// sm_100a kernels (dmx_kernels.cu).  Not part of the ABI.
#pragma once
#include <stdint.h>
#include "../../include/dmx.h"

namespace dmx {

// ---- geometry of the pair scan ---------------------------------------------------------------
constexpr int TILE = 4;        // each consumer lane owns a TILE x TILE block of (j,k) hypotheses
constexpr int BLK = 64;        // samples per block side (16 tiles); nv > 64 is tiled over blocks
constexpr int NB = 8;          // pileup entries per pipeline stage
constexpr int S_PROC = 3;      // stages of the processed (a, w) ring: producers -> consumers
constexpr int S_RAW = 4;       // stages of the raw ring: TMA bulk copies -> producers
constexpr int NPW = 2;         // producer warps (read likelihood + per-sample projections)
constexpr int MAX_CAND = 8;    // near-tie candidates kept per task / per cell
constexpr int MAX_U = DMX_MAX_ALPHA * 9;   // distinct mixing proportions p(n,l,m)

// ---- device layout of one (cell, snp) pileup entry: 16 bytes ------------------------------------
// Built on device by pack_entries_kernel from the CSR arrays of dmx_set_pileup.  Reads with
// allele == 2 carry no likelihood (cmd_cram_demuxlet.cpp:435,:604) and are dropped here; entries
// with more than 3 informative reads keep a pointer into the raw reads array instead.
struct PackedEnt {
  int32_t snp;
  uint16_t nreads;     // informative reads stored inline in r[] (0..3), or 0xFFFF = external
  uint16_t r[3];       // allele<<8 | bq      (external: r[0] | r[1]<<16 = raw read count)
  uint32_t roff;       // external: offset of the entry's first raw read
};
static_assert(sizeof(PackedEnt) == 16, "PackedEnt must be 16 bytes");

// ---- work decomposition -------------------------------------------------------------------------
// One CTA per task = (cell, alpha index n, sample block jb x kb).
enum : uint32_t {
  TF_SYM = 1u,       // alpha_n == 0.5: S[j][k] == S[k][j], only tiles with tj <= tk on diagonal blocks
  TF_DIAG = 2u,      // jb == kb
  TF_EXTRAS = 4u,    // also accumulate, for the samples of this block: singlet product of the
                     // separable n = 0 plane and the pass-1 likelihood (cmd_cram_demuxlet.cpp:455-458)
  TF_SCALARS = 8u,   // also llk0s (:459), llks00[0] when n = 0 is separable, and the k = 0 row-sum factor
  TF_L00 = 16u,      // also llks00[n] for this task's n (:700-709)
  TF_N0 = 32u,       // n == 0 handled densely (alpha[0] != 0): excluded from argmax / sumDouble
};

struct Task {
  int32_t cell;
  int16_t n;           // alpha index
  uint8_t jb, kb;      // block coordinates (units of BLK samples)
  uint32_t flags;
  int32_t part;        // index of this task's Part record
};
static_assert(sizeof(Task) == 16, "Task must be 16 bytes");

// partial reduction of one task over its hypotheses
struct Part {
  double m_all;        // max llk over all hypotheses of the task (incl. j == k)
  double sd;           // sum_{j != k} exp(llk - m_all) (both orders counted for TF_SYM)
  double best_v;       // max llk over j != k
  int32_t ncand;       // hypotheses within eps of best_v (may exceed MAX_CAND: overflow)
  int32_t pad;
  uint32_t cand[MAX_CAND];   // j | k << 12 (global sample ids; unordered j < k for TF_SYM)
};
static_assert(sizeof(Part) == 64, "Part must be 64 bytes");

// per-cell result of merge_kernel
struct CellOut {
  double max_llk, sum_single, sum_double, best_v;
  int32_t ncand;       // > MAX_CAND: overflow, the finaliser recomputes the cell exactly
  int32_t status;      // 0 ok, 1 filtered, 2 empty
  uint32_t cand[MAX_CAND];   // j | k << 12 | n << 24
};
static_assert(sizeof(CellOut) == 72, "CellOut must be 72 bytes");

struct Params {
  int32_t nv, nvp, nA, nu;
  int32_t sep0;                // alpha[0] == 0: the n = 0 plane is separable
  int32_t first_dense;         // first alpha index handled densely
  double eps;                  // candidate window
  double w_single;             // (1 - doublet_prior) / nv
  double w_double[DMX_MAX_ALPHA];   // doublet_prior / nv / (nv-1) / (nA-1) / (alpha == 0.5 ? 2 : 1)
  const float* gp;             // [nsnps][nvp*3], rows zero padded to a multiple of 16 bytes
  const double* gp0;           // [nsnps][4]: cohort mean (x, y, z) and the row sum of sample 0
  const PackedEnt* ent;
  const uint16_t* reads;
  const int64_t* cell_ptr;
  const Task* tasks;
  Part* parts;
  double* row0;                // [ncells][nA][nv]  llksAB[j][0][n]
  double* p1;                  // [ncells][nv]      pass-1 llks
  double* p1_0;                // [ncells]          pass-1 llk0s
  double* l00;                 // [ncells][nA]      llks00
  double* lb0;                 // [ncells]          sum log(row sum of sample 0) (n = 0 factor, Q3)
  double* dense;               // optional [cells][nv][nv][nA]
  int64_t dense_cell_lo;
  double pu[MAX_U];            // distinct p values
  uint8_t umap[DMX_MAX_ALPHA * 9];  // (n*9 + l*3 + m) -> index into pu
};

// host-callable launchers (dmx_kernels.cu)
struct LaunchCfg { int ncw; bool fp32; };
int launch_gp0(const float* gp, double* gp0, int64_t nsnps, int nv, int nvp, void* stream);
int launch_pack(const int32_t* ent_snp, const int64_t* ent_ptr, const uint16_t* reads, PackedEnt* out,
                int64_t E, void* stream);
int launch_pair_scan(const Params& p, int64_t task_lo, int64_t ntasks, int ncw, bool fp32, bool offdiag,
                     void* stream);
int launch_merge(const Params& p, const int64_t* cell_task_ptr, const int32_t* totl, const int32_t* uniq,
                 int32_t min_total, int32_t min_uniq, int32_t min_snp, CellOut* out, int64_t ncells,
                 void* stream);
int set_err_table(const double* err256);
size_t pair_scan_smem_bytes(bool fp32, bool offdiag);

}  // namespace dmx

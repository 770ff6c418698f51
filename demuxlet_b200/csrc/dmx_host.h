// dmx_host.h -- host-side pieces of libdmx that need no CUDA: the exact (reference-arithmetic) scorer
// used by the tie finaliser, the record assembly, and the TSV writers.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "dmx_internal.h"

namespace dmx {

// distinct mixing proportions of an alpha grid: p(n,l,m) = 0.5*l + (m-l)*0.5*alpha[n]
// (cmd_cram_demuxlet.cpp:613), deduplicated bit-wise.  umap[n*9 + l*3 + m] -> index into pu.
struct PGrid {
  int U = 0;
  double pu[MAX_U];
  uint8_t umap[DMX_MAX_ALPHA * 9];
  uint8_t mode[DMX_MAX_ALPHA];   // PM_*
  int sep0 = 0;
};
void build_pgrid(const double* alpha, int nA, PGrid* g);

// read-only view of the host copies of the inputs
struct HostView {
  int32_t nv = 0, nA = 0;
  const double* alpha = nullptr;
  const float* gp = nullptr;         // [nsnps][nv][3]
  int64_t nsnps = 0, ncells = 0;
  const int64_t* cell_ptr = nullptr;
  const int32_t* ent_snp = nullptr;
  const int64_t* ent_ptr = nullptr;
  const uint16_t* reads = nullptr;
};

struct Hyp { int32_t j, k, n; };

// Evaluates llksAB[j][k][n] of one cell with exactly the reference's arithmetic
// (cmd_cram_demuxlet.cpp:597-684: IEEE double, separately rounded mul/add, (l,m) row-major, UMI and SNP
// order, libm log) -- what the reference would print on this host.
class ExactScorer {
 public:
  ExactScorer(const HostView& hv, const double* err256);
  void eval(int64_t cell, const Hyp* hyps, int nh, double* out) const;

 private:
  void entry_pgs(int64_t e, double* pGs) const;
  HostView hv_;
  double err_[256], mat_[256];
  std::vector<double> memo1_;        // [2][256][nA*9]: entries with exactly one informative read
};

// GPU results of one cell, as read back
struct CellGpu {
  const CellOut* co;
  const double* sing;    // [nv]  llksAB[j][0][0]
  const double* row0;    // [nA][nv] llksAB[j][0][n] (dense planes)
  const double* l00;     // [nA]
};

// candidate pairs of a cell whose GPU candidate list overflowed: taken from a dense llksAB block
void candidates_from_dense(const double* dense, int nv, int nA, int first_arg_plane, double eps_rel, double eps_abs,
                           std::vector<Hyp>* out);

// :746-758, :799-857 with near-ties resolved by ExactScorer.  `cands` = ordered candidate hypotheses
// within the window of the GPU's best pair.  Window of a value x: eps_rel * max(1, |x|) + eps_abs, where eps_abs is
// the cell's absolute error allowance (entries x per-entry bound of the fp32 pair scan; 0 for the fp64 scan).
void finalise_cell(const HostView& hv, const ExactScorer& sc, int64_t cell, const CellGpu& g,
                   const std::vector<Hyp>& cands, double doublet_prior, double eps_rel, double eps_abs, bool exact,
                   const RescList* gpu_exact, dmx_best_rec* rec, int32_t* n_from_gpu);

// the host libm's __log_data (ln2hi, ln2lo, poly[5], poly1[11], tab[128]{invc, logc}) read from the
// libm.so.6 image this process has loaded; false if it cannot be located
bool find_libm_log_table(double out[18 + 256]);

}  // namespace dmx

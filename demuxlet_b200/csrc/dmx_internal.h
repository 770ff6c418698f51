// dmx_internal.h -- structures shared by the host side (dmx_api.cu, dmx_finalise.cpp) and the
// sm_100a kernels (dmx_kernels.cu).  Not part of the ABI.
#pragma once
#include <stdint.h>
#include <cuda.h>   // CUtensorMap (type only; the encoder is looked up at run time)
#include "../../include/dmx.h"

namespace dmx {

// ---- geometry of the pair scan ---------------------------------------------------------------
constexpr int TILE = 4;        // each consumer lane owns a TILE x TILE block of (j,k) hypotheses
constexpr int BLK = 64;        // samples per task side (16 tiles); nv > 64 is tiled over tasks
constexpr int BLK_RECT_K = 32; // k-side samples of a rectangular task (16 x 8 tiles = 128 lanes)
constexpr int NB = 8;          // pileup entries per pipeline stage (power of two, <= 8)
constexpr int S_RAW = 4;       // stages of the raw ring: bulk (TMA) copies -> transform warps
constexpr int S_PROC = 2;      // stages of the processed (a, w) ring: transform -> consumer warps
constexpr int NCW = 5;         // consumer warps per slot (136 tiles of the 64 x 64 upper triangle -> 160 lanes)
constexpr int NTW = 2;         // transform warps per slot (one lane per sample of each side)
constexpr int NSLOT = 2;       // independent task pipelines per CTA (one persistent CTA per SM)
constexpr int NTHREADS = NSLOT * (NCW + NTW + 1) * 32;   // + 1 loader warp per slot = 512
constexpr int QHEAD = 8;       // entry record head: 3 pass-1 GLs, 3 values of the separable plane, the entry's llk0s
                               // factor, its llks00 factor of a separable plane 0
constexpr int QPLANE = 10;     // doubles per dense plane in the entry record: 9 Q' values + the entry's llks00[n] factor
constexpr int GRP_DBL = 14;    // doubles per 4-sample group in the processed ring (12 + 2 pad: 112 B
                               // stride spreads 16-byte lane loads over all banks)
constexpr int GRP_F32 = 12;    // floats per 4-sample group in the fp32 processed ring (48 B stride)
constexpr int FLUSH_STAGES = 16 / NB;   // exponent renormalisation period (stages): every 16 entries
// fp32 pair scan (pair_scan_f32_kernel): one task pipeline per CTA; a consumer thread owns an 8 x 8 tile of
// hypotheses for ONE of the NB entries of a stage (lane & 7 = entry, lane >> 3 = tile position inside the warp)
constexpr int F_NCW = 9;       // consumer warps: 36 tile positions of the 64 x 64 upper triangle x 8 entry lanes
#ifndef DMX_F_NTW
#define DMX_F_NTW 4
#endif
constexpr int F_NTW = DMX_F_NTW;   // transform warps (whole stages, taken dynamically)
constexpr int F_NLOAD = 2;     // loader warps (stages round-robin); the raw ring sizes are multiples of it, which keeps one
                               // producer and one consumer warp per ring slot (mbarrier parity waits must not alias)
#ifndef DMX_S_PROC
#define DMX_S_PROC 8
#endif
constexpr int F_S_PROC = DMX_S_PROC;    // stages of the processed ring
constexpr int F_PEEL_STAGES = 4;   // underflow check period (stages = entries per thread)
constexpr int FHEAD = 8;       // floats per entry in the fp32 head table: GL[3], f[3] (separable plane), 2 pad
constexpr int FPLANE = 12;     // floats per entry and dense plane in the fp32 plane tables: Q'[3][3], 3 pad
// Entry classes of the fp32 scan.  An entry with at most one informative read (allele 0 or 1) has a read likelihood
// that is AFFINE in the mixing proportion p(n,l,m) (cmd_cram_demuxlet.cpp:613,:625), hence in (l, m); with the
// genotype vectors normalised to sum 1 the bilinear form collapses to S'[j][k] = X_j + Y_k, X_j = h + cj * D_j,
// Y_k = h + ck * D_k, D = expected dosage of the read's allele (all terms positive).  Such "affine" entries (4 in 5 of
// a droplet pileup) are moved to the front of their cell in the fp32 tables and scanned in stages of their own: one
// add and one multiply per hypothesis instead of a K = 3 dot product.  Their fp32 records:
//   head   h1 c1 (pass 1, :455-458)  hs cs (separable plane)  flag (1: ref read, D = ref dosage)  3 pad
//   plane  h cj ck, 9 pad
constexpr int MEMO_ROWS = 513; // rows of the affine-entry memo tables
constexpr int A_BLK4_FLT = 36; // floats per 4-sample block of an affine stage in the processed ring: 8 entries x 4 + pad
constexpr int MAX_CAND = 8;    // near-tie candidates kept per task / per cell
constexpr int MAX_U = DMX_MAX_ALPHA * 6 + 3;   // distinct mixing proportions p(n,l,m)

// plane modes
enum : uint8_t {
  PM_SEP = 0,   // alpha == 0 at index 0: S[j][k] = A_j * B_k, no pair work
  PM_SYM = 1,   // alpha == 0.5: S[j][k] == S[k][j], only j <= k is computed
  PM_FULL = 2,  // everything else: all nv x nv ordered pairs
};

// ---- work decomposition -------------------------------------------------------------------------
// One CTA per task = (cell, task type); a task type = (alpha index n, sample block j0.. x k0..).
enum : uint32_t {
  TF_TRI = 1u,       // 64 x 64 diagonal block of a PM_SYM plane: tiles with tj <= tk only
  TF_SYM = 2u,       // plane is PM_SYM (unordered pairs count twice in sumDouble)
  TF_EXTRAS = 4u,    // transform lanes also accumulate, for the k-side samples of this task: the singlet
                     // factor of the separable n = 0 plane, the row sum factor, and the pass-1
                     // likelihood (cmd_cram_demuxlet.cpp:455-458)
  TF_SCALARS = 8u,   // also llk0s (:459) and llks00[0] of a separable plane 0
  TF_L00 = 16u,      // also llks00[n] for this task's plane (:700-709)
  TF_NOARG = 32u,    // plane is alpha index 0 handled densely (alpha[0] != 0): excluded from the
                     // best-pair argmax and from sumDouble (:728,:805 start at n = 1)
  TF_ROW0 = 64u,     // the task holds pairs with sample 0 (llksAB[x][0][n]: LLK10 / LLK20, :824-825, and the singlets of
                     // a dense plane 0): symmetric planes as (0, x) in tasks with j0 == 0, full planes as (x, 0) in
                     // tasks with k0 == 0.  The fp32 scan accumulates them in its transform warps.
};

struct TaskType {
  int16_t n;           // alpha index of the plane
  int16_t j0, k0;      // first sample of each side
  int16_t nj, nk;      // samples on each side, padded to a multiple of TILE
  int16_t qoff;        // offset (doubles) of the plane's 9 Q' values inside the entry record
  uint32_t flags;
};
static_assert(sizeof(TaskType) == 16, "TaskType layout");

// partial reduction of one task over its hypotheses
struct Part {
  double m_all;        // max llk over all hypotheses of the task (incl. j == k)
  double sd;           // sum_{j != k} exp(llk - best_v) (both orders counted for TF_SYM)
  double best_v;       // max llk over j != k
  int32_t ncand;       // hypotheses within eps of best_v (may exceed MAX_CAND: overflow)
  int32_t pad;
  uint32_t cand[MAX_CAND];   // j | k << 12 (global sample ids; j < k stands for both orders if TF_SYM)
};
static_assert(sizeof(Part) == 64, "Part must be 64 bytes");

// per-cell result of merge_kernel
struct CellOut {
  double max_llk, sum_single, sum_double, best_v;
  int32_t ncand;       // > MAX_CAND: overflow, the finaliser re-derives the candidates from a dense scan
  int32_t pad;
  uint32_t cand[MAX_CAND];   // j | k << 12 | n << 24
};
static_assert(sizeof(CellOut) == 72, "CellOut must be 72 bytes");

struct ScanParams {
  int32_t nv, nvp, nA, UQ;
  int32_t sep0;                // alpha[0] == 0: the n = 0 plane is separable
  int32_t T;                   // task types per cell
  int32_t want_single;         // accumulate pass 1
  int32_t q_hp, q_po, q_pp;    // raw-stage layout of the entry records: head pitch, plane offset, plane pitch (bytes)
  int32_t dbg;                 // DMX_DEBUG: timing experiments of the fp32 scan (1: skip consumer math, 2: skip
                               // transform math, 4: no per-sample extras, 16: no pairs with sample 0); results are wrong
  int32_t gp_stride;           // floats per GP row in HBM (8 + nvp*3); gp points at the first sample of row 0
  int64_t ntasks;              // nactive * T
  double eps_rel;              // candidate window, relative to max(1, |best|) ...
  double eps_ent;              // ... plus this much per pileup entry of the cell (fp32 pair scan error bound)
  const float* gp;             // [nsnps][gp_stride] (first sample of row 0); the 3 doubles in front of a row are gp0
  const int32_t* ent_snp;      // [E]
  const double* entq;          // [E][UQ]: GL[3], sep-plane f[3], then QPLANE doubles per dense plane
  const float* entf_head;      // [E][FHEAD] fp32 copy of the record heads (fp32 pair scan)
  const float* entf_plane;     // [nd][entf_pstride] fp32 copies of the dense planes, FPLANE floats per entry
  int64_t entf_pstride;        // floats between consecutive planes
  int32_t has_rect;            // the task list has rectangular tasks (k-side rows needed)
  int32_t f_sraw, f_rawslot, f_off_fh;   // fp32 scan raw ring geometry (filled in by launch_pair_scan)
  const int32_t* snp_perm;     // [E] row index of the entry at each position of the class-sorted fp32 tables: SNP id
                               // of a general entry, SNP * 2 + (alt read) of an affine entry (row of the dosage table)
  const double* lm;            // [ncells][nA] sum over the cell's affine entries of log M (scale of the affine form)
  int32_t v_nslot, v_slot;     // second-generation fp32 scan: ring slots and bytes per slot (launch_pair_scan)
  const int32_t* cell_na;      // [ncells] affine entries of the cell (they come first in the fp32 tables)
  const int64_t* cell_ptr;     // [ncells+1]
  const int32_t* cell_order;   // [nactive] cell ids, scheduled order
  const TaskType* tt;          // [T]
  Part* parts;                 // [nactive][T]
  double* row0;                // [ncells][nA][nv]  llksAB[j][0][n] of dense planes
  double* singA;               // [ncells][nv]      separable plane: sum log A_j
  double* lb;                  // [ncells][nv]      separable plane: sum log B_k (row sums)
  double* p1;                  // [ncells][nv]      pass-1 llks
  double* p1_0;                // [ncells]          pass-1 llk0s
  double* l00;                 // [ncells][nA]      llks00
  double* dense;               // optional [nactive][nv][nv][nA]
};

// TMA descriptors of the pair scan's inputs (built by dmx_api.cu:encode_maps, passed as a __grid_constant__).
// A GP row in HBM is [gp0: 3 doubles + 8 B pad][nvp x 3 floats] (gprow_kernel).
struct ScanMaps {
  CUtensorMap gpj;   // float32 [nsnps][8 + nvp*3], box {200, 1}: tile::gather4 of 4 SNP rows x (gp0 | 64 samples)
  CUtensorMap gpk;   // same tensor, box {104, 1}: 4 SNP rows x (8 floats | 32 samples), k side of rectangular tasks
  CUtensorMap q;     // float64 [E][UQ], box {16, NB}: whole records (single dense plane)
  CUtensorMap qh;    // float64 [E][UQ], box {QHEAD, NB}:  record heads of NB consecutive entries
  CUtensorMap qp;    // float64 [E][UQ], box {QPLANE, NB}: one plane of NB consecutive entries
  // second-generation fp32 scan (dmx_scan2.cuh): dosage rows [2 nsnps][nvp + 8] and row sums [nsnps][nvp + 8]
  // (dtab_kernel; sample j at column 4 + j), gather4 boxes of {72, 1} (64 samples) and {40, 1} (32 samples, k side)
  CUtensorMap dj, dk, sj, sk;
};

// exact re-scoring requests of one cell (built by merge_kernel, filled by rescore_kernel)
constexpr int RMAX = 32;
constexpr int RESC_FAST = 8;   // hypotheses per cell and pass of the two-phase re-scorer
struct RescList {
  int32_t n;                   // hypotheses listed; -1: too many (the host scorer handles the cell)
  int32_t pad;
  uint32_t hyp[RMAX];          // j | k << 12 | n << 24
  double val[RMAX];            // llksAB[j][k][n] in the reference's arithmetic
};
static_assert(sizeof(RescList) == 392, "RescList layout");

struct RescoreParams {
  int32_t nv, gp_stride, nA, UQ, sep0;
  int8_t dense_idx[DMX_MAX_ALPHA];
  const float* gp;
  const int32_t* ent_snp;
  const double* entq;
  const uint16_t* ecode;       // fp32 scan: memo row per entry (0xffff: the entry has its own record in entq), or null
  const double* memo_q;
  const int64_t* cell_ptr;
  const int32_t* cell_order;
  RescList* lists;             // [ncells]
  double* terms;               // per cell [RESC_FAST][entries] log(S) terms, cell block at e0 * RESC_FAST
  const int32_t* any_long;     // != 0: some list of this launch set is longer than RESC_FAST (merge_kernel); the passes
                               // after the first return at once otherwise
};

struct MergeParams {
  int32_t nv, nA, T, sep0;
  double eps_rel, eps_ent;     // candidate window: eps_rel * max(1, |x|) + eps_ent * entries of the cell
  const int64_t* cell_ptr;
  double w_single;             // (1 - doublet_prior) / nv
  double w_double[DMX_MAX_ALPHA];   // doublet_prior / nv / (nv-1) / (nA-1) / (alpha == 0.5 ? 2 : 1)
  const int32_t* cell_order;
  const TaskType* tt;
  const Part* parts;
  const double* row0;
  const double* singA;
  const double* lb;
  double* sing;                // [ncells][nv] out: llksAB[j][0][0]
  CellOut* out;                // [ncells]
  RescList* lists;             // [ncells] or null: exact re-scoring requests
  int32_t* any_long;           // set when a list is longer than RESC_FAST hypotheses
  uint8_t mode[DMX_MAX_ALPHA]; // PM_* per alpha index
};

// device half of the decision logic (finalise_kernel): the reference's comparisons (:746-758, :799-814, :837, :844)
// on the values of merge_kernel with the re-scored ones of the RescList in place; a cell whose decision needs a value
// that is neither exact nor clear of its window is left to the host finaliser (flag = 1)
struct FinaliseParams {
  int32_t nv, nA;
  double eps_rel, eps_ent;
  double w_single;             // (1 - doublet_prior) / nv
  uint8_t mode[DMX_MAX_ALPHA];
  const int64_t* cell_ptr;
  const int32_t* cell_order;
  const double* sing;          // [ncells][nv]
  const double* row0;          // [ncells][nA][nv]
  const double* l00;           // [ncells][nA]
  const CellOut* out;          // [ncells]
  const RescList* lists;       // [ncells]
  dmx_best_rec* best;          // [ncells] out: the whole record (PRB.SNG1 with the device's exp())
  uint8_t* flag;               // [ncells] out: 1 = the host finaliser takes the cell
};
int launch_finalise(const FinaliseParams& p, int64_t nactive, void* stream);

struct EntqParams {
  int32_t U, UQ;
  int32_t nd, sep0;            // dense planes; alpha[0] == 0
  int32_t gp_stride;           // floats per GP row
  const float* gprow;          // GP rows with the gp0 prefix (gprow_kernel)
  const int32_t* ent_snp;
  int64_t E0, E;               // entries [E0, E) of this launch
  const int64_t* ent_ptr;
  const uint16_t* reads;
  double* entq;
  float* entf_head;            // [E][FHEAD] or null
  float* entf_plane;           // [nd][entf_pstride]
  int64_t entf_pstride;
  const int32_t* perm;         // [E] position of entry e in the class-sorted fp32 tables (classify_kernel)
  int32_t* snp_perm;           // [E] out: SNP id by position
  // fp32 scan: affine entries (<= 1 informative read) take their records from a 513-row memo table -- every record
  // value depends only on (allele, quality) of that read -- built once per context by the same kernels on a synthetic
  // pileup (rows allele * 256 + quality, row 512 = no informative read); entq_aff_kernel handles them and the general
  // kernels run over the list of the remaining entries
  uint16_t* ecode;             // [E] out: memo row of an affine entry, 0xffff for a general entry
  const double* memo_q;        // [MEMO_ROWS][UQ]
  const float* memo_fh;        // [MEMO_ROWS][FHEAD]
  const float* memo_fp;        // [nd][MEMO_ROWS][FPLANE]
  const int32_t* glist;        // [E] entry at each class-sorted position, sign bit set for a general entry
                               // (classify_kernel): the general kernels then run by position; null: all entries
  double* ents;                // [E][2 + 2 nd]: the entry's factors of llk0s (:459), llks00 of a separable plane 0,
                               // llks00[n] of every dense plane (:700-709), and the scale M of its affine form per
                               // dense plane (1 for a general entry); their logs are summed per cell by cell_scalars_kernel
  double pu[MAX_U];            // distinct p values (cmd_cram_demuxlet.cpp:613)
  uint8_t oslot_ptr[MAX_U + 1];   // CSR: record slots fed by distinct value u
  uint8_t oslot[QHEAD + QPLANE * DMX_MAX_ALPHA];
  uint8_t slot_u[QHEAD + QPLANE * DMX_MAX_ALPHA];   // inverse map: distinct value feeding record slot s
};

struct CellScalarParams {
  int32_t nd, nA, sep0, want_single;
  int8_t plane_n[DMX_MAX_ALPHA];   // alpha index of dense plane d
  const double* ents;              // [E][2 + 2 nd]
  double* lm;                      // [ncells][nA] sum of log M over the affine entries, per dense plane
  const int64_t* cell_ptr;
  const int32_t* cell_order;
  double* p1_0;                    // [ncells]
  double* l00;                     // [ncells][nA]
};

// VCF field -> genotype probabilities (dmx_field.cu; bcf_filtered_reader.cpp:186-320, :360-454)
struct FieldParams {
  int32_t field, nv;           // DMX_FIELD_*
  int64_t nsnps;
  const void* in;              // device: GT int8[nsnps][nv][2], GP float[nsnps][nv][3], PL int32[nsnps][nv][3]
  float* out;                  // device: [nsnps][out_stride], the first nv * 3 floats of a row are written
  int64_t out_stride;
  double* af;                  // device [nsnps] or null: get_af(1) (GT / PL)
  double gt_error;
  const double* prob;          // device [256]: phred2Prob (PhredHelper.cpp:32), PL only
  int* err_flag;               // device: set when a GT allele index is > 1
  uint8_t* pass;               // device [nsnps] or null (GT only): the reference's site filter, with ...
  int32_t min_mac;             // ... --min-mac (cmd_cram_demuxlet.cpp:24, :47) and
  double min_callrate;         // ... --min-callrate (:25, :48)
  int32_t group, pitch;        // filled by launch_field: SNPs per warp pass, shared-memory row pitch (elements)
};
int launch_field(const FieldParams& p, int sm_count, void* stream);

// host-callable launchers (dmx_kernels.cu); all return a cudaError_t as int
int launch_gprow(const float* gp_flat, float* gp_rows, int64_t nsnps, int nv, int nvp, void* stream);
// dosage / row-sum tables of the second-generation fp32 scan from the GP rows: drow [2 nsnps][nvp + 8], srow [nsnps][nvp + 8]
int launch_dtab(const float* gp_rows, float* drow, float* srow, int64_t nsnps, int nvp, void* stream);
int launch_entq(const EntqParams& p, void* stream);
// entry classes of cells [c0, c1): perm[e] = position of entry e inside its cell's block with the affine entries
// (<= 1 informative read) first, both classes in entry order; cell_na[cell] = number of affine entries
// ent_ptr[e] of cells [c0, c1) from the per-entry read counts (what dmx_run sends instead of the 8-byte offsets):
// ent_ptr[e] = cell_r[cell] + sum of cnt over the cell's entries before e; also the end of every cell
int launch_entptr(const int64_t* cell_ptr, const int64_t* cell_r, const uint8_t* cnt, int64_t c0, int64_t c1,
                  int64_t* ent_ptr, void* stream);
int launch_classify(const int64_t* cell_ptr, const int64_t* ent_ptr, const uint16_t* reads, int64_t c0, int64_t c1,
                    int32_t* perm, int32_t* cell_na, int32_t* glist, void* stream);
int launch_entq_aff(const EntqParams& p, void* stream);
int launch_pair_scan(const ScanParams& p, const ScanMaps& tm, int fp32, void* stream);   // fp32: 0 fp64 scan, 1 first-,
                                                                                           // 2 second-generation fp32 scan
int launch_cell_scalars(const CellScalarParams& p, int64_t nactive, void* stream);
int launch_merge(const MergeParams& p, int64_t nactive, void* stream);
int launch_fill_sep_dense(const ScanParams& p, int64_t nactive, void* stream);
int launch_rescore(const RescoreParams& p, int64_t nactive, void* stream);
int set_err_table(const double* err256, const double* err3_256);
int set_log_table(const double* tab274);          // glibc __log_data: ln2hi, ln2lo, A[5], B[11], tab[128][2]
int launch_logtest(const double* x, double* y, int64_t n, void* stream);
int pair_scan_init();          // func attributes (max dynamic smem)
size_t pair_scan_smem_bytes(const ScanParams& p, int fp32);

}  // namespace dmx

/*
 * dmx_oracle.h -- CPU ORACLE for the demuxlet likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker / CPU baseline.
 *
 * What it is: a plain-C restatement, on the flat CSR pileup boundary
 * (include/dmx.h), of the reference's singlet + doublet likelihood scan
 *   /root/reference/cmd_cram_demuxlet.cpp:386-875
 * with the phred tables of /root/reference/PhredHelper.cpp:30-32.  Same
 * operation order, IEEE double, no FMA contraction (build with
 * -O2 -ffp-contract=off), libm log/exp, the reference's printf formats.
 * The only intentional deviation: 64-bit indexing and an on-the-fly
 * gps[j][l]*gps[k][m] product instead of the materialised gpAB table
 * (cmd_cram_demuxlet.cpp:542-560) -- a single rounded product, bit-identical,
 * and free of the reference's int32 overflow at nsnps*nv*nv*9 >= 2^31.
 *
 * Parity pin: the reference has no tests or golden vectors.  The oracle is
 * pinned against oracle/_ref (the reference's own lines 386-881 compiled
 * from /root/reference by oracle/Makefile) byte-for-byte on the four output
 * files; see tests/test_oracle_vs_ref.py and tests/golden/.
 */
#ifndef DMX_ORACLE_H
#define DMX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t nv;                 /* number of VCF samples (>=2)                     */
  int32_t n_alpha;            /* size of the alpha grid (>=2)                    */
  const double* alpha;        /* [n_alpha]; alpha[0] is the "singlet" slot       */
  double doublet_prior;
  int32_t min_total, min_uniq, min_snp;
  int64_t nsnps;
  const float* gp;            /* [nsnps][nv][3], genotype fastest                */
  int64_t ncells;
  const int64_t* cell_ptr;    /* [ncells+1] into entries, cell-id order          */
  const int32_t* ent_snp;     /* [E] snp id, ascending within a cell             */
  const int64_t* ent_ptr;     /* [E+1] into reads                                */
  const uint16_t* reads;      /* [R] allele<<8 | bq, UMI-string order in entry   */
  const int32_t* totl;        /* [ncells] RD.TOTL                                */
  const int32_t* pass;        /* [ncells] RD.PASS                                */
  const int32_t* uniq;        /* [ncells] RD.UNIQ                                */
} orc_input;

/* per-cell result of the doublet scan, cmd_cram_demuxlet.cpp:712-828 */
typedef struct {
  int32_t skipped;            /* 1: filtered (:581) or no SNP (:592); rest undefined */
  int32_t n_snp;
  int32_t i_sing1, i_sing2;   /* :746-758 */
  int32_t j_best, k_best, a_best; /* :799-814 */
  int32_t call;               /* 0 DBL, 1 SNG, 2 AMB  (:837-857) */
  double max_llk, sum_single, sum_double;      /* :713-734 */
  double sing_llk1, sing_llk2, sing_llk0;      /* :816-820 */
  double pair_llk12, pair_llk1, pair_llk2, pair_llk10, pair_llk20, pair_llk00; /* :821-826 */
  double post_doublet, post_singlet;           /* :827-828 */
} orc_best;

/* cohort-mean GP, cmd_cram_demuxlet.cpp:391-401.  gp0[nsnps*3] */
void orc_gp0(const orc_input* in, double* gp0);

/* pass 1 for one cell (cell-major restatement of :415-461; per-cell summation
 * order is SNP-ascending exactly as in the SNP-major reference loop).
 * llks[nv], *llk0 */
void orc_pass1_cell(const orc_input* in, const double* gp0, int64_t cell,
                    double* llks, double* llk0);

/* pass 2 accumulation for one cell, :583-710.
 * llksAB[nv*nv*n_alpha] ([j][k][n]), llksA0[nv*n_alpha], llks00[n_alpha] */
void orc_pass2_cell(const orc_input* in, const double* gp0, int64_t cell,
                    double* llksAB, double* llksA0, double* llks00);

/* exact value of a single hypothesis (j,k,n) -- same arithmetic as
 * orc_pass2_cell restricted to one (j,k,n); used to cross-check the product's
 * tie finaliser. */
double orc_pass2_one(const orc_input* in, int64_t cell, int32_t j, int32_t k, int32_t n);

/* reductions :712-828 on a cell's llksAB/llks00 */
void orc_summarise(const orc_input* in, int64_t cell, const double* llksAB,
                   const double* llks00, orc_best* out);

/* all cells: best[ncells]; optional dense outputs may be NULL.
 *   sing[ncells*nv]        llksAB[j][0][0]
 *   l00[ncells*n_alpha]    llks00
 *   p1[ncells*nv], p1_0[ncells]  pass-1 llks / llk0s
 * nthreads<=1: serial (how the reference runs); >1: OpenMP over cells
 * (the reference's --group-list sharding model). */
void orc_run(const orc_input* in, int32_t nthreads, orc_best* best,
             double* sing, double* l00, double* p1, double* p1_0);

/* run only the listed cells (timing of bounded samples); returns seconds of
 * wall time spent in the per-cell loops (gp0 precompute excluded) */
double orc_time_cells(const orc_input* in, const int64_t* cells, int64_t n,
                      int32_t nthreads, int32_t with_pass1);

/* write <prefix>.single/.sing2/.best[/.pair] exactly as :470-874.
 * barcodes[ncells], samples[nv]: NUL-terminated strings.  Rows are emitted in
 * barcode lexicographic order (std::map<std::string,int> iteration, :472,:576).
 * returns 0, or -1 if a file cannot be created. */
int orc_write_all(const orc_input* in, const char* const* barcodes,
                  const char* const* samples, const char* prefix,
                  int32_t write_pair);

/* diagnostic for SURVEY H1: number of (j<k) pairs at alpha index n whose
 * llksAB[j][k][n] != llksAB[k][j][n] bitwise in this cell */
int64_t orc_count_asym(const orc_input* in, const double* llksAB, int32_t n);

/* ---- VCF field -> genotype probabilities (dmx_field_oracle.c) ----
 * restatement of /root/reference/bcf_filtered_reader.cpp:186-320, :360-454 per biallelic diploid record.
 * field: 0 GT (in = int8[nsnps][nv][2], htslib-decoded allele indices, -1 missing), 1 GP (float[nsnps][nv][3]),
 * 2 PL (int32[nsnps][nv][3]); out float[nsnps][nv][3]; af[nsnps] = get_af(1) or NULL (GT / PL only) */
void orc_field_gt(int32_t nv, const int8_t* al, double gt_error, float* out, double* af);
void orc_field_pl(int32_t nv, const int32_t* pl, float* out, double* af);
void orc_field_gp(int32_t nv, const float* raw, double gt_error, float* out);
void orc_site_filter(int32_t nv, int64_t nsnps, const int8_t* gt, int32_t min_mac, double min_callrate, uint8_t* pass);
int orc_build_gp(int32_t field, int32_t nv, int64_t nsnps, const void* in, double gt_error, float* out, double* af,
                 int32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif

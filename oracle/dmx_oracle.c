/*
 * dmx_oracle.c -- CPU ORACLE (test infrastructure, never shipped; see
 * dmx_oracle.h).  Restates /root/reference/cmd_cram_demuxlet.cpp:386-875 on
 * the flat CSR boundary.  Every block cites the reference lines it follows.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off keeps mul and add separately rounded, as in the
 * reference's -O2 x86-64 (SSE2) build (Makefile.am:6).
 */
#include "dmx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- phred tables: PhredHelper.cpp:30-32 -------------------------------- */
static double g_err[256], g_mat[256];
static int g_tables_ready = 0;

static void init_tables(void) {
  if (g_tables_ready) return;
  for (int i = 0; i <= 255; i++) {
    g_err[i] = (i > 1) ? pow(0.1, i * 0.1) : 0.75;
    g_mat[i] = 1. - g_err[i];
  }
  g_tables_ready = 1;
}

static inline int cell_filtered(const orc_input* in, int64_t c) {
  /* :480 and :581 -- N.SNP is the number of covered SNPs incl. all-allele-2 ones */
  int32_t nsnp = (int32_t)(in->cell_ptr[c + 1] - in->cell_ptr[c]);
  return (in->totl[c] < in->min_total) || (in->uniq[c] < in->min_uniq) ||
         (nsnp < in->min_snp);
}

/* ---- :391-401 ------------------------------------------------------------ */
void orc_gp0(const orc_input* in, double* gp0) {
  const int32_t nv = in->nv;
  for (int64_t s = 0; s < in->nsnps; ++s) {
    const float* g = in->gp + (size_t)s * nv * 3;
    double a = 0, b = 0, c = 0;
    for (int32_t j = 0; j < nv; ++j) {
      a += (double)g[3 * j];
      b += (double)g[3 * j + 1];
      c += (double)g[3 * j + 2];
    }
    a /= nv; b /= nv; c /= nv;
    gp0[s * 3] = a; gp0[s * 3 + 1] = b; gp0[s * 3 + 2] = c;
  }
}

/* ---- :424-459 for one cell ----------------------------------------------- */
void orc_pass1_cell(const orc_input* in, const double* gp0, int64_t cell,
                    double* llks, double* llk0) {
  init_tables();
  const int32_t nv = in->nv;
  for (int32_t k = 0; k < nv; ++k) llks[k] = 0;
  *llk0 = 0;
  for (int64_t e = in->cell_ptr[cell]; e < in->cell_ptr[cell + 1]; ++e) {
    double GLs[3] = {1.0, 1.0, 1.0};
    double tmp;
    for (int64_t r = in->ent_ptr[e]; r < in->ent_ptr[e + 1]; ++r) {
      uint8_t al = (uint8_t)(in->reads[r] >> 8);
      uint8_t bq = (uint8_t)(in->reads[r] & 0xff);
      if (al == 2) continue;                                   /* :435 */
      GLs[0] *= ((al == 0) ? g_mat[bq] : g_err[bq] / 3.0);      /* :437 */
      GLs[1] *= (0.5 - g_err[bq] / 3.0);                        /* :438 */
      GLs[2] *= ((al == 1) ? g_mat[bq] : g_err[bq] / 3.0);      /* :439 */
      tmp = GLs[0] + GLs[1] + GLs[2];
      GLs[0] /= tmp; GLs[1] /= tmp; GLs[2] /= tmp;              /* :440-443 */
    }
    GLs[0] += 1e-6; GLs[1] += 1e-6; GLs[2] += 1e-6;             /* :446-448 */
    tmp = GLs[0] + GLs[1] + GLs[2];
    GLs[0] /= tmp; GLs[1] /= tmp; GLs[2] /= tmp;
    const int64_t s = in->ent_snp[e];
    const float* g = in->gp + (size_t)s * nv * 3;
    for (int32_t k = 0; k < nv; ++k)                            /* :455-456 */
      llks[k] += log(GLs[0] * (double)g[k * 3] + GLs[1] * (double)g[k * 3 + 1] +
                     GLs[2] * (double)g[k * 3 + 2]);
    *llk0 += log(GLs[0] * gp0[s * 3] + GLs[1] * gp0[s * 3 + 1] +
                 GLs[2] * gp0[s * 3 + 2]);                       /* :459 */
  }
}

/* ---- :597-663: read likelihood table pGs[n*9+l*3+m] for one entry ---------- */
static void entry_pGs(const orc_input* in, int64_t e, double* pGs) {
  const int32_t nA = in->n_alpha;
  for (int32_t i = 0; i < nA * 9; ++i) pGs[i] = 1.0;
  for (int64_t r = in->ent_ptr[e]; r < in->ent_ptr[e + 1]; ++r) {
    uint8_t al = (uint8_t)(in->reads[r] >> 8);
    uint8_t bq = (uint8_t)(in->reads[r] & 0xff);
    if (al == 2) continue;                                     /* :604 */
    double pR = (al == 0) ? g_mat[bq] : g_err[bq] / 3.0;        /* :606 */
    double pA = (al == 1) ? g_mat[bq] : g_err[bq] / 3.0;        /* :607 */
    double maxpG = 0;
    for (int32_t k = 0; k < nA; ++k)
      for (int32_t l = 0; l < 3; ++l)
        for (int32_t m = 0; m < 3; ++m) {
          double p = 0.5 * l + (m - l) * 0.5 * in->alpha[k];    /* :613 */
          double* pG = &pGs[k * 9 + l * 3 + m];
          *pG *= (pR * (1.0 - p) + pA * p);                     /* :625 */
          if (maxpG < *pG) maxpG = *pG;
        }
    for (int32_t i = 0; i < nA * 9; ++i) pGs[i] /= maxpG;       /* :632-639 */
  }
  double maxpG = 0;                                            /* :643-654 */
  for (int32_t i = 0; i < nA * 9; ++i) {
    pGs[i] += 1e-6;
    if (maxpG < pGs[i]) maxpG = pGs[i];
  }
  for (int32_t i = 0; i < nA * 9; ++i) pGs[i] /= maxpG;         /* :656-663 */
}

/* ---- :583-710 ------------------------------------------------------------- */
void orc_pass2_cell(const orc_input* in, const double* gp0, int64_t cell,
                    double* llksAB, double* llksA0, double* llks00) {
  init_tables();
  const int32_t nv = in->nv, nA = in->n_alpha;
  memset(llksAB, 0, sizeof(double) * (size_t)nv * nv * nA);
  memset(llksA0, 0, sizeof(double) * (size_t)nv * nA);
  memset(llks00, 0, sizeof(double) * (size_t)nA);
  double* pGs = (double*)malloc(sizeof(double) * nA * 9);
  double* sumPs = (double*)malloc(sizeof(double) * nA);
  double* gd = (double*)malloc(sizeof(double) * nv * 3);

  for (int64_t e = in->cell_ptr[cell]; e < in->cell_ptr[cell + 1]; ++e) {
    entry_pGs(in, e, pGs);
    const int64_t s = in->ent_snp[e];
    const float* g = in->gp + (size_t)s * nv * 3;
    for (int32_t i = 0; i < nv * 3; ++i) gd[i] = (double)g[i];   /* :181-184,:227-231 */
    const double* g0 = gp0 + s * 3;
    double p;
    for (int32_t j = 0; j < nv; ++j) {
      for (int32_t k = 0; k < nv; ++k) {                         /* :673-684 */
        for (int32_t n = 0; n < nA; ++n) sumPs[n] = 0;
        for (int32_t l = 0; l < 3; ++l)
          for (int32_t m = 0; m < 3; ++m) {
            p = gd[j * 3 + l] * gd[k * 3 + m];                   /* gpAB, :553 */
            for (int32_t n = 0; n < nA; ++n) sumPs[n] += (p * pGs[n * 9 + l * 3 + m]);
          }
        double* dst = llksAB + ((size_t)j * nv + k) * nA;
        for (int32_t n = 0; n < nA; ++n) dst[n] += log(sumPs[n]);
      }
      for (int32_t n = 0; n < nA; ++n) sumPs[n] = 0;             /* :687-696 */
      for (int32_t l = 0; l < 3; ++l)
        for (int32_t m = 0; m < 3; ++m) {
          p = gd[j * 3 + l] * g0[m];                             /* gpA0, :554 */
          for (int32_t n = 0; n < nA; ++n) sumPs[n] += (p * pGs[n * 9 + l * 3 + m]);
        }
      for (int32_t n = 0; n < nA; ++n) llksA0[j * nA + n] += log(sumPs[n]);
    }
    for (int32_t n = 0; n < nA; ++n) sumPs[n] = 0;               /* :700-709 */
    for (int32_t l = 0; l < 3; ++l)
      for (int32_t m = 0; m < 3; ++m) {
        p = g0[l] * g0[m];                                       /* gp00, :555 */
        for (int32_t n = 0; n < nA; ++n) sumPs[n] += (p * pGs[n * 9 + l * 3 + m]);
      }
    for (int32_t n = 0; n < nA; ++n) llks00[n] += log(sumPs[n]);
  }
  free(pGs); free(sumPs); free(gd);
}

double orc_pass2_one(const orc_input* in, int64_t cell, int32_t j, int32_t k, int32_t n) {
  init_tables();
  const int32_t nv = in->nv, nA = in->n_alpha;
  double* pGs = (double*)malloc(sizeof(double) * nA * 9);
  double llk = 0;
  for (int64_t e = in->cell_ptr[cell]; e < in->cell_ptr[cell + 1]; ++e) {
    entry_pGs(in, e, pGs);
    const float* g = in->gp + (size_t)in->ent_snp[e] * nv * 3;
    double sum = 0;
    for (int32_t l = 0; l < 3; ++l)
      for (int32_t m = 0; m < 3; ++m) {
        double p = (double)g[j * 3 + l] * (double)g[k * 3 + m];
        sum += (p * pGs[n * 9 + l * 3 + m]);
      }
    llk += log(sum);
  }
  free(pGs);
  return llk;
}

/* ---- :712-828 -------------------------------------------------------------- */
void orc_summarise(const orc_input* in, int64_t cell, const double* llksAB,
                   const double* llks00, orc_best* o) {
  const int32_t nv = in->nv, nA = in->n_alpha;
  const double dp = in->doublet_prior;
  memset(o, 0, sizeof(*o));
  o->n_snp = (int32_t)(in->cell_ptr[cell + 1] - in->cell_ptr[cell]);
#define AB(j, k, n) llksAB[((size_t)(j) * nv + (k)) * nA + (n)]
  double maxLLK = -1e300;                                       /* :713-721 */
  for (int32_t j = 0; j < nv; ++j)
    for (int32_t k = 0; k < nv; ++k)
      for (int32_t n = 0; n < nA; ++n)
        if (maxLLK < AB(j, k, n)) maxLLK = AB(j, k, n);
  double sumSingle = 0, sumDouble = 0;                          /* :724-734 */
  for (int32_t j = 0; j < nv; ++j) {
    sumSingle += (exp(AB(j, 0, 0) - maxLLK) * (1. - dp) / (nv));
    for (int32_t k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (int32_t n = 1; n < nA; ++n)
        sumDouble += (exp(AB(j, k, n) - maxLLK) * dp / (nv) / (nv - 1) / (nA - 1) /
                      (in->alpha[n] == 0.5 ? 2.0 : 1.0));
    }
  }
  int32_t iSing1 = -1, iSing2 = -1;                             /* :746-758 */
  double maxSing1 = -1e300, maxSing2 = -1e300;
  for (int32_t j = 0; j < nv; ++j) {
    if (maxSing1 < AB(j, 0, 0)) {
      maxSing2 = maxSing1; iSing2 = iSing1; iSing1 = j; maxSing1 = AB(j, 0, 0);
    } else if (maxSing2 < AB(j, 0, 0)) {
      iSing2 = j; maxSing2 = AB(j, 0, 0);
    }
  }
  int32_t jBest = -1, kBest = -1, aBest = -1;                   /* :799-814 */
  double maxAB = -1e300;
  for (int32_t j = 0; j < nv; ++j)
    for (int32_t k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (int32_t n = 1; n < nA; ++n)
        if (maxAB < AB(j, k, n)) { jBest = j; kBest = k; aBest = n; maxAB = AB(j, k, n); }
    }
  o->max_llk = maxLLK; o->sum_single = sumSingle; o->sum_double = sumDouble;
  o->i_sing1 = iSing1; o->i_sing2 = iSing2;
  o->j_best = jBest; o->k_best = kBest; o->a_best = aBest;
  o->sing_llk1 = AB(iSing1, 0, 0);                              /* :816-828 */
  o->sing_llk2 = AB(iSing2, 0, 0);
  o->sing_llk0 = llks00[0];
  o->pair_llk12 = AB(jBest, kBest, aBest);
  o->pair_llk1 = AB(jBest, 0, 0);
  o->pair_llk2 = AB(kBest, 0, 0);
  o->pair_llk10 = AB(jBest, 0, aBest);
  o->pair_llk20 = AB(kBest, 0, aBest);
  o->pair_llk00 = llks00[aBest];
  o->post_doublet = sumDouble / (sumSingle + sumDouble);
  o->post_singlet = exp(o->sing_llk1 - maxLLK) * (1. - dp) / (nv) / sumSingle;
  if ((o->pair_llk12 > o->pair_llk1) && (o->pair_llk12 > o->pair_llk2) &&
      (o->pair_llk12 > o->sing_llk1 + 2))                       /* :837 */
    o->call = 0;
  else if (o->sing_llk1 > o->sing_llk2 + 2)                     /* :844 */
    o->call = 1;
  else
    o->call = 2;
#undef AB
}

int64_t orc_count_asym(const orc_input* in, const double* llksAB, int32_t n) {
  const int32_t nv = in->nv, nA = in->n_alpha;
  int64_t c = 0;
  for (int32_t j = 0; j < nv; ++j)
    for (int32_t k = j + 1; k < nv; ++k)
      if (llksAB[((size_t)j * nv + k) * nA + n] != llksAB[((size_t)k * nv + j) * nA + n]) ++c;
  return c;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void orc_run(const orc_input* in, int32_t nthreads, orc_best* best, double* sing,
             double* l00, double* p1, double* p1_0) {
  init_tables();
  const int32_t nv = in->nv, nA = in->n_alpha;
  double* gp0 = (double*)malloc(sizeof(double) * 3 * (size_t)in->nsnps);
  orc_gp0(in, gp0);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    double* AB = (double*)malloc(sizeof(double) * (size_t)nv * nv * nA);
    double* A0 = (double*)malloc(sizeof(double) * (size_t)nv * nA);
    double* L00 = (double*)malloc(sizeof(double) * nA);
    double* P1 = (double*)malloc(sizeof(double) * nv);
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < in->ncells; ++c) {
      double P10 = 0;
      if (p1) {
        orc_pass1_cell(in, gp0, c, P1, &P10);
        memcpy(p1 + c * nv, P1, sizeof(double) * nv);
        if (p1_0) p1_0[c] = P10;
      }
      if (cell_filtered(in, c) || in->cell_ptr[c] == in->cell_ptr[c + 1]) {
        memset(&best[c], 0, sizeof(orc_best));
        best[c].skipped = 1;
        best[c].n_snp = (int32_t)(in->cell_ptr[c + 1] - in->cell_ptr[c]);
        if (sing) for (int32_t j = 0; j < nv; ++j) sing[c * nv + j] = 0;
        if (l00) for (int32_t n = 0; n < nA; ++n) l00[c * nA + n] = 0;
        continue;
      }
      orc_pass2_cell(in, gp0, c, AB, A0, L00);
      orc_summarise(in, c, AB, L00, &best[c]);
      if (sing) for (int32_t j = 0; j < nv; ++j) sing[c * nv + j] = AB[(size_t)j * nv * nA];
      if (l00) memcpy(l00 + c * nA, L00, sizeof(double) * nA);
    }
    free(AB); free(A0); free(L00); free(P1);
  }
  free(gp0);
}

double orc_time_cells(const orc_input* in, const int64_t* cells, int64_t n,
                      int32_t nthreads, int32_t with_pass1) {
  init_tables();
  const int32_t nv = in->nv, nA = in->n_alpha;
  double* gp0 = (double*)malloc(sizeof(double) * 3 * (size_t)in->nsnps);
  orc_gp0(in, gp0);
  if (nthreads < 1) nthreads = 1;
  volatile double sink = 0;
  double t0 = now_s();
#pragma omp parallel num_threads(nthreads)
  {
    double* AB = (double*)malloc(sizeof(double) * (size_t)nv * nv * nA);
    double* A0 = (double*)malloc(sizeof(double) * (size_t)nv * nA);
    double* L00 = (double*)malloc(sizeof(double) * nA);
    double* P1 = (double*)malloc(sizeof(double) * nv);
    double acc = 0;
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
      int64_t c = cells[i];
      double P10 = 0;
      orc_best b;
      if (with_pass1) { orc_pass1_cell(in, gp0, c, P1, &P10); acc += P1[0]; }
      if (cell_filtered(in, c) || in->cell_ptr[c] == in->cell_ptr[c + 1]) continue;
      orc_pass2_cell(in, gp0, c, AB, A0, L00);
      orc_summarise(in, c, AB, L00, &b);
      acc += b.pair_llk12;
    }
#pragma omp critical
    sink += acc;
    free(AB); free(A0); free(L00); free(P1);
  }
  double t1 = now_s();
  free(gp0);
  (void)sink;
  return t1 - t0;
}

/* ---- writers: :470-525 (.single), :533,:760-769 (.sing2), :570,:772-797 (.pair),
 *      :571,:830-874 (.best) ------------------------------------------------- */
typedef struct { const char* s; int64_t id; } bc_key;
static int bc_cmp(const void* a, const void* b) {
  /* std::map<std::string,int32_t> order == byte-wise lexicographic */
  const bc_key* x = (const bc_key*)a; const bc_key* y = (const bc_key*)b;
  const unsigned char* p = (const unsigned char*)x->s;
  const unsigned char* q = (const unsigned char*)y->s;
  while (*p && *p == *q) { ++p; ++q; }
  return (int)*p - (int)*q;
}

int orc_write_all(const orc_input* in, const char* const* barcodes,
                  const char* const* samples, const char* prefix, int32_t write_pair) {
  init_tables();
  const int32_t nv = in->nv, nA = in->n_alpha;
  const double dp = in->doublet_prior;
  char path[4096];
  bc_key* order = (bc_key*)malloc(sizeof(bc_key) * (size_t)in->ncells);
  for (int64_t c = 0; c < in->ncells; ++c) { order[c].s = barcodes[c]; order[c].id = c; }
  qsort(order, (size_t)in->ncells, sizeof(bc_key), bc_cmp);

  double* gp0 = (double*)malloc(sizeof(double) * 3 * (size_t)in->nsnps);
  orc_gp0(in, gp0);
  double* AB = (double*)malloc(sizeof(double) * (size_t)nv * nv * nA);
  double* A0 = (double*)malloc(sizeof(double) * (size_t)nv * nA);
  double* L00 = (double*)malloc(sizeof(double) * nA);
  double* P1 = (double*)malloc(sizeof(double) * nv);

  /* .single */
  snprintf(path, sizeof(path), "%s.single", prefix);
  FILE* ws = fopen(path, "w");
  if (!ws) return -1;
  fprintf(ws, "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n");
  for (int64_t oi = 0; oi < in->ncells; ++oi) {
    const int64_t c = order[oi].id;
    if (cell_filtered(in, c)) continue;                         /* :480 */
    double P10;
    orc_pass1_cell(in, gp0, c, P1, &P10);
    double sumLLK = -1e300;                                     /* :478-489 */
    for (int32_t j = 0; j < nv; ++j) {
      double cur = P1[j];
      if (sumLLK > cur) sumLLK = sumLLK + log(1.0 + exp(cur - sumLLK));
      else sumLLK = cur + log(1.0 + exp(sumLLK - cur));
    }
    const int32_t nsnp = (int32_t)(in->cell_ptr[c + 1] - in->cell_ptr[c]);
    for (int32_t j = 0; j < nv; ++j)
      fprintf(ws, "%s\t%s\t%d\t%d\t%d\t%d\t%.5lf\t%.5lf\t%.3lg\n", barcodes[c], samples[j],
              in->totl[c], in->pass[c], in->uniq[c], nsnp, P1[j], P10, exp(P1[j] - sumLLK));
  }
  fclose(ws);

  snprintf(path, sizeof(path), "%s.sing2", prefix);
  FILE* w2 = fopen(path, "w");
  snprintf(path, sizeof(path), "%s.best", prefix);
  FILE* wb = fopen(path, "w");
  FILE* wp = NULL;
  if (write_pair) { snprintf(path, sizeof(path), "%s.pair", prefix); wp = fopen(path, "w"); }
  if (!w2 || !wb || (write_pair && !wp)) return -1;
  fprintf(w2, "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n");
  if (wp) fprintf(wp, "BARCODE\tSM1.ID\tSM2.ID\tLLK12\tPOSTPRB\n");          /* :570 (5 names, 6 columns) */
  fprintf(wb, "BARCODE\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tBEST\tSNG.1ST\tSNG.LLK1\tSNG.2ND\tSNG.LLK2\tSNG.LLK0\tDBL.1ST\tDBL.2ND\tALPHA\tLLK12\tLLK1\tLLK2\tLLK10\tLLK20\tLLK00\tPRB.DBL\tPRB.SNG1\n");

#define ABv(j, k, n) AB[((size_t)(j) * nv + (k)) * nA + (n)]
  for (int64_t oi = 0; oi < in->ncells; ++oi) {
    const int64_t c = order[oi].id;
    const char* bc = barcodes[c];
    if (cell_filtered(in, c)) continue;                         /* :581 */
    if (in->cell_ptr[c] == in->cell_ptr[c + 1]) continue;       /* :592 */
    orc_pass2_cell(in, gp0, c, AB, A0, L00);
    orc_best b;
    orc_summarise(in, c, AB, L00, &b);
    const int32_t nsnp = b.n_snp;
    for (int32_t j = 0; j < nv; ++j)                            /* :760-769 */
      fprintf(w2, "%s\t%s\t%d\t%d\t%d\t%d\t%.4lf\t%.4lf\t%.3lg\n", bc, samples[j], in->totl[c],
              in->pass[c], in->uniq[c], nsnp, ABv(j, 0, 0), L00[0],
              exp(ABv(j, 0, 0) - b.max_llk) * (1. - dp) / (nv) / b.sum_single);
    if (wp) {                                                   /* :772-797 */
      for (int32_t j = 0; j < nv; ++j) {
        fprintf(wp, "%s\t%s\t%s\t%.3lf\t%.5lf\t%.5lg\n", bc, samples[j], samples[j], in->alpha[0],
                ABv(j, 0, 0),
                exp(ABv(j, 0, 0) - b.max_llk) * (1. - dp) / (nv) / (b.sum_single + b.sum_double));
        for (int32_t k = 0; k < nv; ++k)
          for (int32_t n = 0; n < nA; ++n)
            if ((n > 0) && (j != k)) {
              if ((j > k) && (in->alpha[n] == 0.5)) continue;
              fprintf(wp, "%s\t%s\t%s\t%.3lf\t%.5lf\t%.5lg\n", bc, samples[j], samples[k],
                      in->alpha[n], ABv(j, k, n),
                      exp(ABv(j, k, n) - b.max_llk) * dp / (nv) / (nv - 1) / (nA - 1) /
                          (b.sum_single + b.sum_double));
            }
      }
    }
    fprintf(wb, "%s\t%d\t%d\t%d\t%d\t", bc, in->totl[c], in->pass[c], in->uniq[c], nsnp);
    if (b.call == 0)
      fprintf(wb, "DBL-%s-%s-%.3lf", samples[b.j_best], samples[b.k_best], in->alpha[b.a_best]);
    else if (b.call == 1)
      fprintf(wb, "SNG-%s", samples[b.i_sing1]);
    else
      fprintf(wb, "AMB-%s-%s-%s/%s", samples[b.i_sing1], samples[b.i_sing2], samples[b.j_best],
              samples[b.k_best]);
    fprintf(wb, "\t%s\t%.4lf", samples[b.i_sing1], b.sing_llk1);
    fprintf(wb, "\t%s\t%.4lf\t%.4lf", samples[b.i_sing2], b.sing_llk2, b.sing_llk0);
    fprintf(wb, "\t%s\t%s\t%.3lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.3lg\t%.3lg\n",
            samples[b.j_best], samples[b.k_best], in->alpha[b.a_best], b.pair_llk12, b.pair_llk1,
            b.pair_llk2, b.pair_llk10, b.pair_llk20, b.pair_llk00, b.post_doublet, b.post_singlet);
  }
#undef ABv
  fclose(w2); fclose(wb); if (wp) fclose(wp);
  free(order); free(gp0); free(AB); free(A0); free(L00); free(P1);
  return 0;
}

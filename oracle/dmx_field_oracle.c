/*
 * dmx_field_oracle.c -- CPU ORACLE for the VCF-field -> genotype-probability transform.
 *
 * TEST INFRASTRUCTURE ONLY (see dmx_oracle.h): loaded by tests/ and by the cpu legs of tools/field_bench.py,
 * never by demuxlet_b200/.
 *
 * Plain-C restatement, per biallelic diploid VCF record, of what demuxlet stores in scl.snps[i].gps
 * (/root/reference/cmd_cram_demuxlet.cpp:227-231 reads it back through get_posterior_at):
 *   --field GT  /root/reference/bcf_filtered_reader.cpp:186-242 (allele counts acs / an over the non-missing alleles)
 *               and :368-404 (called genotype: 1 - gt_error on it, gt_error / 2 elsewhere; missing genotype:
 *               Hardy-Weinberg from (acs + 1/2) / (an + 1))
 *   --field PL  :244-320 (ten EM iterations of the allele frequency under HWE, phred table PhredHelper.cpp:32,
 *               negative / > 255 phred read as 255 through the uint32_t parameter of PhredHelper.h:40)
 *   --field GP  :408-450 (per-sample sum-normalisation in float, running float sums seeded with the pseudo sample
 *               (1/4, 1/2, 1/4), mean over nv + 1, blend with weight gt_error in double, rounded to float)
 * Same operation order and operand widths (float where the reference computes in float, double where it promotes),
 * no FMA contraction (-ffp-contract=off).  All samples selected (sm_icols[i] = i), ploidy 2 (:152-153).
 *
 * Parity pin: oracle/_ref/ref_field runs the reference's own three member functions (compiled from /root/reference)
 * on the same records; tests/test_field.py compares bit for bit, live and against tests/golden/field_*.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "dmx_oracle.h"

static double g_prob[256];
static int g_prob_ready = 0;

static void prob_init(void) {
  if (g_prob_ready) return;
  for (int i = 0; i <= 255; ++i) g_prob[i] = pow(0.1, i * 0.1);   /* PhredHelper.cpp:32 */
  g_prob_ready = 1;
}

static double to_prob(int32_t pl) {                                /* PhredHelper.h:40: uint32_t phred */
  uint32_t p = (uint32_t)pl;
  return p > 255 ? g_prob[255] : g_prob[p];
}

void orc_field_gt(int32_t nv, const int8_t* al, double gt_error, float* out, double* af) {
  double acs[2] = {0.0, 0.0};
  int32_t an = 0;
  for (int32_t i = 0; i < nv; ++i)                                  /* :231-240 */
    for (int j = 0; j < 2; ++j)
      if (al[2 * i + j] >= 0) { ++an; acs[al[2 * i + j]] += 1.0; }
  for (int32_t i = 0; i < nv; ++i) {
    const int a1 = al[2 * i], a2 = al[2 * i + 1];
    float* o = out + 3 * i;
    if (a1 < 0 || a2 < 0) {                                         /* :381-388 */
      int l = 0;
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k <= j; ++k, ++l)
          o[l] = (float)((j == k ? 1.0 : 2.0) * (acs[j] + 1.0 / 2) / (an + 1.0) * (acs[k] + 1.0 / 2) / (an + 1.0));
    } else {                                                        /* :398-402 */
      const int g = a1 > a2 ? a1 * (a1 + 1) / 2 + a2 : a2 * (a2 + 1) / 2 + a1;
      for (int j = 0; j < 3; ++j) o[j] = (g == j) ? 1.0 - gt_error : gt_error / (3 - 1.0);
    }
  }
  if (af) *af = (acs[1] + 1.0 / 2) / (an + 1.0);                    /* bcf_filtered_reader.h:177-179 */
}

void orc_field_pl(int32_t nv, const int32_t* pl, float* out, double* af) {
  prob_init();
  double acs[2] = {0.5, 0.5}, gp[3];
  int32_t an = 0;
  for (int it = 0; it < 10; ++it) {                                 /* :271-310 */
    double newacs[2] = {0.0, 0.0};
    an = 0;
    for (int32_t i = 0; i < nv; ++i) {
      double sumgp = 0;
      int l = 0;
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k <= j; ++k, ++l)
          sumgp += (gp[l] = (j == k ? 1 : 2) * acs[j] * acs[k] * to_prob(pl[3 * i + l]));
      l = 0;
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k <= j; ++k, ++l) {
          gp[l] /= sumgp;
          newacs[j] += gp[l];
          newacs[k] += gp[l];
        }
      an += 2;
      if (it + 1 == 10)
        for (l = 0; l < 3; ++l) out[3 * i + l] = (float)gp[l];
    }
    for (int i = 0; i < 2; ++i) acs[i] = newacs[i] / an;
  }
  for (int i = 0; i < 2; ++i) acs[i] *= an;                         /* :312-313 */
  if (af) *af = (acs[1] + 1.0 / 2) / (an + 1.0);
}

void orc_field_gp(int32_t nv, const float* raw, double gt_error, float* out) {
  float gp_sums[3] = {0.25f, 0.5f, 0.25f};                          /* :417-421 */
  for (int32_t i = 0; i < nv; ++i) {                                /* :424-435 */
    float sumgp = 0;
    for (int j = 0; j < 3; ++j) sumgp += raw[3 * i + j];
    for (int j = 0; j < 3; ++j) {
      out[3 * i + j] = raw[3 * i + j] / sumgp;
      gp_sums[j] += out[3 * i + j];
    }
  }
  for (int j = 0; j < 3; ++j) gp_sums[j] /= (int32_t)(nv + 1.0);    /* :437-438 */
  for (int32_t i = 0; i < nv; ++i)                                  /* :441-446 */
    for (int j = 0; j < 3; ++j)
      out[3 * i + j] = ((1.0 - gt_error) * out[3 * i + j] + gt_error * gp_sums[j]);
}

/* passed_vfilter's genotype-count block, bcf_filtered_reader.cpp:540-573, with every threshold but --min-mac and
 * --min-callrate at its default (bcf_filter_arg.h:52-64), where the other comparisons cannot fail */
void orc_site_filter(int32_t nv, int64_t nsnps, const int8_t* gt, int32_t min_mac, double min_callrate, uint8_t* pass) {
  for (int64_t s = 0; s < nsnps; ++s) {
    const int8_t* al = gt + s * nv * 2;
    double acs[2] = {0.0, 0.0};
    int32_t an = 0;
    for (int32_t i = 0; i < nv * 2; ++i)
      if (al[i] >= 0) { ++an; acs[al[i]] += 1.0; }
    pass[s] = 1;
    if (min_callrate > (double)an / (2.0 * (double)nv)) { pass[s] = 0; continue; }   /* :553 */
    const int32_t ac = an - acs[0];                                                  /* :555 */
    if ((ac < min_mac) || (an - ac < min_mac)) pass[s] = 0;                          /* :564 */
  }
}

/* field: 0 GT (in = int8[nsnps][nv][2]), 1 GP (float[nsnps][nv][3]), 2 PL (int32[nsnps][nv][3]);
 * out float[nsnps][nv][3]; af double[nsnps] or NULL (GT / PL only) */
int orc_build_gp(int32_t field, int32_t nv, int64_t nsnps, const void* in, double gt_error, float* out, double* af,
                 int32_t nthreads) {
  if (field < 0 || field > 2) return -1;
  prob_init();
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t s = 0; s < nsnps; ++s) {
    float* o = out + s * nv * 3;
    if (field == 0) orc_field_gt(nv, (const int8_t*)in + s * nv * 2, gt_error, o, af ? af + s : 0);
    else if (field == 1) orc_field_gp(nv, (const float*)in + s * nv * 3, gt_error, o);
    else orc_field_pl(nv, (const int32_t*)in + s * nv * 3, o, af ? af + s : 0);
  }
  return 0;
}

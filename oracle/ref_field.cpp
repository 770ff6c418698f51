/*
 * ref_field.cpp -- drives the REFERENCE's own VCF-field -> genotype-probability code on flat arrays.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product.
 *
 * The matrix `float gps[nsamples*3]` that demuxlet widens into scl.snps[i].gps
 * (/root/reference/cmd_cram_demuxlet.cpp:181-184, :227-231) is produced per VCF record by
 * BCFFilteredReader::parse_posteriors (/root/reference/bcf_filtered_reader.cpp:360-454), which
 * for --field GT calls parse_genotypes (:186-242) and for --field PL parse_likelihoods (:244-320,
 * a 10-iteration EM).  That translation unit needs htslib, which this image does not have, so
 * oracle/Makefile extracts exactly those three member-function definitions (and the six-line
 * inline get_genotype_at, bcf_filtered_reader.h:144-149) from the reference source WHERE IT LIES
 * into oracle/_ref/ (git-ignored build output) and this harness #includes them unchanged around a
 * class that declares only the members they touch.  The htslib names they use are supplied here:
 *   bcf_gt_allele / bcf_gt_missing / bcf_alleles2gt   htslib's published vcf.h definitions
 *   bcf_get_genotypes / bcf_get_format_int32 / _float copy the current record of the input file
 *   bcf_unpack, bcf_hdr_id2name, parse_int_fields     no-ops (minDP = minGQ = 0, the CLI default)
 * phredConv.toProb is the reference's own PhredHelper.cpp, compiled in place.
 *
 * File format (native endian): int32 field (0 GT, 1 GP, 2 PL), int32 nv, int64 nsnps,
 * double gt_error, then per SNP and sample: GT int8[2] htslib-decoded allele indices (-1 missing),
 * GP float[3], PL int32[3].  Output: float[nsnps][nv][3] (get_posterior_at order), then
 * double[nsnps] get_af(1) (bcf_filtered_reader.h:177-179).
 * field 3 = the genotype-count tests of passed_vfilter (:540-573) on GT records: the header is followed by
 * int32 min_mac, double min_callrate; output uint8[nsnps] pass flags.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "Error.h"
#include "PhredHelper.h"

struct bcf_hdr_t { int32_t nsamples; };
struct bcf1_t {
  int32_t rid, pos, n_allele;
  struct { char** allele; } d;
};
#define BCF_UN_ALL 15
static inline void bcf_unpack(bcf1_t*, int) {}
static inline const char* bcf_hdr_id2name(bcf_hdr_t*, int32_t) { return "chr"; }
static inline int32_t bcf_hdr_nsamples(bcf_hdr_t* h) { return h->nsamples; }
/* htslib vcf.h */
#define bcf_gt_missing 0
#define bcf_gt_unphased(idx) (((idx) + 1) << 1)
#define bcf_gt_allele(val) (((val) >> 1) - 1)
#define bcf_alleles2gt(a, b) ((a) > (b) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a)))

/* the current record, as the harness' stand-ins for htslib's FORMAT accessors hand it out */
static const int8_t* g_rec_gt;
static const float* g_rec_gp;
static const int32_t* g_rec_pl;
static int32_t g_nv;

static int bcf_get_genotypes(bcf_hdr_t*, bcf1_t*, int32_t** dst, int32_t* ndst) {
  if (!g_rec_gt) return -1;
  if (*ndst < g_nv * 2) { *dst = (int32_t*)realloc(*dst, sizeof(int32_t) * g_nv * 2); *ndst = g_nv * 2; }
  for (int32_t i = 0; i < g_nv * 2; ++i) (*dst)[i] = g_rec_gt[i] < 0 ? bcf_gt_missing : bcf_gt_unphased(g_rec_gt[i]);
  return g_nv * 2;
}
static int bcf_get_format_int32(bcf_hdr_t*, bcf1_t*, const char*, int32_t** dst, int32_t* ndst) {
  if (!g_rec_pl) return -1;
  if (*ndst < g_nv * 3) { *dst = (int32_t*)realloc(*dst, sizeof(int32_t) * g_nv * 3); *ndst = g_nv * 3; }
  memcpy(*dst, g_rec_pl, sizeof(int32_t) * g_nv * 3);
  return g_nv * 3;
}
static int bcf_get_format_float(bcf_hdr_t*, bcf1_t*, const char*, float** dst, int32_t* ndst) {
  if (!g_rec_gp) return -1;
  if (*ndst < g_nv * 3) { *dst = (float*)realloc(*dst, sizeof(float) * g_nv * 3); *ndst = g_nv * 3; }
  memcpy(*dst, g_rec_gp, sizeof(float) * g_nv * 3);
  return g_nv * 3;
}

/* only the members bcf_filtered_reader.cpp:186-320,360-454 touch (declarations as bcf_filtered_reader.h:68-96) */
class BCFFilteredReader {
 public:
  struct { int32_t minDP, minGQ; } gfilt;
  struct {                                   /* bcf_filter_arg.h:30-47, defaults :52-64 */
    bool require_GT;
    int32_t minAC, minMAC, maxAC, maxMAC;
    double minAF, minMAF, maxAF, maxMAF, minCallRate;
  } vfilt;
  struct { bcf_hdr_t* hdr; } cdr;
  bcf1_t* cur;
  int32_t* gts; int32_t n_gts;
  int32_t* pls; int32_t n_pls;
  float* gps; int32_t n_gps;
  void* flds;
  int8_t* ploidies;
  std::vector<double> acs;
  int32_t an;
  std::vector<int32_t> sm_icols;
  BCFFilteredReader() : cur(NULL), gts(NULL), n_gts(0), pls(NULL), n_pls(0), gps(NULL), n_gps(0), flds(NULL),
                        ploidies(NULL), an(0) {
    gfilt.minDP = gfilt.minGQ = 0;
    vfilt.require_GT = true;
    vfilt.minAC = vfilt.minMAC = 0; vfilt.maxAC = vfilt.maxMAC = INT_MAX;
    vfilt.minAF = vfilt.minMAF = 0; vfilt.maxAF = vfilt.maxMAF = 1.; vfilt.minCallRate = 0;
  }
  bcf1_t* cursor() { return cur; }
  bool parse_int_fields(const char*, bcf_hdr_t*, bcf1_t*) { return false; }
  bool parse_genotypes(bcf_hdr_t* hdr = NULL, bcf1_t* v = NULL);
  bool parse_likelihoods(bcf_hdr_t* hdr = NULL, bcf1_t* v = NULL, const char* name = "PL");
  bool parse_posteriors(bcf_hdr_t* hdr = NULL, bcf1_t* v = NULL, const char* name = "GP", double gt_error = 1e-4);
  /* the tail of passed_vfilter (bcf_filtered_reader.cpp:540-573): the genotype-count tests */
  bool passed_gt_counts(bcf_hdr_t* hdr, bcf1_t* v) {
#include "_ref/bfr_540_573.inc"
  }
#include "_ref/bfr_h_144_149.inc"
  inline float get_posterior_at(int32_t i, int32_t ngenotypes = 3) {   /* bcf_filtered_reader.h:159-161 */
    return gps[sm_icols[i / ngenotypes] * ngenotypes + (i % ngenotypes)];
  }
  inline double get_af(int32_t allele, double pseudocount = 1.0) {      /* bcf_filtered_reader.h:177-179 */
    return (acs[allele] + pseudocount / acs.size()) / (an + pseudocount);
  }
};

#include "_ref/bfr_186_242.inc"
#include "_ref/bfr_244_320.inc"
#include "_ref/bfr_360_454.inc"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_field <in.fld> <out.gp>\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t field, nv; int64_t nsnps; double gt_error;
  if (fread(&field, 4, 1, f) != 1 || fread(&nv, 4, 1, f) != 1 || fread(&nsnps, 8, 1, f) != 1 ||
      fread(&gt_error, 8, 1, f) != 1) { fprintf(stderr, "short header\n"); return 2; }
  int32_t min_mac = 0; double min_callrate = 0;
  if (field == 3 && (fread(&min_mac, 4, 1, f) != 1 || fread(&min_callrate, 8, 1, f) != 1)) { fprintf(stderr, "short header\n"); return 2; }
  const size_t per = (field == 0 || field == 3) ? (size_t)nv * 2 : (size_t)nv * 12;
  std::vector<unsigned char> rec(per);
  g_nv = nv;
  bcf_hdr_t hdr; hdr.nsamples = nv;
  char a0[] = "A", a1[] = "C"; char* alleles[2] = {a0, a1};
  bcf1_t v; v.rid = 0; v.pos = 0; v.n_allele = 2; v.d.allele = alleles;
  BCFFilteredReader vr;
  vr.cdr.hdr = &hdr; vr.cur = &v;
  vr.ploidies = new int8_t[nv];
  memset(vr.ploidies, 2, nv);                       /* bcf_filtered_reader.cpp:152-153 */
  for (int32_t i = 0; i < nv; ++i) vr.sm_icols.push_back(i);
  vr.acs.assign(2, 0.0);
  vr.vfilt.minMAC = min_mac; vr.vfilt.minCallRate = min_callrate;   /* cmd_cram_demuxlet.cpp:24-25, :47-48 */
  const char* name = field == 0 ? "GT" : (field == 1 ? "GP" : "PL");
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 2; }
  std::vector<float> out((size_t)nv * 3);
  std::vector<double> afs((size_t)nsnps);
  for (int64_t s = 0; s < nsnps; ++s) {
    if (fread(rec.data(), 1, per, f) != per) { fprintf(stderr, "short record\n"); return 2; }
    g_rec_gt = (field == 0 || field == 3) ? (const int8_t*)rec.data() : NULL;
    if (field == 3) {
      const unsigned char ok = vr.passed_gt_counts(&hdr, &v) ? 1 : 0;
      fwrite(&ok, 1, 1, o);
      continue;
    }
    g_rec_gp = field == 1 ? (const float*)rec.data() : NULL;
    g_rec_pl = field == 2 ? (const int32_t*)rec.data() : NULL;
    v.pos = (int32_t)s;
    if (!vr.parse_posteriors(&hdr, &v, name, gt_error)) { fprintf(stderr, "parse_posteriors failed\n"); return 3; }
    for (int32_t i = 0; i < nv * 3; ++i) out[i] = vr.get_posterior_at(i);   /* cmd_cram_demuxlet.cpp:227-229 */
    fwrite(out.data(), sizeof(float), out.size(), o);
    afs[s] = vr.get_af(1);
  }
  if (field != 3) fwrite(afs.data(), sizeof(double), afs.size(), o);
  fclose(o);
  fclose(f);
  return 0;
}

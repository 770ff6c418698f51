/*
 * ref_harness.cpp -- drives the REFERENCE's own hot-path code on a flat pileup.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product.
 *
 * The reference keeps the whole likelihood scan inline in main()
 * (/root/reference/cmd_cram_demuxlet.cpp:386-881) and the rest of that file
 * needs htslib, which this image does not have.  oracle/Makefile therefore
 * extracts exactly those line ranges from the reference source WHERE IT LIES
 * into oracle/_ref/ (git-ignored build output, never committed) and this
 * harness #includes them unchanged inside a function that provides the
 * handful of names they use:
 *   scl      a real sc_dropseq_lib_t filled via the reference's own
 *            add_snp/add_cell/add_read (sc_drop_seq.cpp, compiled in place)
 *   vr       {verbose, get_sample_id_at()}   (bcf_filtered_reader.h:166)
 *   htsFile/hts_open/hts_close/hprintf -> stdio (hts_utils.cpp:1013 formats
 *            with kvsprintf == vsnprintf semantics)
 *   phredConv, notice, error: the reference's PhredHelper.cpp / Error.cpp
 * Nothing of the reference is copied into the repository.
 *
 * Limits inherited from the reference: int32 indexing of the gpAB table
 * (nsnps*nv*nv*9 must stay < 2^31) -- callers compact SNP ids.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "Error.h"
#include "PhredHelper.h"
#include "sc_drop_seq.h"

typedef FILE htsFile;
static htsFile* hts_open(const char* fn, const char* mode) { return fopen(fn, mode); }
static int hts_close(htsFile* fp) { return fclose(fp); }
#define hprintf fprintf

struct vr_shim {
  int32_t verbose;
  std::vector<std::string> ids;
  const char* get_sample_id_at(int32_t i) { return ids[i].c_str(); }
};

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

template <class T>
static void rd(FILE* f, std::vector<T>& v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
static std::string rd_str(FILE* f) {
  std::string s; int c;
  while ((c = fgetc(f)) > 0) s.push_back((char)c);
  return s;
}

static double g_t_pass1 = 0, g_t_table = 0, g_t_loop = 0;

static int32_t run_reference_block(sc_dropseq_lib_t& scl, std::vector<int32_t>& snpids, int32_t nv,
                                   vr_shim& vr, std::string outPrefix, bool writePair,
                                   std::vector<double>& gridAlpha, double doublet_prior,
                                   int32_t minTotalReads, int32_t minUniqReads,
                                   int32_t minCoveredSNPs) {
  int32_t nAlpha = (int32_t)gridAlpha.size();     /* cmd_cram_demuxlet.cpp:348 */
  double* gps = NULL;                             /* cmd_cram_demuxlet.cpp:181 */
  double t0 = now_s();
#include "_ref/hot_386_538.inc"
  double t1 = now_s();
#include "_ref/hot_539_560.inc"
  double t2 = now_s();
#include "_ref/hot_561_881.inc"
  double t3 = now_s();
  g_t_pass1 = t1 - t0; g_t_table = t2 - t1; g_t_loop = t3 - t2;
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s in.dmxp outPrefix [--alpha A]... [--write-pair] [--doublet-prior P] "
                    "[--min-total N] [--min-uniq N] [--min-snp N]\n", argv[0]);
    return 1;
  }
  std::vector<double> gridAlpha;
  bool writePair = false;
  double doublet_prior = 0.5;
  int32_t minTotal = 0, minUniq = 0, minSnp = 0;
  for (int i = 3; i < argc; ++i) {
    std::string a(argv[i]);
    if (a == "--alpha" && i + 1 < argc) gridAlpha.push_back(atof(argv[++i]));
    else if (a == "--write-pair") writePair = true;
    else if (a == "--doublet-prior" && i + 1 < argc) doublet_prior = atof(argv[++i]);
    else if (a == "--min-total" && i + 1 < argc) minTotal = atoi(argv[++i]);
    else if (a == "--min-uniq" && i + 1 < argc) minUniq = atoi(argv[++i]);
    else if (a == "--min-snp" && i + 1 < argc) minSnp = atoi(argv[++i]);
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 1; }
  }
  if (gridAlpha.empty()) { gridAlpha.push_back(0); gridAlpha.push_back(0.5); }  /* :78-90 */

  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DMXPLP01", 8) != 0) { fprintf(stderr, "bad magic\n"); return 2; }
  int64_t hdr[5];
  if (fread(hdr, 8, 5, f) != 5) return 2;
  const int64_t nv = hdr[0], nsnps = hdr[1], ncells = hdr[2], E = hdr[3], R = hdr[4];
  std::vector<float> gp; std::vector<int64_t> cell_ptr, ent_ptr; std::vector<int32_t> ent_snp, totl, pass, uniq;
  std::vector<uint16_t> reads;
  rd(f, gp, (size_t)nsnps * nv * 3); rd(f, cell_ptr, ncells + 1); rd(f, ent_snp, E); rd(f, ent_ptr, E + 1);
  rd(f, reads, R); rd(f, totl, ncells); rd(f, pass, ncells); rd(f, uniq, ncells);
  std::vector<std::string> barcodes(ncells);
  for (int64_t c = 0; c < ncells; ++c) barcodes[c] = rd_str(f);
  vr_shim vr; vr.verbose = 10000;
  for (int64_t j = 0; j < nv; ++j) vr.ids.push_back(rd_str(f));
  fclose(f);

  if ((double)nsnps * nv * nv * 9 >= 2147483648.0) {
    fprintf(stderr, "reference int32 table index would overflow (SURVEY Q1); compact the SNP ids\n");
    return 3;
  }

  /* fill the reference's pileup container through its own API */
  sc_dropseq_lib_t scl;
  std::vector<int32_t> snpids;
  for (int64_t s = 0; s < nsnps; ++s) {
    double* g = new double[nv * 3];
    for (int64_t i = 0; i < nv * 3; ++i) g[i] = gp[s * nv * 3 + i];   /* float -> double, :227-231 */
    snpids.push_back(scl.add_snp(0, (int32_t)s, 'A', 'C', 0.0, g));
  }
  char umi[32];
  for (int64_t c = 0; c < ncells; ++c) {
    int32_t ic = scl.add_cell(barcodes[c].c_str());
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e)
      for (int64_t r = ent_ptr[e]; r < ent_ptr[e + 1]; ++r) {
        snprintf(umi, sizeof(umi), "U%09lld", (long long)(r - ent_ptr[e]));   /* string order == flat order */
        scl.add_read(ent_snp[e], ic, umi, (char)(reads[r] >> 8), (char)(reads[r] & 0xff));
      }
  }
  for (int64_t c = 0; c < ncells; ++c) {
    scl.cell_totl_reads[c] = totl[c];
    scl.cell_pass_reads[c] = pass[c];
    scl.cell_uniq_reads[c] = uniq[c];
  }

  run_reference_block(scl, snpids, (int32_t)nv, vr, std::string(argv[2]), writePair, gridAlpha,
                      doublet_prior, minTotal, minUniq, minSnp);
  printf("{\"ncells\": %lld, \"nv\": %lld, \"nsnps\": %lld, \"t_pass1_s\": %.6f, \"t_table_s\": %.6f, \"t_loop_s\": %.6f}\n",
         (long long)ncells, (long long)nv, (long long)nsnps, g_t_pass1, g_t_table, g_t_loop);
  return 0;
}

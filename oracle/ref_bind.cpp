/*
 * ref_bind.cpp -- the call site INTEGRATION.md section 2 gives a demuxlet maintainer, compiled for real.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product.
 *
 * A real sc_dropseq_lib_t is filled through the reference's own add_snp / add_cell / add_read
 * (/root/reference/sc_drop_seq.cpp, compiled in place), the variables the reference's main() has in scope at
 * cmd_cram_demuxlet.cpp:383 are set up under their own names, and then oracle/integration_block.inc -- the very
 * text of the documentation's code block (tests/test_host.py checks they are identical) -- runs against libdmx.
 * tests/test_gpu_configs.py diffs the four files it writes with the ones the reference's lines 386-881 wrote for
 * the same pileups (tests/golden, made by ref_harness).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "Error.h"
#include "sc_drop_seq.h"
#include "dmx.h"

struct vr_shim {   /* BCFFilteredReader::get_sample_id_at, bcf_filtered_reader.h:166 */
  std::vector<std::string> ids;
  const char* get_sample_id_at(int32_t i) { return ids[i].c_str(); }
};

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

/* the reference's main() from :383 on, with the integration block in front of the (absent) CPU loop */
static int32_t main_from_383(sc_dropseq_lib_t& scl, int32_t nv, vr_shim& vr, std::string outPrefix, bool writePair,
                             std::vector<double>& gridAlpha, double doublet_prior, int32_t minTotalReads,
                             int32_t minUniqReads, int32_t minCoveredSNPs, int32_t gpuDevice) {
  int32_t nAlpha = (int32_t)gridAlpha.size();     /* cmd_cram_demuxlet.cpp:348 */
#include "integration_block.inc"
  error("[E:%s] this harness has no CPU loop: --gpu is required", __FUNCTION__);
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s in.dmxp outPrefix [--alpha A]... [--write-pair] [--doublet-prior P] "
                    "[--min-total N] [--min-uniq N] [--min-snp N] [--gpu D]\n", argv[0]);
    return 1;
  }
  std::vector<double> gridAlpha;
  bool writePair = false;
  double doublet_prior = 0.5;
  int32_t minTotal = 0, minUniq = 0, minSnp = 0, gpuDevice = 0;
  for (int i = 3; i < argc; ++i) {
    std::string a(argv[i]);
    if (a == "--alpha" && i + 1 < argc) gridAlpha.push_back(atof(argv[++i]));
    else if (a == "--write-pair") writePair = true;
    else if (a == "--doublet-prior" && i + 1 < argc) doublet_prior = atof(argv[++i]);
    else if (a == "--min-total" && i + 1 < argc) minTotal = atoi(argv[++i]);
    else if (a == "--min-uniq" && i + 1 < argc) minUniq = atoi(argv[++i]);
    else if (a == "--min-snp" && i + 1 < argc) minSnp = atoi(argv[++i]);
    else if (a == "--gpu" && i + 1 < argc) gpuDevice = atoi(argv[++i]);
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
  vr_shim vr;
  for (int64_t j = 0; j < nv; ++j) vr.ids.push_back(rd_str(f));
  fclose(f);

  /* fill the reference's pileup container through its own API (as ref_harness.cpp does) */
  sc_dropseq_lib_t scl;
  for (int64_t s = 0; s < nsnps; ++s) {
    double* g = new double[nv * 3];
    for (int64_t i = 0; i < nv * 3; ++i) g[i] = gp[s * nv * 3 + i];   /* float -> double, :227-231 */
    scl.add_snp(0, (int32_t)s, 'A', 'C', 0.0, g);
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
  return main_from_383(scl, (int32_t)nv, vr, std::string(argv[2]), writePair, gridAlpha, doublet_prior, minTotal,
                       minUniq, minSnp, gpuDevice);
}

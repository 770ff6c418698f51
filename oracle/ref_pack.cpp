/*
 * ref_pack.cpp -- drives the REFERENCE's own pileup container on a list of reads and flattens the result.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product.  Compiled by oracle/Makefile together with the
 * reference's sc_drop_seq.cpp / Error.cpp from /root/reference where they lie (nothing is copied).
 *
 * in : "DMXRDS01" n ncells nsnps | int32 cell[n] | int32 snp[n] | uint8 allele[n] | uint8 bq[n] |
 *      int64 umi_off[n+1] | char umi_buf[umi_off[n]]            (reads in arrival order)
 * out: "DMXCSR01" ncells E R | int64 cell_ptr[ncells+1] | int32 ent_snp[E] | int64 ent_ptr[E+1] | uint16 reads[R] |
 *      int32 pass[ncells] | int32 uniq[ncells]
 * The reads go through sc_dropseq_lib_t::add_read (sc_drop_seq.cpp:34-77) one by one; the container is then walked
 * exactly like the likelihood loops walk it (cmd_cram_demuxlet.cpp:421-430, :594-602).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <string>
#include <vector>

#include "Error.h"
#include "sc_drop_seq.h"

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
template <class T>
static void wr(FILE* f, const std::vector<T>& v) {
  if (!v.empty() && fwrite(v.data(), sizeof(T), v.size(), f) != v.size()) { fprintf(stderr, "short write\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.rds out.csr\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  char magic[8];
  int64_t h[3];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DMXRDS01", 8) != 0 || fread(h, 8, 3, f) != 3) return 2;
  const int64_t n = h[0], ncells = h[1], nsnps = h[2];
  std::vector<int32_t> cell, snp; std::vector<uint8_t> al, bq; std::vector<int64_t> off; std::vector<char> buf;
  rd(f, cell, n); rd(f, snp, n); rd(f, al, n); rd(f, bq, n); rd(f, off, n + 1); rd(f, buf, (size_t)off[n]);
  fclose(f);

  sc_dropseq_lib_t scl;
  static double gps3[3] = {1. / 3, 1. / 3, 1. / 3};
  for (int64_t s = 0; s < nsnps; ++s) scl.add_snp(0, (int32_t)s, 'A', 'C', 0.0, gps3);
  char name[32];
  for (int64_t c = 0; c < ncells; ++c) { snprintf(name, sizeof(name), "C%09lld", (long long)c); scl.add_cell(name); }
  const double t0 = now_s();
  std::string umi;
  for (int64_t i = 0; i < n; ++i) {
    umi.assign(buf.data() + off[i], (size_t)(off[i + 1] - off[i]));
    scl.add_read(snp[i], cell[i], umi.c_str(), (char)al[i], (char)bq[i]);
  }
  const double t1 = now_s();
  std::vector<int64_t> cell_ptr(ncells + 1, 0), ent_ptr(1, 0);
  std::vector<int32_t> ent_snp;
  std::vector<uint16_t> reads;
  for (int64_t c = 0; c < ncells; ++c) {
    for (std::map<int32_t, sc_snp_droplet_t*>::iterator it = scl.cell_umis[c].begin(); it != scl.cell_umis[c].end(); ++it) {
      ent_snp.push_back(it->first);
      for (sc_snp_droplet_it_t u = it->second->begin(); u != it->second->end(); ++u)
        reads.push_back((uint16_t)((((u->second >> 24) & 0xff) << 8) | ((u->second >> 16) & 0xff)));
      ent_ptr.push_back((int64_t)reads.size());
    }
    cell_ptr[c + 1] = (int64_t)ent_snp.size();
  }
  const double t2 = now_s();
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 2; }
  const int64_t oh[3] = {ncells, (int64_t)ent_snp.size(), (int64_t)reads.size()};
  fwrite("DMXCSR01", 1, 8, o);
  fwrite(oh, 8, 3, o);
  wr(o, cell_ptr); wr(o, ent_snp); wr(o, ent_ptr); wr(o, reads);
  std::vector<int32_t> pass(scl.cell_pass_reads.begin(), scl.cell_pass_reads.end()),
      uniq(scl.cell_uniq_reads.begin(), scl.cell_uniq_reads.end());
  wr(o, pass); wr(o, uniq);
  fclose(o);
  printf("{\"n_reads\": %lld, \"t_add_read_s\": %.6f, \"t_flatten_s\": %.6f}\n", (long long)n, t1 - t0, t2 - t1);
  return 0;
}

#include "../include/dmx.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  if (!strcmp(argv[1], "vcf")) {
    for (int field = 0; field < 3; ++field) {
      dmx_vcf* v = nullptr;
      int rc = dmx_vcf_read(argv[2], field, &v);
      int32_t nv = 0, nch = 0; int64_t nrec = 0, nskip = 0;
      if (rc == 0) { dmx_vcf_sizes(v, &nv, &nrec, &nskip, &nch);
        std::vector<int8_t> gt((size_t)nrec * nv * 2 + 1); std::vector<char> vals((size_t)nrec * nv * 12 + 1);
        std::vector<int32_t> rid(nrec + 1), pos(nrec + 1); std::vector<char> ref(nrec + 1), alt(nrec + 1); std::vector<uint8_t> nf(nrec + 1);
        dmx_vcf_copy(v, gt.data(), vals.data(), rid.data(), pos.data(), ref.data(), alt.data()); dmx_vcf_no_field(v, nf.data()); }
      printf("vcf field %d rc %d nrec %lld nskip %lld nv %d %s\n", field, rc, (long long)nrec, (long long)nskip, nv, rc ? dmx_vcf_last_error() : "");
      dmx_vcf_free(v);
    }
    return 0;
  }
  FILE* f = fopen(argv[3], "rb");
  int64_t n; int32_t nch;
  if (fread(&n, 8, 1, f) != 1 || fread(&nch, 4, 1, f) != 1) return 3;
  std::vector<std::string> names(nch);
  for (auto& s : names) { int c; while ((c = fgetc(f)) > 0) s.push_back((char)c); }
  std::vector<int32_t> rid(n), pos(n); std::vector<char> ref(n), alt(n);
  if (fread(rid.data(), 4, n, f) != (size_t)n || fread(pos.data(), 4, n, f) != (size_t)n || fread(ref.data(), 1, n, f) != (size_t)n || fread(alt.data(), 1, n, f) != (size_t)n) return 3;
  fclose(f);
  std::vector<const char*> cn; for (auto& s : names) cn.push_back(s.c_str());
  for (int mode = 0; mode < 2; ++mode) {
    dmx_sweep_opts o; dmx_sweep_opts_default(&o);
    if (mode == 1) o.tag_umi = "";
    dmx_sweep* sw = nullptr;
    int rc = dmx_sweep_bam(argv[2], &o, nch, cn.data(), n, rid.data(), pos.data(), ref.data(), alt.data(), &sw);
    int64_t nc = 0, nr = 0, ub = 0;
    if (rc == 0) { dmx_sweep_sizes(sw, &nc, &nr, &ub);
      std::vector<int32_t> cell(nr + 1), snp(nr + 1), totl(nc + 1); std::vector<uint8_t> al(nr + 1), bq(nr + 1); std::vector<char> ubuf(ub + 1); std::vector<int64_t> uo(nr + 2);
      dmx_sweep_copy(sw, cell.data(), snp.data(), al.data(), bq.data(), ubuf.data(), uo.data(), totl.data());
      // ... and through the packer (N1), on several threads
      dmx_pileup* pl = nullptr;
      const int prc = dmx_pack_reads(nr, cell.data(), snp.data(), al.data(), bq.data(), ubuf.data(), uo.data(), nc, 4, &pl);
      int64_t pc = 0, pe = 0, pr = 0;
      if (prc == 0) { dmx_pileup_sizes(pl, &pc, &pe, &pr);
        std::vector<int64_t> cp(pc + 1), ep(pe + 1); std::vector<int32_t> es(pe + 1), pa(pc + 1), un(pc + 1); std::vector<uint16_t> rd(pr + 1);
        dmx_pileup_copy(pl, cp.data(), es.data(), ep.data(), rd.data(), pa.data(), un.data()); }
      printf("pack rc %d cells %lld entries %lld reads %lld\n", prc, (long long)pc, (long long)pe, (long long)pr);
      dmx_pileup_free(pl); }
    printf("bam mode %d rc %d cells %lld tuples %lld %s\n", mode, rc, (long long)nc, (long long)nr, rc ? dmx_sweep_last_error() : "");
    dmx_sweep_free(sw);
  }
  return 0;
}

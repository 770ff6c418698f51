// dmx_main.cpp -- command line driver of the hot path: demuxlet's options, libdmx underneath.
//
// The reference's main() (cmd_cram_demuxlet.cpp:15-883) reads a BAM and a VCF through htslib, builds the pileup
// (sc_dropseq_lib_t) and then runs lines 386-881.  This driver keeps the option names and the four output files of
// that program for the part that is in scope: it takes the finished pileup as a flat file (--plp, the .dmxp
// container of demuxlet_b200/plp.py: GP matrix + CSR pileup + barcodes + sample ids, i.e. exactly what crosses the
// C ABI) and calls dmx_set_gp / dmx_set_pileup / dmx_run / dmx_write_outputs.  --sam FILE --vcf FILE run the whole reference
// command line through the library's own BGZF / BAM / VCF / BCF readers (no htslib: no CRAM).
//
//   dmx-demuxlet --plp in.dmxp --out prefix [--alpha A]... [--doublet-prior P] [--min-total N] [--min-uniq N]
//                [--min-snp N] [--group-list FILE] [--write-pair] [--gpu D | --gpus N | --gpu-list a,b,..]
//                [--fp64-pairs] [--verbose]
//
// Options and defaults follow cmd_cram_demuxlet.cpp:29-33,:54-56,:64-71,:78-90 (alpha grid {0, 0.5} when none is
// given; barcodes restricted to --group-list like :68,:131-140).  Exit status 0 on success, 1 on error with the
// reference's "[E:...]" style message on stderr (Error.cpp:27-41).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <set>
#include <string>
#include <vector>

#include "../../include/dmx.h"

namespace {
[[noreturn]] void die(const char* fmt, const char* a = "", const char* b = "") {
  fprintf(stderr, "\nFATAL ERROR - \n[E:dmx-demuxlet] ");
  fprintf(stderr, fmt, a, b);
  fprintf(stderr, "\n");
  exit(1);
}

struct Plp {
  int64_t nv = 0, nsnps = 0, ncells = 0, E = 0, R = 0;
  std::vector<float> gp;
  std::vector<int64_t> cell_ptr, ent_ptr;
  std::vector<int32_t> ent_snp, totl, npass, uniq;
  std::vector<uint16_t> reads;
  std::vector<std::string> barcodes, samples;
};

template <class T>
void rd(FILE* f, std::vector<T>* v, size_t n, const char* path) {
  v->resize(n);
  if (n && fread(v->data(), sizeof(T), n, f) != n) die("Truncated pileup file %s", path);
}

void load_plp(const char* path, Plp* p) {
  FILE* f = fopen(path, "rb");
  if (!f) die("Cannot open file %s for reading", path);
  char magic[8];
  int64_t h[5];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DMXPLP01", 8) != 0) die("%s is not a DMXPLP01 pileup file", path);
  if (fread(h, 8, 5, f) != 5) die("Truncated pileup file %s", path);
  p->nv = h[0]; p->nsnps = h[1]; p->ncells = h[2]; p->E = h[3]; p->R = h[4];
  if (p->nv < 0 || p->nsnps < 0 || p->ncells < 0 || p->E < 0 || p->R < 0) die("Corrupt header in %s", path);
  rd(f, &p->gp, (size_t)p->nsnps * p->nv * 3, path);
  rd(f, &p->cell_ptr, (size_t)p->ncells + 1, path);
  rd(f, &p->ent_snp, (size_t)p->E, path);
  rd(f, &p->ent_ptr, (size_t)p->E + 1, path);
  rd(f, &p->reads, (size_t)p->R, path);
  rd(f, &p->totl, (size_t)p->ncells, path);
  rd(f, &p->npass, (size_t)p->ncells, path);
  rd(f, &p->uniq, (size_t)p->ncells, path);
  std::string cur;
  int c;
  std::vector<std::string> names;
  while ((c = fgetc(f)) != EOF) {
    if (c == 0) { names.push_back(cur); cur.clear(); }
    else cur.push_back((char)c);
  }
  fclose(f);
  if ((int64_t)names.size() != p->ncells + p->nv) die("Barcode / sample id table of %s has the wrong length", path);
  p->barcodes.assign(names.begin(), names.begin() + p->ncells);
  p->samples.assign(names.begin() + p->ncells, names.end());
}

// --field-matrix: the FORMAT values of the VCF records, flat (int32 field code 0 GT / 1 GP / 2 PL, int32 nv, int64 nsnps,
// double geno_error, then int8[nsnps][nv][2] allele indices, float[nsnps][nv][3] or int32[nsnps][nv][3]); the
// genotype-probability matrix is then built on the device (dmx_set_field) instead of read from the pileup file
struct FieldFile {
  int32_t field = -1, nv = 0;
  int64_t nsnps = 0;
  double geno_error = 0.01;
  std::vector<unsigned char> data;
};

void load_field(const char* path, FieldFile* ff) {
  FILE* f = fopen(path, "rb");
  if (!f) die("Cannot open file %s for reading", path);
  if (fread(&ff->field, 4, 1, f) != 1 || fread(&ff->nv, 4, 1, f) != 1 || fread(&ff->nsnps, 8, 1, f) != 1 ||
      fread(&ff->geno_error, 8, 1, f) != 1)
    die("Truncated field matrix %s", path);
  if (ff->field < DMX_FIELD_GT || ff->field > DMX_FIELD_PL || ff->nv < 1 || ff->nsnps < 0) die("Corrupt header in %s", path);
  const size_t n = (size_t)ff->nsnps * ff->nv * (ff->field == DMX_FIELD_GT ? 2 : 12);
  ff->data.resize(n);
  if (n && fread(ff->data.data(), 1, n, f) != n) die("Truncated field matrix %s", path);
  fclose(f);
}

// --group-list (:68, :131-140): keep only the listed barcodes; cell order is preserved
void restrict_to_groups(const char* path, Plp* p) {
  FILE* f = fopen(path, "r");
  if (!f) die("Cannot open file %s for reading", path);
  std::set<std::string> keep;
  char line[4096];
  while (fgets(line, sizeof(line), f)) {
    size_t n = strcspn(line, " \t\r\n");
    if (n) keep.insert(std::string(line, n));
  }
  fclose(f);
  Plp q;
  q.nv = p->nv; q.nsnps = p->nsnps;
  q.gp.swap(p->gp);
  q.samples = p->samples;
  q.cell_ptr.push_back(0);
  q.ent_ptr.push_back(0);
  for (int64_t i = 0; i < p->ncells; ++i) {
    if (!keep.count(p->barcodes[i])) continue;
    for (int64_t e = p->cell_ptr[i]; e < p->cell_ptr[i + 1]; ++e) {
      q.ent_snp.push_back(p->ent_snp[e]);
      for (int64_t r = p->ent_ptr[e]; r < p->ent_ptr[e + 1]; ++r) q.reads.push_back(p->reads[r]);
      q.ent_ptr.push_back((int64_t)q.reads.size());
    }
    q.cell_ptr.push_back((int64_t)q.ent_snp.size());
    q.totl.push_back(p->totl[i]); q.npass.push_back(p->npass[i]); q.uniq.push_back(p->uniq[i]);
    q.barcodes.push_back(p->barcodes[i]);
  }
  q.ncells = (int64_t)q.barcodes.size();
  q.E = (int64_t)q.ent_snp.size();
  q.R = (int64_t)q.reads.size();
  *p = std::move(q);
}

void usage() {
  fprintf(stderr,
          "dmx-demuxlet: demuxlet's genotype-likelihood scan on B200 (libdmx)\n"
          "  --plp FILE          flat pileup (.dmxp): GP matrix + (barcode, snp, allele, BQ) table   [required]\n"
          "  --out PREFIX        output prefix: PREFIX.best .single .sing2 [.pair]                     [required]\n"
          "  --alpha A           grid of alpha, repeatable (default 0 and 0.5)\n"
          "  --doublet-prior P   prior of doublet (default 0.5)\n"
          "  --min-total N / --min-uniq N / --min-snp N   per-barcode filters (default 0)\n"
          "  --group-list FILE   only the barcodes listed in FILE\n"
          "  --write-pair        also write the (huge) PREFIX.pair\n"
          "  --gpu D             CUDA device ordinal (default 0)\n"
          "  --gpus N            shard the barcodes over the first N GPUs of this host (contiguous ranges balanced by\n"
          "                      pileup entries, the reference's --group-list model); one set of output files\n"
          "  --gpu-list a,b,..   the same over the listed device ordinals\n"
          "  --fp64-pairs        pair scan in fp64 (default: fp32 products, exact re-score of the decisive values)\n"
          "  --field GT|GP|PL --field-matrix FILE [--geno-error E]\n"
          "                      build the genotype probabilities on the GPU from the flat FORMAT values in FILE\n"
          "                      (as demuxlet's --field / --geno-error do per VCF record) instead of the matrix in --plp\n"
          "  --vcf FILE [--field F] [--geno-error E] [--min-mac N] [--min-callrate R]\n"
          "                      (F: GT, PL, or the name of a float FORMAT field with genotype posteriors, GP by default)\n"
          "                      the same from a VCF (plain / gzip / bgzip text) or BCF (F defaults to GP): biallelic records that\n"
          "                      pass the genotype-count filter must be the SNPs of the pileup, in order\n"
          "  --sam FILE --vcf FILE   instead of --plp: pile the reads of a coordinate-sorted BAM (or SAM text) up on the VCF's records\n"
          "                      [--tag-group CB] [--tag-UMI UB] [--min-MQ 20] [--excl-flag 0x0f04] [--min-BQ 13] [--cap-BQ 40]\n"
          "                      [--min-TD 0] [--sm ID]... [--sm-list FILE]; own BGZF / BAM / VCF / BCF readers (no htslib; no CRAM)\n");
}
}  // namespace

int main(int argc, char** argv) {
  const char *plp_path = nullptr, *out = nullptr, *groups = nullptr, *field = nullptr, *field_path = nullptr,
             *vcf_path = nullptr, *sam_path = nullptr;
  dmx_sweep_opts so;
  dmx_sweep_opts_default(&so);
  std::vector<std::string> sm_ids;
  const char* sm_list = nullptr;
  double geno_error = -1, min_callrate = 0.5;     // :25
  int min_mac = 1;                                // :24
  std::vector<double> alpha;
  double prior = 0.5;
  int min_total = 0, min_uniq = 0, min_snp = 0, gpu = 0, write_pair = 0, verbose = 0, fp64 = 0;
  std::vector<int32_t> gpu_list;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&]() -> const char* {
      if (i + 1 >= argc) die("Missing value for option %s", argv[i]);
      return argv[++i];
    };
    if (a == "--plp") plp_path = val();
    else if (a == "--out") out = val();
    else if (a == "--alpha") alpha.push_back(atof(val()));
    else if (a == "--doublet-prior") prior = atof(val());
    else if (a == "--min-total") min_total = atoi(val());
    else if (a == "--min-uniq") min_uniq = atoi(val());
    else if (a == "--min-snp") min_snp = atoi(val());
    else if (a == "--group-list") groups = val();
    else if (a == "--write-pair") write_pair = 1;
    else if (a == "--gpu") gpu = atoi(val());
    else if (a == "--gpus") {
      const int n = atoi(val());
      if (n < 1) die("--gpus needs a positive count", "");
      gpu_list.clear();
      for (int d = 0; d < n; ++d) gpu_list.push_back(d);
    } else if (a == "--gpu-list") {
      gpu_list.clear();
      for (const char* q = val(); *q;) {
        char* end = nullptr;
        gpu_list.push_back((int32_t)strtol(q, &end, 10));
        if (end == q) die("--gpu-list takes device ordinals separated by commas", "");
        q = *end ? end + 1 : end;
      }
    }
    else if (a == "--fp64-pairs") fp64 = 1;
    else if (a == "--verbose") verbose = 1;
    else if (a == "--help" || a == "-h") { usage(); return 0; }
    else if (a == "--field") field = val();
    else if (a == "--field-matrix") field_path = val();
    else if (a == "--geno-error") geno_error = atof(val());
    else if (a == "--vcf") vcf_path = val();
    else if (a == "--min-mac") min_mac = atoi(val());              // :47
    else if (a == "--min-callrate") min_callrate = atof(val());    // :48
    else if (a == "--sam") sam_path = val();                       // :39
    else if (a == "--tag-group") so.tag_group = val();             // :40
    else if (a == "--tag-UMI") so.tag_umi = val();                 // :41
    else if (a == "--cap-BQ") so.cap_bq = atoi(val());             // :60-64
    else if (a == "--min-BQ") so.min_bq = atoi(val());
    else if (a == "--min-MQ") so.min_mq = atoi(val());
    else if (a == "--min-TD") so.min_td = atoi(val());
    else if (a == "--excl-flag") so.exclude_flag = (uint32_t)strtoul(val(), nullptr, 0);
    else if (a == "--sm") sm_ids.push_back(val());                 // :50
    else if (a == "--sm-list") sm_list = val();                    // :51
    else die("Unrecognized option %s", argv[i]);
  }
  if ((!plp_path && !sam_path) || !out) { usage(); die("--plp (or --sam with --vcf) and --out are required"); }
  if (plp_path && sam_path) die("--plp and --sam are two sources of the same pileup; give one");
  if (sam_path && !vcf_path) die("--sam needs --vcf (the SNPs the reads are piled up on)");
  if ((sm_list || !sm_ids.empty()) && !vcf_path) die("--sm / --sm-list select samples of --vcf", "");
  if (alpha.empty()) { alpha.push_back(0.0); alpha.push_back(0.5); }   // :78-90
  // :121-139: a tag is empty or exactly two characters
  if (so.tag_group && so.tag_group[0] && strlen(so.tag_group) != 2)
    die("Cannot recognize group tag %s. It is suppose to be a length 2 string", so.tag_group);
  if (so.tag_umi && so.tag_umi[0] && strlen(so.tag_umi) != 2)
    die("Cannot recognize UMI tag %s. It is suppose to be a length 2 string", so.tag_umi);
  if (!gpu_list.empty()) gpu = gpu_list[0];   // the VCF-side device work (site filter) runs on the first one

  if (field && !field_path && !vcf_path)
    die("--field %s needs --vcf FILE or --field-matrix FILE (the FORMAT values as flat arrays)", field);
  if (field_path && vcf_path) die("--vcf and --field-matrix are two sources of the same matrix; give one");
  Plp p;
  if (plp_path) load_plp(plp_path, &p);
  FieldFile ff;
  if (field_path) {
    load_field(field_path, &ff);
    static const char* const names[3] = {"GT", "GP", "PL"};
    if (field && strcmp(field, names[ff.field]) != 0) die("--field %s does not match the field matrix (%s)", field, names[ff.field]);
    if (ff.nv != p.nv || ff.nsnps != p.nsnps) die("Field matrix %s does not have the samples / SNPs of the pileup", field_path);
    if (geno_error >= 0) ff.geno_error = geno_error;   // :46
  }
  if (vcf_path) {
    // the VCF side of the reference's sweep (:142-231): records in file order, the genotype-count filter
    // (--min-mac / --min-callrate, always on GT), the survivors are SNP ids 0, 1, ... of the pileup (scl.add_snp)
    static const char* const names[3] = {"GT", "GP", "PL"};
    if (!field) field = "GP";                          // :9
    ff.field = DMX_FIELD_GP;                           // any other name is read as posteriors like GP
    for (int f = 0; f < 3; ++f)                        // (parse_posteriors' last branch, bcf_filtered_reader.cpp:408)
      if (strcmp(field, names[f]) == 0) ff.field = f;
    dmx_vcf* v = nullptr;
    if (dmx_vcf_read_key(vcf_path, ff.field, ff.field == DMX_FIELD_GP ? field : nullptr, &v) != DMX_OK)
      die("Cannot read VCF: %s", dmx_vcf_last_error());
    if (sm_list) {                                     // bcf_filtered_reader.cpp:95-101: first field of every line
      FILE* sf = fopen(sm_list, "r");
      if (!sf) die("Cannot open file %s for reading", sm_list);
      char line[4096];
      while (fgets(line, sizeof(line), sf)) {
        const size_t n = strcspn(line, " \t\r\n");
        if (n) sm_ids.emplace_back(line, n);
      }
      fclose(sf);
    }
    if (!sm_ids.empty()) {
      std::vector<const char*> nm;
      for (const std::string& x : sm_ids) nm.push_back(x.c_str());
      if (dmx_vcf_select(v, (int32_t)nm.size(), nm.data()) != DMX_OK) die("%s", dmx_vcf_last_error());
    }
    int32_t vnv = 0;
    int64_t nrec = 0, nskip = 0;
    dmx_vcf_sizes(v, &vnv, &nrec, &nskip, nullptr);
    if (plp_path && vnv != p.nv) die("VCF %s does not have the samples of the pileup", vcf_path);
    std::vector<int8_t> gt((size_t)nrec * vnv * 2);
    const size_t per = (size_t)vnv * (ff.field == DMX_FIELD_GT ? 2 : 12);
    std::vector<unsigned char> vals((size_t)nrec * per);
    dmx_vcf_copy(v, gt.data(), vals.data(), nullptr, nullptr, nullptr, nullptr);
    std::vector<uint8_t> pass((size_t)nrec);
    const int frc = dmx_site_filter(gpu, vnv, nrec, gt.data(), min_mac, min_callrate, pass.data());
    if (frc != DMX_OK) die("libdmx: %s", dmx_strerror(frc));
    {
      // a record without the requested FORMAT key stops the reference only when it has passed the filter (:153,:211-212)
      std::vector<uint8_t> no_field((size_t)nrec);
      std::vector<int32_t> rid((size_t)nrec), pos((size_t)nrec);
      dmx_vcf_no_field(v, no_field.data());
      dmx_vcf_copy(v, nullptr, nullptr, rid.data(), pos.data(), nullptr, nullptr);
      for (int64_t r = 0; r < nrec; ++r)
        if (pass[(size_t)r] && no_field[(size_t)r])
          die("Cannot parse posterior probability at %s:%s", dmx_vcf_chrom(v, rid[(size_t)r]),
              std::to_string(pos[(size_t)r] + 1).c_str());
    }
    int64_t keep = 0;
    for (int64_t r = 0; r < nrec; ++r)
      if (pass[(size_t)r]) {
        if (keep != r) memmove(vals.data() + (size_t)keep * per, vals.data() + (size_t)r * per, per);
        ++keep;
      }
    vals.resize((size_t)keep * per);
    if (verbose)
      fprintf(stderr, "NOTICE - VCF: %lld biallelic records, %lld others skipped, %lld pass --min-mac %d --min-callrate %g\n",
              (long long)nrec, (long long)nskip, (long long)keep, min_mac, min_callrate);
    if (plp_path && keep != p.nsnps) die("The records of %s that pass the filters are not the SNPs of the pileup", vcf_path);
    p.samples.resize((size_t)vnv);
    for (int32_t j = 0; j < vnv; ++j) p.samples[(size_t)j] = dmx_vcf_sample(v, j);   // :506: names from the VCF header
    if (sam_path) {
      // the BAM side (:186-338): sweep the reads over the surviving records, then the sort-based pileup packer in
      // place of scl.add_read; --group-list acts inside the sweep like bcdSet (:131-140, :258)
      std::vector<int32_t> rid((size_t)nrec), pos((size_t)nrec);
      std::vector<char> ref((size_t)nrec), alt((size_t)nrec);
      dmx_vcf_copy(v, nullptr, nullptr, rid.data(), pos.data(), ref.data(), alt.data());
      int64_t k2 = 0;
      for (int64_t r = 0; r < nrec; ++r)
        if (pass[(size_t)r]) { rid[(size_t)k2] = rid[(size_t)r]; pos[(size_t)k2] = pos[(size_t)r]; ref[(size_t)k2] = ref[(size_t)r]; alt[(size_t)k2] = alt[(size_t)r]; ++k2; }
      int32_t nchrom = 0;
      dmx_vcf_sizes(v, nullptr, nullptr, nullptr, &nchrom);
      std::vector<const char*> chroms((size_t)nchrom);
      for (int32_t c = 0; c < nchrom; ++c) chroms[(size_t)c] = dmx_vcf_chrom(v, c);
      std::vector<std::string> keep_bcs;
      std::vector<const char*> keep_ptr;
      if (groups) {
        FILE* gf = fopen(groups, "r");
        if (!gf) die("Cannot open file %s for reading", groups);
        char line[4096];
        while (fgets(line, sizeof(line), gf)) {
          const size_t n = strcspn(line, " \t\r\n");
          if (n) keep_bcs.emplace_back(line, n);
        }
        fclose(gf);
        for (const std::string& b : keep_bcs) keep_ptr.push_back(b.c_str());
        so.n_groups = (int64_t)keep_ptr.size();
        so.groups = keep_ptr.data();
        groups = nullptr;
      }
      dmx_sweep* sw = nullptr;
      if (dmx_sweep_bam(sam_path, &so, nchrom, chroms.data(), keep, rid.data(), pos.data(), ref.data(), alt.data(), &sw) != DMX_OK)
        die("Cannot sweep %s: %s", sam_path, dmx_sweep_last_error());
      int64_t nc = 0, nr = 0, ub = 0, cnt[4] = {0, 0, 0, 0};
      dmx_sweep_sizes(sw, &nc, &nr, &ub);
      dmx_sweep_counts(sw, &cnt[0], &cnt[1], &cnt[2], &cnt[3]);
      if (cnt[0] > 0 && cnt[2] + cnt[1] == cnt[0] && cnt[2] > 0)   // :176-178
        die("Your VCF/BCF files and SAM/BAM/CRAM files does not have any matching chromosomes%s", "");
      std::vector<int32_t> rc_cell((size_t)nr), rc_snp((size_t)nr);
      std::vector<uint8_t> rc_al((size_t)nr), rc_bq((size_t)nr);
      std::vector<char> umi_buf((size_t)std::max<int64_t>(ub, 1));
      std::vector<int64_t> umi_off((size_t)nr + 1);
      p.totl.resize((size_t)nc);
      dmx_sweep_copy(sw, rc_cell.data(), rc_snp.data(), rc_al.data(), rc_bq.data(), umi_buf.data(), umi_off.data(), p.totl.data());
      p.barcodes.resize((size_t)nc);
      for (int64_t i = 0; i < nc; ++i) p.barcodes[(size_t)i] = dmx_sweep_barcode(sw, i);
      dmx_sweep_free(sw);
      dmx_pileup* pk = nullptr;
      if (dmx_pack_reads(nr, rc_cell.data(), rc_snp.data(), rc_al.data(), rc_bq.data(), umi_buf.data(), umi_off.data(), nc, 0, &pk) != DMX_OK)
        die("dmx_pack_reads failed", "");
      dmx_pileup_sizes(pk, &p.ncells, &p.E, &p.R);
      p.cell_ptr.resize((size_t)p.ncells + 1); p.ent_snp.resize((size_t)p.E); p.ent_ptr.resize((size_t)p.E + 1);
      p.reads.resize((size_t)p.R); p.npass.resize((size_t)p.ncells); p.uniq.resize((size_t)p.ncells);
      dmx_pileup_copy(pk, p.cell_ptr.data(), p.ent_snp.data(), p.ent_ptr.data(), p.reads.data(), p.npass.data(), p.uniq.data());
      dmx_pileup_free(pk);
      p.nv = vnv; p.nsnps = keep;
      if (verbose)
        fprintf(stderr, "NOTICE - BAM: %lld records, %lld fail --min-MQ / --excl-flag, %lld off the VCF's chromosomes, %lld outside "
                "--group-list; %lld (read, SNP) observations\n", (long long)cnt[0], (long long)cnt[1], (long long)cnt[2],
                (long long)cnt[3], (long long)nr);
      if (p.nv < 2) die("demuxlet needs at least two samples in the VCF", "");
    }
    dmx_vcf_free(v);
    ff.nv = vnv; ff.nsnps = keep;
    ff.geno_error = geno_error >= 0 ? geno_error : 0.01;   // :10
    ff.data.swap(vals);
    field_path = vcf_path;
  }
  if (groups) restrict_to_groups(groups, &p);
  if (verbose)
    fprintf(stderr, "NOTICE - %lld droplets, %lld SNPs, %lld samples, %lld (droplet, SNP) entries, %lld reads\n",
            (long long)p.ncells, (long long)p.nsnps, (long long)p.nv, (long long)p.E, (long long)p.R);

  if (p.ncells == 0) {
    // no droplet at all (empty BAM, --group-list matching nothing): the reference's loops run over an empty bc_map and
    // leave the header lines (:470, :533, :570-571)
    const std::string pre(out);
    struct { const char* ext; const char* hdr; } files[4] = {
        {".single", "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n"},
        {".sing2", "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n"},
        {".best", "BARCODE\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tBEST\tSNG.1ST\tSNG.LLK1\tSNG.2ND\tSNG.LLK2\tSNG.LLK0\tDBL.1ST\t"
                  "DBL.2ND\tALPHA\tLLK12\tLLK1\tLLK2\tLLK10\tLLK20\tLLK00\tPRB.DBL\tPRB.SNG1\n"},
        {".pair", "BARCODE\tSM1.ID\tSM2.ID\tLLK12\tPOSTPRB\n"}};
    for (int f = 0; f < (write_pair ? 4 : 3); ++f) {
      FILE* w = fopen((pre + files[f].ext).c_str(), "w");
      if (!w) die("Cannot open file %s for writing", (pre + files[f].ext).c_str());
      fputs(files[f].hdr, w);
      fclose(w);
    }
    fprintf(stderr, "NOTICE - no droplet to score: header-only output files written\n");
    return 0;
  }
  dmx_config cfg;
  dmx_config_default(&cfg);
  cfg.nv = (int32_t)p.nv;
  cfg.n_alpha = (int32_t)alpha.size();
  cfg.alpha = alpha.data();
  cfg.doublet_prior = prior;
  cfg.min_total = min_total; cfg.min_uniq = min_uniq; cfg.min_snp = min_snp;
  cfg.device = gpu;
  if (!fp64) cfg.flags |= DMX_FP32_PAIRS;
  dmx_ctx* ctx = nullptr;
  int rc = gpu_list.size() > 1 ? dmx_create_sharded(&ctx, &cfg, (int32_t)gpu_list.size(), gpu_list.data())
                               : dmx_create(&ctx, &cfg);
  if (rc == DMX_OK)
    rc = field_path ? dmx_set_field(ctx, ff.field, ff.data.data(), ff.nsnps, ff.geno_error)
                    : dmx_set_gp(ctx, p.gp.data(), p.nsnps);
  if (rc == DMX_OK)
    rc = dmx_set_pileup(ctx, p.ncells, p.cell_ptr.data(), p.ent_snp.data(), p.ent_ptr.data(), p.reads.data(),
                        p.totl.data(), p.npass.data(), p.uniq.data());
  if (rc == DMX_OK) rc = dmx_run(ctx);
  if (rc == DMX_OK) {
    std::vector<const char*> bcs(p.ncells), sms(p.nv);
    for (int64_t i = 0; i < p.ncells; ++i) bcs[i] = p.barcodes[i].c_str();
    for (int64_t j = 0; j < p.nv; ++j) sms[j] = p.samples[j].c_str();
    rc = dmx_write_outputs(ctx, bcs.data(), sms.data(), out, write_pair);
  }
  if (rc != DMX_OK) {
    const std::string why = std::string(dmx_strerror(rc)) + (rc == DMX_E_CUDA ? std::string(": ") + dmx_last_error() : "");
    dmx_destroy(ctx);
    die("libdmx: %s", why.c_str());
  }
  if (verbose) {
    dmx_stats st;
    dmx_get_stats(ctx, &st);
    fprintf(stderr, "NOTICE - %lld kernel launches, copies %.2f ms, kernels %.2f ms (pair scan %.2f ms), finaliser %.2f ms\n",
            (long long)st.gpu_launches, st.h2d_ms, st.kernel_ms, st.pair_kernel_ms, st.finalise_ms);
    for (int sh = 0; sh < dmx_shard_count(ctx) && dmx_shard_count(ctx) > 1; ++sh) {
      int64_t lo = 0, hi = 0;
      int32_t dev = 0;
      dmx_stats ss;
      dmx_shard_range(ctx, sh, &lo, &hi, &dev);
      dmx_shard_stats(ctx, sh, &ss);
      fprintf(stderr, "NOTICE - shard %d: GPU %d, droplets [%lld, %lld), %lld entries, kernels %.2f ms\n", sh, (int)dev,
              (long long)lo, (long long)hi, (long long)ss.n_entries, ss.kernel_ms);
    }
  }
  dmx_destroy(ctx);
  return 0;
}

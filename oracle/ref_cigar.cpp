/*
 * ref_cigar.cpp -- drives the REFERENCE's own base look-up through the CIGAR on flat read records.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product.
 *
 * bam_get_base_and_qual_and_read_and_qual (/root/reference/hts_utils.cpp:279-359) with its helpers
 * bam_get_seq_string / bam_get_qual_string (:174-196) is what the BAM sweep calls per (read, SNP)
 * (cmd_cram_demuxlet.cpp:304).  hts_utils.cpp as a whole needs htslib; oracle/Makefile extracts exactly those line
 * ranges from the reference source WHERE IT LIES into oracle/_ref/ (git-ignored) and this harness #includes them
 * unchanged after supplying the htslib names they use -- the bam1_t layout and the accessor macros of htslib's
 * published sam.h, kstring's kputc / kputw, and the two macros of the reference's own hts_utils.h (:125, :234).
 *
 * Input (native endian): int32 n; per case int32 pos, int32 n_cigar, uint32 cigar[n_cigar] (len << 4 | op),
 * int32 l_seq, uint8 seq4[(l_seq + 1) / 2], uint8 qual[l_seq], int32 nq, uint32 qpos[nq].
 * Output: per query int8 base, int8 qual (as the function returns them: qual + 33 in a char), int32 rpos.
 */
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

/* htslib kstring.h */
typedef struct { size_t l, m; char* s; } kstring_t;
static inline void ks_grow(kstring_t* s, size_t need) {
  if (s->l + need + 1 > s->m) { s->m = (s->l + need + 1) * 2 + 16; s->s = (char*)realloc(s->s, s->m); }
}
static inline int kputc(int c, kstring_t* s) { ks_grow(s, 1); s->s[s->l++] = (char)c; s->s[s->l] = 0; return c; }
static inline int kputw(int c, kstring_t* s) {
  char buf[16];
  const int n = snprintf(buf, sizeof(buf), "%d", c);
  ks_grow(s, (size_t)n);
  memcpy(s->s + s->l, buf, (size_t)n);
  s->l += (size_t)n;
  s->s[s->l] = 0;
  return 0;
}
/* htslib sam.h */
typedef struct {
  int32_t tid, pos;
  uint16_t bin;
  uint8_t qual, l_qname;
  uint16_t flag, n_cigar;
  int32_t l_qseq, mtid, mpos, isize;
} bam1_core_t;
typedef struct { bam1_core_t core; int l_data; uint32_t m_data; uint8_t* data; } bam1_t;
#define BAM_CIGAR_STR "MIDNSHP=XB"
#define bam_cigar_op(c) ((c) & 0xf)
#define bam_cigar_oplen(c) ((c) >> 4)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR[bam_cigar_op(c)])
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)
/* the reference's hts_utils.h:125, :234 */
#define BAM_READ_INDEX_NA -1
#define bam_get_l_qseq(s) ((s)->core.l_qseq)

#include "_ref/hts_174_196.inc"
#include "_ref/hts_279_359.inc"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_cigar <in> <out>\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  FILE* o = fopen(argv[2], "wb");
  if (!f || !o) { perror("open"); return 2; }
  int32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) return 2;
  kstring_t rs = {0, 0, 0}, rq = {0, 0, 0};
  for (int32_t c = 0; c < n; ++c) {
    int32_t pos, ncig, lseq, nq;
    if (fread(&pos, 4, 1, f) != 1 || fread(&ncig, 4, 1, f) != 1) return 2;
    std::vector<uint8_t> data(4 * (size_t)ncig);     /* l_qname = 0: data = cigar | seq | qual */
    if (ncig && fread(data.data(), 4, (size_t)ncig, f) != (size_t)ncig) return 2;
    if (fread(&lseq, 4, 1, f) != 1) return 2;
    const size_t off = data.size(), sb = ((size_t)lseq + 1) / 2;
    data.resize(off + sb + (size_t)lseq + 8);
    if (sb && fread(data.data() + off, 1, sb, f) != sb) return 2;
    if (lseq && fread(data.data() + off + sb, 1, (size_t)lseq, f) != (size_t)lseq) return 2;
    if (fread(&nq, 4, 1, f) != 1) return 2;
    bam1_t b;
    memset(&b, 0, sizeof(b));
    b.core.pos = pos; b.core.n_cigar = (uint16_t)ncig; b.core.l_qseq = lseq; b.core.l_qname = 0;
    b.data = data.data();
    for (int32_t q = 0; q < nq; ++q) {
      uint32_t qp;
      if (fread(&qp, 4, 1, f) != 1) return 2;
      char base = 0, qual = 0;
      int32_t rpos = 0;
      bam_get_base_and_qual_and_read_and_qual(&b, qp, base, qual, rpos, &rs, &rq);
      fwrite(&base, 1, 1, o); fwrite(&qual, 1, 1, o); fwrite(&rpos, 4, 1, o);
    }
  }
  fclose(o);
  fclose(f);
  return 0;
}

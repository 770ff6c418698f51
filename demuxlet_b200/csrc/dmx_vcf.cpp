// dmx_vcf.cpp -- htslib-free VCF / BCF reader for the arrays the field kernels take (VCF half; host only).
//
// The reference reads VCF / BCF records through htslib (BCFChunkedReader::read, /root/reference/bcf_chunked_reader.cpp:
// 265-305: bcf_read / vcf_parse1) and asks the record for its FORMAT values with bcf_get_genotypes /
// bcf_get_format_float / bcf_get_format_int32 (bcf_filtered_reader.cpp:194, :250, :408).  htslib is not in this image,
// so this file parses the text form itself -- plain, gzip or bgzip -- and the binary form, BCF2 (recognised by its magic
// after decompression; section "BCF2" below), and delivers, for every record, what those accessors would:
//   GT   two allele indices per sample as bcf_gt_allele() decodes them ("." -> -1, a haploid call "1" -> (1, -1):
//        htslib pads it with bcf_int32_vector_end, which decodes to a negative allele)
//   GP   three floats per sample, "." or a short vector -> NaN (htslib: bcf_float_missing / vector_end, both NaNs)
//   PL   three int32 per sample, "." or a short vector -> INT32_MIN (bcf_int32_missing; vector_end = INT32_MIN + 1 reads
//        the same through PhredHelper.h:40's uint32_t cast: > 255)
// A record whose FORMAT lacks the requested key is delivered with all values missing and flagged (dmx_vcf_no_field): the
// reference fails on it only if it passes the genotype filter (parse_posteriors, cmd_cram_demuxlet.cpp:153,:211).
// Records are delivered in file order; records whose allele count is not 2 are skipped and counted (demuxlet sets
// maxAlleles = 2, cmd_cram_demuxlet.cpp:26; a record without ALT cannot pass --min-mac 1).  All samples are selected
// (dmx_vcf_select applies --sm / --sm-list afterwards).  The chromosome ids follow the header's ##contig lines, then first
// appearance in the records (htslib's dictionary: bcf_hdr_name2id, cmd_cram_demuxlet.cpp:197-199).
// The file is streamed window by window (dmx_bgzf.h): a window's lines / records are parsed in contiguous ranges on the
// host threads straight into the arrays; memory is one window plus the arrays.
// Pins: the reference's tutorial VCF against an independent parse and the example file of the VCF specification
// (tests/test_vcf.py); beyond those the python restatement (tests/vcfio.py) on generated files -- "parity unpinned"
// against htslib itself.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <zlib.h>

#include <algorithm>
#include <limits>
#include <functional>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/dmx.h"
#include "dmx_bgzf.h"

struct dmx_vcf {
  int32_t field = DMX_FIELD_GT;
  std::vector<std::string> samples, chroms;
  std::vector<int32_t> rid, pos;      // pos 0-based like bcf1_t::pos
  std::vector<char> ref, alt;         // first base of REF / ALT (what add_snp keeps, cmd_cram_demuxlet.cpp:230)
  std::vector<int8_t> gt;             // [nrec][nv][2]
  std::vector<float> gp;              // [nrec][nv][3] (field GP)
  std::vector<int32_t> pl;            // [nrec][nv][3] (field PL)
  std::vector<uint8_t> no_field;      // [nrec] 1: the record's FORMAT lacks the requested key (values all missing)
  int64_t n_skipped = 0;              // records with n_allele != 2
  std::string error;
};

namespace {

thread_local std::string g_vcf_error;

struct Tok {
  const char* p;
  size_t n;
};

// split [p, e) at `sep`
void split(const char* p, const char* e, char sep, std::vector<Tok>* out) {
  out->clear();
  const char* s = p;
  for (const char* q = p; q <= e; ++q)
    if (q == e || *q == sep) {
      out->push_back(Tok{s, (size_t)(q - s)});
      s = q + 1;
    }
}

const float kNaN = std::numeric_limits<float>::quiet_NaN();

// the idx-th ':'-separated subfield of a sample column, or an empty token
Tok subfield(const Tok& col, int idx) {
  const char *p = col.p, *e = col.p + col.n;
  for (int i = 0; i < idx; ++i) {
    const char* c = (const char*)memchr(p, ':', (size_t)(e - p));
    if (!c) return Tok{e, 0};
    p = c + 1;
  }
  const char* c = (const char*)memchr(p, ':', (size_t)(e - p));
  return Tok{p, (size_t)((c ? c : e) - p)};
}

void parse_gt(const Tok& t, int8_t* out) {
  out[0] = out[1] = -1;
  const char *p = t.p, *e = t.p + t.n;
  for (int a = 0; a < 2 && p < e; ++a) {
    if (*p == '.') {
      ++p;
    } else {
      int v = 0;
      bool any = false;
      while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; }
      if (any) out[a] = (int8_t)(v > 126 ? 126 : v);
    }
    if (p < e && (*p == '/' || *p == '|')) ++p;
    else break;
  }
}

template <class T, class F>
void parse_vec3(const Tok& t, T* out, T missing, F conv) {
  out[0] = out[1] = out[2] = missing;
  const char *p = t.p, *e = t.p + t.n;
  for (int k = 0; k < 3 && p < e; ++k) {
    const char* c = (const char*)memchr(p, ',', (size_t)(e - p));
    const char* q = c ? c : e;
    if (!(q - p == 1 && *p == '.') && q > p) out[k] = conv(p, q);
    if (!c) break;
    p = c + 1;
  }
}

// htslib parses a FORMAT float with strtod and stores it in a float (vcf.c, vcf_parse_format): (float)strtod(text).
// Plain decimal text -- [-+]digits[.digits][e[-+]digits] with a mantissa below 2^53 and a decimal exponent within
// +-22 -- is converted exactly as strtod does by ONE correctly rounded double operation on two exactly representable
// numbers (Clinger's fast path); everything else (inf, nan, hex, long mantissas, big exponents) goes to strtod itself.
float conv_float(const char* p, const char* q) {
  static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char* s = p;
  bool neg = false;
  if (s < q && (*s == '-' || *s == '+')) neg = *s++ == '-';
  uint64_t m = 0;
  int digits = 0, frac = 0;
  bool any = false, ok = true;
  for (; s < q && *s >= '0' && *s <= '9'; ++s) { m = m * 10 + (uint64_t)(*s - '0'); any = true; if (m) ++digits; }
  if (s < q && *s == '.') {
    ++s;
    for (; s < q && *s >= '0' && *s <= '9'; ++s) { m = m * 10 + (uint64_t)(*s - '0'); any = true; ++frac; if (m) ++digits; }
  }
  int ex = 0;
  if (any && s < q && (*s == 'e' || *s == 'E')) {
    const char* t = s + 1;
    bool eneg = false;
    if (t < q && (*t == '-' || *t == '+')) eneg = *t++ == '-';
    if (t < q && *t >= '0' && *t <= '9') {
      for (; t < q && *t >= '0' && *t <= '9'; ++t) { ex = ex * 10 + (*t - '0'); if (ex > 9999) ok = false; }
      if (eneg) ex = -ex;
      s = t;
    }
  }
  const int k = ex - frac;
  if (any && ok && s == q && digits <= 15 && k >= -22 && k <= 22) {
    const double d = k < 0 ? (double)m / kPow10[-k] : (double)m * kPow10[k];
    return (float)(neg ? -d : d);
  }
  char tmp[128];
  const size_t n = std::min<size_t>((size_t)(q - p), sizeof(tmp) - 1);
  memcpy(tmp, p, n);
  tmp[n] = 0;
  return (float)strtod(tmp, nullptr);
}

int32_t conv_int(const char* p, const char* q) {
  long v = 0;
  bool neg = false;
  if (p < q && *p == '-') { neg = true; ++p; }
  for (; p < q && *p >= '0' && *p <= '9'; ++p) v = v * 10 + (*p - '0');
  return (int32_t)(neg ? -v : v);
}

}  // namespace

namespace {

struct Line { const char *p, *e; int64_t no; };

// Records of one contiguous range of lines [lo, hi), parsed by one host thread straight into the arrays of the whole
// file at record slots base + lo, base + lo + 1, ... behind the `base` records read so far (a skipped record leaves its
// slot to the next one; the ranges are closed up afterwards).  Chromosome ids are local to the part until the merge.
struct Part {
  std::vector<std::string> chroms;
  size_t n = 0;                      // records delivered into slots [lo, lo + n)
  int64_t n_skipped = 0;
  std::string error;
  int64_t err_line = 0;
};

void parse_range(const std::vector<Line>& lines, size_t lo, size_t hi, size_t base, int32_t field, const char* key, size_t nv,
                 dmx_vcf* v, Part* part) {
  const char* want = field == DMX_FIELD_GT ? nullptr : key;
  const size_t want_n = want ? strlen(want) : 0;
  std::vector<Tok> cols, keys, alts;
  auto fail = [&](int64_t no, const std::string& why) { part->error = why; part->err_line = no; };
  size_t slot = base + lo;
  for (size_t li = lo; li < hi; ++li) {
    const char *p = lines[li].p, *e = lines[li].e;
    split(p, e, '\t', &cols);
    if (cols.size() != 9 + nv) return fail(lines[li].no, "wrong number of columns");
    // n_allele = 1 + number of ALT alleles ("." = none)
    split(cols[4].p, cols[4].p + cols[4].n, ',', &alts);
    const bool no_alt = cols[4].n == 1 && cols[4].p[0] == '.';
    size_t chrom_id = part->chroms.size();
    auto is_chrom = [&](size_t c) { return part->chroms[c].size() == cols[0].n && memcmp(part->chroms[c].data(), cols[0].p, cols[0].n) == 0; };
    if (!part->chroms.empty() && is_chrom(part->chroms.size() - 1)) chrom_id = part->chroms.size() - 1;
    else {
      for (chrom_id = 0; chrom_id < part->chroms.size(); ++chrom_id)
        if (is_chrom(chrom_id)) break;
      if (chrom_id == part->chroms.size()) part->chroms.emplace_back(cols[0].p, cols[0].n);   // htslib adds an undeclared contig on sight
    }
    if (no_alt || alts.size() != 1) { ++part->n_skipped; continue; }
    split(cols[8].p, cols[8].p + cols[8].n, ':', &keys);
    int i_gt = -1, i_f = -1;
    for (size_t k = 0; k < keys.size(); ++k) {
      if (keys[k].n == 2 && memcmp(keys[k].p, "GT", 2) == 0) i_gt = (int)k;
      if (want && keys[k].n == want_n && memcmp(keys[k].p, want, want_n) == 0) i_f = (int)k;
    }
    if (i_gt < 0) return fail(lines[li].no, "no GT in FORMAT (demuxlet's record filter needs it)");
    // a record without the requested key is an error in the reference only if it gets as far as parse_posteriors,
    // i.e. passes the genotype filter (cmd_cram_demuxlet.cpp:153,:211): flag it, the caller decides after the filter
    v->no_field[slot] = (uint8_t)(want && i_f < 0);
    v->rid[slot] = (int32_t)chrom_id;
    v->pos[slot] = conv_int(cols[1].p, cols[1].p + cols[1].n) - 1;
    v->ref[slot] = cols[3].n ? cols[3].p[0] : 'N';
    v->alt[slot] = cols[4].n ? cols[4].p[0] : 'N';
    int8_t* gt = &v->gt[slot * nv * 2];
    float* gp = field == DMX_FIELD_GP ? &v->gp[slot * nv * 3] : nullptr;
    int32_t* pl = field == DMX_FIELD_PL ? &v->pl[slot * nv * 3] : nullptr;
    for (size_t s = 0; s < nv; ++s) {
      const Tok& col = cols[9 + s];
      parse_gt(subfield(col, i_gt), gt + s * 2);
      const Tok fcol = i_f >= 0 ? subfield(col, i_f) : Tok{col.p, 0};
      if (gp) parse_vec3<float>(fcol, gp + s * 3, kNaN, conv_float);
      if (pl) parse_vec3<int32_t>(fcol, pl + s * 3, std::numeric_limits<int32_t>::min(), conv_int);
    }
    ++slot;
  }
  part->n = slot - base - lo;
}


// ---- BCF2 (the binary form of VCF, BGZF-compressed; VCF specification 6.3) ------------------------------------------
// Header: "BCF\2\2", uint32 l_text, the VCF header text.  Strings are numbered by the header's dictionaries: FILTER /
// INFO / FORMAT ids in order of appearance behind PASS = 0 (or their IDX=), contigs in order (or IDX=).  Record: uint32
// l_shared, l_indiv; shared = CHROM, POS (0-based), rlen, QUAL, n_allele << 16 | n_info, n_fmt << 24 | n_sample, ID,
// alleles, FILTER, INFO; indiv = per FORMAT key: typed int (dictionary index), type byte, n_sample x length values.
// A typed value starts with a byte (length << 4 | type; length 15 = a typed int follows), types 1 / 2 / 3 = int8 / 16 /
// 32, 5 = float, 7 = char; the smallest value of an int type is "missing", the next one "end of vector", as are the
// float bit patterns 0x7F800001 / 0x7F800002.  What bcf_get_genotypes / bcf_get_format_float / bcf_get_format_int32
// hand to the reference: alleles (x >> 1) - 1 (missing and padding negative), floats (missing / padding NaN), int32
// (missing / padding INT32_MIN); a field stored with the wrong type is an error there (-2) -> flagged like a missing key.
struct BcfRec { const uint8_t* p; int64_t no; };

int bcf_type_size(int t) { return t == 1 ? 1 : (t == 2 ? 2 : (t == 3 || t == 5 ? 4 : (t == 7 ? 1 : 0))); }

bool bcf_int(const uint8_t*& p, const uint8_t* e, int t, int64_t* out) {
  const int sz = bcf_type_size(t);
  if ((t != 1 && t != 2 && t != 3) || p + sz > e) return false;
  if (t == 1) *out = (int8_t)p[0];
  else if (t == 2) { int16_t x; memcpy(&x, p, 2); *out = x; }
  else { int32_t x; memcpy(&x, p, 4); *out = x; }
  p += sz;
  return true;
}

bool bcf_typed(const uint8_t*& p, const uint8_t* e, int* type, int64_t* len) {
  if (p >= e) return false;
  const uint8_t b = *p++;
  *type = b & 15;
  *len = b >> 4;
  if (*len == 15) {
    if (p >= e) return false;
    const int t2 = *p++ & 15;
    if (!bcf_int(p, e, t2, len) || *len < 0) return false;
  }
  return true;
}

// one int of a vector of type t at q; missing / end of vector -> INT32_MIN
inline int32_t bcf_vec_int(const uint8_t* q, int t) {
  if (t == 1) { const int8_t x = (int8_t)q[0]; return x <= INT8_MIN + 1 ? INT32_MIN : x; }
  if (t == 2) { int16_t x; memcpy(&x, q, 2); return x <= INT16_MIN + 1 ? INT32_MIN : x; }
  int32_t x; memcpy(&x, q, 4); return x <= INT32_MIN + 1 ? INT32_MIN : x;
}

void parse_range_bcf(const std::vector<BcfRec>& recs, size_t lo, size_t hi, size_t base, int32_t field, int64_t id_gt,
                     int64_t id_key, size_t nv, size_t n_contig, dmx_vcf* v, Part* part) {
  auto fail = [&](int64_t no, const std::string& why) { part->error = why; part->err_line = no; };
  size_t slot = base + lo;
  for (size_t ri = lo; ri < hi; ++ri) {
    const uint8_t* r = recs[ri].p;
    uint32_t l_shared, l_indiv;
    memcpy(&l_shared, r, 4);
    memcpy(&l_indiv, r + 4, 4);
    const uint8_t *p = r + 8, *se = r + 8 + l_shared, *ie = se + l_indiv;
    if (l_shared < 24) return fail(recs[ri].no, "corrupt BCF record");
    int32_t chrom, pos;
    uint32_t n_allele_info, n_fmt_sample;
    memcpy(&chrom, p, 4);
    memcpy(&pos, p + 4, 4);
    memcpy(&n_allele_info, p + 16, 4);
    memcpy(&n_fmt_sample, p + 20, 4);
    p += 24;
    const uint32_t n_allele = n_allele_info >> 16, n_fmt = n_fmt_sample >> 24, n_sample = n_fmt_sample & 0xffffffu;
    if (chrom < 0 || (size_t)chrom >= n_contig) return fail(recs[ri].no, "CHROM is not in the header's contig dictionary");
    if (n_allele != 2) { ++part->n_skipped; continue; }
    if (n_sample != nv) return fail(recs[ri].no, "wrong number of samples");
    int type;
    int64_t len;
    if (!bcf_typed(p, se, &type, &len) || p + len > se) return fail(recs[ri].no, "corrupt BCF record");   // ID
    p += len;
    char a0[2] = {'N', 'N'};
    for (int a = 0; a < 2; ++a) {
      if (!bcf_typed(p, se, &type, &len) || type != 7 || p + len > se) return fail(recs[ri].no, "corrupt BCF record");
      if (len > 0) a0[a] = (char)p[0];
      p += len;
    }
    // FORMAT fields
    const uint8_t *gt_p = nullptr, *f_p = nullptr;
    int gt_t = 0, f_t = 0;
    int64_t gt_n = 0, f_n = 0;
    p = se;
    for (uint32_t f = 0; f < n_fmt; ++f) {
      int kt;
      int64_t kl, key;
      if (!bcf_typed(p, ie, &kt, &kl) || kl != 1 || !bcf_int(p, ie, kt, &key)) return fail(recs[ri].no, "corrupt BCF record");
      if (!bcf_typed(p, ie, &type, &len)) return fail(recs[ri].no, "corrupt BCF record");
      const int64_t bytes = (int64_t)bcf_type_size(type) * len * (int64_t)nv;
      if (p + bytes > ie) return fail(recs[ri].no, "corrupt BCF record");
      if (key == id_gt) { gt_p = p; gt_t = type; gt_n = len; }
      if (field != DMX_FIELD_GT && key == id_key) { f_p = p; f_t = type; f_n = len; }
      p += bytes;
    }
    if (!gt_p || gt_t < 1 || gt_t > 3) return fail(recs[ri].no, "no GT in FORMAT (demuxlet's record filter needs it)");
    const bool f_ok = field == DMX_FIELD_GT || (f_p && (field == DMX_FIELD_GP ? f_t == 5 : (f_t >= 1 && f_t <= 3)));
    v->no_field[slot] = (uint8_t)!f_ok;
    v->rid[slot] = chrom;
    v->pos[slot] = pos;
    v->ref[slot] = a0[0];
    v->alt[slot] = a0[1];
    int8_t* gt = &v->gt[slot * nv * 2];
    float* gp = field == DMX_FIELD_GP ? &v->gp[slot * nv * 3] : nullptr;
    int32_t* pl = field == DMX_FIELD_PL ? &v->pl[slot * nv * 3] : nullptr;
    const int gsz = bcf_type_size(gt_t), fsz = bcf_type_size(f_t);
    for (size_t s = 0; s < nv; ++s) {
      for (int a = 0; a < 2; ++a) {
        int32_t x = INT32_MIN;
        if (a < gt_n) x = bcf_vec_int(gt_p + ((int64_t)s * gt_n + a) * gsz, gt_t);
        const int32_t al = x == INT32_MIN ? -1 : (x >> 1) - 1;            // bcf_gt_allele
        gt[s * 2 + a] = (int8_t)(al > 126 ? 126 : (al < -1 ? -1 : al));
      }
      if (gp)
        for (int k = 0; k < 3; ++k) {
          float x = kNaN;
          if (f_ok && k < f_n) {
            uint32_t bits;
            memcpy(&bits, f_p + ((int64_t)s * f_n + k) * fsz, 4);
            if (bits != 0x7F800001u && bits != 0x7F800002u) memcpy(&x, &bits, 4);
          }
          gp[s * 3 + k] = x;
        }
      if (pl)
        for (int k = 0; k < 3; ++k)
          pl[s * 3 + k] = (f_ok && k < f_n) ? bcf_vec_int(f_p + ((int64_t)s * f_n + k) * fsz, f_t) : INT32_MIN;
    }
    ++slot;
  }
  part->n = slot - base - lo;
}

// value of `key` inside the <...> of a header line (quoted values may hold commas)
bool meta_value(const char* q, const char* e, const char* key, std::string* out) {
  const size_t kn = strlen(key);
  while (q < e) {
    const char* k = q;
    while (q < e && *q != '=' && *q != ',' && *q != '>') ++q;
    const bool hit = (size_t)(q - k) == kn && memcmp(k, key, kn) == 0;
    if (q < e && *q == '=') ++q;
    const char* val = q;
    if (q < e && *q == '"') { ++q; while (q < e && *q != '"') ++q; if (q < e) ++q; }
    else while (q < e && *q != ',' && *q != '>') ++q;
    if (hit && q > val) { out->assign(val, (size_t)(q - val)); return true; }
    if (q < e && *q == '>') break;
    if (q < e) ++q;
  }
  return false;
}

// registers `id` in a dictionary (header order, or at IDX=): htslib's bcf_hdr_register_hrec
void dict_add(std::vector<std::string>* dict, const std::string& id, const char* q, const char* e) {
  if (std::find(dict->begin(), dict->end(), id) != dict->end()) return;
  std::string idx;
  size_t at = dict->size();
  if (meta_value(q, e, "IDX", &idx)) at = (size_t)std::max(0, atoi(idx.c_str()));
  if (at >= dict->size()) dict->resize(at + 1);
  (*dict)[at] = id;
}

}  // namespace

extern "C" {

int dmx_vcf_read(const char* path, int32_t field, dmx_vcf** out) {
  return dmx_vcf_read_key(path, field, nullptr, out);
}

int dmx_vcf_read_key(const char* path, int32_t field, const char* key, dmx_vcf** out) {
  if (!path || !out || field < DMX_FIELD_GT || field > DMX_FIELD_PL) return DMX_E_INVALID;
  if (!key || !key[0]) key = field == DMX_FIELD_PL ? "PL" : "GP";
  *out = nullptr;
  const int nthreads = dmx_io::host_threads("DMX_VCF_THREADS");
  // The file is read window by window (a batch of inflated BGZF blocks, or 32 MB of any other stream, behind the
  // unconsumed tail of the last one): memory is the window plus the arrays, and inflating runs ahead of the parsing.
  dmx_io::ByteStream in;
  if (!in.open(path, nthreads)) {
    g_vcf_error = std::string("cannot read ") + path;
    return DMX_E_IO;
  }
  dmx_vcf* v = new dmx_vcf();
  v->field = field;
  auto fail = [&](int64_t lineno, const std::string& why) {
    g_vcf_error = std::string(path) + ": line " + std::to_string(lineno) + ": " + why;
    delete v;
    return DMX_E_INVALID;
  };
  auto fail_io = [&]() {
    g_vcf_error = std::string("cannot read ") + path;
    delete v;
    return DMX_E_IO;
  };
  size_t cur = 0;                                     // read position inside the window
  auto need = [&](size_t n) {                         // at least n bytes from cur on, if the file has them
    while (in.size() - cur < n) {
      const bool got = in.more(cur);
      cur = 0;
      if (!got) return in.size() >= n;
    }
    return true;
  };
  need(9);
  if (in.bad()) return fail_io();
  const bool is_bcf = in.size() >= 9 && memcmp(in.data(), "BCF\2\2", 5) == 0;
  std::vector<std::string> ids(1, "PASS");            // BCF: the FILTER / INFO / FORMAT string dictionary
  std::vector<Tok> cols;
  bool have_header = false;
  int64_t lineno = 0;
  // one header line: contig and (BCF) string dictionaries, sample names
  auto header_line = [&](const char* p, const char* e) -> bool {
    std::string id;
    if (e - p > 13 && memcmp(p, "##contig=<", 10) == 0) {
      // the header's contig dictionary defines the chromosome ids (bcf_hdr_name2id, cmd_cram_demuxlet.cpp:197-199):
      // a read on a declared contig counts for its droplet even when no record of the file lies there
      if (meta_value(p + 10, e, "ID", &id)) dict_add(&v->chroms, id, p + 10, e);
    } else if (is_bcf && ((e - p > 11 && memcmp(p, "##FORMAT=<", 10) == 0) || (e - p > 9 && memcmp(p, "##INFO=<", 8) == 0) ||
                          (e - p > 11 && memcmp(p, "##FILTER=<", 10) == 0))) {
      const char* q = (const char*)memchr(p, '<', (size_t)(e - p)) + 1;
      if (meta_value(q, e, "ID", &id)) dict_add(&ids, id, q, e);
    } else if (e - p > 6 && memcmp(p, "#CHROM", 6) == 0) {
      split(p, e, '\t', &cols);
      if (cols.size() < 10) return false;
      for (size_t i = 9; i < cols.size(); ++i) v->samples.emplace_back(cols[i].p, cols[i].n);
      have_header = true;
    }
    return true;
  };
  // one batch of records (lines of text, or BCF records, all inside the window): contiguous ranges on the host
  // threads, each range into its own slots behind the records read so far; then, in file order, chromosome ids by
  // first appearance and the ranges closed up over the slots of skipped records
  std::vector<Line> lines;
  std::vector<BcfRec> recs;
  int64_t id_gt = -1, id_key = -1;
  dmx_io::Pool pool(nthreads);
  std::vector<Part> parts;
  auto run_batch = [&]() -> int64_t {                 // 0, or the failing line / record number
    const size_t nv = v->samples.size(), nl = is_bcf ? recs.size() : lines.size(), base = v->rid.size();
    if (nl == 0) return 0;
    const size_t nth = std::max<size_t>(1, std::min<size_t>((size_t)nthreads, nl / 256 + 1));
    v->rid.resize(base + nl);
    v->pos.resize(base + nl);
    v->ref.resize(base + nl);
    v->alt.resize(base + nl);
    v->no_field.resize(base + nl);
    v->gt.resize((base + nl) * nv * 2);
    if (field == DMX_FIELD_GP) v->gp.resize((base + nl) * nv * 3);
    if (field == DMX_FIELD_PL) v->pl.resize((base + nl) * nv * 3);
    parts.assign(nth, Part());
    auto lo_of = [&](size_t t) { return nl * t / nth; };
    const size_t n_contig = v->chroms.size();
    pool.run((int)nth, [&](int t) {
      if (is_bcf) parse_range_bcf(recs, lo_of((size_t)t), lo_of((size_t)t + 1), base, field, id_gt, id_key, nv, n_contig, v, &parts[(size_t)t]);
      else parse_range(lines, lo_of((size_t)t), lo_of((size_t)t + 1), base, field, key, nv, v, &parts[(size_t)t]);
    });
    for (const Part& pt : parts)
      if (!pt.error.empty()) { g_vcf_error = pt.error; return pt.err_line; }   // the first failing range = the first failing line
    size_t nrec = base;
    for (size_t t = 0; t < nth; ++t) {
      const Part& pt = parts[t];
      const size_t lo = base + lo_of(t);
      if (!is_bcf) {       // (a BCF record carries the index into the header's contig dictionary itself)
        std::vector<int32_t> remap(pt.chroms.size());
        for (size_t c = 0; c < pt.chroms.size(); ++c) {
          size_t r = 0;
          for (; r < v->chroms.size(); ++r)
            if (v->chroms[r] == pt.chroms[c]) break;
          if (r == v->chroms.size()) v->chroms.push_back(pt.chroms[c]);
          remap[c] = (int32_t)r;
        }
        for (size_t i = 0; i < pt.n; ++i) v->rid[lo + i] = remap[(size_t)v->rid[lo + i]];
      }
      if (nrec != lo && pt.n) {
        auto close_up = [&](auto& vec, size_t per) {
          if (!vec.empty()) memmove(vec.data() + nrec * per, vec.data() + lo * per, pt.n * per * sizeof(vec[0]));
        };
        close_up(v->rid, 1); close_up(v->pos, 1); close_up(v->ref, 1); close_up(v->alt, 1); close_up(v->no_field, 1);
        close_up(v->gt, nv * 2); close_up(v->gp, nv * 3); close_up(v->pl, nv * 3);
      }
      nrec += pt.n;
      v->n_skipped += pt.n_skipped;
    }
    v->rid.resize(nrec);
    v->pos.resize(nrec);
    v->ref.resize(nrec);
    v->alt.resize(nrec);
    v->no_field.resize(nrec);
    v->gt.resize(nrec * nv * 2);
    if (field == DMX_FIELD_GP) v->gp.resize(nrec * nv * 3);
    if (field == DMX_FIELD_PL) v->pl.resize(nrec * nv * 3);
    return 0;
  };

  if (is_bcf) {
    uint32_t l_text;
    memcpy(&l_text, in.data() + 5, 4);
    if (!need((size_t)l_text + 9)) return in.bad() ? fail_io() : fail(0, "truncated BCF header");
    const char *p = (const char*)in.data() + 9, *end = p + l_text;
    while (end > p && end[-1] == 0) --end;
    while (p < end) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
      const char* e = nl ? nl : end;
      ++lineno;
      if (e > p && e[-1] == '\r') --e;
      if (e > p) {
        if (*p != '#') return fail(lineno, "text records inside a BCF header");
        if (!header_line(p, e)) return fail(lineno, "no sample columns");
      }
      p = nl ? nl + 1 : end;
    }
    if (!have_header) return fail(0, "no #CHROM line");
    for (size_t i = 0; i < ids.size(); ++i) {
      if (ids[i] == "GT") id_gt = (int64_t)i;
      if (ids[i] == key) id_key = (int64_t)i;
    }
    cur = 9 + (size_t)l_text;
    int64_t no = 0;
    for (;;) {
      recs.clear();
      const uint8_t* w = in.data();
      const size_t wsize = in.size();
      size_t q = cur;
      while (q + 8 <= wsize) {
        uint32_t l_shared, l_indiv;
        memcpy(&l_shared, w + q, 4);
        memcpy(&l_indiv, w + q + 4, 4);
        if ((uint64_t)l_shared + l_indiv + 8 > (uint64_t)(wsize - q)) break;
        recs.push_back(BcfRec{w + q, ++no});
        q += 8 + (size_t)l_shared + l_indiv;
      }
      const int64_t bad = run_batch();
      if (bad) return fail(bad, g_vcf_error);
      const bool got = in.more(q);
      cur = 0;
      if (!got) {
        if (in.bad()) return fail_io();
        if (in.size() > 0) return fail(no + 1, "truncated BCF record");
        break;
      }
    }
  } else {
    for (bool last = false; !last;) {
      lines.clear();
      const char* w = (const char*)in.data();
      const size_t wsize = in.size();
      size_t q = cur;
      auto take = [&](const char* p, const char* e) -> int {     // 0 ok, else an error code for the messages below
        if (e > p && e[-1] == '\r') --e;
        ++lineno;
        if (e <= p) return 0;
        if (*p == '#') return header_line(p, e) ? 0 : 1;
        if (!have_header) return 2;
        lines.push_back(Line{p, e, lineno});
        return 0;
      };
      int err = 0;
      while (q < wsize && !err) {
        const char* nl = (const char*)memchr(w + q, '\n', wsize - q);
        if (!nl) break;
        err = take(w + q, nl);
        q = (size_t)(nl - w) + 1;
      }
      if (!err) {
        // the rest of the window is an incomplete line -- or, at the end of the file, a last line without a newline
        const int64_t bad = run_batch();
        if (bad) return fail(bad, g_vcf_error);
        const bool got = in.more(q);
        cur = 0;
        if (!got) {
          if (in.bad()) return fail_io();
          last = true;
          if (in.size() > 0) {
            lines.clear();
            err = take((const char*)in.data(), (const char*)in.data() + in.size());
            if (!err) {
              const int64_t bad2 = run_batch();
              if (bad2) return fail(bad2, g_vcf_error);
            }
          }
        }
      }
      if (err) return fail(lineno, err == 1 ? "no sample columns" : "record before the #CHROM line");
    }
    if (!have_header) return fail(0, "no #CHROM line");
  }
  *out = v;
  return DMX_OK;
}

// --sm / --sm-list (bcf_filtered_reader.cpp:95-121): keep the named samples only.  The reference collects the ids in a
// std::set<std::string>, so the kept columns come out in byte-lexicographic order of their names, duplicates merged;
// an id that is not in the header is an error there (:107-108) and here.
int dmx_vcf_select(dmx_vcf* v, int32_t n, const char* const* names) {
  if (!v || n < 1 || !names) return DMX_E_INVALID;
  std::vector<std::string> want(names, names + n);
  std::sort(want.begin(), want.end());
  want.erase(std::unique(want.begin(), want.end()), want.end());
  std::vector<size_t> col;
  for (const std::string& w : want) {
    size_t c = 0;
    for (; c < v->samples.size(); ++c)
      if (v->samples[c] == w) break;
    if (c == v->samples.size()) {
      g_vcf_error = "Cannot find sample ID " + w + " from the VCF file";
      return DMX_E_INVALID;
    }
    col.push_back(c);
  }
  const size_t nv0 = v->samples.size(), nv1 = col.size(), nrec = v->pos.size();
  auto gather = [&](auto& vec, size_t per) {
    if (vec.empty()) return;
    typename std::remove_reference<decltype(vec)>::type out(nrec * nv1 * per);
    for (size_t r = 0; r < nrec; ++r)
      for (size_t k = 0; k < nv1; ++k)
        for (size_t x = 0; x < per; ++x) out[(r * nv1 + k) * per + x] = vec[(r * nv0 + col[k]) * per + x];
    vec.swap(out);
  };
  gather(v->gt, 2);
  gather(v->gp, 3);
  gather(v->pl, 3);
  v->samples = want;
  return DMX_OK;
}

int dmx_vcf_sizes(const dmx_vcf* v, int32_t* nv, int64_t* nrec, int64_t* n_skipped, int32_t* n_chrom) {
  if (!v) return DMX_E_INVALID;
  if (nv) *nv = (int32_t)v->samples.size();
  if (nrec) *nrec = (int64_t)v->pos.size();
  if (n_skipped) *n_skipped = v->n_skipped;
  if (n_chrom) *n_chrom = (int32_t)v->chroms.size();
  return DMX_OK;
}

const char* dmx_vcf_sample(const dmx_vcf* v, int32_t i) {
  return (v && i >= 0 && (size_t)i < v->samples.size()) ? v->samples[(size_t)i].c_str() : nullptr;
}

const char* dmx_vcf_chrom(const dmx_vcf* v, int32_t rid) {
  return (v && rid >= 0 && (size_t)rid < v->chroms.size()) ? v->chroms[(size_t)rid].c_str() : nullptr;
}

int dmx_vcf_copy(const dmx_vcf* v, int8_t* gt, void* field_values, int32_t* rid, int32_t* pos, char* ref, char* alt) {
  if (!v) return DMX_E_INVALID;
  const size_t n = v->pos.size();
  if (gt && !v->gt.empty()) memcpy(gt, v->gt.data(), v->gt.size());
  if (field_values) {
    if (v->field == DMX_FIELD_GT && !v->gt.empty()) memcpy(field_values, v->gt.data(), v->gt.size());
    if (v->field == DMX_FIELD_GP && !v->gp.empty()) memcpy(field_values, v->gp.data(), v->gp.size() * sizeof(float));
    if (v->field == DMX_FIELD_PL && !v->pl.empty()) memcpy(field_values, v->pl.data(), v->pl.size() * sizeof(int32_t));
  }
  if (n) {
    if (rid) memcpy(rid, v->rid.data(), n * sizeof(int32_t));
    if (pos) memcpy(pos, v->pos.data(), n * sizeof(int32_t));
    if (ref) memcpy(ref, v->ref.data(), n);
    if (alt) memcpy(alt, v->alt.data(), n);
  }
  return DMX_OK;
}

int dmx_vcf_no_field(const dmx_vcf* v, uint8_t* flags) {
  if (!v) return DMX_E_INVALID;
  if (flags && !v->no_field.empty()) memcpy(flags, v->no_field.data(), v->no_field.size());
  return DMX_OK;
}

void dmx_vcf_free(dmx_vcf* v) { delete v; }

const char* dmx_vcf_last_error(void) { return g_vcf_error.c_str(); }

}  // extern "C"

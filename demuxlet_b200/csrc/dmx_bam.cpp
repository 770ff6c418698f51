// dmx_bam.cpp -- htslib-free BAM reader and the reference's BAM sweep (BAM half; host only).
//
// The reference walks a position-sorted BAM next to the position-sorted VCF (cmd_cram_demuxlet.cpp:186-338): every read
// that passes the alignment filter (SAMFilteredReader::passed_filter, sam_filtered_reader.cpp:284-296: MAPQ >= 20 and
// none of the flags 0x0f04, cmd_cram_demuxlet.cpp:18-19) and lies on a chromosome of the VCF gets its cell barcode
// (tag CB, :239-268; --group-list keeps only listed barcodes) and UMI (tag UB, :272-293), counts towards RD.TOTL
// (:295), and is tested against every buffered SNP: bam_get_base_and_qual_and_read_and_qual (hts_utils.cpp:279-359)
// walks the CIGAR to the read offset of the SNP position -- only inside an M operation; D / N give no base; S / I advance
// the read offset; other operations are ignored, as there --, then base 'N' is dropped, the quality is thresholded
// (--min-BQ 13) and capped (--cap-BQ 40), --min-TD trims read ends, allele = 0 ref / 1 alt / 2 other (:304-325), and
// scl.add_read receives (snp, cell, umi, allele, bq).  Here the same sweep emits those tuples, in the same order, as the
// flat arrays dmx_pack_reads takes (the packer reproduces add_read's first-seen semantics), the barcodes in first-seen
// order (add_cell, sc_drop_seq.cpp:21-32) and cell_totl_reads.
//
// Which SNPs a read is tested against is the reference's SNP buffer, restated exactly (:197-232, :306-308,
// bcf_filtered_reader.cpp:649-669): before a read is tested the buffer drops its leading SNPs that lie on an earlier
// chromosome or more than one base before the read start, then reads ahead until the newest SNP is at or behind the
// read end (or on a later chromosome); the read is tested against everything buffered BY POSITION ALONE.  For a read
// whose CIGAR accounts for all its bases with M / S / I that is the SNPs under its aligned span -- plus, rarely, a
// SNP of another chromosome that sits in the buffer at the same coordinate (the look-ahead SNP of the next chromosome,
// or the tail of the previous one right after a chromosome change); for a read with '=' / 'X' operations (which the
// reference's walk ignores) every buffered SNP outside its M / D / N segments yields the base at the offset where the
// walk ended.  Both are reproduced: the buffer state [s, hi] is a pair of running maxima over the reads in file order.
//
// Execution: the file is mapped, BGZF blocks (gzip members of <= 64 KB with the 'BC' extra field, SAM spec 4.1) are
// inflated on all host threads one batch ahead, and the records of a batch are walked on all host threads in contiguous
// chunks.  A chunk is walked under the assumption that only the SNPs under the aligned span can give a base; a cheap
// serial pass then runs the buffer recurrence over the chunk's reads and, where a read could see more than that (the
// cases above), the chunk is walked again with the exact buffer state.  Barcode ids are assigned per chunk and merged in
// file order, so that ids, tuples and counts do not depend on the number of threads.  Any other gzip file (or a plain
// one) is read through zlib's gzread and walked the same way.
// SAM text (with its @SQ header; plain, gzip or bgzip) is read too: every line becomes the BAM record sam_parse1 would
// build of it, on the walking threads.  Not handled: CRAM.  --tag-UMI "" reproduces the reference's unseeded rand() UMIs (one walking
// thread).  The BAM must be coordinate-sorted, as for the reference.  Parity: the checker is the python restatement in
// tests/bamio.py (a literal transcription of the buffer loop) on generated BAM files; the CIGAR look-up is pinned on
// the reference's own function (oracle/ref_cigar) -- BGZF / BAM decoding itself is "parity unpinned" against htslib.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <future>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/dmx.h"
#include "dmx_bgzf.h"

struct dmx_sweep {
  std::vector<std::string> barcodes;     // cell id -> barcode (first-seen order)
  std::vector<int32_t> totl;             // cell_totl_reads
  std::vector<int32_t> cell, snp;        // one element per add_read call
  std::vector<uint8_t> allele, bq;
  std::vector<char> umi_buf;
  std::vector<int64_t> umi_off;
  int64_t n_records = 0, n_filtered = 0, n_off_vcf = 0, n_skip_bcd = 0;
};

namespace {

using dmx_io::ByteStream;
using dmx_io::Pool;

thread_local std::string g_bam_error;

int env_threads() { return dmx_io::host_threads("DMX_BAM_THREADS"); }

int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

// bam_aux_get for a 'Z' tag: walks the typed tag list of a record (SAM spec 4.2.4); returns the NUL-terminated string or
// null when the tag is absent or has another type (:245-247 then keeps ".")
const char* aux_get_z(const uint8_t* p, const uint8_t* e, const char* tag) {
  while (p + 3 <= e) {
    const bool hit = p[0] == (uint8_t)tag[0] && p[1] == (uint8_t)tag[1];
    const char t = (char)p[2];
    p += 3;
    size_t sz = 0;
    switch (t) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        const uint8_t* z = (const uint8_t*)memchr(p, 0, (size_t)(e - p));
        if (!z) return nullptr;
        if (hit) return t == 'Z' ? (const char*)p : nullptr;
        p = z + 1;
        continue;
      }
      case 'B': {
        if (p + 5 > e) return nullptr;
        const char st = (char)p[0];
        const uint32_t cnt = rd_u32(p + 1);
        const size_t es = (st == 'c' || st == 'C') ? 1 : ((st == 's' || st == 'S') ? 2 : 4);
        sz = 5 + es * cnt;
        break;
      }
      default: return nullptr;   // corrupt tag list
    }
    if (hit) return nullptr;
    p += sz;
  }
  return nullptr;
}

// The SNP list and the options one sweep works with.
struct Sweep {
  const dmx_sweep_opts* o;
  int64_t nsnps;
  const int32_t *snp_rid, *snp_pos;
  const char *snp_ref, *snp_alt;
  int32_t n_ref;
  std::vector<int32_t> tid2rid;
  std::vector<int64_t> rid_end;                         // rid -> index behind the last SNP with that or a smaller rid
  std::unordered_set<std::string_view> groups;
  bool use_group, random_umi;
  // first SNP index i >= from with (rid, pos)[i] >= (c, p); the SNPs are sorted, reads mostly too: gallop, then bisect
  int64_t lower(int32_t c, int64_t p, int64_t from) const {
    auto less = [&](int64_t i) { return snp_rid[i] < c || (snp_rid[i] == c && (int64_t)snp_pos[i] < p); };
    if (from < 0 || from > nsnps || (from > 0 && !less(from - 1))) from = 0;
    int64_t lo = from, step = 1;
    while (lo + step <= nsnps && less(lo + step - 1)) { lo += step; step <<= 1; }
    int64_t hi = std::min(nsnps, lo + step);           // answer in [lo, hi]
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (less(mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

// The reference's SNP buffer = SNPs [s, hi] (empty when s == hi + 1); cmd_cram_demuxlet.cpp:150,:202-237
struct BufState { int64_t s, hi; };

// What the buffer recurrence needs to know of one read that reached :202
struct ReadKey {
  int32_t c, pos;
  int64_t endpos, lo, la;     // lo: first SNP not cleared by clear_buffer_before; la: where the read-ahead stops
  bool plain;                 // M + S + I cover the read: no base outside [pos, endpos)
};

inline void advance_buffer(const Sweep& S, const ReadKey& k, BufState* b) {
  b->s = std::max(b->s, std::min(k.lo, b->hi + 1));          // bcf_filtered_reader.cpp:660-667
  b->hi = std::max(b->hi, std::min(k.la, S.nsnps - 1));      // :209-237 (vr.read() fails at the end: nothing is added)
}

// True when read k, tested against the buffer b (already advanced), can see a SNP outside the ones under its aligned
// span on its own chromosome: then the fast walk of its chunk is not the reference's.
inline bool needs_exact(const Sweep& S, const ReadKey& k, const BufState& b) {
  if (!k.plain) return true;
  for (int64_t i = b.s; i <= b.hi && i < k.lo && S.snp_rid[i] < k.c; ++i)       // tail of earlier chromosomes
    if (S.snp_pos[i] >= k.pos && S.snp_pos[i] < k.endpos) return true;
  const int64_t x = S.rid_end[(size_t)k.c];                                     // look-ahead SNP of a later chromosome
  return x < S.nsnps && x <= b.hi && S.snp_pos[x] >= k.pos && S.snp_pos[x] < k.endpos;
}

struct ChunkOut {
  std::vector<int32_t> lcell, snp;         // tuples with chunk-local barcode ids
  std::vector<uint8_t> allele, bq;
  std::vector<char> umi;
  std::vector<int64_t> umi_end;
  std::vector<std::string_view> bcs;       // chunk-local barcode id -> string (inside the window)
  std::vector<int32_t> bc_totl;
  std::vector<ReadKey> keys;
  std::unordered_map<std::string_view, int32_t> bc_map;
  int64_t n_filtered = 0, n_off_vcf = 0, n_skip_bcd = 0;
  bool corrupt = false;
  void clear() {
    lcell.clear(); snp.clear(); allele.clear(); bq.clear(); umi.clear(); umi_end.clear();
    bcs.clear(); bc_totl.clear(); keys.clear(); bc_map.clear();
    n_filtered = n_off_vcf = n_skip_bcd = 0;
    corrupt = false;
  }
};

// Walks records [r0, r1) of the window.  exact == nullptr: every read is tested against the SNPs under its aligned
// span (and its ReadKey is recorded); otherwise against the reference's buffer, whose state *exact carries along.
void walk_chunk(const Sweep& S, const uint8_t* win, const size_t* offs, size_t r0, size_t r1, BufState* exact,
                struct random_data* rnd, ChunkOut* out) {
  static const char kBase[] = "=ACMGRSVTWYHKDBN";
  const dmx_sweep_opts* o = S.o;
  out->clear();
  int64_t hint_lo = 0;
  for (size_t r = r0; r < r1; ++r) {
    if (r + 6 < r1) {                       // the records a few steps ahead: fixed fields, CIGAR, start of the tags
      __builtin_prefetch(win + offs[r + 6]);
      __builtin_prefetch(win + offs[r + 6] + 64);
      __builtin_prefetch(win + offs[r + 6] + 128);
    }
    const uint8_t* b = win + offs[r] + 4;
    const int32_t block = rd_i32(b - 4);
    const int32_t tid = rd_i32(b), pos = rd_i32(b + 4);
    const uint32_t l_name = b[8], mapq = b[9], n_cigar = rd_u16(b + 12), flag = rd_u16(b + 14);
    const int32_t l_seq = rd_i32(b + 16);
    const uint8_t* cigar = b + 32 + l_name;
    const uint8_t* seq = cigar + 4 * (size_t)n_cigar;
    const uint8_t* qual = seq + ((size_t)std::max(l_seq, 0) + 1) / 2;
    const uint8_t* aux = qual + (size_t)std::max(l_seq, 0);
    const uint8_t* end = b + block;
    if (l_seq < 0 || aux > end) { out->corrupt = true; return; }
    // sam_filtered_reader.cpp:284-296
    if ((int32_t)mapq < o->min_mq || (o->exclude_flag & flag)) { ++out->n_filtered; continue; }
    const int32_t rid = (tid >= 0 && tid < S.n_ref) ? S.tid2rid[(size_t)tid] : -1;
    if (rid < 0) { ++out->n_off_vcf; continue; }                                      // :197-200
    // bam_endpos (htslib >= 1.10): start + reference span of the CIGAR (M, D, N, =, X), at least one base
    int64_t span = 0, covered = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
      const uint32_t v = rd_u32(cigar + 4 * c), op = v & 15u;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += v >> 4;
      if (op == 0 || op == 1 || op == 4) covered += v >> 4;
    }
    // (htslib: a read flagged unmapped ends one base behind its start whatever its CIGAR says -- it is still walked
    // through that CIGAR below, so for it the aligned span does not bound where a base can be found)
    const bool unmapped_cigar = (flag & 4u) && n_cigar > 0;
    if (unmapped_cigar) span = 0;
    ReadKey k;
    k.c = rid;
    k.pos = pos;
    k.endpos = (int64_t)pos + (span > 0 ? span : 1);
    k.plain = n_cigar == 0 || (covered >= l_seq && !unmapped_cigar);
    k.lo = hint_lo = S.lower(rid, (int64_t)pos - 1, hint_lo);
    k.la = S.lower(rid, k.endpos, k.lo);
    int64_t i0, i1;              // SNPs to test: [i0, i1)
    if (exact) {
      advance_buffer(S, k, exact);                                                   // :202-237
      i0 = exact->s;
      i1 = exact->hi + 1;
    } else {
      out->keys.push_back(k);
      i0 = k.lo;
      i1 = k.la;
    }
    // barcode (:239-268)
    std::string_view sbcd(".");
    if (S.use_group) {
      const char* z = aux_get_z(aux, end, o->tag_group);
      if (z) sbcd = std::string_view(z);
      if (!S.groups.empty() && S.groups.find(sbcd) == S.groups.end()) { ++out->n_skip_bcd; continue; }
    }
    int32_t ibcd;
    {
      auto it = out->bc_map.find(sbcd);
      if (it == out->bc_map.end()) {
        ibcd = (int32_t)out->bcs.size();
        out->bc_map.emplace(sbcd, ibcd);
        out->bcs.push_back(sbcd);
        out->bc_totl.push_back(0);
      } else {
        ibcd = it->second;
      }
    }
    // UMI (:272-293): a missing tag leaves "."; --tag-UMI "" draws "." + hex(rand()) per read (:274-276).  The
    // reference never seeds rand(), so its UMIs are glibc's default sequence (TYPE_3 generator, seed 1): the same
    // numbers come from a private random_r state here, whatever else the process does with rand().
    const char* sumi = ".";
    char rnd_umi[16];
    if (S.random_umi) {
      int32_t rv = 0;
      random_r(rnd, &rv);
      snprintf(rnd_umi, sizeof(rnd_umi), ".%x", (unsigned)rv);
      sumi = rnd_umi;
    } else {
      const char* z = aux_get_z(aux, end, o->tag_umi);
      if (z) sumi = z;
    }
    ++out->bc_totl[(size_t)ibcd];                                                    // :295
    if (!n_cigar) continue;            // no walk: base 'N' (hts_utils.cpp:285-288)
    const size_t umi_len = strlen(sumi);
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t sp = S.snp_pos[i];
      if (k.plain && (sp < pos || sp >= k.endpos)) continue;     // the walk ends at rpos >= l_seq: no base
      // hts_utils.cpp:279-359, by position alone
      const uint32_t spos = (uint32_t)sp;
      uint32_t cpos = (uint32_t)pos;
      int32_t rpos = 0;
      for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t v = rd_u32(cigar + 4 * c), op = v & 15u, len = v >> 4;
        if (op == 0) {                            // M
          if (spos >= cpos && spos <= cpos + len - 1) { rpos += (int32_t)(spos - cpos); break; }
          cpos += len;
          rpos += (int32_t)len;
        } else if (op == 2 || op == 3) {          // D, N
          if (spos >= cpos && spos <= cpos + len - 1) { rpos = -1; break; }
          cpos += len;
        } else if (op == 4 || op == 1) {          // S, I
          rpos += (int32_t)len;
        }
      }
      if (rpos < 0 || rpos >= l_seq) continue;    // BAM_READ_INDEX_NA
      const char base = kBase[(seq[rpos >> 1] >> ((~rpos & 1) << 2)) & 15];
      if (base == 'N') continue;                                                      // :314
      // the reference carries the quality as a (signed) char qual + 33 and subtracts 33 again (:318, :323): values
      // >= 95, and the 0xff of a record without qualities, wrap to negative numbers and are dropped by --min-BQ
      const int32_t q = (int32_t)(int8_t)(uint8_t)(qual[rpos] + 33) - 33;
      if (q < o->min_bq) continue;                                                   // :318
      if (rpos < o->min_td - 1) continue;                                            // :319
      if (rpos + o->min_td > l_seq) continue;                                        // :320
      out->lcell.push_back(ibcd);
      out->snp.push_back((int32_t)i);
      out->allele.push_back((uint8_t)(base == S.snp_ref[i] ? 0 : (base == S.snp_alt[i] ? 1 : 2)));   // :322
      out->bq.push_back((uint8_t)(q > o->cap_bq ? o->cap_bq : q));                   // :323
      out->umi.insert(out->umi.end(), sumi, sumi + umi_len);
      out->umi_end.push_back((int64_t)out->umi.size());
    }
  }
}


// One SAM text line as the BAM record htslib's sam_parse1 would build of it (SAM spec 1.4 -> 4.2), appended to *out with
// its block_size in front: the walker then reads text and binary input alike.  Mate fields are not needed (-1 / 0).
bool sam_to_bam(const char* p, const char* e, const std::unordered_map<std::string_view, int32_t>& name2tid,
                std::vector<uint8_t>* out) {
  static const char kOps[] = "MIDNSHP=X";
  const char* f[12];
  size_t n[12];
  int nf = 0;
  const char* s = p;
  for (const char* q = p; q <= e && nf < 11; ++q)
    if (q == e || *q == '\t') { f[nf] = s; n[nf] = (size_t)(q - s); ++nf; s = q + 1; }
  if (nf < 11) return false;
  const char* tags = s <= e ? s : e;            // behind the 11th field
  auto num = [](const char* a, size_t len) {
    long v = 0;
    size_t i = (len && a[0] == '-') ? 1 : 0;
    for (; i < len && a[i] >= '0' && a[i] <= '9'; ++i) v = v * 10 + (a[i] - '0');
    return (len && a[0] == '-') ? -v : v;
  };
  if (n[0] == 0 || n[0] > 254) return false;
  int32_t tid = -1;
  if (!(n[2] == 1 && f[2][0] == '*')) {
    auto it = name2tid.find(std::string_view(f[2], n[2]));
    if (it != name2tid.end()) tid = it->second;
  }
  const size_t at = out->size();
  out->resize(at + 36 + n[0] + 1);
  uint8_t* b = out->data() + at + 4;
  const int32_t pos = (int32_t)num(f[3], n[3]) - 1, l_seq = (n[9] == 1 && f[9][0] == '*') ? 0 : (int32_t)n[9];
  const int32_t m1 = -1, zero = 0;
  memcpy(b, &tid, 4);
  memcpy(b + 4, &pos, 4);
  b[8] = (uint8_t)(n[0] + 1);
  b[9] = (uint8_t)num(f[4], n[4]);
  b[10] = b[11] = 0;                              // bin: not used
  const uint16_t flag = (uint16_t)num(f[1], n[1]);
  memcpy(b + 14, &flag, 2);
  memcpy(b + 16, &l_seq, 4);
  memcpy(b + 20, &m1, 4);
  memcpy(b + 24, &m1, 4);
  memcpy(b + 28, &zero, 4);
  memcpy(b + 32, f[0], n[0]);
  b[32 + n[0]] = 0;
  uint32_t n_cigar = 0;
  if (!(n[5] == 1 && f[5][0] == '*')) {
    const char *c = f[5], *ce = f[5] + n[5];
    while (c < ce) {
      uint32_t len = 0;
      bool any = false;
      while (c < ce && *c >= '0' && *c <= '9') { len = len * 10 + (uint32_t)(*c - '0'); ++c; any = true; }
      const char* op = c < ce ? strchr(kOps, *c) : nullptr;
      if (!any || !op || !*c) return false;
      const uint32_t v = (len << 4) | (uint32_t)(op - kOps);
      out->insert(out->end(), (const uint8_t*)&v, (const uint8_t*)&v + 4);
      ++n_cigar;
      ++c;
    }
    if (n_cigar > 65535) return false;
  }
  {
    const uint16_t nc = (uint16_t)n_cigar;
    memcpy(out->data() + at + 4 + 12, &nc, 2);
  }
  // sequence (4 bits per base) and qualities
  const size_t sq = out->size();
  out->resize(sq + ((size_t)l_seq + 1) / 2 + (size_t)l_seq);
  uint8_t* q4 = out->data() + sq;
  memset(q4, 0, ((size_t)l_seq + 1) / 2);
  for (int32_t i = 0; i < l_seq; ++i) {
    static const char kBase[] = "=ACMGRSVTWYHKDBN";
    const char ch = f[9][i] >= 'a' && f[9][i] <= 'z' ? (char)(f[9][i] - 32) : f[9][i];
    const char* w = strchr(kBase, ch);
    const uint8_t code = (w && ch) ? (uint8_t)(w - kBase) : 15;
    q4[i >> 1] |= (uint8_t)(code << ((~i & 1) << 2));
  }
  uint8_t* ql = q4 + ((size_t)l_seq + 1) / 2;
  if (n[10] == 1 && f[10][0] == '*') memset(ql, 0xff, (size_t)l_seq);
  else {
    if ((int32_t)n[10] != l_seq) return false;
    for (int32_t i = 0; i < l_seq; ++i) ql[i] = (uint8_t)(f[10][i] - 33);
  }
  // tags TAG:TYPE:VALUE
  for (const char* t = tags; t < e;) {
    const char* te = (const char*)memchr(t, '\t', (size_t)(e - t));
    if (!te) te = e;
    if (te - t >= 5 && t[2] == ':' && t[4] == ':') {
      const char ty = t[3];
      const char *val = t + 5;
      const size_t vl = (size_t)(te - val);
      out->push_back((uint8_t)t[0]);
      out->push_back((uint8_t)t[1]);
      if (ty == 'A') { out->push_back('A'); out->push_back(vl ? (uint8_t)val[0] : 0); }
      else if (ty == 'i') { const int32_t v = (int32_t)num(val, vl); out->push_back('i'); out->insert(out->end(), (const uint8_t*)&v, (const uint8_t*)&v + 4); }
      else if (ty == 'f') { const float v = strtof(std::string(val, vl).c_str(), nullptr); out->push_back('f'); out->insert(out->end(), (const uint8_t*)&v, (const uint8_t*)&v + 4); }
      else if (ty == 'Z' || ty == 'H') { out->push_back((uint8_t)ty); out->insert(out->end(), val, val + vl); out->push_back(0); }
      else if (ty == 'B' && vl >= 1) {
        const char st = val[0];
        const size_t es = (st == 'c' || st == 'C') ? 1 : ((st == 's' || st == 'S') ? 2 : 4);
        out->push_back('B');
        out->push_back((uint8_t)st);
        const size_t cnt_at = out->size();
        out->resize(cnt_at + 4);
        uint32_t cnt = 0;
        for (const char* c = val + 1; c < te;) {
          if (*c == ',') ++c;
          const char* ce = (const char*)memchr(c, ',', (size_t)(te - c));
          if (!ce) ce = te;
          if (st == 'f') { const float v = strtof(std::string(c, (size_t)(ce - c)).c_str(), nullptr); out->insert(out->end(), (const uint8_t*)&v, (const uint8_t*)&v + 4); }
          else { const int32_t v = (int32_t)num(c, (size_t)(ce - c)); out->insert(out->end(), (const uint8_t*)&v, (const uint8_t*)&v + es); }
          ++cnt;
          c = ce;
        }
        memcpy(out->data() + cnt_at, &cnt, 4);
      } else return false;
    } else if (te > t) return false;
    t = te + 1;
  }
  const int32_t block = (int32_t)(out->size() - at - 4);
  memcpy(out->data() + at, &block, 4);
  return true;
}

}  // namespace

extern "C" {

void dmx_sweep_opts_default(dmx_sweep_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->tag_group = "CB";        // cmd_cram_demuxlet.cpp:12
  o->tag_umi = "UB";          // :13
  o->cap_bq = 40;             // :14
  o->min_bq = 13;             // :15
  o->min_td = 0;              // :16
  o->exclude_flag = 0x0f04;   // :18
  o->min_mq = 20;             // :19
}

int dmx_sweep_bam(const char* path, const dmx_sweep_opts* o, int32_t n_chrom, const char* const* chroms, int64_t nsnps,
                  const int32_t* snp_rid, const int32_t* snp_pos, const char* snp_ref, const char* snp_alt,
                  dmx_sweep** out) {
  if (!path || !o || !out || n_chrom < 0 || nsnps < 0 || (nsnps > 0 && (!snp_rid || !snp_pos || !snp_ref || !snp_alt)) ||
      (n_chrom > 0 && !chroms) || o->n_groups < 0 || (o->n_groups > 0 && !o->groups))
    return DMX_E_INVALID;
  *out = nullptr;
  if (!o->tag_umi) {
    g_bam_error = "tag_umi is null";
    return DMX_E_UNSUPPORTED;
  }
  for (int64_t i = 0; i < nsnps; ++i) {
    if (snp_rid[i] < 0 || snp_rid[i] >= n_chrom) {
      g_bam_error = "SNP on a chromosome that is not in the list";
      return DMX_E_INVALID;
    }
    if (i > 0 && (snp_rid[i] < snp_rid[i - 1] || (snp_rid[i] == snp_rid[i - 1] && snp_pos[i] < snp_pos[i - 1]))) {
      g_bam_error = "SNPs are not sorted by chromosome and position";
      return DMX_E_INVALID;
    }
  }
  Sweep S;
  S.o = o;
  S.nsnps = nsnps;
  S.snp_rid = snp_rid;
  S.snp_pos = snp_pos;
  S.snp_ref = snp_ref;
  S.snp_alt = snp_alt;
  S.use_group = o->tag_group && o->tag_group[0];
  S.random_umi = !o->tag_umi[0];
  S.rid_end.assign((size_t)n_chrom, 0);
  {
    int64_t i = 0;
    for (int32_t c = 0; c < n_chrom; ++c) {
      while (i < nsnps && snp_rid[i] <= c) ++i;
      S.rid_end[(size_t)c] = i;
    }
  }
  for (int64_t i = 0; i < o->n_groups; ++i) S.groups.insert(std::string_view(o->groups[i]));
  const int nth = S.random_umi ? 1 : env_threads();    // the rand() sequence is drawn in file order
  size_t min_chunk = 2048;                              // records per walking thread, at least
  if (const char* env = getenv("DMX_BAM_MIN_CHUNK")) min_chunk = (size_t)std::max(1, atoi(env));

  ByteStream gz;
  if (!gz.open(path, env_threads())) {
    g_bam_error = std::string("cannot open ") + path;
    return DMX_E_IO;
  }
  size_t cur = 0;             // read position inside the window
  auto need = [&](size_t n) {
    while (gz.size() - cur < n) {
      const bool got = gz.more(cur);
      cur = 0;
      if (!got) return gz.size() >= n;
    }
    return true;
  };
  // BAM, or SAM text with its header (recognised by the leading '@')
  const bool is_sam = need(1) && gz.data()[cur] == '@';
  if (!is_sam && (!need(8) || memcmp(gz.data() + cur, "BAM\1", 4) != 0)) {
    g_bam_error = std::string(path) + " is not a BAM file";
    return DMX_E_INVALID;
  }
  std::vector<std::string> ref_names;
  if (is_sam) {
    // header lines: @SQ SN:name in order are the reference ids (sam_hdr_parse)
    for (;;) {
      while (!memchr(gz.data() + cur, '\n', gz.size() - cur)) {
        const bool got = gz.more(cur);
        cur = 0;
        if (!got) break;
      }
      if (gz.size() == cur || gz.data()[cur] != '@') break;
      const char* lp = (const char*)gz.data() + cur;
      const char* nl = (const char*)memchr(lp, '\n', gz.size() - cur);
      const char* le = nl ? nl : (const char*)gz.data() + gz.size();
      if (le - lp > 4 && memcmp(lp, "@SQ\t", 4) == 0) {
        for (const char* q = lp + 4; q < le;) {
          const char* qe = (const char*)memchr(q, '\t', (size_t)(le - q));
          if (!qe) qe = le;
          if (qe - q > 3 && memcmp(q, "SN:", 3) == 0) {
            const char* ve = qe;
            while (ve > q + 3 && (ve[-1] == '\r')) --ve;
            ref_names.emplace_back(q + 3, (size_t)(ve - q - 3));
          }
          q = qe + 1;
        }
      }
      cur = (size_t)(le - (const char*)gz.data()) + (nl ? 1 : 0);
    }
    S.n_ref = (int32_t)ref_names.size();
  } else {
    const size_t l_text = (size_t)std::max(0, rd_i32(gz.data() + cur + 4));
    cur += 8;
    if (!need(l_text + 4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
    cur += l_text;
    S.n_ref = rd_i32(gz.data() + cur);
    cur += 4;
    for (int32_t t = 0; t < S.n_ref; ++t) {
      if (!need(4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
      const int32_t l = rd_i32(gz.data() + cur);
      cur += 4;
      if (l < 1 || !need((size_t)l + 4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
      ref_names.emplace_back((const char*)gz.data() + cur, strnlen((const char*)gz.data() + cur, (size_t)l));
      cur += (size_t)l + 4;
    }
  }
  // tid -> VCF rid by name (bcf_hdr_name2id, :197); the reference insists on the same chromosome order (:160-176)
  std::unordered_map<std::string, int32_t> name2rid;
  for (int32_t r = 0; r < n_chrom; ++r) name2rid[chroms[r]] = r;
  S.tid2rid.assign((size_t)std::max(0, S.n_ref), -1);
  std::unordered_map<std::string_view, int32_t> name2tid;       // SAM text: RNAME -> tid
  int32_t prev = -1;
  for (int32_t t = 0; t < S.n_ref; ++t) {
    name2tid.emplace(std::string_view(ref_names[(size_t)t]), t);
    auto it = name2rid.find(ref_names[(size_t)t]);
    if (it != name2rid.end()) {
      if (prev >= it->second) {
        g_bam_error = "VCF and BAM have different orderings of chromosomes";   // :166-169
        return DMX_E_INVALID;
      }
      S.tid2rid[(size_t)t] = prev = it->second;
    }
  }
  struct random_data rnd_buf;                 // glibc's rand() of an unseeded process: TYPE_3 state, seed 1
  char rnd_state[128];
  memset(&rnd_buf, 0, sizeof(rnd_buf));
  initstate_r(1u, rnd_state, sizeof(rnd_state), &rnd_buf);

  dmx_sweep* sw = new dmx_sweep();
  sw->umi_off.push_back(0);
  std::deque<std::string> bc_store;                        // stable storage for the keys of bc_map
  std::unordered_map<std::string_view, int32_t> bc_map;
  BufState buf{0, nsnps > 0 ? 0 : -1};        // the first SNP is read before the loop (:150)
  std::vector<size_t> offs;
  Pool pool(nth);
  std::vector<ChunkOut> chunks((size_t)nth);
  std::vector<std::vector<int32_t>> l2g((size_t)nth);
  auto fail = [&](const char* msg) {
    g_bam_error = msg;
    delete sw;
    return DMX_E_INVALID;
  };
  // One batch of records: walked in contiguous chunks on the host threads (walk(lo, hi, exact, out) walks records
  // [lo, hi) of the batch), merged in file order, scattered into the output arrays.
  using WalkFn = std::function<void(int, size_t, size_t, BufState*, ChunkOut*)>;
  auto run_batch = [&](size_t nrec, const WalkFn& walk) -> bool {
    sw->n_records += (int64_t)nrec;
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)nth, nrec / min_chunk));
    auto lo_of = [&](int t) { return nrec * (size_t)t / (size_t)nt; };
    if (S.random_umi) {
      walk(0, 0, nrec, &buf, &chunks[0]);
    } else {
      pool.run(nt, [&](int t) { walk(t, lo_of(t), lo_of(t + 1), nullptr, &chunks[(size_t)t]); });
    }
    // in file order: buffer recurrence (exact second walk where a read sees more than its own span), barcode ids
    size_t total = sw->cell.size(), umi_total = sw->umi_buf.size();
    std::vector<size_t> at((size_t)nt + 1), umi_at((size_t)nt + 1);
    for (int t = 0; t < nt; ++t) {
      ChunkOut& c = chunks[(size_t)t];
      if (c.corrupt) return false;
      if (!S.random_umi) {
        BufState b = buf;
        bool redo = false;
        for (const ReadKey& k : c.keys) {
          advance_buffer(S, k, &b);
          if (needs_exact(S, k, b)) { redo = true; break; }
        }
        if (redo) {
          walk(t, lo_of(t), lo_of(t + 1), &buf, &c);
        } else {
          buf = b;
        }
      }
      std::vector<int32_t>& map = l2g[(size_t)t];
      map.resize(c.bcs.size());
      for (size_t i = 0; i < c.bcs.size(); ++i) {
        auto it = bc_map.find(c.bcs[i]);
        if (it == bc_map.end()) {
          bc_store.emplace_back(c.bcs[i]);
          it = bc_map.emplace(std::string_view(bc_store.back()), (int32_t)bc_store.size() - 1).first;
          sw->totl.push_back(0);
        }
        map[i] = it->second;
        sw->totl[(size_t)it->second] += c.bc_totl[i];
      }
      sw->n_filtered += c.n_filtered;
      sw->n_off_vcf += c.n_off_vcf;
      sw->n_skip_bcd += c.n_skip_bcd;
      at[(size_t)t] = total;
      umi_at[(size_t)t] = umi_total;
      total += c.lcell.size();
      umi_total += c.umi.size();
    }
    at[(size_t)nt] = total;
    if (total > sw->cell.capacity()) {          // grow once to the size the decoded share of the file suggests
      const double frac = gz.decoded_share();
      const size_t want = frac > 0 ? (size_t)((double)total / frac * 1.05) + 4096 : 0;
      const size_t umi_want = frac > 0 ? (size_t)((double)umi_total / frac * 1.05) + 65536 : 0;
      sw->cell.reserve(std::max(want, total));
      sw->snp.reserve(std::max(want, total));
      sw->allele.reserve(std::max(want, total));
      sw->bq.reserve(std::max(want, total));
      sw->umi_off.reserve(std::max(want, total) + 1);
      sw->umi_buf.reserve(std::max(umi_want, umi_total));
    }
    sw->cell.resize(total);
    sw->snp.resize(total);
    sw->allele.resize(total);
    sw->bq.resize(total);
    sw->umi_buf.resize(umi_total);
    sw->umi_off.resize(total + 1);
    pool.run(nt, [&](int t) {
      const ChunkOut& c = chunks[(size_t)t];
      const size_t a = at[(size_t)t], n = c.lcell.size();
      const std::vector<int32_t>& map = l2g[(size_t)t];
      for (size_t i = 0; i < n; ++i) sw->cell[a + i] = map[(size_t)c.lcell[i]];
      if (n) {
        memcpy(sw->snp.data() + a, c.snp.data(), n * 4);
        memcpy(sw->allele.data() + a, c.allele.data(), n);
        memcpy(sw->bq.data() + a, c.bq.data(), n);
        memcpy(sw->umi_buf.data() + umi_at[(size_t)t], c.umi.data(), c.umi.size());
      }
      for (size_t i = 0; i < n; ++i) sw->umi_off[a + i + 1] = (int64_t)umi_at[(size_t)t] + c.umi_end[i];
    });
    return true;
  };
  if (is_sam) {
    // text: the complete lines of the window; every walking thread first turns its lines into BAM records
    struct TLine { const char *p, *e; };
    std::vector<TLine> tl;
    std::vector<std::vector<uint8_t>> rec_buf((size_t)nth);
    std::vector<std::vector<size_t>> rec_off((size_t)nth);
    for (bool last = false; !last;) {
      tl.clear();
      const char* w = (const char*)gz.data();
      const size_t wsize = gz.size();
      size_t p = cur;
      while (p < wsize) {
        const char* nl = (const char*)memchr(w + p, '\n', wsize - p);
        if (!nl) break;
        const char* le = nl;
        if (le > w + p && le[-1] == '\r') --le;
        if (le > w + p) tl.push_back(TLine{w + p, le});
        p = (size_t)(nl - w) + 1;
      }
      const WalkFn walk = [&](int t, size_t lo, size_t hi, BufState* exact, ChunkOut* co) {
        std::vector<uint8_t>& rb = rec_buf[(size_t)t];
        std::vector<size_t>& ro = rec_off[(size_t)t];
        if (!exact || S.random_umi) {           // a second, exact walk finds the records already built
          rb.clear();
          ro.clear();
          for (size_t i = lo; i < hi; ++i) {
            ro.push_back(rb.size());
            if (!sam_to_bam(tl[i].p, tl[i].e, name2tid, &rb)) { co->clear(); co->corrupt = true; return; }
          }
        }
        walk_chunk(S, rb.data(), ro.data(), 0, ro.size(), exact, &rnd_buf, co);
      };
      if (!run_batch(tl.size(), walk)) return fail("malformed SAM record");
      const bool got = gz.more(p);
      cur = 0;
      if (!got) {
        if (gz.bad()) return fail("truncated SAM file");
        last = true;
        if (gz.size() > 0) {                     // a last line without a newline
          tl.clear();
          const char* lp = (const char*)gz.data();
          tl.push_back(TLine{lp, lp + gz.size()});
          if (!run_batch(1, walk)) return fail("malformed SAM record");
        }
      }
    }
  } else {
    for (;;) {
      // complete records of the window
      offs.clear();
      const uint8_t* win = gz.data();
      const size_t wsize = gz.size();
      size_t p = cur;
      while (p + 4 <= wsize) {
        const int32_t block = rd_i32(win + p);
        // the hop is a chain of dependent loads over memory another thread wrote: fetching a few records ahead takes
        // it from ~65 to ~25 ns per record where the hardware prefetcher does not follow (virtual machines)
        __builtin_prefetch(win + p + 2048);
        if (block < 32) return fail("corrupt BAM record");
        if (p + 4 + (size_t)block > wsize) break;
        offs.push_back(p);
        p += 4 + (size_t)block;
      }
      const WalkFn walk = [&](int, size_t lo, size_t hi, BufState* exact, ChunkOut* co) {
        walk_chunk(S, win, offs.data(), lo, hi, exact, &rnd_buf, co);
      };
      if (!run_batch(offs.size(), walk)) return fail("corrupt BAM record");
      // next window
      const bool got = gz.more(p);
      cur = 0;
      if (!got) {
        if (gz.bad() || gz.size() > 0) return fail("truncated BAM record");
        break;
      }
    }
  }
  sw->barcodes.assign(bc_store.begin(), bc_store.end());
  *out = sw;
  return DMX_OK;
}

int dmx_sweep_sizes(const dmx_sweep* s, int64_t* ncells, int64_t* nreads, int64_t* umi_bytes) {
  if (!s) return DMX_E_INVALID;
  if (ncells) *ncells = (int64_t)s->barcodes.size();
  if (nreads) *nreads = (int64_t)s->cell.size();
  if (umi_bytes) *umi_bytes = (int64_t)s->umi_buf.size();
  return DMX_OK;
}

int dmx_sweep_counts(const dmx_sweep* s, int64_t* n_records, int64_t* n_filtered, int64_t* n_off_vcf, int64_t* n_skip_bcd) {
  if (!s) return DMX_E_INVALID;
  if (n_records) *n_records = s->n_records;
  if (n_filtered) *n_filtered = s->n_filtered;
  if (n_off_vcf) *n_off_vcf = s->n_off_vcf;
  if (n_skip_bcd) *n_skip_bcd = s->n_skip_bcd;
  return DMX_OK;
}

int dmx_sweep_copy(const dmx_sweep* s, int32_t* cell, int32_t* snp, uint8_t* allele, uint8_t* bq, char* umi_buf,
                   int64_t* umi_off, int32_t* totl) {
  if (!s) return DMX_E_INVALID;
  const size_t n = s->cell.size();
  if (n) {
    if (cell) memcpy(cell, s->cell.data(), n * 4);
    if (snp) memcpy(snp, s->snp.data(), n * 4);
    if (allele) memcpy(allele, s->allele.data(), n);
    if (bq) memcpy(bq, s->bq.data(), n);
  }
  if (umi_buf && !s->umi_buf.empty()) memcpy(umi_buf, s->umi_buf.data(), s->umi_buf.size());
  if (umi_off) memcpy(umi_off, s->umi_off.data(), s->umi_off.size() * 8);
  if (totl && !s->totl.empty()) memcpy(totl, s->totl.data(), s->totl.size() * 4);
  return DMX_OK;
}

const char* dmx_sweep_barcode(const dmx_sweep* s, int64_t i) {
  return (s && i >= 0 && (size_t)i < s->barcodes.size()) ? s->barcodes[(size_t)i].c_str() : nullptr;
}

void dmx_sweep_free(dmx_sweep* s) { delete s; }

const char* dmx_sweep_last_error(void) { return g_bam_error.c_str(); }

}  // extern "C"

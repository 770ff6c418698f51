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
// The SNP buffer of the reference (clear_buffer_before / read-ahead, :197-232) only bounds which SNPs are *tested*; a SNP
// outside the aligned span yields no base, so testing exactly the SNPs with read start <= pos < read end on the read's
// chromosome, in VCF order, gives the same tuples.  BAM decoding (BGZF = a series of gzip members, which zlib's gzread
// walks through; record layout of the SAM specification section 4.2) replaces htslib's sam_read1 / bam_aux_get.
// Not handled: CRAM, SAM text.  --tag-UMI "" reproduces the reference's unseeded rand() UMIs.  Parity: the checker is
// the python restatement in tests/bamio.py on generated BAM files -- "parity unpinned" against htslib itself.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/dmx.h"

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

thread_local std::string g_bam_error;

// Byte stream over a BAM file.  A BGZF file (gzip members of <= 64 KB with the 'BC' extra field that gives each
// block's compressed size, SAM spec 4.1) is inflated block-parallel: the compressed file is indexed by walking the block
// headers, batches of blocks are inflated on the host threads into one buffer at their prefix-sum offsets (one batch
// ahead of the record walk), and the records are then read in file order -- htslib's bgzf_mt idea.  Any other gzip file
// (or a plain one) goes through zlib's gzread.
class ByteStream {
 public:
  ~ByteStream() {
    if (next_.valid()) next_.wait();
    if (gz_) gzclose(gz_);
  }
  bool open(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    uint8_t h[18];
    const size_t got = fread(h, 1, 18, f);
    const bool bgzf = got == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 &&
                      h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
    if (!bgzf) {
      fclose(f);
      gz_ = gzopen(path, "rb");
      if (gz_) gzbuffer(gz_, 1 << 20);
      return gz_ != nullptr;
    }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    comp_.resize((size_t)std::max(0L, sz));
    const bool ok = fread(comp_.data(), 1, comp_.size(), f) == comp_.size();
    fclose(f);
    if (!ok) return false;
    // index: every block starts with the 18-byte header checked above; BSIZE = total block size - 1
    for (size_t off = 0; off + 28 <= comp_.size();) {
      const uint8_t* b = comp_.data() + off;
      if (!(b[0] == 0x1f && b[1] == 0x8b && b[12] == 'B' && b[13] == 'C')) { bad_ = true; break; }
      const size_t bsize = (size_t)(b[16] | (b[17] << 8)) + 1;
      if (bsize < 26 || off + bsize > comp_.size()) { bad_ = true; break; }
      uint32_t isize;
      memcpy(&isize, b + bsize - 4, 4);
      blocks_.push_back(Blk{off, bsize, isize});
      off += bsize;
    }
    nth_ = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* env = getenv("DMX_BAM_THREADS")) nth_ = std::max(1, std::min(atoi(env), 256));
    parallel_ = true;
    schedule();
    return true;
  }
  int threads() const { return parallel_ ? nth_ : 1; }
  // n bytes or fail; read_some-style EOF probe through `at_eof`
  bool read(void* dst, size_t n) {
    if (!parallel_) {
      char* p = (char*)dst;
      while (n) {
        const int k = gzread(gz_, p, (unsigned)std::min<size_t>(n, 1u << 30));
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
      }
      return true;
    }
    char* p = (char*)dst;
    while (n) {
      if (pos_ == cur_.size() && !advance()) return false;
      const size_t k = std::min(n, cur_.size() - pos_);
      memcpy(p, cur_.data() + pos_, k);
      pos_ += k;
      p += k;
      n -= k;
    }
    return true;
  }
  // 1: four bytes read, 0: clean end of file, -1: truncated
  int read4(uint8_t* dst) {
    if (!parallel_) {
      const int k = gzread(gz_, dst, 4);
      return k == 4 ? 1 : (k == 0 ? 0 : -1);
    }
    if (pos_ == cur_.size() && !advance()) return bad_ ? -1 : 0;
    if (cur_.size() - pos_ >= 4) {
      memcpy(dst, cur_.data() + pos_, 4);
      pos_ += 4;
      return 1;
    }
    return read(dst, 4) ? 1 : -1;
  }

 private:
  struct Blk { size_t off, size; uint32_t isize; };
  static constexpr size_t kBatch = 1024;    // blocks per batch (<= 64 MB of BAM)
  bool inflate_batch(size_t b0, size_t b1, std::vector<uint8_t>* out) {
    std::vector<size_t> at(b1 - b0 + 1, 0);
    for (size_t i = b0; i < b1; ++i) at[i - b0 + 1] = at[i - b0] + blocks_[i].isize;
    out->resize(at.back());
    std::atomic<size_t> nextb{b0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
      z_stream zs;
      for (;;) {
        const size_t i = nextb.fetch_add(1);
        if (i >= b1) break;
        const Blk& bl = blocks_[i];
        if (bl.isize == 0) continue;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { ok = false; break; }
        zs.next_in = comp_.data() + bl.off + 18;
        zs.avail_in = (uInt)(bl.size - 26);
        zs.next_out = out->data() + at[i - b0];
        zs.avail_out = bl.isize;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        uint32_t crc;
        memcpy(&crc, comp_.data() + bl.off + bl.size - 8, 4);
        if (rc != Z_STREAM_END || zs.avail_out != 0 ||
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), out->data() + at[i - b0], bl.isize) != crc)
          ok = false;
      }
    };
    const int nt = (int)std::min<size_t>((size_t)nth_, b1 - b0);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
    return ok;
  }
  void schedule() {      // start inflating the next batch in the background
    if (next_blk_ >= blocks_.size()) return;
    const size_t b0 = next_blk_, b1 = std::min(blocks_.size(), b0 + kBatch);
    next_blk_ = b1;
    next_ = std::async(std::launch::async, [this, b0, b1]() { return inflate_batch(b0, b1, &ahead_); });
  }
  bool advance() {       // the batch that was inflated ahead becomes the current one
    if (!next_.valid()) return false;
    if (!next_.get()) { bad_ = true; return false; }
    cur_.swap(ahead_);
    pos_ = 0;
    schedule();
    return !cur_.empty() || advance();
  }
  gzFile gz_ = nullptr;
  bool parallel_ = false, bad_ = false;
  int nth_ = 1;
  std::vector<uint8_t> comp_, cur_, ahead_;
  std::vector<Blk> blocks_;
  size_t next_blk_ = 0, pos_ = 0;
  std::future<bool> next_;
};

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
      (n_chrom > 0 && !chroms))
    return DMX_E_INVALID;
  *out = nullptr;
  if (!o->tag_umi) {
    g_bam_error = "tag_umi is null";
    return DMX_E_UNSUPPORTED;
  }
  for (int64_t i = 1; i < nsnps; ++i)
    if (snp_rid[i] < snp_rid[i - 1] || (snp_rid[i] == snp_rid[i - 1] && snp_pos[i] < snp_pos[i - 1])) {
      g_bam_error = "SNPs are not sorted by chromosome and position";
      return DMX_E_INVALID;
    }
  ByteStream gz;
  if (!gz.open(path)) {
    g_bam_error = std::string("cannot open ") + path;
    return DMX_E_IO;
  }
  uint8_t hdr[12];
  if (!gz.read(hdr, 8) || memcmp(hdr, "BAM\1", 4) != 0) {
    g_bam_error = std::string(path) + " is not a BAM file";
    return DMX_E_INVALID;
  }
  std::vector<uint8_t> buf((size_t)std::max(0, rd_i32(hdr + 4)));
  if (!buf.empty() && !gz.read(buf.data(), buf.size())) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
  if (!gz.read(hdr, 4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
  const int32_t n_ref = rd_i32(hdr);
  // BAM tid -> VCF rid by name (bcf_hdr_name2id, :188); the reference insists on the same chromosome order (:160-176)
  std::unordered_map<std::string, int32_t> name2rid;
  for (int32_t r = 0; r < n_chrom; ++r) name2rid[chroms[r]] = r;
  std::vector<int32_t> tid2rid((size_t)std::max(0, n_ref), -1);
  int32_t prev = -1;
  for (int32_t t = 0; t < n_ref; ++t) {
    if (!gz.read(hdr, 4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
    const int32_t l = rd_i32(hdr);
    std::vector<char> nm((size_t)std::max(0, l) + 4);
    if (l < 1 || !gz.read(nm.data(), (size_t)l + 4)) { g_bam_error = "truncated BAM header"; return DMX_E_INVALID; }
    auto it = name2rid.find(std::string(nm.data()));
    if (it != name2rid.end()) {
      if (prev >= it->second) {
        g_bam_error = "VCF and BAM have different orderings of chromosomes";   // :169-171
        return DMX_E_INVALID;
      }
      tid2rid[(size_t)t] = prev = it->second;
    }
  }
  std::unordered_set<std::string> groups;
  for (int64_t i = 0; i < o->n_groups; ++i) groups.insert(o->groups[i]);
  const bool use_group = o->tag_group && o->tag_group[0];
  struct random_data rnd_buf;                 // glibc's rand() of an unseeded process: TYPE_3 state, seed 1
  char rnd_state[128];
  memset(&rnd_buf, 0, sizeof(rnd_buf));
  initstate_r(1u, rnd_state, sizeof(rnd_state), &rnd_buf);

  dmx_sweep* sw = new dmx_sweep();
  sw->umi_off.push_back(0);
  std::unordered_map<std::string, int32_t> bc_map;
  int64_t ibeg = 0;                       // first SNP that can still overlap a read (reads are position-sorted)
  static const char kBase[] = "=ACMGRSVTWYHKDBN";
  for (;;) {
    uint8_t bs[4];
    const int k = gz.read4(bs);
    if (k == 0) break;                    // end of file
    if (k != 1) { g_bam_error = "truncated BAM record"; delete sw; return DMX_E_INVALID; }
    const int32_t block = rd_i32(bs);
    if (block < 32) { g_bam_error = "corrupt BAM record"; delete sw; return DMX_E_INVALID; }
    buf.resize((size_t)block);
    if (!gz.read(buf.data(), buf.size())) { g_bam_error = "truncated BAM record"; delete sw; return DMX_E_INVALID; }
    ++sw->n_records;
    const uint8_t* b = buf.data();
    const int32_t tid = rd_i32(b), pos = rd_i32(b + 4);
    const uint32_t l_name = b[8], mapq = b[9], n_cigar = rd_u16(b + 12), flag = rd_u16(b + 14);
    const int32_t l_seq = rd_i32(b + 16);
    const uint8_t* cigar = b + 32 + l_name;
    const uint8_t* seq = cigar + 4 * (size_t)n_cigar;
    const uint8_t* qual = seq + ((size_t)l_seq + 1) / 2;
    const uint8_t* aux = qual + (size_t)l_seq;
    const uint8_t* end = b + block;
    if (l_seq < 0 || aux > end) { g_bam_error = "corrupt BAM record"; delete sw; return DMX_E_INVALID; }
    // sam_filtered_reader.cpp:284-296
    if ((int32_t)mapq < o->min_mq || (o->exclude_flag & flag)) { ++sw->n_filtered; continue; }
    const int32_t rid = (tid >= 0 && tid < n_ref) ? tid2rid[(size_t)tid] : -1;
    if (rid < 0) { ++sw->n_off_vcf; continue; }                                      // :188-191
    // bam_endpos: start + reference span of the CIGAR (M, D, N, =, X), at least one base
    int64_t span = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
      const uint32_t v = rd_u32(cigar + 4 * c), op = v & 15u;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += v >> 4;
    }
    const int64_t endpos = (int64_t)pos + (span > 0 ? span : 1);
    // barcode (:239-268)
    int32_t ibcd = 0;
    const char* sbcd = ".";
    if (use_group) {
      const char* z = aux_get_z(aux, end, o->tag_group);
      if (z) sbcd = z;
      if (!groups.empty() && groups.find(sbcd) == groups.end()) { ++sw->n_skip_bcd; continue; }
    }
    {
      auto it = bc_map.find(sbcd);
      if (it == bc_map.end()) {
        ibcd = (int32_t)sw->barcodes.size();
        bc_map.emplace(sbcd, ibcd);
        sw->barcodes.emplace_back(sbcd);
        sw->totl.push_back(0);
      } else {
        ibcd = it->second;
      }
    }
    // UMI (:272-293): a missing tag leaves "."; --tag-UMI "" draws "." + hex(rand()) per read (:274-276).  The
    // reference never seeds rand(), so its UMIs are glibc's default sequence (TYPE_3 generator, seed 1): the same
    // numbers come from a private random_r state here, whatever else the process does with rand().
    const char* sumi = ".";
    char rnd_umi[16];
    if (!o->tag_umi[0]) {
      int32_t rv = 0;
      random_r(&rnd_buf, &rv);
      snprintf(rnd_umi, sizeof(rnd_umi), ".%x", (unsigned)rv);
      sumi = rnd_umi;
    } else {
      const char* z = aux_get_z(aux, end, o->tag_umi);
      if (z) sumi = z;
    }
    ++sw->totl[(size_t)ibcd];                                                        // :295
    // SNPs of this chromosome from the read start on
    while (ibeg < nsnps && (snp_rid[ibeg] < rid || (snp_rid[ibeg] == rid && snp_pos[ibeg] < pos))) ++ibeg;
    const size_t umi_len = strlen(sumi);
    for (int64_t i = ibeg; i < nsnps && snp_rid[i] == rid && snp_pos[i] < endpos; ++i) {
      // hts_utils.cpp:279-359
      const uint32_t spos = (uint32_t)snp_pos[i];
      uint32_t cpos = (uint32_t)pos;
      int32_t rpos = 0;
      if (!n_cigar) continue;
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
      if (base == 'N') continue;                                                      // :306
      // the reference carries the quality as a (signed) char qual + 33 and subtracts 33 again (:310, :315): values
      // >= 95, and the 0xff of a record without qualities, wrap to negative numbers and are dropped by --min-BQ
      const int32_t q = (int32_t)(int8_t)(uint8_t)(qual[rpos] + 33) - 33;
      if (q < o->min_bq) continue;                                                   // :310
      if (rpos < o->min_td - 1) continue;                                            // :311
      if (rpos + o->min_td > l_seq) continue;                                        // :312
      sw->cell.push_back(ibcd);
      sw->snp.push_back((int32_t)i);
      sw->allele.push_back((uint8_t)(base == snp_ref[i] ? 0 : (base == snp_alt[i] ? 1 : 2)));   // :314
      sw->bq.push_back((uint8_t)(q > o->cap_bq ? o->cap_bq : q));                    // :315
      sw->umi_buf.insert(sw->umi_buf.end(), sumi, sumi + umi_len);
      sw->umi_off.push_back((int64_t)sw->umi_buf.size());
    }
  }
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

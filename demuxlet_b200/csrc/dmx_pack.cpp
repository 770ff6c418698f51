// dmx_pack.cpp -- sort-based pileup packer: replaces the reference's read-by-read insertion into
// std::map<snp, std::map<cell, std::map<umi, uint32>>> (sc_dropseq_lib_t::add_read, sc_drop_seq.cpp:34-77) and the
// flattening of that container by one bucket pass and one sort per cell.  Host code, no device.
//
// Semantics kept from add_read: every call counts as a pass-filtered read of its cell (:40); a (snp, cell, umi) seen for
// the first time stores that read's allele and quality and counts as a unique read (:43-48, :53-55, :74); later reads of
// the same key only bump a counter the likelihood path never looks at (:58).  Order kept from the containers: cells in
// id order, SNPs ascending inside a cell (std::map<int32_t, ...>), UMIs in std::string order inside an entry.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "../../include/dmx.h"

struct dmx_pileup {
  int64_t ncells = 0;
  std::vector<int64_t> cell_ptr, ent_ptr;
  std::vector<int32_t> ent_snp, pass, uniq;
  std::vector<uint16_t> reads;
};

namespace {
template <class F>
void for_cells(int64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 64) {
    for (int64_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> th;
  const int64_t chunk = std::max<int64_t>(1, n / (nthreads * 32));
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const int64_t lo = next.fetch_add(chunk);
        if (lo >= n) break;
        for (int64_t i = lo; i < std::min(n, lo + chunk); ++i) f(i);
      }
    });
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

int dmx_pack_reads(int64_t n, const int32_t* cell, const int32_t* snp, const uint8_t* allele, const uint8_t* bq,
                   const char* umi_buf, const int64_t* umi_off, int64_t ncells, int32_t n_threads, dmx_pileup** out) {
  if (!out || n < 0 || ncells < 0 || (n > 0 && (!cell || !snp || !allele || !bq || !umi_buf || !umi_off)))
    return DMX_E_INVALID;
  *out = nullptr;
  for (int64_t i = 0; i < n; ++i)
    if (cell[i] < 0 || cell[i] >= ncells || snp[i] < 0 || umi_off[i + 1] < umi_off[i]) return DMX_E_INVALID;
  dmx_pileup* p = new (std::nothrow) dmx_pileup();
  if (!p) return DMX_E_NOMEM;
  int nth = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nth = std::max(1, nth);
  try {
    p->ncells = ncells;
    // 1. bucket the read indices by cell, arrival order kept inside a bucket
    std::vector<int64_t> start(ncells + 1, 0);
    for (int64_t i = 0; i < n; ++i) ++start[cell[i] + 1];
    for (int64_t c = 0; c < ncells; ++c) start[c + 1] += start[c];
    // sort keys: SNP id, the first 8 UMI bytes as a big-endian integer (its order is memcmp's), arrival index; the
    // full strings are compared only when two 8-byte prefixes tie
    struct Key {
      uint64_t pre;
      int64_t idx;
      int32_t snp;
    };
    auto prefix = [&](int64_t i) {
      const int64_t len = umi_off[i + 1] - umi_off[i];
      const unsigned char* u = reinterpret_cast<const unsigned char*>(umi_buf + umi_off[i]);
      uint64_t v = 0;
      for (int b = 0; b < 8; ++b) v = (v << 8) | (b < len ? u[b] : 0u);
      return v;
    };
    std::vector<Key> keys(n);
    {
      std::vector<int64_t> fill(start.begin(), start.end() - 1);
      for (int64_t i = 0; i < n; ++i) {
        Key& k = keys[fill[cell[i]]++];
        k.idx = i;
        k.snp = snp[i];
      }
    }
    // 2. per cell: sort by (snp, umi string, arrival); count entries and distinct (snp, umi)
    auto umi_cmp = [&](const Key& a, const Key& b) {   // std::string::compare order: <0, 0, >0
      if (a.pre != b.pre) return a.pre < b.pre ? -1 : 1;
      const int64_t la = umi_off[a.idx + 1] - umi_off[a.idx], lb = umi_off[b.idx + 1] - umi_off[b.idx];
      if (la <= 8 && lb <= 8) return la < lb ? -1 : (la > lb ? 1 : 0);   // equal prefixes: only zero padding differs
      const int r = memcmp(umi_buf + umi_off[a.idx], umi_buf + umi_off[b.idx], (size_t)std::min(la, lb));
      return r != 0 ? r : (la < lb ? -1 : (la > lb ? 1 : 0));
    };
    std::vector<int64_t> n_ent(ncells + 1, 0), n_rd(ncells + 1, 0);
    p->pass.assign(ncells, 0);
    p->uniq.assign(ncells, 0);
    for_cells(ncells, nth, [&](int64_t c) {
      Key* b = keys.data() + start[c];
      Key* e = keys.data() + start[c + 1];
      for (Key* q = b; q < e; ++q) q->pre = prefix(q->idx);
      std::sort(b, e, [&](const Key& x, const Key& y) {
        if (x.snp != y.snp) return x.snp < y.snp;
        const int r = umi_cmp(x, y);
        return r != 0 ? r < 0 : x.idx < y.idx;
      });
      int64_t ne = 0, nr = 0;
      for (Key* q = b; q < e; ++q) {
        const bool new_snp = q == b || q->snp != (q - 1)->snp;
        if (new_snp) ++ne;
        if (new_snp || umi_cmp(*q, *(q - 1)) != 0) ++nr;
      }
      n_ent[c + 1] = ne;
      n_rd[c + 1] = nr;
      p->pass[c] = (int32_t)(e - b);          // :40  ++cell_pass_reads on every call
      p->uniq[c] = (int32_t)nr;               // :74  ++cell_uniq_reads on every new (snp, cell, umi)
    });
    // 3. offsets, then the flat arrays
    for (int64_t c = 0; c < ncells; ++c) { n_ent[c + 1] += n_ent[c]; n_rd[c + 1] += n_rd[c]; }
    const int64_t E = n_ent[ncells], R = n_rd[ncells];
    p->cell_ptr.assign(n_ent.begin(), n_ent.end());
    p->ent_snp.resize(E);
    p->ent_ptr.resize(E + 1);
    p->reads.resize(R);
    p->ent_ptr[E] = R;
    for_cells(ncells, nth, [&](int64_t c) {
      const Key* b = keys.data() + start[c];
      const Key* e = keys.data() + start[c + 1];
      int64_t ie = n_ent[c], ir = n_rd[c];
      for (const Key* q = b; q < e; ++q) {
        const bool new_snp = q == b || q->snp != (q - 1)->snp;
        if (new_snp) {
          p->ent_snp[ie] = q->snp;
          p->ent_ptr[ie] = ir;
          ++ie;
        }
        if (new_snp || umi_cmp(*q, *(q - 1)) != 0)   // first arrival of this (snp, umi): its allele / quality stay (:44, :53)
          p->reads[ir++] = (uint16_t)(((uint16_t)allele[q->idx] << 8) | bq[q->idx]);
      }
    });
  } catch (const std::bad_alloc&) {
    delete p;
    return DMX_E_NOMEM;
  }
  *out = p;
  return DMX_OK;
}

int dmx_pileup_sizes(const dmx_pileup* p, int64_t* ncells, int64_t* n_entries, int64_t* n_reads) {
  if (!p) return DMX_E_INVALID;
  if (ncells) *ncells = p->ncells;
  if (n_entries) *n_entries = (int64_t)p->ent_snp.size();
  if (n_reads) *n_reads = (int64_t)p->reads.size();
  return DMX_OK;
}

int dmx_pileup_copy(const dmx_pileup* p, int64_t* cell_ptr, int32_t* ent_snp, int64_t* ent_ptr, uint16_t* reads,
                    int32_t* pass, int32_t* uniq) {
  if (!p) return DMX_E_INVALID;
  if (cell_ptr) memcpy(cell_ptr, p->cell_ptr.data(), p->cell_ptr.size() * sizeof(int64_t));
  if (ent_snp && !p->ent_snp.empty()) memcpy(ent_snp, p->ent_snp.data(), p->ent_snp.size() * sizeof(int32_t));
  if (ent_ptr) memcpy(ent_ptr, p->ent_ptr.data(), p->ent_ptr.size() * sizeof(int64_t));
  if (reads && !p->reads.empty()) memcpy(reads, p->reads.data(), p->reads.size() * sizeof(uint16_t));
  if (pass && p->ncells) memcpy(pass, p->pass.data(), p->pass.size() * sizeof(int32_t));
  if (uniq && p->ncells) memcpy(uniq, p->uniq.data(), p->uniq.size() * sizeof(int32_t));
  return DMX_OK;
}

void dmx_pileup_free(dmx_pileup* p) { delete p; }

}  // extern "C"

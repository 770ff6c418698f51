// dmx_finalise.cpp -- host half of the path: the exact scorer and the per-cell decision logic.
//
// Why this exists (DESIGN.md "tie finaliser"): at alpha = 0.5 the reference's llksAB[j][k] and
// llksAB[k][j] are the same nine terms summed in row-major vs column-major order
// (/root/reference/cmd_cram_demuxlet.cpp:675-681), so which of the two its strict-< argmax (:802-814)
// keeps is decided by rounding noise of IEEE double arithmetic and of the host libm's log().  No GPU
// arithmetic reproduces that noise, so the kernels return every hypothesis within a relative window of
// each maximum and this file re-scores just those few with the reference's own operation order on the
// host, then applies the reference's comparisons.  Build with -ffp-contract=off.
#include "dmx_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace dmx {

void build_pgrid(const double* alpha, int nA, PGrid* g) {
  g->U = 0;
  for (int n = 0; n < nA; ++n) {
    for (int l = 0; l < 3; ++l)
      for (int m = 0; m < 3; ++m) {
        const double p = 0.5 * l + (m - l) * 0.5 * alpha[n];   // :613
        int u = -1;
        for (int i = 0; i < g->U; ++i)
          if (memcmp(&g->pu[i], &p, sizeof(double)) == 0) { u = i; break; }
        if (u < 0) { u = g->U++; g->pu[u] = p; }
        g->umap[n * 9 + l * 3 + m] = (uint8_t)u;
      }
    g->mode[n] = (n == 0 && alpha[n] == 0.0) ? PM_SEP : (alpha[n] == 0.5 ? PM_SYM : PM_FULL);
  }
  g->sep0 = g->mode[0] == PM_SEP;
}

ExactScorer::ExactScorer(const HostView& hv, const double* err256) : hv_(hv) {
  for (int i = 0; i < 256; ++i) {
    err_[i] = err256[i];
    mat_[i] = 1. - err256[i];                                  // PhredHelper.cpp:32
  }
  // memo for entries with a single informative read (about 4 in 5 of them)
  const int nq = hv_.nA * 9;
  memo1_.assign((size_t)2 * 256 * nq, 0.0);
  for (int al = 0; al < 2; ++al)
    for (int bq = 0; bq < 256; ++bq) {
      double* pGs = &memo1_[((size_t)al * 256 + bq) * nq];
      for (int i = 0; i < nq; ++i) pGs[i] = 1.0;
      const double pR = (al == 0) ? mat_[bq] : err_[bq] / 3.0;
      const double pA = (al == 1) ? mat_[bq] : err_[bq] / 3.0;
      double mx = 0;
      for (int n = 0; n < hv_.nA; ++n)
        for (int l = 0; l < 3; ++l)
          for (int m = 0; m < 3; ++m) {
            const double p = 0.5 * l + (m - l) * 0.5 * hv_.alpha[n];
            double& x = pGs[n * 9 + l * 3 + m];
            x *= (pR * (1.0 - p) + pA * p);
            if (mx < x) mx = x;
          }
      for (int i = 0; i < nq; ++i) pGs[i] /= mx;
      mx = 0;
      for (int i = 0; i < nq; ++i) {
        pGs[i] += 1e-6;
        if (mx < pGs[i]) mx = pGs[i];
      }
      for (int i = 0; i < nq; ++i) pGs[i] /= mx;
    }
}

// :597-663 for one entry
void ExactScorer::entry_pgs(int64_t e, double* pGs) const {
  const int nA = hv_.nA, nq = nA * 9;
  const int64_t r0 = hv_.ent_ptr[e], r1 = hv_.ent_ptr[e + 1];
  int ninf = 0;
  uint16_t first = 0;
  for (int64_t r = r0; r < r1; ++r)
    if ((hv_.reads[r] >> 8) != 2) {
      if (ninf == 0) first = hv_.reads[r];
      ++ninf;
    }
  if (ninf == 0) {
    for (int i = 0; i < nq; ++i) pGs[i] = 1.0;
    return;
  }
  if (ninf == 1) {
    memcpy(pGs, &memo1_[((size_t)(first >> 8) * 256 + (first & 0xff)) * nq], sizeof(double) * nq);
    return;
  }
  for (int i = 0; i < nq; ++i) pGs[i] = 1.0;
  for (int64_t r = r0; r < r1; ++r) {
    const int al = hv_.reads[r] >> 8, bq = hv_.reads[r] & 0xff;
    if (al == 2) continue;                                       // :604
    const double pR = (al == 0) ? mat_[bq] : err_[bq] / 3.0;     // :606
    const double pA = (al == 1) ? mat_[bq] : err_[bq] / 3.0;     // :607
    double mx = 0;
    for (int n = 0; n < nA; ++n)
      for (int l = 0; l < 3; ++l)
        for (int m = 0; m < 3; ++m) {
          const double p = 0.5 * l + (m - l) * 0.5 * hv_.alpha[n];   // :613
          double& x = pGs[n * 9 + l * 3 + m];
          x *= (pR * (1.0 - p) + pA * p);                            // :625
          if (mx < x) mx = x;
        }
    for (int i = 0; i < nq; ++i) pGs[i] /= mx;                   // :632-639
  }
  double mx = 0;
  for (int i = 0; i < nq; ++i) {                                 // :643-654
    pGs[i] += 1e-6;
    if (mx < pGs[i]) mx = pGs[i];
  }
  for (int i = 0; i < nq; ++i) pGs[i] /= mx;                     // :656-663
}

void ExactScorer::eval(int64_t cell, const Hyp* hyps, int nh, double* out) const {
  const int nv = hv_.nv;
  double pGs[DMX_MAX_ALPHA * 9];
  for (int h = 0; h < nh; ++h) out[h] = 0.0;
  for (int64_t e = hv_.cell_ptr[cell]; e < hv_.cell_ptr[cell + 1]; ++e) {
    entry_pgs(e, pGs);
    const float* g = hv_.gp + (size_t)hv_.ent_snp[e] * nv * 3;
    for (int h = 0; h < nh; ++h) {
      const float* gj = g + 3 * hyps[h].j;
      const float* gk = g + 3 * hyps[h].k;
      const double* q = pGs + hyps[h].n * 9;
      double sum = 0;
      for (int l = 0; l < 3; ++l)
        for (int m = 0; m < 3; ++m) {
          const double p = (double)gj[l] * (double)gk[m];        // gpAB, :553
          sum += (p * q[l * 3 + m]);                             // :678
        }
      out[h] += log(sum);                                        // :683
    }
  }
}

void candidates_from_dense(const double* dense, int nv, int nA, int first_arg_plane, double eps_rel, double eps_abs,
                           std::vector<Hyp>* out) {
  out->clear();
  double best = -INFINITY;
  for (int j = 0; j < nv; ++j)
    for (int k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (int n = first_arg_plane; n < nA; ++n) best = fmax(best, dense[((size_t)j * nv + k) * nA + n]);
    }
  const double thr = best - (eps_rel * fmax(1.0, fabs(best)) + eps_abs);
  for (int j = 0; j < nv; ++j)
    for (int k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (int n = first_arg_plane; n < nA; ++n)
        if (dense[((size_t)j * nv + k) * nA + n] >= thr) out->push_back(Hyp{j, k, n});
    }
}

namespace {
struct ExactCache {
  std::vector<Hyp> h;
  std::vector<double> v;
  bool find(int j, int k, int n, double* out) const {
    for (size_t i = 0; i < h.size(); ++i)
      if (h[i].j == j && h[i].k == k && h[i].n == n) { *out = v[i]; return true; }
    return false;
  }
};
inline bool near(double a, double b, double eps_rel, double eps_abs) {
  return fabs(a - b) <= eps_rel * (fmax(1.0, fabs(a)) + fmax(1.0, fabs(b))) + 2.0 * eps_abs;
}
}  // namespace

void finalise_cell(const HostView& hv, const ExactScorer& sc, int64_t cell, const CellGpu& g,
                   const std::vector<Hyp>& cands, double doublet_prior, double eps_rel, double eps_abs, bool exact,
                   const RescList* gl, dmx_best_rec* rec, int32_t* n_from_gpu) {
  const int nv = hv.nv;
  // per-thread scratch, reused from cell to cell (a run finalises tens of thousands of cells per thread)
  thread_local std::vector<double> vals;
  thread_local std::vector<uint8_t> is_exact;
  thread_local ExactCache cache;
  thread_local std::vector<Hyp> need;
  thread_local std::vector<double> out, mo;
  thread_local std::vector<Hyp> miss;
  thread_local std::vector<size_t> miss_at;
  vals.assign(g.sing, g.sing + nv);
  is_exact.assign(nv, 0);
  cache.h.clear();
  cache.v.clear();
  need.clear();

  // ---- which singlets are too close to call (:746-758 compares llksAB[j][0][0]) ---------------
  {
    int i1 = 0;
    for (int j = 1; j < nv; ++j)
      if (vals[j] > vals[i1]) i1 = j;
    double m2 = -INFINITY;
    for (int j = 0; j < nv; ++j)
      if (j != i1) m2 = fmax(m2, vals[j]);
    const double w = 2.0 * (eps_rel * fmax(1.0, fabs(m2)) + eps_abs);
    int close2 = 0;
    for (int j = 0; j < nv; ++j)
      if (j != i1 && vals[j] >= m2 - w) ++close2;
    if (exact && (vals[i1] - m2 <= w || close2 >= 2))
      for (int j = 0; j < nv; ++j)
        if (vals[j] >= m2 - w) need.push_back(Hyp{j, 0, 0});
  }
  const size_t n_sing_need = need.size();
  if (exact && cands.size() > 1) need.insert(need.end(), cands.begin(), cands.end());

  auto run_batch = [&](std::vector<Hyp>& hs) {
    if (hs.empty()) return;
    out.assign(hs.size(), 0.0);
    // values the device re-scorer already produced (bit-identical to the host arithmetic, verified at
    // dmx_create); the rest is scored here
    miss.clear();
    miss_at.clear();
    for (size_t i = 0; i < hs.size(); ++i) {
      bool hit = false;
      if (gl && gl->n > 0) {
        const uint32_t key = (uint32_t)hs[i].j | ((uint32_t)hs[i].k << 12) | ((uint32_t)hs[i].n << 24);
        for (int q = 0; q < gl->n; ++q)
          if (gl->hyp[q] == key) { out[i] = gl->val[q]; hit = true; break; }
      }
      if (hit) ++*n_from_gpu;
      else { miss.push_back(hs[i]); miss_at.push_back(i); }
    }
    if (!miss.empty()) {
      mo.assign(miss.size(), 0.0);
      sc.eval(cell, miss.data(), (int)miss.size(), mo.data());
      for (size_t i = 0; i < miss.size(); ++i) out[miss_at[i]] = mo[i];
    }
    for (size_t i = 0; i < hs.size(); ++i) {
      cache.h.push_back(hs[i]);
      cache.v.push_back(out[i]);
      if (hs[i].k == 0 && hs[i].n == 0) { vals[hs[i].j] = out[i]; is_exact[hs[i].j] = 1; }
    }
    rec->n_rescored += (int32_t)hs.size();
    hs.clear();
  };
  rec->n_rescored = 0;
  rec->exact_mask = 0;
  (void)n_sing_need;
  run_batch(need);

  // ---- top-2 singlets, :746-758 ------------------------------------------------------------------
  int iS1 = -1, iS2 = -1;
  {
    double max1 = -1e300, max2 = -1e300;
    for (int j = 0; j < nv; ++j) {
      if (max1 < vals[j]) { max2 = max1; iS2 = iS1; iS1 = j; max1 = vals[j]; }
      else if (max2 < vals[j]) { iS2 = j; max2 = vals[j]; }
    }
  }
  // ---- best pair, :799-814 -----------------------------------------------------------------------
  int jB = -1, kB = -1, nB = -1;
  double llk12 = g.co->best_v;
  bool llk12_exact = false;
  if (cands.size() == 1) {
    jB = cands[0].j; kB = cands[0].k; nB = cands[0].n;
  } else if (!cands.empty()) {
    double mx = -1e300;
    for (size_t i = 0; i < cands.size(); ++i) {   // cands are in (j,k,n) lexicographic order
      double v = g.co->best_v;
      const bool have = cache.find(cands[i].j, cands[i].k, cands[i].n, &v);
      if (!have && i > 0) continue;               // non-exact mode: first candidate wins
      if (mx < v) { mx = v; jB = cands[i].j; kB = cands[i].k; nB = cands[i].n; llk12_exact = have; }
    }
    llk12 = mx;
  }
  if (iS1 < 0 || iS2 < 0 || jB < 0) {   // NaN inputs: nothing compares true; keep indices valid
    if (iS1 < 0) iS1 = 0;
    if (iS2 < 0) iS2 = iS1 == 0 && nv > 1 ? 1 : 0;
    if (jB < 0) { jB = 0; kB = nv > 1 ? 1 : 0; nB = hv.nA > 1 ? 1 : 0; }
  }

  // ---- decision thresholds :837, :844: exactify any comparison that is too close ------------------
  auto exactify_sing = [&](int j) {
    if (!is_exact[j]) need.push_back(Hyp{j, 0, 0});
  };
  if (exact) {
    const double s1 = vals[iS1], s2 = vals[iS2], l1 = vals[jB], l2 = vals[kB];
    bool want12 = false;
    if (near(llk12, l1, eps_rel, eps_abs)) { want12 = true; exactify_sing(jB); }
    if (near(llk12, l2, eps_rel, eps_abs)) { want12 = true; exactify_sing(kB); }
    if (near(llk12, s1 + 2, eps_rel, eps_abs)) { want12 = true; exactify_sing(iS1); }
    if (near(s1, s2 + 2, eps_rel, eps_abs)) { exactify_sing(iS1); exactify_sing(iS2); }
    if (want12 && !llk12_exact) need.push_back(Hyp{jB, kB, nB});
    // de-duplicate
    std::sort(need.begin(), need.end(), [](const Hyp& a, const Hyp& b) {
      return a.j != b.j ? a.j < b.j : (a.k != b.k ? a.k < b.k : a.n < b.n);
    });
    need.erase(std::unique(need.begin(), need.end(),
                           [](const Hyp& a, const Hyp& b) { return a.j == b.j && a.k == b.k && a.n == b.n; }),
               need.end());
    if (!need.empty()) {
      run_batch(need);
      double v;
      if (cache.find(jB, kB, nB, &v)) { llk12 = v; llk12_exact = true; }
    }
  }

  rec->sng1 = iS1; rec->sng2 = iS2;
  rec->dbl1 = jB; rec->dbl2 = kB; rec->alpha_idx = nB;
  rec->sng_llk1 = vals[iS1];
  rec->sng_llk2 = vals[iS2];
  rec->sng_llk0 = g.l00[0];
  rec->llk12 = llk12;
  rec->llk1 = vals[jB];
  rec->llk2 = vals[kB];
  rec->llk10 = g.row0[(size_t)nB * nv + jB];   // :824 (nB >= 1: always a dense plane)
  rec->llk20 = g.row0[(size_t)nB * nv + kB];   // :825
  rec->llk00 = g.l00[nB];
  rec->max_llk = g.co->max_llk;
  rec->sum_single = g.co->sum_single;
  rec->sum_double = g.co->sum_double;
  rec->prb_dbl = rec->sum_double / (rec->sum_single + rec->sum_double);                     // :827
  rec->prb_sng1 = exp(rec->sng_llk1 - rec->max_llk) * (1. - doublet_prior) / nv / rec->sum_single;   // :828
  if (is_exact[iS1]) rec->exact_mask |= DMX_X_SNG1 | (jB == iS1 ? DMX_X_LLK1 : 0) | (kB == iS1 ? DMX_X_LLK2 : 0);
  if (is_exact[iS2]) rec->exact_mask |= DMX_X_SNG2;
  if (is_exact[jB]) rec->exact_mask |= DMX_X_LLK1;
  if (is_exact[kB]) rec->exact_mask |= DMX_X_LLK2;
  if (llk12_exact) rec->exact_mask |= DMX_X_LLK12;
  if ((rec->llk12 > rec->llk1) && (rec->llk12 > rec->llk2) && (rec->llk12 > rec->sng_llk1 + 2))   // :837
    rec->call = 0;
  else if (rec->sng_llk1 > rec->sng_llk2 + 2)                                                     // :844
    rec->call = 1;
  else
    rec->call = 2;
}

}  // namespace dmx

// ---- locating the host libm's log table -------------------------------------------------------------
#include <stdio.h>
namespace dmx {
bool find_libm_log_table(double out[18 + 256]) {
  // DMX_NO_LIBM_TABLE: behave as on a host whose libm has no such table (tests of the host-scorer fallback)
  if (getenv("DMX_NO_LIBM_TABLE")) return false;
  // path of the libm this process has mapped
  std::string path;
  if (FILE* m = fopen("/proc/self/maps", "r")) {
    char line[1024];
    while (fgets(line, sizeof(line), m)) {
      const char* p = strstr(line, "/");
      if (p && strstr(p, "libm.so")) {
        path = p;
        while (!path.empty() && (path.back() == '\n' || path.back() == ' ')) path.pop_back();
        break;
      }
    }
    fclose(m);
  }
  const char* fallbacks[] = {"/lib/x86_64-linux-gnu/libm.so.6", "/usr/lib/x86_64-linux-gnu/libm.so.6",
                             "/lib64/libm.so.6", "/usr/lib64/libm.so.6"};
  std::vector<std::string> tries;
  if (!path.empty()) tries.push_back(path);
  for (const char* fb : fallbacks) tries.push_back(fb);
  // struct log_data starts with ln2hi, ln2lo (N = 128 variant), then poly[0] ~ -0.5, ...
  const double ln2hi = 0x1.62e42fefa3800p-1, ln2lo = 0x1.ef35793c76730p-45;
  for (const std::string& t : tries) {
    FILE* fp = fopen(t.c_str(), "rb");
    if (!fp) continue;
    std::vector<unsigned char> buf;
    unsigned char chunk[65536];
    size_t n;
    while ((n = fread(chunk, 1, sizeof(chunk), fp)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    fclose(fp);
    const size_t need = (18 + 256) * sizeof(double);
    for (size_t off = 0; off + need <= buf.size(); off += 8) {
      double a, b;
      memcpy(&a, &buf[off], 8);
      if (a != ln2hi) continue;
      memcpy(&b, &buf[off + 8], 8);
      if (b != ln2lo) continue;
      double tab[18 + 256];
      memcpy(tab, &buf[off], need);
      // sanity: polynomial heads and a monotone invc column that brackets [1/1.42, 1/0.69]
      bool ok = fabs(tab[2] + 0.5) < 1e-9 && tab[7] == -0.5 && fabs(tab[8] - 1.0 / 3) < 1e-9;
      for (int i = 0; ok && i < 128; ++i) {
        const double invc = tab[18 + 2 * i], logc = tab[19 + 2 * i];
        if (!(invc > 0.69 && invc < 1.46) || fabs(logc + log(invc)) > 1e-9) ok = false;
      }
      if (ok) {
        memcpy(out, tab, need);
        return true;
      }
    }
  }
  return false;
}
}  // namespace dmx

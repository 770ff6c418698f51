// dmx_scan2.cuh -- the fp32 pair scan, second generation (pair_scan_v2_kernel).  Textually included by dmx_kernels.cu
// inside namespace dmx (it shares the mbarrier / TMA helpers, the task walk and the per-task reductions with the
// first-generation kernels there).
//
// What changed against pair_scan_f32_kernel, and why (profiles/r2n_prof.txt): that kernel moved
// every stage through two hand-overs (TMA -> transform warps -> processed ring -> consumers, plus a products ring to an
// accumulator warp), and the seven warps serving those hand-overs -- each a single dependent instruction stream of
// 500-3000 cycles per stage -- cost 45 % of the issued instructions and left the nine consumer warps waiting a quarter of
// the time.  Here the operands the consumers need are laid out in HBM so that the TMA gathers deliver them READY TO USE:
//
//   affine entries (<= 1 informative read, 4 in 5): the expected dosage D of the read's allele under the normalised
//     genotype vector, one float per sample, from the per-(SNP, allele) table `drow` (dtab_kernel); the likelihood of a
//     pair is S' = M (alpha + bl D_j + bm D_k) with three per-entry constants from the fp32 plane record, so a consumer
//     thread forms X_i = alpha + bl D_j[i] (8 FFMA), Y_c = bm D_k[c] (8 FMUL) and acc[i][c] *= X_i + Y_c (64 FADD + 64
//     FMUL) straight from the gathered rows.  The factor M is common to all hypotheses of the plane: its logs are summed
//     per cell in fp64 by cell_scalars_kernel (ScanParams::lm).  The row sums s the normalisation divided out come from
//     a second per-SNP table `srow`; their logs (lb) are added back in the epilogue as before.
//   general entries: the raw GP rows as gathered; a consumer thread forms w_k = Q' g_k for its eight k samples itself
//     (72 FFMA) and then the K = 3 dot products as before, on the UNnormalised vectors (no lb term for these entries).
//
// So there is one hand-over (TMA -> everybody), no transform warps, no processed ring and no products ring:
//   warps 0..8    consumers, a thread owns an 8 x 8 tile of hypotheses for entry lane & 7 of every stage (as before)
//   warps 9..11   "singles": the per-sample running products (singlet factor of the separable plane, row sums, pass 1,
//                 pairs with sample 0), stage i of a task by warp i % 3, every lane with its own products and exponents;
//                 combined in a fixed order at the end of a task (bit-reproducible)
//   warps 13..15  loaders, stage g by warp g % 3: tile::gather4 copies of the rows + two bulk copies of the records
// Bank conflicts: the eight entry rows of a stage are two gather4 boxes of four rows; row pitches of 18 / 10 / 50 / 26
// sixteen-byte chunks (= 2 mod 8) spread four rows over the even bank groups, and the second box is fetched four floats
// further left in the table so that its rows land one chunk to the right, on the odd groups: every LDS.128 of a quarter
// warp (eight entry lanes, same column) is conflict free.
#pragma once

namespace {
constexpr int V_NCW = 9, V_NSW = 3, V_NLD = 3;
constexpr int V_W_SING0 = 9, V_W_LOAD0 = 13;    // warp 12 is idle: sub-partition 0 already runs three consumers
constexpr uint32_t V_DPJ = 72 * 4, V_DPK = 40 * 4;          // D / s row pitches in a stage (j side, k side of rectangular tasks)
constexpr uint32_t V_GPJ = 200 * 4, V_GPK = 104 * 4;        // raw GP row pitches
constexpr uint32_t V_OFF_HEAD = 0, V_OFF_PLANE = NB * FHEAD * 4, V_OFF_ROWS = NB * (FHEAD + FPLANE) * 4;   // 0, 256, 640
constexpr uint32_t V_OFF_DJ = V_OFF_ROWS, V_OFF_SJ = V_OFF_DJ + NB * V_DPJ, V_OFF_DK = V_OFF_SJ + NB * V_DPJ,
                   V_OFF_SK = V_OFF_DK + NB * V_DPK;
constexpr uint32_t V_OFF_GJ = V_OFF_ROWS, V_OFF_GK = V_OFF_GJ + NB * V_GPJ;
constexpr uint32_t V_SLOT_TRI = V_OFF_GK, V_SLOT_RECT = V_OFF_GK + NB * V_GPK;                            // 7040, 10368
constexpr int V_NSLOT_TRI = 30, V_NSLOT_RECT = 18;
static_assert(V_SLOT_TRI % 128 == 0 && V_SLOT_RECT % 128 == 0 && V_OFF_SJ % 128 == 0 && V_OFF_DK % 128 == 0 &&
              V_OFF_SK % 128 == 0 && (4 * V_DPJ) % 128 == 0 && (4 * V_DPK) % 128 == 0 && (4 * V_GPJ) % 128 == 0 &&
              (4 * V_GPK) % 128 == 0 && V_OFF_SK + NB * V_DPK <= V_SLOT_RECT && V_OFF_DK <= V_SLOT_TRI,
              "v2 stage layout");
static_assert(V_NSLOT_TRI % V_NLD == 0 && V_NSLOT_RECT % V_NLD == 0 && V_W_SING0 + V_NSW <= 12 && V_W_LOAD0 >= 12 && V_W_LOAD0 + V_NLD == 16 && 2 * V_NSLOT_TRI + 4 <= 64,
              "v2 ring ownership / barrier area");
#ifndef DMX_V_LOAD_SLEEP_NS
#define DMX_V_LOAD_SLEEP_NS 800
#endif
constexpr unsigned V_LOAD_SLEEP_NS = DMX_V_LOAD_SLEEP_NS;
constexpr int VQ = 6;   // per-sample quantities of the singles warps: PA, PBa, PBg, P1, P0, PJa
struct VScratch {
  double lb[2][2][BLK];              // [task parity][j side, k side][sample]: sum of log s over the affine entries
  double comb[V_NSW - 1][VQ][BLK];   // partial sums of the singles warps 1, 2 -> warp 0
};
constexpr uint32_t V_TS_BYTES = (sizeof(VScratch) + 127) / 128 * 128;
constexpr uint32_t V_SMEM_TRI = SLOT_CTRL + V_NSLOT_TRI * V_SLOT_TRI + V_TS_BYTES;
constexpr uint32_t V_SMEM_RECT = SLOT_CTRL + V_NSLOT_RECT * V_SLOT_RECT + V_TS_BYTES;
static_assert(V_SMEM_TRI <= 232448 && V_SMEM_RECT <= 232448, "v2 shared memory");

// byte offset of entry row e inside a two-box row region of pitch P: the second box sits one 16-byte chunk to the right
__device__ __forceinline__ uint32_t vrow(uint32_t e, uint32_t pitch) { return e * pitch + ((e & 4u) ? 16u : 0u); }

__device__ __forceinline__ void ld8(const unsigned char* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 16);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// eight samples x three genotypes (96 B) -> g[8][3]
__device__ __forceinline__ void ld24(const unsigned char* p, float (&g)[8][3]) {
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(p + q * 16);
    const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) g[(q * 4 + c) / 3][(q * 4 + c) % 3] = xs[c];
  }
}

// mbarrier helpers on shared-window addresses computed once per role (the generic-pointer forms re-derive the window
// base in every loop iteration).  try_wait without a suspend-time hint: the hinted form compiles to TRYWAIT +
// NANOSLEEP.SYNCS + PHASECHK and the sleep returns at once on this part (profiles/r2s: 57 polls per stage).
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  while (!mbar_try_a(bar, parity)) {}
}
// non-blocking probe (the result is consumed after the stage's arithmetic: its latency is hidden)
__device__ __forceinline__ bool mbar_test_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// DMX_ARRIVE_ALL (sanitizer builds): every lane arrives instead of one lane after __syncwarp, so that racecheck sees an
// arrival by each thread whose accesses the barrier orders (tools/run_profile.sh part s)
#ifndef DMX_ARRIVE_ALL
#define DMX_ARRIVE_ALL 0
#endif
constexpr uint32_t V_ARR = DMX_ARRIVE_ALL ? 32u : 1u;   // arrivals per warp
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bulk copy of streamed data (the entry records: read once per launch) with an L2 evict-first policy, so that the
// stream does not push the table rows the gathers re-use out of the 126 MB L2
__device__ __forceinline__ void bulk_g2s_stream_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                      uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t}" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// A loader runs ten stages ahead of the consumers, so its ring slot is practically never free when it first looks.
// nanosleep returns within tens of cycles here, and every poll costs issue slots of the SM sub-partition the loader
// shares with two consumer warps; a dependent uncached load parks the warp on the scoreboard for an L2 round trip
// (~600 cycles) at the price of two instructions.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, const int32_t* probe, unsigned& sink) {
  while (!mbar_try_a(bar, parity)) {
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)   // six dependent round trips (the address of each load depends on the previous value)
      asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(probe + (v & 1u)) : "memory");
    sink += v;
  }
}

// tile arithmetic of the second-generation scan.  The products of a diagonal tile below the diagonal are not wanted:
// they start at zero (and stay there), and the affine stages of the two warps that own the eight diagonal tiles skip
// them (DIAG); the general stages and the range check are one code path for every warp (instruction cache).
template <bool DIAG>
__device__ __forceinline__ void v2_tile_gen(const float (&a)[8][3], const float (&w)[8][3], float (&acc)[64]) {
#pragma unroll
  for (int i0 = 0; i0 < 8; i0 += 2) {
    float s[2][8];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = a[i0 + r][0] * w[c][0];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = fmaf(a[i0 + r][1], w[c][1], s[r][c]);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) s[r][c] = fmaf(a[i0 + r][2], w[c][2], s[r][c]);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i0 + r) acc[(i0 + r) * 8 + c] *= s[r][c];
  }
}
template <bool DIAG>
__device__ __forceinline__ void v2_tile_aff(const float (&x)[8], const float (&y)[8], float (&acc)[64]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (!DIAG || c >= i) acc[i * 8 + c] *= x[i] + y[c];
}

// Range of the running products (block floating point, one integer exponent per thread).  Every factor is in
// [~2^-21, 1 + a few ulp], so the products only shrink: they start at 2^V_TOP and every V_PEEL stages the warp
// checks whether some thread's LARGEST product has fallen below 2^V_LOW -- from there V_PEEL more factors cannot
// take it out of the normal range -- and if so every thread brings its largest product back to [2^V_TOP, 2^(V_TOP+1)).
// A hypothesis more than ~2^160 below the best of its tile underflows to zero (llk = -inf).
constexpr int V_PEEL = 8, V_TOP = 120, V_LOW = -126 + 21 * V_PEEL + 4;
static_assert(V_LOW < V_TOP - 40, "range check window");
template <bool DIAG>
__device__ __forceinline__ void v2_range_check(float (&acc)[64], int& ex, bool active) {
  float m8[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) m8[c] = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (!DIAG || c >= i) m8[c] = fmaxf(m8[c], acc[i * 8 + c]);
  const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
  if (__any_sync(0xffffffffu, active && mx < __uint_as_float((uint32_t)(V_LOW + 127) << 23))) {
    const int eb = (int)(__float_as_uint(mx) >> 23);          // biased exponent of the largest product
    int k = (eb == 0 || eb >= 254) ? 0 : V_TOP + 127 - eb;    // all-zero tile (padding samples only): leave it alone
    ex -= k;
    const int k1 = min(k, 120);                               // 2^k in two representable steps
    const float s1 = __uint_as_float((uint32_t)(127 + k1) << 23), s2 = __uint_as_float((uint32_t)(127 + k - k1) << 23);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (!DIAG || c >= i) acc[i * 8 + c] = (acc[i * 8 + c] * s1) * s2;
  }
}

// 16-byte shared-memory load from a shared-window address (volatile: stays behind the barrier wait it follows)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// keeps a loop-invariant value in its register (at the register limit ptxas otherwise re-derives thread-id arithmetic
// and the shared window base in every iteration of the stage loop)
#define V_KEEP(x) asm volatile("" : "+r"(x))

struct VFac { float q[VQ][4]; };
}  // namespace

// =============================== loaders ==========================================================
__device__ __forceinline__ void v2_loader(const ScanParams& p, const ScanMaps& tm, unsigned char* ring, uint64_t* full,
                                          uint64_t* empty, int lane, int u, int NS, int lid) {
  const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(full), empty0 = smem_u32(empty);
  const uint32_t nslot = (uint32_t)p.v_nslot, slotb = (uint32_t)p.v_slot;
  uint32_t gbase = 0;
  unsigned sink = 0;
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  TaskInfoF ti;
  PROF_DECL
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const int j0 = tt.j0, k0 = tt.k0;
    const float* fplane = p.entf_plane + (int64_t)((tt.qoff - QHEAD) / QPLANE) * p.entf_pstride;
    int it = (int)((V_NLD + lid - gbase % V_NLD) % V_NLD);
    auto fetch = [&](int st) -> int {
      int64_t eb;
      int nvalid;
      bool aff;
      stage_span(ti, st, &eb, &nvalid, &aff);
      return lane < nvalid ? __ldg(p.snp_perm + eb + lane) : 0;
    };
    int idx_nxt = it < ti.nst ? fetch(it) : 0;
    uint32_t slot = (gbase + (uint32_t)it) % nslot, ph = ((gbase + (uint32_t)it) / nslot) & 1u;
    for (; it < ti.nst; it += V_NLD, slot += V_NLD) {
      if (slot >= nslot) { slot -= nslot; ph ^= 1u; }
      int64_t eb;
      int nvalid;
      bool aff;
      stage_span(ti, it, &eb, &nvalid, &aff);
      const int idx_cur = idx_nxt;
      if (it + V_NLD < ti.nst) idx_nxt = fetch(it + V_NLD);
      int s[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) s[i] = __shfl_sync(0xffffffffu, idx_cur, i);
      PROF_WAIT(1, mbar_wait_parked(empty0 + slot * 8u, ph ^ 1u, p.snp_perm, sink));
      const uint32_t rs = ring0 + slot * slotb, bar = full0 + slot * 8u;
      if (aff) {   // rows of the dosage table (SNP * 2 + allele) and of the row-sum table (SNP)
        mbar_expect_tx_elect(bar, V_OFF_ROWS + 2u * NB * V_DPJ + (same ? 0u : 2u * NB * V_DPK));
        tma_gather4_elect(rs + V_OFF_DJ, &tm.dj, 4 + j0, s[0], s[1], s[2], s[3], bar);
        tma_gather4_elect(rs + V_OFF_DJ + 4 * V_DPJ, &tm.dj, j0, s[4], s[5], s[6], s[7], bar);
        tma_gather4_elect(rs + V_OFF_SJ, &tm.sj, 4 + j0, s[0] >> 1, s[1] >> 1, s[2] >> 1, s[3] >> 1, bar);
        tma_gather4_elect(rs + V_OFF_SJ + 4 * V_DPJ, &tm.sj, j0, s[4] >> 1, s[5] >> 1, s[6] >> 1, s[7] >> 1, bar);
        if (!same) {
          tma_gather4_elect(rs + V_OFF_DK, &tm.dk, 4 + k0, s[0], s[1], s[2], s[3], bar);
          tma_gather4_elect(rs + V_OFF_DK + 4 * V_DPK, &tm.dk, k0, s[4], s[5], s[6], s[7], bar);
          tma_gather4_elect(rs + V_OFF_SK, &tm.sk, 4 + k0, s[0] >> 1, s[1] >> 1, s[2] >> 1, s[3] >> 1, bar);
          tma_gather4_elect(rs + V_OFF_SK + 4 * V_DPK, &tm.sk, k0, s[4] >> 1, s[5] >> 1, s[6] >> 1, s[7] >> 1, bar);
        }
      } else {     // raw GP rows
        mbar_expect_tx_elect(bar, V_OFF_ROWS + NB * V_GPJ + (same ? 0u : NB * V_GPK));
        tma_gather4_elect(rs + V_OFF_GJ, &tm.gpj, 3 * j0 + 4, s[0], s[1], s[2], s[3], bar);
        tma_gather4_elect(rs + V_OFF_GJ + 4 * V_GPJ, &tm.gpj, 3 * j0, s[4], s[5], s[6], s[7], bar);
        if (!same) {
          tma_gather4_elect(rs + V_OFF_GK, &tm.gpk, 3 * k0 + 4, s[0], s[1], s[2], s[3], bar);
          tma_gather4_elect(rs + V_OFF_GK + 4 * V_GPK, &tm.gpk, 3 * k0, s[4], s[5], s[6], s[7], bar);
        }
      }
      // entry slots beyond the stage's last entry read the records that follow (the tables have NB records of slack)
      bulk_g2s_stream_elect(rs + V_OFF_HEAD, p.entf_head + eb * FHEAD, NB * FHEAD * 4, bar, pol_stream);
      bulk_g2s_stream_elect(rs + V_OFF_PLANE, fplane + eb * FPLANE, NB * FPLANE * 4, bar, pol_stream);
    }
    gbase += (uint32_t)ti.nst;
  }
  if (sink == 0x9e3779b9u && p.dbg == -1) p.parts[0].pad = (int32_t)sink;   // keeps the parking loads alive; never true
  PROF_END(0, lane)
}

// =============================== singles warps =====================================================
// Lane = (4-sample block q = lane & 15, entry half eh = lane >> 4): entries eh, eh + 2, eh + 4, eh + 6 of its stages.
// Affine entries enter in normalised form (the consumers' convention: lb = sum of log s is added at the end),
// general entries in raw form.  ROW0: 0 none, 1 symmetric plane (pairs (0, x), x on the k side), 2 full plane
// (pairs (x, 0), x on the j side).
// affine stage: the four entries of a lane unrolled and branch free (their loads go out together; an entry slot beyond
// the stage's last entry multiplies by exact ones)
template <bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void v2_singles_aff(const unsigned char* __restrict__ rs, int eh, int q, int nvalid, bool jq,
                                               bool kq, VFac& F) {
#pragma unroll
  for (int i4 = 0; i4 < NB / 2; ++i4) {
    const int e = eh + 2 * i4;
    const bool live = e < nvalid;
    float4 hd = *reinterpret_cast<const float4*>(rs + V_OFF_HEAD + e * (FHEAD * 4));            // h1 c1 hs cs
    float4 pl = *reinterpret_cast<const float4*>(rs + V_OFF_PLANE + e * (FPLANE * 4) + 16);     // alpha bl bm
    hd.x = live ? hd.x : 1.0f; hd.y = live ? hd.y : 0.0f; hd.z = live ? hd.z : 1.0f; hd.w = live ? hd.w : 0.0f;
    pl.x = live ? pl.x : 1.0f; pl.y = live ? pl.y : 0.0f; pl.z = live ? pl.z : 0.0f;
    if (kq) {
      const unsigned char* dk = SAME ? rs + V_OFF_DJ + vrow(e, V_DPJ) : rs + V_OFF_DK + vrow(e, V_DPK);
      const unsigned char* sk = SAME ? rs + V_OFF_SJ + vrow(e, V_DPJ) : rs + V_OFF_SK + vrow(e, V_DPK);
      const float4 d4 = *reinterpret_cast<const float4*>(dk + q * 16), s4 = *reinterpret_cast<const float4*>(sk + q * 16);
      const float D[4] = {d4.x, d4.y, d4.z, d4.w};
      const float S[4] = {live ? s4.x : 1.0f, live ? s4.y : 1.0f, live ? s4.z : 1.0f, live ? s4.w : 1.0f};
      float x0 = 0.f;
      if (ROW0 == 1) x0 = fmaf(pl.y, *reinterpret_cast<const float*>(rs + V_OFF_DJ + vrow(e, V_DPJ)), pl.x);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        F.q[1][i] *= S[i];
        if (EXTRAS) {
          F.q[3][i] *= fmaf(hd.y, D[i], hd.x);
          F.q[0][i] *= fmaf(hd.w, D[i], hd.z);
        }
        if (ROW0 == 1) F.q[4][i] *= fmaf(pl.z, D[i], x0);
      }
    }
    if (!SAME && jq) {
      const float4 s4 = *reinterpret_cast<const float4*>(rs + V_OFF_SJ + vrow(e, V_DPJ) + q * 16);
      F.q[5][0] *= live ? s4.x : 1.0f; F.q[5][1] *= live ? s4.y : 1.0f;
      F.q[5][2] *= live ? s4.z : 1.0f; F.q[5][3] *= live ? s4.w : 1.0f;
      if (ROW0 == 2) {
        const float4 d4 = *reinterpret_cast<const float4*>(rs + V_OFF_DJ + vrow(e, V_DPJ) + q * 16);
        const float y0 = pl.z * *reinterpret_cast<const float*>(rs + V_OFF_DK + vrow(e, V_DPK));
        F.q[4][0] *= fmaf(pl.y, d4.x, pl.x) + y0; F.q[4][1] *= fmaf(pl.y, d4.y, pl.x) + y0;
        F.q[4][2] *= fmaf(pl.y, d4.z, pl.x) + y0; F.q[4][3] *= fmaf(pl.y, d4.w, pl.x) + y0;
      }
    }
  }
}

// general stage (one in five): rolled, the instruction cache is the scarcer resource
template <bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void v2_singles_gen(const unsigned char* __restrict__ rs, int eh, int q, int nvalid, bool jq,
                                               bool kq, VFac& F) {
#pragma unroll 1
  for (int i4 = 0; i4 < NB / 2; ++i4) {
    const int e = eh + 2 * i4;
    if (e >= nvalid) break;
    const float4* hp = reinterpret_cast<const float4*>(rs + V_OFF_HEAD + e * (FHEAD * 4));
    const float4* pp = reinterpret_cast<const float4*>(rs + V_OFF_PLANE + e * (FPLANE * 4));
    {
      const float4 qa = pp[0], qb = pp[1];       // Q'[l][m] row-major: (00 01 02 10) (11 12 20 21) (22 . . .)
      const float q22 = reinterpret_cast<const float*>(pp + 2)[0];
      if (kq) {
        float g[12];
        load_quad(SAME ? rs + V_OFF_GJ + vrow(e, V_GPJ) + 16 + q * 48 : rs + V_OFF_GK + vrow(e, V_GPK) + 16 + q * 48, g);
        const float4 h0 = hp[0];                  // GL0 GL1 GL2 f0
        const float2 h1 = reinterpret_cast<const float2*>(hp + 1)[0];   // f1 f2
        float a00 = 0.f, a01 = 0.f, a02 = 0.f;
        if (ROW0 == 1) {
          const float* a0 = reinterpret_cast<const float*>(rs + V_OFF_GJ + vrow(e, V_GPJ) + 16);
          a00 = a0[0]; a01 = a0[1]; a02 = a0[2];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f0 = g[3 * i], f1 = g[3 * i + 1], f2 = g[3 * i + 2];
          F.q[2][i] *= (f0 + f1) + f2;
          if (EXTRAS) {
            F.q[0][i] *= fmaf(h1.y, f2, fmaf(h1.x, f1, h0.w * f0));
            F.q[3][i] *= fmaf(h0.z, f2, fmaf(h0.y, f1, h0.x * f0));
          }
          if (ROW0 == 1) {
            const float w0 = fmaf(qa.z, f2, fmaf(qa.y, f1, qa.x * f0));
            const float w1 = fmaf(qb.y, f2, fmaf(qb.x, f1, qa.w * f0));
            const float w2 = fmaf(q22, f2, fmaf(qb.w, f1, qb.z * f0));
            F.q[4][i] *= fmaf(a02, w2, fmaf(a01, w1, a00 * w0));
          }
        }
      }
      if (!SAME && jq && ROW0 == 2) {
        float a[12];
        load_quad(rs + V_OFF_GJ + vrow(e, V_GPJ) + 16 + q * 48, a);
        const float* g0p = reinterpret_cast<const float*>(rs + V_OFF_GK + vrow(e, V_GPK) + 16);
        const float f0 = g0p[0], f1 = g0p[1], f2 = g0p[2];
        const float w0 = fmaf(qa.z, f2, fmaf(qa.y, f1, qa.x * f0));
        const float w1 = fmaf(qb.y, f2, fmaf(qb.x, f1, qa.w * f0));
        const float w2 = fmaf(q22, f2, fmaf(qb.w, f1, qb.z * f0));
#pragma unroll
        for (int i = 0; i < 4; ++i) F.q[4][i] *= fmaf(a[3 * i + 2], w2, fmaf(a[3 * i + 1], w1, a[3 * i] * w0));
      }
    }
  }
}

template <bool SAME, bool EXTRAS, int ROW0>
__device__ __forceinline__ void v2_singles_stage(const unsigned char* __restrict__ rs, bool aff, int eh, int q,
                                                 int nvalid, bool jq, bool kq, VFac& F) {
  if (aff) v2_singles_aff<SAME, EXTRAS, ROW0>(rs, eh, q, nvalid, jq, kq, F);
  else v2_singles_gen<SAME, EXTRAS, ROW0>(rs, eh, q, nvalid, jq, kq, F);
}

__device__ __forceinline__ void v2_singles_dispatch(bool same, bool extras_t, int row0, const unsigned char* rs, bool aff,
                                                    int eh, int q, int nvalid, bool jq, bool kq, VFac& F) {
  if (same) {
    if (extras_t) {
      if (row0) v2_singles_stage<true, true, 1>(rs, aff, eh, q, nvalid, jq, kq, F);
      else v2_singles_stage<true, true, 0>(rs, aff, eh, q, nvalid, jq, kq, F);
    } else {
      if (row0) v2_singles_stage<true, false, 1>(rs, aff, eh, q, nvalid, jq, kq, F);
      else v2_singles_stage<true, false, 0>(rs, aff, eh, q, nvalid, jq, kq, F);
    }
  } else {
    if (extras_t) {
      if (row0 == 2) v2_singles_stage<false, true, 2>(rs, aff, eh, q, nvalid, jq, kq, F);
      else if (row0 == 1) v2_singles_stage<false, true, 1>(rs, aff, eh, q, nvalid, jq, kq, F);
      else v2_singles_stage<false, true, 0>(rs, aff, eh, q, nvalid, jq, kq, F);
    } else {
      if (row0 == 2) v2_singles_stage<false, false, 2>(rs, aff, eh, q, nvalid, jq, kq, F);
      else if (row0 == 1) v2_singles_stage<false, false, 1>(rs, aff, eh, q, nvalid, jq, kq, F);
      else v2_singles_stage<false, false, 0>(rs, aff, eh, q, nvalid, jq, kq, F);
    }
  }
}

__device__ __forceinline__ void v2_singles(const ScanParams& p, unsigned char* ring, uint64_t* full, uint64_t* empty,
                                           VScratch* ts, uint64_t* lb_full, uint64_t* lb_empty, int lane, int u, int NS,
                                           int sid) {
  const int eh = lane >> 4, q = lane & 15;
  const uint32_t nslot = (uint32_t)p.v_nslot, slotb = (uint32_t)p.v_slot;
  const double LN2 = 0.693147180559945309417232121458;
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  uint32_t gslot = 0, gph = 0;   // ring position of the current task's first stage
  TaskInfoF ti;
  PROF_DECL
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool same = tt.flags & TF_TRI;
    const bool extras_t = (tt.flags & TF_EXTRAS) != 0;
    const bool jq = 4 * q < tt.nj;
    const bool kq = same ? jq : (4 * q < tt.nk);
    const int row0 = !(tt.flags & TF_ROW0) ? 0 : ((tt.flags & TF_SYM) ? 1 : (same ? 0 : 2));
    VFac R;                      // running products of this lane (mantissas in [1, 2) after every stage)
    int X[VQ][4];                // their exponents
#pragma unroll
    for (int a = 0; a < VQ; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) { R.q[a][i] = 1.0f; X[a][i] = 0; }
    // ring position of this warp's first stage of the task
    uint32_t slot = (gslot + (uint32_t)sid) % nslot, ph = gph ^ (gslot + (uint32_t)sid >= nslot ? 1u : 0u);
    for (int it = sid; it < ti.nst; it += V_NSW) {
      const bool aff = it < ti.nstA;
      const int lastA = ti.nstA - 1, lastG = ti.nst - 1;
      const int nvalid = it == lastA ? ti.na - lastA * NB : (it == lastG ? (int)(ti.e1 - ti.e0) - ti.na - (lastG - ti.nstA) * NB : NB);
      PROF_WAIT(1, mbar_wait_a(full0 + slot * 8u, ph));
      // a stage multiplies <= 4 entry factors, each >= ~2^-21, into a mantissa in [1, 2): no underflow before the peel
      PROF_WAIT(5, v2_singles_dispatch(same, extras_t, row0, ring + (size_t)slot * slotb, aff, eh, q, nvalid, jq, kq, R));
      __syncwarp();
      if (DMX_ARRIVE_ALL || lane == 0) mbar_arrive_a(empty0 + slot * 8u);
#pragma unroll
      for (int a = 0; a < VQ; ++a) {
        if ((a == 4 && !row0) || (a == 5 && same) || ((a == 0 || a == 3) && !extras_t)) continue;   // factors are all 1
#pragma unroll
        for (int i = 0; i < 4; ++i) peelf_unbiased(R.q[a][i], X[a][i]);   // exact zeros (padding samples) stay zero
      }
      slot += V_NSW;
      if (slot >= nslot) { slot -= nslot; ph ^= 1u; }
    }
    {   // ring position of the next task's first stage
      const uint32_t adv = gslot + (uint32_t)ti.nst;
      gph ^= (adv / nslot) & 1u;
      gslot = adv % nslot;
    }
    // logs of this lane's products; the two entry halves of the warp, then the three warps, in a fixed order
    double v[VQ][4];
#pragma unroll
    for (int a = 0; a < VQ; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double x = (double)X[a][i] * LN2 + (double)logf(R.q[a][i]);
        v[a][i] = x + __shfl_down_sync(0xffffffffu, x, 16);   // lanes 0..15: entry half 0 + entry half 1
      }
    if (sid != 0 && eh == 0) {
#pragma unroll
      for (int a = 0; a < VQ; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i) ts->comb[sid - 1][a][4 * q + i] = v[a][i];
    }
    consumer_bar_sync<V_NSW * 32>(2);
    if (sid == 0) {
      const int par = r & 1;
      if (eh == 0) {
#pragma unroll
        for (int a = 0; a < VQ; ++a)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int w = 0; w < V_NSW - 1; ++w) v[a][i] += ts->comb[w][a][4 * q + i];
      }
      mbar_wait(&lb_empty[par], ((uint32_t)(r >> 1) & 1u) ^ 1u);
      if (eh == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ts->lb[par][1][4 * q + i] = v[1][i];
          ts->lb[par][0][4 * q + i] = same ? v[1][i] : v[5][i];
        }
      }
      __syncwarp();
      if (DMX_ARRIVE_ALL || lane == 0) mbar_arrive(&lb_full[par]);
      if (eh == 0) {
        // per-sample results: the k-side samples of tasks with TF_EXTRAS; pairs with sample 0 where the task holds them
        const double lb0 = row0 == 2 ? ts->lb[par][1][0] : ts->lb[par][0][0];   // sample 0's share of a pair with it
        const double lmn = row0 ? __ldg(p.lm + (int64_t)ti.cell * p.nA + tt.n) : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int smp = 4 * q + i;
          if (extras_t && smp < tt.nk && tt.k0 + smp < p.nv) {
            const int64_t o = (int64_t)ti.cell * p.nv + tt.k0 + smp;
            if (p.sep0) p.singA[o] = v[0][i] + v[1][i];
            p.lb[o] = v[1][i] + v[2][i];
            if (p.want_single) p.p1[o] = v[3][i] + v[1][i];
          }
          if (row0) {
            const bool on_j = row0 == 2;
            const int gs = (on_j ? tt.j0 : tt.k0) + smp;
            if (smp < (on_j ? tt.nj : tt.nk) && gs < p.nv)
              p.row0[((int64_t)ti.cell * p.nA + tt.n) * p.nv + gs] = v[4][i] + lmn + lb0 + (on_j ? (same ? v[1][i] : v[5][i]) : v[1][i]);
          }
        }
      }
    }
    consumer_bar_sync<V_NSW * 32>(2);     // comb may be overwritten by the next task now
  }
  PROF_END(1, lane)
}

// =============================== the kernel =========================================================
__global__ void __launch_bounds__(F_LAUNCH_THREADS, 1) pair_scan_v2_kernel(const __grid_constant__ ScanParams p,
                                                                         const __grid_constant__ ScanMaps tm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslot = p.v_nslot;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + V_NSLOT_TRI;
  uint64_t* lb_full = empty + V_NSLOT_TRI;
  uint64_t* lb_empty = lb_full + 2;
  Scratch* scr = reinterpret_cast<Scratch*>(smem + SCRATCH_OFF);
  unsigned char* ring = smem + SLOT_CTRL;
  VScratch* ts = reinterpret_cast<VScratch*>(ring + (size_t)nslot * p.v_slot);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nslot; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (V_NCW + 1) * V_ARR);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&lb_full[s], V_ARR);
      mbar_init(&lb_empty[s], V_NCW * V_ARR);
    }
    scr->ncand = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int NS = (int)gridDim.x, u = (int)blockIdx.x;
  if (warp >= 12) {   // the last warpgroup gives registers to the other three (consumers: 64 accumulators + operands)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F_REGS_LO));
    if (warp >= V_W_LOAD0) v2_loader(p, tm, ring, full, empty, lane, u, NS, warp - V_W_LOAD0);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F_REGS_HI));
  if (warp >= V_W_SING0) {
    v2_singles(p, ring, full, empty, ts, lb_full, lb_empty, lane, u, NS, warp - V_W_SING0);
    return;
  }
  // =============================== consumer warps ============================================
  const int rw = warp;
  const uint32_t el = (uint32_t)lane & 7u;
  const int pos = rw * 4 + (lane >> 3);
  const uint32_t slotb = (uint32_t)p.v_slot;
  uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), ring_a = smem_u32(ring);
  int lane_k = lane;
  V_KEEP(full0); V_KEEP(empty0); V_KEEP(ring_a); V_KEEP(lane_k);
  uint32_t slot = 0, ph = 0;       // ring position of the next stage (continues across tasks)
  bool ready = false;              // the next stage's barrier has been seen complete
  TaskInfoF ti;
  PROF_DECL
  for (int r = 0; next_task_f32(p, r, u, NS, &ti); ++r) {
    const TaskType tt = p.tt[ti.tti];
    const bool tri = tt.flags & TF_TRI;
    const int nbj = (tt.nj + 7) >> 3, nbk = (tt.nk + 7) >> 3;
    int bj, bk;
    const bool diag = tri && pos >= 28;
    if (tri) {   // tile positions 0..27: off-diagonal tiles (bj < bk, row-major), 28..35: the diagonal tiles
      if (diag) {
        bj = bk = pos - 28;
      } else {
        int q = pos;
        bj = 0;
        while (q >= 7 - bj) { q -= 7 - bj; ++bj; }
        bk = bj + 1 + q;
      }
    } else {
      bj = pos >> 2;
      bk = pos & 3;
    }
    const bool active = (tri || pos < 32) && bj < nbj && bk < nbk;
    // operand addresses inside a stage
    const uint32_t xoff = V_OFF_DJ + vrow(el, V_DPJ) + (uint32_t)bj * 32u;
    const uint32_t yoff = tri ? V_OFF_DJ + vrow(el, V_DPJ) + (uint32_t)bk * 32u : V_OFF_DK + vrow(el, V_DPK) + (uint32_t)bk * 32u;
    const uint32_t aoff = V_OFF_GJ + vrow(el, V_GPJ) + 16u + (uint32_t)bj * 96u;
    const uint32_t goff = tri ? V_OFF_GJ + vrow(el, V_GPJ) + 16u + (uint32_t)bk * 96u : V_OFF_GK + vrow(el, V_GPK) + 16u + (uint32_t)bk * 96u;
    const uint32_t poff = V_OFF_PLANE + el * (FPLANE * 4);
    float acc[64];
#pragma unroll
    for (int h = 0; h < 64; ++h) acc[h] = (diag && (h & 7) < (h >> 3)) ? 0.0f : __uint_as_float((uint32_t)(V_TOP + 127) << 23);
    int ex = -V_TOP;                          // true products = acc * 2^ex
    // stages with fewer than NB entries: the last affine and the last general one
    const int lastA = ti.nstA - 1, nvA = ti.na - lastA * NB;
    const int lastG = ti.nst - 1, nvG = (int)(ti.e1 - ti.e0) - ti.na - (lastG - ti.nstA) * NB;
    const bool act = active && !(p.dbg & 1);
    const unsigned char* rs;
    uint32_t slot2, ph2;
    // Few taken branches per stage (every one costs an instruction-fetch bubble): the affine stages and the general
    // stages are loops of their own, in groups of V_PEEL stages with the range check between groups; entry lanes beyond
    // the last entry of an affine stage, and threads without a tile, multiply by an exact 1 instead of branching.
    auto begin_stage = [&]() {
      if (!ready) { PROF_WAIT(1, mbar_wait_a(full0 + slot * 8u, ph)); }
      rs = ring + (size_t)slot * slotb;
      slot2 = slot + 1; ph2 = ph;
      if (slot2 == (uint32_t)nslot) { slot2 = 0; ph2 ^= 1u; }
      ready = mbar_test_a(full0 + slot2 * 8u, ph2);
    };
    auto end_stage = [&]() {
      __syncwarp();
      if (DMX_ARRIVE_ALL || lane_k == 0) mbar_arrive_a(empty0 + slot * 8u);
      slot = slot2;
      ph = ph2;
    };
    auto affine_stages = [&](auto diag_c) {
      constexpr bool DIAG = decltype(diag_c)::value;
      // shared-window addresses of this thread's operands inside slot 0, pinned to registers
      uint32_t ax = ring_a + xoff, ay = ring_a + yoff, ap = ring_a + poff + 16u;
      int nlive = act ? ((int)el < nvA ? 2 : ((int)el < NB ? 1 : 0)) : 0;   // 2: live in every stage, 1: not in the last one
      V_KEEP(ax); V_KEEP(ay); V_KEEP(ap); V_KEEP(nlive);
      for (int it = 0; it < ti.nstA;) {
        const int lim = min(ti.nstA, it + V_PEEL);
#pragma unroll 1
        for (; it < lim; ++it) {
          if (!ready) { PROF_WAIT(1, mbar_wait_a(full0 + slot * 8u, ph)); }
          const uint32_t so = slot * slotb;
          slot2 = slot + 1; ph2 = ph;
          if (slot2 == (uint32_t)nslot) { slot2 = 0; ph2 ^= 1u; }
          ready = mbar_test_a(full0 + slot2 * 8u, ph2);
          const bool live = nlive > (it == lastA ? 1 : 0);
          float4 pl = lds128(ap + so);   // alpha bl bm
          const float4 x0 = lds128(ax + so), x1 = lds128(ax + so + 16u), y0 = lds128(ay + so), y1 = lds128(ay + so + 16u);
          pl.x = live ? pl.x : 1.0f; pl.y = live ? pl.y : 0.0f; pl.z = live ? pl.z : 0.0f;
          float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { x[i] = fmaf(pl.y, x[i], pl.x); y[i] = pl.z * y[i]; }
          v2_tile_aff<DIAG>(x, y, acc);
          end_stage();
        }
        v2_range_check<false>(acc, ex, active);
      }
    };
    if (diag) affine_stages(std::true_type{});
    else affine_stages(std::false_type{});
    for (int it = ti.nstA; it < ti.nst;) {
      const int lim = min(ti.nst, it + V_PEEL);
#pragma unroll 1
      for (; it < lim; ++it) {
        begin_stage();
        if (act && (int)el < (it == lastG ? nvG : NB)) {
          float w[8][3];
          {
            const float4 qa = *reinterpret_cast<const float4*>(rs + poff), qb = *reinterpret_cast<const float4*>(rs + poff + 16);
            const float q22 = *reinterpret_cast<const float*>(rs + poff + 32);
            float g[8][3];
            ld24(rs + goff, g);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              w[c][0] = fmaf(qa.z, g[c][2], fmaf(qa.y, g[c][1], qa.x * g[c][0]));
              w[c][1] = fmaf(qb.y, g[c][2], fmaf(qb.x, g[c][1], qa.w * g[c][0]));
              w[c][2] = fmaf(q22, g[c][2], fmaf(qb.w, g[c][1], qb.z * g[c][0]));
            }
          }
          float a[8][3];
          ld24(rs + aoff, a);
          v2_tile_gen<false>(a, w, acc);
        }
        end_stage();
      }
      v2_range_check<false>(acc, ex, active);
    }
    // ---- fold the eight entry lanes of each tile (reduce-scatter over lane bits 2, 1, 0): lane e ends up with the
    // products of tile row e.  Values of different lanes are multiplied in fp64 (no range problem).
    double d3[8];
    const bool b2 = el & 4, b1 = el & 2, b0 = el & 1;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {          // columns 4 cg .. 4 cg + 3
      double d1[16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float lo = acc[i * 8 + cg * 4 + c], hi = acc[(i + 4) * 8 + cg * 4 + c];
          const double keep = (double)(b2 ? hi : lo), send = (double)(b2 ? lo : hi);
          d1[i * 4 + c] = keep * __shfl_xor_sync(0xffffffffu, send, 4);
        }
      double d2[8];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double lo = d1[i * 4 + c], hi = d1[(i + 2) * 4 + c];
          d2[i * 4 + c] = (b1 ? hi : lo) * __shfl_xor_sync(0xffffffffu, b1 ? lo : hi, 2);
        }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double lo = d2[c], hi = d2[4 + c];
        d3[cg * 4 + c] = (b0 ? hi : lo) * __shfl_xor_sync(0xffffffffu, b0 ? lo : hi, 1);
      }
    }
    int exs = ex;
    exs += __shfl_xor_sync(0xffffffffu, exs, 4);
    exs += __shfl_xor_sync(0xffffffffu, exs, 2);
    exs += __shfl_xor_sync(0xffffffffu, exs, 1);
    // what the affine entries left out: the row sums of the normalisation (lb[j] + lb[k], from the singles warps) and
    // the per-entry scale M (sum of logs per cell and plane, cell_scalars_kernel)
    const int par = r & 1;
    const double lmn = __ldg(p.lm + (int64_t)ti.cell * p.nA + tt.n);
    PROF_WAIT(2, mbar_wait(&lb_full[par], (uint32_t)(r >> 1) & 1u));
    const double lbj = ts->lb[par][0][bj * 8 + el] + lmn;
    double v[8];
    int vj[8], vk[8];
    uint32_t vmask = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = tt.j0 + bj * 8 + (int)el, k = tt.k0 + bk * 8 + c;
      const bool valid = active && j < p.nv && k < p.nv && (!tri || j <= k);
      v[c] = valid ? (double)exs * 0.693147180559945309417232121458 + log_pos_fast(d3[c]) + (lbj + ts->lb[par][1][bk * 8 + c])
                   : -INFINITY;
      vj[c] = j;
      vk[c] = k;
      if (valid) vmask |= 1u << c;
    }
    __syncwarp();
    if (DMX_ARRIVE_ALL || lane == 0) mbar_arrive(&lb_empty[par]);
    PROF_WAIT(3, (task_epilogue<8, V_NCW, false>(p, ti, tt, scr, 0, rw, lane, v, vj, vk, vmask)));
  }
  PROF_END(3, lane)
}

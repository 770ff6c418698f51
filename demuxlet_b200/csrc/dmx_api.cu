// dmx_api.cu -- the C ABI of include/dmx.h: context, uploads, kernel orchestration, read-back,
// tie finaliser threads and the TSV writers.  No CPU fallback: without an sm_100 device dmx_create
// fails with DMX_E_NODEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dmx_host.h"
#include "dmx_internal.h"

using namespace dmx;

namespace {
thread_local std::string g_last_error;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool view = false;   // p belongs to another buffer (a shard's slice of the sharded context's result arrays)
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    cudaError_t e = cudaMallocHost((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void adopt(T* ptr, size_t n) {
    release();
    p = ptr;
    cap = n;
    view = true;
  }
  void release() {
    if (p && !view) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    view = false;
  }
};
}  // namespace

namespace {
class WorkerPool;
}

struct Chunk {
  int64_t c0, c1;   // cell ids [c0, c1)
  int64_t o0, o1;   // its active cells, largest first: order[o0, o1)
  int idx = 0;           // position in its schedule
  bool packed = false;   // this run sent the chunk's read counts (1 B per entry) instead of ent_ptr (8 B): the kernels of
                         // the chunk start with entptr_kernel
};

struct dmx_ctx {
  dmx_config cfg;
  std::vector<double> alpha;
  PGrid pg;
  int UQ = 0;
  int dense_idx[DMX_MAX_ALPHA];
  std::vector<TaskType> tt;
  double err256[256];
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;   // kernels; host -> device copies of dmx_run;
                                                                      // read back of finished chunks
  std::vector<cudaEvent_t> cev;     // per-chunk events (run_path)
  std::vector<Chunk> chunks;        // contiguous cell-id ranges, balanced by entries (dmx_set_pileup): one chunk for
                                    // dmx_compute (inputs resident) ...
  std::vector<Chunk> chunks_p;      // ... several for dmx_run, whose host -> device copies overlap the kernels
  std::vector<int32_t> order_p;     // schedule of chunks_p
  WorkerPool* pool = nullptr;       // host threads of the finaliser
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // inputs (host): borrowed or owned
  int nvp = 0;
  int64_t nsnps = 0, ncells = 0, E = 0, R = 0;
  const float* h_gp = nullptr;
  const int64_t* h_cell_ptr = nullptr;
  const int32_t* h_ent_snp = nullptr;
  const int64_t* h_ent_ptr = nullptr;
  const uint16_t* h_reads = nullptr;
  PinBuf<float> o_gp;
  PinBuf<int64_t> o_cell_ptr, o_ent_ptr;
  PinBuf<int32_t> o_ent_snp;
  PinBuf<uint16_t> o_reads;
  std::vector<int32_t> totl, npass, uniq;
  bool have_gp = false, have_plp = false, gp_dirty = true, plp_dirty = true, uploaded = false, computed = false;
  std::vector<int64_t> sched_cell_ptr;   // cell_ptr the current schedule was built for
  int gptab_state = 0;              // tables derived from the device GP matrix: 0 stale, 1 GP rows, 2 + dosage / row-sum tables

  // device
  DevBuf<float> d_gp;              // uploaded [nsnps][nvp*3]
  DevBuf<float> d_gprow;           // [nsnps][8 + nvp*3]: gp0 | samples (gprow_kernel)
  DevBuf<float> d_drow, d_srow;    // second-generation fp32 scan: dosage rows [2 nsnps][nvp + 8], row sums [nsnps][nvp + 8]
  PinBuf<uint8_t> h_cnt;           // reads per entry, staged by dmx_run for the upload (pack_chunk)
  PinBuf<int64_t> h_cell_r;        // ent_ptr at the first entry of every cell
  DevBuf<uint8_t> d_cnt;
  DevBuf<int64_t> d_cell_r;
  DevBuf<double> d_lm;             // [ncells][nA] sum of log M over the affine entries (cell_scalars_kernel)
  int scan_gen = 2;                // fp32 scan generation (DMX_SCAN_GEN=1 selects pair_scan_f32_kernel)
  DevBuf<double> d_entq, d_row0, d_singA, d_lb, d_p1, d_p1_0, d_l00, d_sing, d_dense, d_dense2;
  DevBuf<double> d_terms;          // [E][RESC_FAST] log terms of the two-phase exact re-scorer
  DevBuf<double> d_ents;           // [E][2 + nd] per-entry factors of llk0s / llks00 (entq -> cell_scalars)
  DevBuf<float> d_entf_head, d_entf_plane;   // fp32 copies of the entry records (fp32 pair scan), class-sorted per cell
  DevBuf<int32_t> d_perm, d_snp_perm, d_cell_na;   // entry -> position, SNP id by position, affine entries per cell
  DevBuf<double> d_memo_q;         // memo tables of the affine entries (build_memo)
  DevBuf<float> d_memo_fh, d_memo_fp;
  DevBuf<uint16_t> d_ecode;        // memo row per entry (0xffff: general entry with its own fp64 record)
  DevBuf<int32_t> d_glist;         // entry by class-sorted position (sign bit: general entry)
  bool memo_ready = false;
  bool entq_full = false;          // the fp64 entry table holds every entry (fp32 mode: only after ensure_entq_full)
  double err3_256[256];
  DevBuf<int64_t> d_cell_ptr, d_ent_ptr;
  DevBuf<int32_t> d_ent_snp, d_order;
  DevBuf<uint16_t> d_reads;
  DevBuf<TaskType> d_tt;
  DevBuf<Part> d_parts;
  DevBuf<CellOut> d_out;
  DevBuf<RescList> d_lists;
  PinBuf<RescList> r_lists;
  DevBuf<dmx_best_rec> d_best;      // finalise_kernel: records of the cells it could decide ...
  DevBuf<uint8_t> d_flag;           // ... and 1 for the cells left to the host finaliser
  DevBuf<int32_t> d_any_long;       // per chunk: some re-score list is longer than one pass
  PinBuf<uint8_t> r_flag;
  bool gpu_exact = false;          // device log() verified bit-identical to the host libm

  // schedule
  std::vector<int32_t> order;       // active cells, largest first
  std::vector<int32_t> status;      // per cell: 0 active, 1 filtered, 2 empty

  // read-back (pinned)
  PinBuf<CellOut> r_out;
  PinBuf<double> r_sing, r_row0, r_l00, r_p1, r_p1_0, r_dense, r_dense2;

  PinBuf<dmx_best_rec> best;        // [ncells] the records (pinned: finalise_kernel's records are copied straight into it)
  dmx_stats st;
  ScanMaps maps;                    // TMA descriptors of d_gprow / d_entq (encode_maps)

  // sharded context (dmx_create_sharded): one child context per device over the barcode range
  // [kid_lo[i], kid_lo[i + 1]); the children's per-barcode result arrays are views into this context's
  std::vector<dmx_ctx*> kids;
  std::vector<int64_t> kid_lo;
};

namespace {

int fail_cuda(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return DMX_E_CUDA;
}
#define CU(x)                                                  \
  do {                                                         \
    cudaError_t e__ = (cudaError_t)(x);                        \
    if (e__ != cudaSuccess) return fail_cuda(e__, #x);         \
  } while (0)
#define RC(x)                                                  \
  do {                                                         \
    const int rc__ = (x);                                      \
    if (rc__ != DMX_OK) return rc__;                           \
  } while (0)

// Upload the host libm's log table and check the device transcription of glibc's log() against the
// host's own log() on a large sample of the values the path produces (S in (1e-7, 1.1]) plus every
// table-interval boundary.  Only a 100 % bit-exact match enables the device re-scorer.
int verify_device_log(dmx_ctx* c) {
  c->gpu_exact = false;
  double tab[18 + 256];
  if (!find_libm_log_table(tab)) return DMX_OK;
  CU(set_log_table(tab));
  const int64_t N = 1 << 19;
  std::vector<double> x(N), y(N);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    return (double)(st >> 11) * (1.0 / 9007199254740992.0);
  };
  int64_t i = 0;
  for (int t = 0; t < 128; ++t)               // interval boundaries of the 128-entry table, +- a few ulps
    for (int d = -4; d <= 4; ++d) {
      uint64_t bits = 0x3fe6000000000000ull + ((uint64_t)t << 45) + (uint64_t)(int64_t)d;
      double v;
      memcpy(&v, &bits, 8);
      x[i++] = v;
      x[i++] = v * 0.0009765625;
    }
  for (int d = -2000; d <= 2000; ++d) {       // both edges of the near-1 branch
    uint64_t b0 = 0x3fee000000000000ull + (uint64_t)(int64_t)d, b1 = 0x3ff1090000000000ull + (uint64_t)(int64_t)d,
             b2 = 0x3ff0000000000000ull + (uint64_t)(int64_t)d;
    memcpy(&x[i++], &b0, 8); memcpy(&x[i++], &b1, 8); memcpy(&x[i++], &b2, 8);
  }
  for (; i < N; ++i) {
    const int kind = (int)(i & 3);
    if (kind == 0) x[i] = exp(-16.2 * rnd());                 // log-uniform down to 1e-7
    else if (kind == 1) x[i] = 0.9 + 0.2 * rnd();             // around 1
    else if (kind == 2) x[i] = 1.0 - 1e-5 * rnd() + 3e-7;     // the uninformative-entry cluster
    else x[i] = 1e-6 * (1.0 + 50.0 * rnd());                  // the floor cluster
  }
  DevBuf<double> dx, dy;
  CU(dx.reserve(N)); CU(dy.reserve(N));
  CU(cudaMemcpyAsync(dx.p, x.data(), N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(launch_logtest(dx.p, dy.p, N, c->stream));
  CU(cudaMemcpyAsync(y.data(), dy.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  dx.release(); dy.release();
  bool same = true;
  for (int64_t q = 0; q < N && same; ++q) {
    volatile double xv = x[q];
    const double h = log(xv);
    if (memcmp(&h, &y[q], 8) != 0) same = false;
  }
  c->gpu_exact = same;
  return DMX_OK;
}

// ---- TMA descriptors ----------------------------------------------------------------------------------
// cuTensorMapEncodeTiled lives in libcuda; it is looked up through the runtime so the library needs no -lcuda.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encoder() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
  }
  return fn;
}

int encode_2d(CUtensorMap* m, CUtensorMapDataType dt, size_t esize, void* base, uint64_t cols, uint64_t rows,
              uint32_t box_cols, uint32_t box_rows) {
  encode_tiled_fn enc = get_encoder();
  if (!enc) {
    g_last_error = "cuTensorMapEncodeTiled not available";
    return DMX_E_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * esize};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return DMX_E_CUDA;
  }
  return DMX_OK;
}

// descriptors of the pair scan's inputs; call after d_gprow / d_entq have their final addresses
int encode_maps(dmx_ctx* c) {
  memset(&c->maps, 0, sizeof(c->maps));
  if (c->nsnps == 0 || c->E == 0) return DMX_OK;   // nothing is launched
  const uint64_t rowlen = (uint64_t)c->nvp * 3 + 8;
  int rc = encode_2d(&c->maps.gpj, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_gprow.p, rowlen, (uint64_t)c->nsnps,
                     8 + BLK * 3, 1);
  if (rc == DMX_OK)
    rc = encode_2d(&c->maps.gpk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_gprow.p, rowlen, (uint64_t)c->nsnps,
                   8 + BLK_RECT_K * 3, 1);
  if (rc == DMX_OK && c->UQ == QHEAD + QPLANE)
    rc = encode_2d(&c->maps.q, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, c->d_entq.p, (uint64_t)c->UQ, (uint64_t)c->E,
                   QHEAD + QPLANE, NB);
  if (rc == DMX_OK)
    rc = encode_2d(&c->maps.qh, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, c->d_entq.p, (uint64_t)c->UQ, (uint64_t)c->E,
                   QHEAD, NB);
  if (rc == DMX_OK)
    rc = encode_2d(&c->maps.qp, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, c->d_entq.p, (uint64_t)c->UQ, (uint64_t)c->E,
                   QPLANE, NB);
  if (c->d_drow.p && c->d_srow.p) {
    const uint64_t dp = (uint64_t)c->nvp + 8;
    if (rc == DMX_OK) rc = encode_2d(&c->maps.dj, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_drow.p, dp, 2 * (uint64_t)c->nsnps, BLK + 8, 1);
    if (rc == DMX_OK) rc = encode_2d(&c->maps.dk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_drow.p, dp, 2 * (uint64_t)c->nsnps, BLK_RECT_K + 8, 1);
    if (rc == DMX_OK) rc = encode_2d(&c->maps.sj, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_srow.p, dp, (uint64_t)c->nsnps, BLK + 8, 1);
    if (rc == DMX_OK) rc = encode_2d(&c->maps.sk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->d_srow.p, dp, (uint64_t)c->nsnps, BLK_RECT_K + 8, 1);
  }
  return rc;
}

void build_tasks(dmx_ctx* c) {
  const int nA = c->cfg.n_alpha, nvp = c->nvp;
  const PGrid& g = c->pg;
  c->tt.clear();
  int nd = 0;
  for (int n = 0; n < nA; ++n) c->dense_idx[n] = (g.mode[n] == PM_SEP) ? -1 : nd++;
  c->UQ = QHEAD + QPLANE * nd;
  int fd = -1;
  for (int n = 0; n < nA; ++n)
    if (g.mode[n] != PM_SEP) { fd = n; break; }
  const int nblk = (nvp + BLK - 1) / BLK;
  for (int n = 0; n < nA; ++n) {
    if (g.mode[n] == PM_SEP) continue;
    const bool sym = g.mode[n] == PM_SYM;
    bool first = true;
    auto push = [&](int j0, int nj, int k0, int nk, uint32_t fl) {
      TaskType t;
      memset(&t, 0, sizeof(t));
      t.n = (int16_t)n;
      t.j0 = (int16_t)j0; t.nj = (int16_t)nj; t.k0 = (int16_t)k0; t.nk = (int16_t)nk;
      t.flags = fl | (sym ? TF_SYM : 0u) | (n == 0 ? TF_NOARG : 0u) | ((sym ? j0 == 0 : k0 == 0) ? TF_ROW0 : 0u);
      if (first) {
        t.flags |= TF_L00;
        if (n == fd) t.flags |= TF_SCALARS;
        first = false;
      }
      t.qoff = (int16_t)(QHEAD + QPLANE * c->dense_idx[n]);
      c->tt.push_back(t);
    };
    for (int jb = 0; jb < nblk; ++jb) {
      const int j0 = jb * BLK, nj = std::min(BLK, nvp - j0);
      if (sym) {
        push(j0, nj, j0, nj, TF_TRI | (n == fd ? TF_EXTRAS : 0u));
        for (int k0 = j0 + BLK; k0 < nvp; k0 += BLK_RECT_K) push(j0, nj, k0, std::min(BLK_RECT_K, nvp - k0), 0u);
      } else {
        for (int k0 = 0; k0 < nvp; k0 += BLK_RECT_K)
          push(j0, nj, k0, std::min(BLK_RECT_K, nvp - k0), (n == fd && jb == 0) ? TF_EXTRAS : 0u);
      }
    }
  }
}

// Persistent host threads for the finaliser (a run calls parallel_for once per chunk: creating threads each time
// would cost more than the work).
class WorkerPool {
 public:
  ~WorkerPool() { stop(); }
  void ensure(int nthreads) {
    if ((int)th_.size() == nthreads - 1) return;
    stop();
    quit_ = false;
    for (int t = 1; t < nthreads; ++t) th_.emplace_back([this, t]() { loop(t); });
  }
  // f(i, tid) for i in [0, n); the caller is worker 0
  void run(int64_t n, int nthreads, const std::function<void(int64_t, int)>& f) {
    if (nthreads <= 1 || n < 2) {
      for (int64_t i = 0; i < n; ++i) f(i, 0);
      return;
    }
    ensure(nthreads);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &f; n_ = n; next_.store(0);
      chunk_ = std::max<int64_t>(1, n / (nthreads * 16));
      pending_ = (int)th_.size();
      ++gen_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this]() { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work(int tid) {
    for (;;) {
      const int64_t lo = next_.fetch_add(chunk_);
      if (lo >= n_) break;
      const int64_t hi = std::min(n_, lo + chunk_);
      for (int64_t i = lo; i < hi; ++i) (*fn_)(i, tid);
    }
  }
  void loop(int tid) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_;
      }
      work(tid);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
    th_.clear();
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t, int)>* fn_ = nullptr;
  std::atomic<int64_t> next_{0};
  int64_t n_ = 0, chunk_ = 1;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool quit_ = false;
};

int host_threads(const dmx_ctx* c) {
  int n = c->cfg.n_host_threads;
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  return std::max(1, n);
}

HostView host_view(const dmx_ctx* c) {
  HostView hv;
  hv.nv = c->cfg.nv; hv.nA = c->cfg.n_alpha; hv.alpha = c->alpha.data();
  hv.gp = c->h_gp; hv.nsnps = c->nsnps; hv.ncells = c->ncells;
  hv.cell_ptr = c->h_cell_ptr; hv.ent_snp = c->h_ent_snp; hv.ent_ptr = c->h_ent_ptr; hv.reads = c->h_reads;
  return hv;
}

void fill_scan_params(dmx_ctx* c, ScanParams* p) {
  memset(p, 0, sizeof(*p));
  p->nv = c->cfg.nv; p->nvp = c->nvp; p->nA = c->cfg.n_alpha; p->UQ = c->UQ;
  p->sep0 = c->pg.sep0; p->T = (int)c->tt.size();
  p->want_single = (c->cfg.flags & DMX_SKIP_SINGLE) ? 0 : 1;
  p->ntasks = (int64_t)c->order.size() * p->T;
  p->eps_rel = 1e-9;
  // fp32 pair scan: |error of a llk| <= entries * (a few fp32 ulps); 2^-20 per entry leaves a wide margin
  p->eps_ent = (c->cfg.flags & DMX_FP32_PAIRS) ? 0x1p-20 : 0.0;
  if (c->UQ == QHEAD + QPLANE) { p->q_hp = (QHEAD + QPLANE) * 8; p->q_po = QHEAD * 8; p->q_pp = (QHEAD + QPLANE) * 8; }
  else { p->q_hp = QHEAD * 8; p->q_po = NB * QHEAD * 8; p->q_pp = QPLANE * 8; }
  p->gp_stride = c->nvp * 3 + 8;
  p->dbg = getenv("DMX_DEBUG") ? atoi(getenv("DMX_DEBUG")) : 0;
  p->gp = c->d_gprow.p + 8; p->ent_snp = c->d_ent_snp.p; p->entq = c->d_entq.p;
  p->has_rect = 0;
  for (const TaskType& t : c->tt)
    if (!(t.flags & TF_TRI)) p->has_rect = 1;
  p->entf_head = c->d_entf_head.p; p->entf_plane = c->d_entf_plane.p; p->entf_pstride = (c->E + NB) * FPLANE;
  p->snp_perm = c->d_snp_perm.p; p->cell_na = c->d_cell_na.p;
  p->cell_ptr = c->d_cell_ptr.p; p->cell_order = c->d_order.p; p->tt = c->d_tt.p; p->parts = c->d_parts.p;
  p->row0 = c->d_row0.p; p->singA = c->d_singA.p; p->lb = c->d_lb.p; p->p1 = c->d_p1.p;
  p->p1_0 = c->d_p1_0.p; p->l00 = c->d_l00.p; p->dense = nullptr;
  p->lm = c->d_lm.p;
}

// candidate list of a cell -> ordered hypotheses in (j,k,n) lexicographic order
void expand_cands(const dmx_ctx* c, const CellOut& co, std::vector<Hyp>* out) {
  out->clear();
  const int nc = std::min<int>(co.ncand, MAX_CAND);
  for (int i = 0; i < nc; ++i) {
    const int j = co.cand[i] & 0xfff, k = (co.cand[i] >> 12) & 0xfff, n = co.cand[i] >> 24;
    out->push_back(Hyp{j, k, n});
    if (c->pg.mode[n] == PM_SYM) out->push_back(Hyp{k, j, n});
  }
  std::sort(out->begin(), out->end(), [](const Hyp& a, const Hyp& b) {
    return a.j != b.j ? a.j < b.j : (a.k != b.k ? a.k < b.k : a.n < b.n);
  });
}

int for_each_kid(dmx_ctx* c, const std::function<int(int)>& f);
int ensure_entq_full(dmx_ctx* c);

// dense llksAB for a list of active cells (device order buffer is temporarily replaced)
// Dense llksAB of a list of active cells: block i of the list lands at host_out + dest[i] * nv * nv * nA (dest null:
// i).  Two batches in flight: the fp64 scan of batch b runs while batch b - 1 comes back into pinned memory and the
// host threads move batch b - 2 ... b - 1 to its (usually pageable) destination.
int dense_scan(dmx_ctx* c, const std::vector<int32_t>& cells, double* host_out, const int64_t* dest = nullptr) {
  const int nv = c->cfg.nv, nA = c->cfg.n_alpha;
  const size_t per = (size_t)nv * nv * nA;
  if (cells.empty()) return DMX_OK;
  if (!c->kids.empty()) {
    // sharded context: every shard scans its own cells of the list (on its own device, in parallel)
    const int nk = (int)c->kids.size();
    std::vector<std::vector<int32_t>> mine(nk);
    std::vector<std::vector<int64_t>> at(nk);
    for (size_t i = 0; i < cells.size(); ++i) {
      const int k = (int)(std::upper_bound(c->kid_lo.begin(), c->kid_lo.end(), (int64_t)cells[i]) - c->kid_lo.begin()) - 1;
      mine[k].push_back((int32_t)(cells[i] - c->kid_lo[k]));
      at[k].push_back(dest ? dest[i] : (int64_t)i);
    }
    return for_each_kid(c, [&](int k) { return dense_scan(c->kids[k], mine[k], host_out, at[k].data()); });
  }
  CU(cudaSetDevice(c->cfg.device));
  RC(ensure_entq_full(c));
  const size_t batch = std::min(cells.size(), std::max<size_t>(1, ((size_t)256 << 20) / (per * sizeof(double))));
  const size_t nb = (cells.size() + batch - 1) / batch;
  const int depth = nb > 1 ? 2 : 1;
  DevBuf<int32_t> d_cells[2];
  DevBuf<Part> d_parts;
  DevBuf<double>& d_dense2 = c->d_dense2;   // second device / pinned buffers (kept by the context between calls)
  PinBuf<double>& pin2 = c->r_dense2;
  cudaEvent_t ev_scan[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  struct Rel {
    DevBuf<int32_t>* a; DevBuf<Part>* b; cudaEvent_t* e1; cudaEvent_t* e2;
    ~Rel() {
      a[0].release(); a[1].release(); b->release();
      for (int i = 0; i < 2; ++i) { if (e1[i]) cudaEventDestroy(e1[i]); if (e2[i]) cudaEventDestroy(e2[i]); }
    }
  } rel{d_cells, &d_parts, ev_scan, ev_copy};
  CU(d_parts.reserve(batch * c->tt.size()));
  CU(c->d_dense.reserve(batch * per)); CU(c->r_dense.reserve(batch * per));
  if (depth == 2) { CU(d_dense2.reserve(batch * per)); CU(pin2.reserve(batch * per)); }
  for (int i = 0; i < depth; ++i) {
    CU(d_cells[i].reserve(batch));
    CU(cudaEventCreateWithFlags(&ev_scan[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
  }
  double* dbuf[2] = {c->d_dense.p, d_dense2.p};
  double* pbuf[2] = {c->r_dense.p, pin2.p};
  const int nth = host_threads(c);
  auto scatter = [&](size_t b) -> int {      // batch b: pinned buffer -> destination blocks
    CU(cudaEventSynchronize(ev_copy[b % depth]));
    const size_t lo = b * batch, n = std::min(batch, cells.size() - lo);
    const double* src = pbuf[b % depth];
    c->pool->run((int64_t)n, nth, [&](int64_t i, int) {
      const int64_t d = dest ? dest[lo + (size_t)i] : (int64_t)(lo + (size_t)i);
      memcpy(host_out + (size_t)d * per, src + (size_t)i * per, per * sizeof(double));
    });
    return DMX_OK;
  };
  for (size_t b = 0; b < nb; ++b) {
    const int s = (int)(b % depth);
    const size_t lo = b * batch, n = std::min(batch, cells.size() - lo);
    if (b >= (size_t)depth) CU(cudaStreamWaitEvent(c->stream, ev_copy[s], 0));   // its device buffer has been read back
    CU(cudaMemcpyAsync(d_cells[s].p, cells.data() + lo, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    ScanParams p;
    fill_scan_params(c, &p);
    p.cell_order = d_cells[s].p;
    p.parts = d_parts.p;
    p.dense = dbuf[s];
    p.ntasks = (int64_t)n * p.T;
    // dense blocks always come from the fp64 scan: every hypothesis keeps its own exponent there, whereas the fp32
    // scan lets hypotheses that are hopeless against the best of their tile underflow
    CU(launch_pair_scan(p, c->maps, 0, c->stream));
    CU(launch_fill_sep_dense(p, (int64_t)n, c->stream));
    c->st.gpu_launches += 1 + (p.sep0 ? 1 : 0);
    CU(cudaEventRecord(ev_scan[s], c->stream));
    if (b >= 1) RC(scatter(b - 1));            // frees the pinned buffer batch b + 1 will use; overlaps scan b
    CU(cudaStreamWaitEvent(c->stream3, ev_scan[s], 0));
    CU(cudaMemcpyAsync(pbuf[s], dbuf[s], n * per * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
    CU(cudaEventRecord(ev_copy[s], c->stream3));
  }
  RC(scatter(nb - 1));
  CU(cudaStreamSynchronize(c->stream));
  return DMX_OK;
}


// ---- sharded contexts -----------------------------------------------------------------------------------------------
// f(shard index) on one host thread per shard (the caller's thread takes shard 0); first non-zero status wins
int for_each_kid(dmx_ctx* c, const std::function<int(int)>& f) {
  const int n = (int)c->kids.size();
  std::vector<int> rc(n, 0);
  std::vector<std::string> err(n);
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i)
    th.emplace_back([&, i]() {
      rc[i] = f(i);
      if (rc[i]) err[i] = g_last_error;   // thread-local text of the failing shard
    });
  if (n > 0) {
    rc[0] = f(0);
    if (rc[0]) err[0] = g_last_error;
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) {
      g_last_error = err[i];
      return rc[i];
    }
  return DMX_OK;
}

int sharded_set_gp(dmx_ctx* c, const float* gp, int64_t nsnps) {
  c->nsnps = nsnps;
  const size_t n = (size_t)nsnps * c->cfg.nv * 3;
  if (c->cfg.flags & DMX_BORROW_INPUTS) {
    c->h_gp = gp;
  } else {
    CU(c->o_gp.reserve(n));
    memcpy(c->o_gp.p, gp, n * sizeof(float));
    c->h_gp = c->o_gp.p;
  }
  for (dmx_ctx* k : c->kids) RC(dmx_set_gp(k, c->h_gp, nsnps));   // children borrow: one host copy, n device copies
  c->have_gp = true;
  c->computed = false;
  return DMX_OK;
}

int sharded_set_pileup(dmx_ctx* c, int64_t ncells, const int64_t* cell_ptr, const int32_t* ent_snp,
                       const int64_t* ent_ptr, const uint16_t* reads, const int32_t* totl, const int32_t* pass,
                       const int32_t* uniq) {
  const int nk = (int)c->kids.size();
  const int nv = c->cfg.nv, nA = c->cfg.n_alpha;
  const int64_t E = cell_ptr[ncells], R = ent_ptr[E];
  c->ncells = ncells; c->E = E; c->R = R;
  CU(c->o_cell_ptr.reserve(ncells + 1));
  memcpy(c->o_cell_ptr.p, cell_ptr, (ncells + 1) * sizeof(int64_t));
  c->h_cell_ptr = c->o_cell_ptr.p;
  c->totl.assign(totl, totl + ncells);
  c->npass.assign(pass, pass + ncells);
  c->uniq.assign(uniq, uniq + ncells);
  c->status.assign(ncells, 0);
  for (int64_t i = 0; i < ncells; ++i) {
    const int64_t ns = cell_ptr[i + 1] - cell_ptr[i];
    if (totl[i] < c->cfg.min_total || uniq[i] < c->cfg.min_uniq || ns < c->cfg.min_snp) c->status[i] = 1;
    else if (ns == 0) c->status[i] = 2;
  }
  // contiguous barcode ranges with (nearly) equal numbers of pileup entries
  c->kid_lo.assign(nk + 1, 0);
  for (int k = 1; k < nk; ++k) {
    const int64_t target = (int64_t)((double)E * k / nk);
    int64_t cut = std::lower_bound(cell_ptr, cell_ptr + ncells + 1, target) - cell_ptr;
    cut = std::min(std::max(cut, c->kid_lo[k - 1]), ncells);
    c->kid_lo[k] = cut;
  }
  c->kid_lo[nk] = ncells;
  // per-barcode result arrays of the whole job; each shard reads its records back into its own slice
  CU(c->r_p1.reserve((size_t)ncells * nv + nk + 1)); CU(c->r_p1_0.reserve(ncells + nk + 1));
  CU(c->r_sing.reserve((size_t)ncells * nv + nk + 1)); CU(c->r_l00.reserve((size_t)ncells * nA + nk + 1));
  const bool borrow = (c->cfg.flags & DMX_BORROW_INPUTS) != 0;
  RC(for_each_kid(c, [&](int k) {
    dmx_ctx* kid = c->kids[k];
    const int64_t lo = c->kid_lo[k], hi = c->kid_lo[k + 1], n = hi - lo;
    const int64_t e0 = cell_ptr[lo], e1 = cell_ptr[hi], r0 = ent_ptr[e0], r1 = ent_ptr[e1];
    CU(cudaSetDevice(kid->cfg.device));
    CU(kid->o_cell_ptr.reserve(n + 1)); CU(kid->o_ent_ptr.reserve(e1 - e0 + 1));
    for (int64_t i = 0; i <= n; ++i) kid->o_cell_ptr.p[i] = cell_ptr[lo + i] - e0;
    for (int64_t i = 0; i <= e1 - e0; ++i) kid->o_ent_ptr.p[i] = ent_ptr[e0 + i] - r0;
    const int32_t* snp = ent_snp + e0;
    const uint16_t* rd = reads + r0;
    if (!borrow) {
      CU(kid->o_ent_snp.reserve(e1 - e0)); CU(kid->o_reads.reserve(r1 - r0));
      if (e1 > e0) memcpy(kid->o_ent_snp.p, snp, (e1 - e0) * sizeof(int32_t));
      if (r1 > r0) memcpy(kid->o_reads.p, rd, (r1 - r0) * sizeof(uint16_t));
      snp = kid->o_ent_snp.p;
      rd = kid->o_reads.p;
    }
    RC(dmx_set_pileup(kid, n, kid->o_cell_ptr.p, snp, kid->o_ent_ptr.p, rd, totl + lo, pass + lo, uniq + lo));
    kid->r_p1.adopt(c->r_p1.p + lo * nv, (size_t)n * nv + 1);
    kid->r_p1_0.adopt(c->r_p1_0.p + lo, (size_t)n + 1);
    kid->r_sing.adopt(c->r_sing.p + lo * nv, (size_t)n * nv + 1);
    kid->r_l00.adopt(c->r_l00.p + lo * nA, (size_t)n * nA + 1);
    return (int)DMX_OK;
  }));
  c->have_plp = true;
  c->computed = false;
  return DMX_OK;
}

// mode 0: dmx_run, 1: dmx_upload, 2: dmx_compute
int sharded_run(dmx_ctx* c, int mode) {
  if (!c->have_gp || !c->have_plp) return DMX_E_STATE;
  const double t0 = now_ms();
  if (mode != 1) {
    CU(c->best.reserve((size_t)c->ncells + 1));
    if (c->ncells) memset(c->best.p, 0, (size_t)c->ncells * sizeof(dmx_best_rec));
  }
  RC(for_each_kid(c, [&](int k) {
    dmx_ctx* kid = c->kids[k];
    const int rc = mode == 0 ? dmx_run(kid) : (mode == 1 ? dmx_upload(kid) : dmx_compute(kid));
    if (rc != DMX_OK || mode == 1) return rc;
    if (kid->ncells) memcpy(c->best.p + c->kid_lo[k], kid->best.p, kid->ncells * sizeof(dmx_best_rec));
    return (int)DMX_OK;
  }));
  dmx_stats a;
  memset(&a, 0, sizeof(a));
  a.gpu_exact = 1;
  for (const dmx_ctx* kid : c->kids) {
    const dmx_stats& s = kid->st;
    a.h2d_ms = std::max(a.h2d_ms, s.h2d_ms); a.prep_ms = std::max(a.prep_ms, s.prep_ms);
    a.kernel_ms = std::max(a.kernel_ms, s.kernel_ms); a.d2h_ms = std::max(a.d2h_ms, s.d2h_ms);
    a.finalise_ms = std::max(a.finalise_ms, s.finalise_ms); a.pair_kernel_ms = std::max(a.pair_kernel_ms, s.pair_kernel_ms);
    a.h2d_bytes += s.h2d_bytes; a.d2h_bytes += s.d2h_bytes; a.n_entries += s.n_entries; a.n_reads += s.n_reads;
    a.n_tasks += s.n_tasks; a.gpu_launches += s.gpu_launches; a.n_rescored += s.n_rescored; a.n_fullcell += s.n_fullcell;
    a.n_rescored_gpu += s.n_rescored_gpu; a.algo_bytes += s.algo_bytes;
    a.gpu_exact = std::min(a.gpu_exact, s.gpu_exact);
  }
  a.total_ms = now_ms() - t0;
  if (mode == 1) {
    c->st.h2d_ms = a.h2d_ms;
    c->st.h2d_bytes = a.h2d_bytes;
    c->uploaded = true;
  } else {
    c->st = a;
    c->computed = true;
  }
  return DMX_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int dmx_abi_version(void) { return DMX_ABI_VERSION; }

const char* dmx_strerror(int s) {
  switch (s) {
    case DMX_OK: return "ok";
    case DMX_E_INVALID: return "invalid argument";
    case DMX_E_CUDA: return "CUDA error";
    case DMX_E_NODEVICE: return "no sm_100 CUDA device (libdmx has no CPU fallback)";
    case DMX_E_STATE: return "call order violated";
    case DMX_E_NOMEM: return "out of memory";
    case DMX_E_IO: return "cannot create output file";
    case DMX_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

const char* dmx_last_error(void) { return g_last_error.c_str(); }

void dmx_config_default(dmx_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->doublet_prior = 0.5;          // cmd_cram_demuxlet.cpp:29
  cfg->flags = DMX_EXACT_TIES;
}

int dmx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

void* dmx_alloc_host(int64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, (size_t)std::max<int64_t>(bytes, 1)) != cudaSuccess) return nullptr;
  return p;
}
void dmx_free_host(void* p) {
  if (p) cudaFreeHost(p);
}

int dmx_create(dmx_ctx** out, const dmx_config* cfg) {
  if (!out || !cfg || !cfg->alpha) return DMX_E_INVALID;
  *out = nullptr;
  // the reference indexes out of bounds for nv < 2 or n_alpha < 2
  if (cfg->nv < 2 || cfg->nv > 4095 || cfg->n_alpha < 2 || cfg->n_alpha > DMX_MAX_ALPHA) return DMX_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
    g_last_error = "no CUDA device";
    return DMX_E_NODEVICE;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device);
  if (major != 10) {
    g_last_error = "device is not compute capability 10.x";
    return DMX_E_NODEVICE;
  }
  CU(cudaSetDevice(cfg->device));
  dmx_ctx* c = new (std::nothrow) dmx_ctx();
  if (!c) return DMX_E_NOMEM;
  c->cfg = *cfg;
  c->pool = new WorkerPool();
  c->alpha.assign(cfg->alpha, cfg->alpha + cfg->n_alpha);
  c->cfg.alpha = c->alpha.data();
  memset(&c->st, 0, sizeof(c->st));
  build_pgrid(c->alpha.data(), cfg->n_alpha, &c->pg);
  c->nvp = (cfg->nv + TILE - 1) / TILE * TILE;
  build_tasks(c);
  for (int i = 0; i <= 255; ++i) c->err256[i] = (i > 1) ? pow(0.1, i * 0.1) : 0.75;   // PhredHelper.cpp:30
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking);
  for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
  for (int i = 0; i <= 255; ++i) c->err3_256[i] = c->err256[i] / 3.0;   // the quotient of :606-607
  if (e == cudaSuccess) e = (cudaError_t)set_err_table(c->err256, c->err3_256);
  if (e == cudaSuccess) e = (cudaError_t)pair_scan_init();
  if (e == cudaSuccess) e = c->d_tt.reserve(c->tt.size());
  if (e == cudaSuccess)
    e = cudaMemcpy(c->d_tt.p, c->tt.data(), c->tt.size() * sizeof(TaskType), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    dmx_destroy(c);
    return fail_cuda(e, "dmx_create");
  }
  if (cfg->flags & DMX_EXACT_TIES) {
    const int rc = verify_device_log(c);
    if (rc != DMX_OK) {
      dmx_destroy(c);
      return rc;
    }
  }
  *out = c;
  return DMX_OK;
}

int dmx_create_sharded(dmx_ctx** out, const dmx_config* cfg, int32_t n_devices, const int32_t* devices) {
  if (!out || !cfg || !cfg->alpha || n_devices < 1 || n_devices > 64 || !devices) return DMX_E_INVALID;
  *out = nullptr;
  if (cfg->nv < 2 || cfg->nv > 4095 || cfg->n_alpha < 2 || cfg->n_alpha > DMX_MAX_ALPHA) return DMX_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_last_error = "no CUDA device";
    return DMX_E_NODEVICE;
  }
  for (int i = 0; i < n_devices; ++i)
    if (devices[i] < 0 || devices[i] >= ndev) {
      g_last_error = "device ordinal out of range";
      return DMX_E_NODEVICE;
    }
  dmx_ctx* c = new (std::nothrow) dmx_ctx();
  if (!c) return DMX_E_NOMEM;
  c->cfg = *cfg;
  c->cfg.device = devices[0];
  c->pool = new WorkerPool();
  c->alpha.assign(cfg->alpha, cfg->alpha + cfg->n_alpha);
  c->cfg.alpha = c->alpha.data();
  memset(&c->st, 0, sizeof(c->st));
  build_pgrid(c->alpha.data(), cfg->n_alpha, &c->pg);
  c->nvp = (cfg->nv + TILE - 1) / TILE * TILE;
  c->kids.assign((size_t)n_devices, nullptr);
  int total = cfg->n_host_threads;
  if (total <= 0) total = (int)std::thread::hardware_concurrency();
  const int rc = for_each_kid(c, [&](int k) {
    dmx_config kc = *cfg;
    kc.alpha = c->alpha.data();
    kc.device = devices[k];
    kc.flags |= DMX_BORROW_INPUTS;                         // the sharded context owns (or borrows) the host inputs
    kc.n_host_threads = std::max(1, total / n_devices);    // the shards share the host's cores
    return dmx_create(&c->kids[(size_t)k], &kc);
  });
  if (rc != DMX_OK) {
    const std::string why = g_last_error;
    dmx_destroy(c);
    g_last_error = why;
    return rc;
  }
  *out = c;
  return DMX_OK;
}

int dmx_shard_count(const dmx_ctx* c) {
  if (!c) return 0;
  return c->kids.empty() ? 1 : (int)c->kids.size();
}

int dmx_shard_range(const dmx_ctx* c, int32_t shard, int64_t* cell_lo, int64_t* cell_hi, int32_t* device) {
  if (!c || shard < 0 || shard >= dmx_shard_count(c)) return DMX_E_INVALID;
  if (!c->have_plp) return DMX_E_STATE;
  if (c->kids.empty()) {
    if (cell_lo) *cell_lo = 0;
    if (cell_hi) *cell_hi = c->ncells;
    if (device) *device = c->cfg.device;
    return DMX_OK;
  }
  if (cell_lo) *cell_lo = c->kid_lo[(size_t)shard];
  if (cell_hi) *cell_hi = c->kid_lo[(size_t)shard + 1];
  if (device) *device = c->kids[(size_t)shard]->cfg.device;
  return DMX_OK;
}

int dmx_shard_stats(dmx_ctx* c, int32_t shard, dmx_stats* out) {
  if (!c || !out || shard < 0 || shard >= dmx_shard_count(c)) return DMX_E_INVALID;
  *out = c->kids.empty() ? c->st : c->kids[(size_t)shard]->st;
  return DMX_OK;
}

void dmx_destroy(dmx_ctx* c) {
  if (!c) return;
  for (dmx_ctx* k : c->kids) dmx_destroy(k);   // their result views point into this context's arrays
  c->kids.clear();
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->d_terms.release(); c->d_ents.release(); c->d_entf_head.release(); c->d_entf_plane.release();
  c->d_perm.release(); c->d_snp_perm.release(); c->d_cell_na.release();
  c->d_memo_q.release(); c->d_memo_fh.release(); c->d_memo_fp.release(); c->d_ecode.release();
  c->d_glist.release();
  c->d_drow.release(); c->d_srow.release(); c->d_lm.release();
  c->h_cnt.release(); c->h_cell_r.release(); c->d_cnt.release(); c->d_cell_r.release();
  c->d_gp.release(); c->d_gprow.release(); c->d_entq.release(); c->d_row0.release(); c->d_singA.release();
  c->d_lb.release(); c->d_p1.release(); c->d_p1_0.release(); c->d_l00.release(); c->d_sing.release();
  c->d_dense.release(); c->d_cell_ptr.release(); c->d_ent_ptr.release(); c->d_ent_snp.release();
  c->d_order.release(); c->d_reads.release(); c->d_tt.release(); c->d_parts.release(); c->d_out.release();
  c->d_lists.release(); c->r_lists.release();
  c->d_best.release(); c->d_flag.release(); c->best.release(); c->r_flag.release(); c->d_any_long.release();
  c->o_gp.release(); c->o_cell_ptr.release(); c->o_ent_ptr.release(); c->o_ent_snp.release(); c->o_reads.release();
  c->r_out.release(); c->r_sing.release(); c->r_row0.release(); c->r_l00.release(); c->r_p1.release();
  c->r_p1_0.release(); c->r_dense.release(); c->r_dense2.release(); c->d_dense2.release();
  for (int i = 0; i < 6; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  delete c->pool;
  for (cudaEvent_t e2 : c->cev)
    if (e2) cudaEventDestroy(e2);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int dmx_set_gp(dmx_ctx* c, const float* gp, int64_t nsnps) {
  if (!c || !gp || nsnps < 0) return DMX_E_INVALID;
  CU(cudaSetDevice(c->cfg.device));
  if (!c->kids.empty()) return sharded_set_gp(c, gp, nsnps);
  c->nsnps = nsnps;
  const size_t n = (size_t)nsnps * c->cfg.nv * 3;
  if (c->cfg.flags & DMX_BORROW_INPUTS) {
    c->h_gp = gp;
  } else {
    CU(c->o_gp.reserve(n));
    memcpy(c->o_gp.p, gp, n * sizeof(float));
    c->h_gp = c->o_gp.p;
  }
  c->have_gp = true;
  c->gp_dirty = true;
  c->gptab_state = 0;
  c->uploaded = false;
  c->computed = false;
  return DMX_OK;
}

int dmx_set_pileup(dmx_ctx* c, int64_t ncells, const int64_t* cell_ptr, const int32_t* ent_snp,
                   const int64_t* ent_ptr, const uint16_t* reads, const int32_t* totl, const int32_t* pass,
                   const int32_t* uniq) {
  if (!c || ncells < 0 || !cell_ptr || !ent_ptr || !totl || !pass || !uniq) return DMX_E_INVALID;
  if (ncells > 0x7fffffffLL) return DMX_E_UNSUPPORTED;
  if (cell_ptr[0] != 0) return DMX_E_INVALID;
  const int64_t E = cell_ptr[ncells];
  if (E < 0 || (E > 0 && (!ent_snp))) return DMX_E_INVALID;
  if (ent_ptr[0] != 0) return DMX_E_INVALID;
  const int64_t R = ent_ptr[E];
  if (R < 0 || (R > 0 && !reads)) return DMX_E_INVALID;
  for (int64_t i = 0; i < ncells; ++i)
    if (cell_ptr[i + 1] < cell_ptr[i]) return DMX_E_INVALID;
  CU(cudaSetDevice(c->cfg.device));
  if (!c->kids.empty()) return sharded_set_pileup(c, ncells, cell_ptr, ent_snp, ent_ptr, reads, totl, pass, uniq);
  c->ncells = ncells; c->E = E; c->R = R;
  if (c->cfg.flags & DMX_BORROW_INPUTS) {
    c->h_cell_ptr = cell_ptr; c->h_ent_snp = ent_snp; c->h_ent_ptr = ent_ptr; c->h_reads = reads;
  } else {
    CU(c->o_cell_ptr.reserve(ncells + 1)); CU(c->o_ent_snp.reserve(E)); CU(c->o_ent_ptr.reserve(E + 1));
    CU(c->o_reads.reserve(R));
    memcpy(c->o_cell_ptr.p, cell_ptr, (ncells + 1) * sizeof(int64_t));
    if (E) memcpy(c->o_ent_snp.p, ent_snp, E * sizeof(int32_t));
    memcpy(c->o_ent_ptr.p, ent_ptr, (E + 1) * sizeof(int64_t));
    if (R) memcpy(c->o_reads.p, reads, R * sizeof(uint16_t));
    c->h_cell_ptr = c->o_cell_ptr.p; c->h_ent_snp = c->o_ent_snp.p; c->h_ent_ptr = c->o_ent_ptr.p;
    c->h_reads = c->o_reads.p;
  }
  // schedule: drop filtered (:581) and empty (:592) cells, largest cells first.  It depends on the cell sizes and the
  // read counts only: a caller that hands in the same droplets again (new reads, new qualities) keeps it
  const bool same_sched = c->have_plp && (int64_t)c->sched_cell_ptr.size() == ncells + 1 &&
                          (int64_t)c->totl.size() == ncells && (int64_t)c->uniq.size() == ncells &&
                          memcmp(c->sched_cell_ptr.data(), cell_ptr, (ncells + 1) * sizeof(int64_t)) == 0 &&
                          memcmp(c->totl.data(), totl, ncells * sizeof(int32_t)) == 0 &&
                          memcmp(c->uniq.data(), uniq, ncells * sizeof(int32_t)) == 0;
  c->npass.assign(pass, pass + ncells);
  if (same_sched) {
    c->plp_dirty = true;
    c->uploaded = false;
    c->computed = false;
    return DMX_OK;
  }
  c->totl.assign(totl, totl + ncells);
  c->uniq.assign(uniq, uniq + ncells);
  c->sched_cell_ptr.assign(cell_ptr, cell_ptr + ncells + 1);
  c->status.assign(ncells, 0);
  c->order.clear();
  for (int64_t i = 0; i < ncells; ++i) {
    const int64_t ns = cell_ptr[i + 1] - cell_ptr[i];
    if (totl[i] < c->cfg.min_total || uniq[i] < c->cfg.min_uniq || ns < c->cfg.min_snp) c->status[i] = 1;
    else if (ns == 0) c->status[i] = 2;
    else c->order.push_back((int32_t)i);
  }
  // chunks: contiguous cell-id ranges with about the same number of entries; inside a chunk the active cells are
  // scheduled largest first.  Copies, kernels and the host finaliser of different chunks overlap (run_path).
  int K = (E >= (2 << 20) && ncells >= 256) ? 2 : 1;
  if (const char* ek = getenv("DMX_CHUNKS")) K = std::max(1, std::min(16, atoi(ek)));
  // one sort of the active cells (largest first, ties by cell id) serves every schedule: a chunked schedule is a
  // stable partition of it by chunk
  std::vector<uint64_t> keys(c->order.size());
  for (size_t i = 0; i < keys.size(); ++i) {
    const int32_t cell = c->order[i];
    const uint64_t sz = (uint64_t)std::min<int64_t>(cell_ptr[cell + 1] - cell_ptr[cell], 0xffffffffLL);
    keys[i] = ((0xffffffffull - sz) << 32) | (uint32_t)cell;
  }
  std::sort(keys.begin(), keys.end());
  auto build = [&](int nk, std::vector<int32_t>* ord, std::vector<Chunk>* chunks) {
    chunks->clear();
    ord->clear();
    ord->reserve(keys.size());
    int64_t cb = 0;
    for (int k = 0; k < nk; ++k) {
      Chunk ck;
      ck.c0 = cb;
      int64_t ce = cb;
      // a small first chunk lets the kernels start early (two chunks by default: 14 % / 86 % of the entries)
      // two chunks: the first one's kernels should end about when the second one's copies do -- sooner when the offsets
      // travel as one-byte read counts (pack_chunk; an entry of a droplet pileup has a handful of reads)
      double f2 = (R <= 8 * E && E >= (4 << 20)) ? 0.14 : 0.24;
      if (const char* ef = getenv("DMX_CHUNK_F0")) f2 = std::min(0.9, std::max(0.02, atof(ef)));
      const double frac = nk == 3 ? (k == 0 ? 0.15 : 0.70) : (nk == 2 ? f2 : (double)(k + 1) / nk);
      if (k == nk - 1) ce = ncells;
      else
        while (ce < ncells && (double)cell_ptr[ce] < frac * (double)E) ++ce;
      ck.c1 = ce;
      ck.o0 = (int64_t)ord->size();
      for (const uint64_t key : keys) {
        const int32_t cell = (int32_t)(uint32_t)key;
        if (cell >= cb && cell < ce) ord->push_back(cell);
      }
      ck.o1 = (int64_t)ord->size();
      ck.idx = k;
      chunks->push_back(ck);
      cb = ce;
    }
  };
  // resident inputs: one chunk (every chunk costs ~0.3 ms of latency-bound per-cell kernels and a pipeline fill of the
  // scan; the decision logic runs on the device, so the host has nothing to overlap)
  build(1, &c->order, &c->chunks);
  build(K, &c->order_p, &c->chunks_p);
  c->have_plp = true;
  c->plp_dirty = true;
  c->uploaded = false;
  c->computed = false;
  return DMX_OK;
}

namespace {
// ---- host -> device copies: the inputs that every chunk needs, then one chunk's slices of the CSR arrays -----------
int check_inputs(dmx_ctx* c) {
  if (c->E > 0 && !(c->cfg.flags & DMX_TRUST_INPUTS)) {
    // SNP ids must index the GP matrix
    // (checked lazily here rather than in dmx_set_pileup so either setter can come first)
    int32_t mx = -1, mn = 0;
    for (int64_t e = 0; e < c->E; ++e) {
      mx = std::max(mx, c->h_ent_snp[e]);
      mn = std::min(mn, c->h_ent_snp[e]);
    }
    if (mn < 0 || mx >= c->nsnps) return DMX_E_INVALID;
  }
  return DMX_OK;
}

int reserve_inputs(dmx_ctx* c) {
  const int nvp = c->nvp;
  CU(c->d_gp.reserve((size_t)c->nsnps * nvp * 3 + 4));
  CU(c->d_gprow.reserve((size_t)c->nsnps * (nvp * 3 + 8) + 256));
  CU(c->d_cell_ptr.reserve(c->ncells + 1)); CU(c->d_ent_snp.reserve(c->E + 1)); CU(c->d_ent_ptr.reserve(c->E + 1));
  CU(c->d_reads.reserve(c->R + 1)); CU(c->d_order.reserve(c->order.size() + 1));
  return DMX_OK;
}

int upload_common(dmx_ctx* c, cudaStream_t st, int64_t* bytes, const std::vector<int32_t>& order) {
  const int nv = c->cfg.nv, nvp = c->nvp;
  if (c->gp_dirty) {
    if (nvp == nv) {
      CU(cudaMemcpyAsync(c->d_gp.p, c->h_gp, (size_t)c->nsnps * nv * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
      CU(cudaMemsetAsync(c->d_gp.p, 0, (size_t)c->nsnps * nvp * 3 * sizeof(float), st));
      CU(cudaMemcpy2DAsync(c->d_gp.p, (size_t)nvp * 12, c->h_gp, (size_t)nv * 12, (size_t)nv * 12, (size_t)c->nsnps,
                           cudaMemcpyHostToDevice, st));
    }
    *bytes += c->nsnps * nv * 3 * (int64_t)sizeof(float);
  }
  if (c->plp_dirty) {
    CU(cudaMemcpyAsync(c->d_cell_ptr.p, c->h_cell_ptr, (c->ncells + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    *bytes += (c->ncells + 1) * 8;
  }
  // the schedule goes with every run: dmx_compute and dmx_run use different ones
  if (!order.empty())
    CU(cudaMemcpyAsync(c->d_order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  *bytes += (int64_t)order.size() * 4;
  return DMX_OK;
}

// Host side of the packed upload: ent_ptr is 8 bytes per entry, a third of everything dmx_run sends, yet an entry has a
// handful of reads.  The pool writes the counts of the chunk's entries (1 byte each) and the first offset of every cell
// into pinned staging buffers while earlier copies are in flight; entptr_kernel rebuilds the offsets on the device.
// Returns false (chunk goes unpacked) when an entry has more than 255 reads.
bool pack_chunk(dmx_ctx* c, const Chunk& ck) {
  const int64_t e0 = c->h_cell_ptr[ck.c0], e1 = c->h_cell_ptr[ck.c1];
  // worth it when the copies saved (7 bytes per entry over PCIe) outweigh the host pass, and when there is a GP upload to
  // hide that pass under: large pileups only (C3: 11.3 M entries; C2's 2.3 M are faster unpacked)
  int64_t min_entries = 4 << 20;
  if (const char* em = getenv("DMX_PTR_PACK_MIN")) min_entries = atoll(em);
  if (c->E < min_entries || e1 - e0 < (1 << 16) || getenv("DMX_NO_PTR_PACK")) return false;
  if (c->h_cnt.reserve((size_t)c->E + 1) != cudaSuccess || c->h_cell_r.reserve((size_t)c->ncells + 2) != cudaSuccess ||
      c->d_cnt.reserve((size_t)c->E + 1) != cudaSuccess || c->d_cell_r.reserve((size_t)c->ncells + 2) != cudaSuccess)
    return false;
  const int nth = host_threads(c);
  const int64_t n = e1 - e0, blk = 1 << 15, nb = (n + blk - 1) / blk;
  std::atomic<int> big{0};
  const int64_t* ep = c->h_ent_ptr;
  uint8_t* cnt = c->h_cnt.p;
  c->pool->run(nb, nth, [&](int64_t b, int) {
    const int64_t lo = e0 + b * blk, hi = std::min(e1, lo + blk);
    int64_t mx = 0;
    for (int64_t e = lo; e < hi; ++e) {
      const int64_t d = ep[e + 1] - ep[e];
      mx = std::max(mx, d);
      cnt[e] = (uint8_t)d;
    }
    if (mx > 255) big.store(1);
  });
  if (big.load()) return false;
  for (int64_t cl = ck.c0; cl <= ck.c1; ++cl) c->h_cell_r.p[cl] = ep[c->h_cell_ptr[cl]];
  return true;
}

int upload_chunk(dmx_ctx* c, Chunk& ck, cudaStream_t st, int64_t* bytes, bool may_pack) {
  if (!c->plp_dirty) return DMX_OK;
  const int64_t e0 = c->h_cell_ptr[ck.c0], e1 = c->h_cell_ptr[ck.c1];
  const int64_t r0 = c->h_ent_ptr[e0], r1 = c->h_ent_ptr[e1];
  ck.packed = may_pack && pack_chunk(c, ck);
  if (e1 > e0)
    CU(cudaMemcpyAsync(c->d_ent_snp.p + e0, c->h_ent_snp + e0, (e1 - e0) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (ck.packed) {
    CU(cudaMemcpyAsync(c->d_cnt.p + e0, c->h_cnt.p + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_cell_r.p + ck.c0, c->h_cell_r.p + ck.c0, (size_t)(ck.c1 - ck.c0 + 1) * sizeof(int64_t),
                       cudaMemcpyHostToDevice, st));
    *bytes += (e1 - e0) + (ck.c1 - ck.c0 + 1) * 8;
  } else {
    CU(cudaMemcpyAsync(c->d_ent_ptr.p + e0, c->h_ent_ptr + e0, (e1 - e0 + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    *bytes += (e1 - e0 + 1) * 8;
  }
  if (r1 > r0)
    CU(cudaMemcpyAsync(c->d_reads.p + r0, c->h_reads + r0, (r1 - r0) * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  *bytes += (e1 - e0) * 4 + (r1 - r0) * 2;
  return DMX_OK;
}

extern "C++" {
// parameters of the entry-table kernels (everything but the entry range)
void fill_entq_params(dmx_ctx* c, EntqParams* eqp, bool fp32) {
  EntqParams& eq = *eqp;
  memset(&eq, 0, sizeof(eq));
  const int nA = c->cfg.n_alpha;
  const int nd = (c->UQ - QHEAD) / QPLANE;
  eq.U = c->pg.U; eq.UQ = c->UQ;
  eq.nd = nd; eq.sep0 = c->pg.sep0; eq.gp_stride = c->nvp * 3 + 8; eq.gprow = c->d_gprow.p;
  eq.ent_snp = c->d_ent_snp.p; eq.ent_ptr = c->d_ent_ptr.p; eq.reads = c->d_reads.p;
  eq.entq = c->d_entq.p;
  eq.ents = c->d_ents.p;
  eq.entf_head = fp32 ? c->d_entf_head.p : nullptr;
  eq.entf_plane = fp32 ? c->d_entf_plane.p : nullptr;
  eq.entf_pstride = (c->E + NB) * FPLANE;
  eq.perm = fp32 ? c->d_perm.p : nullptr;
  eq.snp_perm = fp32 ? c->d_snp_perm.p : nullptr;
  eq.ecode = fp32 ? c->d_ecode.p : nullptr;
  eq.memo_q = c->d_memo_q.p; eq.memo_fh = c->d_memo_fh.p; eq.memo_fp = c->d_memo_fp.p;
  memcpy(eq.pu, c->pg.pu, sizeof(double) * c->pg.U);
  // record layout: GL[3] (written by the kernel), f[3] = Q'[0][l][0] of a separable plane 0, then per dense
  // plane 9 values Q'[n][l][m] (+1 pad); oslot lists, per distinct p value, the record slots it feeds
  std::vector<std::vector<int>> slots(c->pg.U);
  if (c->pg.sep0)
    for (int l = 0; l < 3; ++l) slots[c->pg.umap[l * 3]].push_back(3 + l);
  for (int n = 0; n < nA; ++n)
    if (c->dense_idx[n] >= 0)
      for (int i = 0; i < 9; ++i) slots[c->pg.umap[n * 9 + i]].push_back(QHEAD + QPLANE * c->dense_idx[n] + i);
  int o = 0;
  for (int u2 = 0; u2 < c->pg.U; ++u2) {
    eq.oslot_ptr[u2] = (uint8_t)o;
    for (int sl : slots[u2]) { eq.oslot[o++] = (uint8_t)sl; eq.slot_u[sl] = (uint8_t)u2; }
  }
  eq.oslot_ptr[c->pg.U] = (uint8_t)o;
}

// Memo tables of the affine entries (fp32 scan): the entry-table kernels themselves run once on a synthetic pileup of
// MEMO_ROWS entries -- one read of allele a and quality q in row a * 256 + q, no informative read in the last row --
// so a memo record is bit for bit what the kernels would write for such an entry.  (The SNP-dependent factors of a
// record are not memo material: entq_aff_kernel recomputes them.)  Needs the GP rows of SNP 0 to exist.
int build_memo(dmx_ctx* c) {
  const int nd = (c->UQ - QHEAD) / QPLANE;
  std::vector<int64_t> ptr(MEMO_ROWS + 1);
  std::vector<uint16_t> rd(MEMO_ROWS, 0);
  for (int i = 0; i < MEMO_ROWS - 1; ++i) { ptr[i] = i; rd[i] = (uint16_t)i; }   // allele << 8 | quality
  ptr[MEMO_ROWS - 1] = MEMO_ROWS - 1;
  ptr[MEMO_ROWS] = MEMO_ROWS - 1;          // the last row has no read at all
  DevBuf<int64_t> d_ptr;
  DevBuf<uint16_t> d_rd;
  DevBuf<int32_t> d_snp;
  DevBuf<double> d_ents;
  struct Rel { DevBuf<int64_t>* a; DevBuf<uint16_t>* b; DevBuf<int32_t>* s; DevBuf<double>* e;
               ~Rel() { a->release(); b->release(); s->release(); e->release(); } } rel{&d_ptr, &d_rd, &d_snp, &d_ents};
  CU(d_ptr.reserve(MEMO_ROWS + 1)); CU(d_rd.reserve(MEMO_ROWS)); CU(d_snp.reserve(MEMO_ROWS));
  CU(d_ents.reserve((size_t)MEMO_ROWS * (2 + 2 * nd)));
  CU(c->d_memo_q.reserve((size_t)MEMO_ROWS * c->UQ)); CU(c->d_memo_fh.reserve((size_t)MEMO_ROWS * FHEAD));
  CU(c->d_memo_fp.reserve((size_t)std::max(nd, 1) * MEMO_ROWS * FPLANE));
  CU(cudaMemcpyAsync(d_ptr.p, ptr.data(), ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_rd.p, rd.data(), rd.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(d_snp.p, 0, MEMO_ROWS * sizeof(int32_t), c->stream));
  EntqParams eq;
  fill_entq_params(c, &eq, true);
  eq.E0 = 0; eq.E = MEMO_ROWS;
  eq.ent_snp = d_snp.p; eq.ent_ptr = d_ptr.p; eq.reads = d_rd.p;
  eq.entq = c->d_memo_q.p; eq.ents = d_ents.p;
  eq.entf_head = c->d_memo_fh.p; eq.entf_plane = c->d_memo_fp.p; eq.entf_pstride = (int64_t)MEMO_ROWS * FPLANE;
  eq.perm = nullptr; eq.snp_perm = nullptr; eq.ecode = nullptr;
  CU(launch_entq(eq, c->stream));
  CU(cudaStreamSynchronize(c->stream));     // the staging vectors live on this stack frame
  c->st.gpu_launches += 1;
  c->memo_ready = true;
  return DMX_OK;
}

// the full fp64 entry table (every entry through the general kernels): what the fp64 scan streams.  In fp32 mode
// only the general entries have their records there after a run; dense llksAB blocks (dmx_get_pair, --write-pair,
// candidate-list overflow) come from the fp64 scan and fill the rest in first.
int ensure_entq_full(dmx_ctx* c) {
  if (c->entq_full || !(c->cfg.flags & DMX_FP32_PAIRS) || c->E == 0) return DMX_OK;
  EntqParams eq;
  fill_entq_params(c, &eq, false);
  eq.E0 = 0; eq.E = c->E;
  CU(launch_entq(eq, c->stream));
  c->st.gpu_launches += 1;
  c->entq_full = true;
  return DMX_OK;
}
}  // extern "C++"

// ---- the kernels of one chunk (a contiguous cell-id range; its active cells are order[o0, o1)) ----------------------
int enqueue_chunk(dmx_ctx* c, const Chunk& ck, const ScanParams& sp0, bool dev_rescore, cudaEvent_t* ev /* [5] */) {
  const int nv = c->cfg.nv, nA = c->cfg.n_alpha;
  const int T = (int)c->tt.size();
  const int nd = (c->UQ - QHEAD) / QPLANE;
  const bool fp32 = (c->cfg.flags & DMX_FP32_PAIRS) != 0;
  const int64_t e0 = c->h_cell_ptr[ck.c0], e1 = c->h_cell_ptr[ck.c1];
  const int64_t nact = ck.o1 - ck.o0;
  CU(cudaEventRecord(ev[0], c->stream));
  if (ck.packed) {   // the offsets of this chunk came as read counts
    CU(launch_entptr(c->d_cell_ptr.p, c->d_cell_r.p, c->d_cnt.p, ck.c0, ck.c1, c->d_ent_ptr.p, c->stream));
    c->st.gpu_launches += 1;
  }
  EntqParams eq;
  fill_entq_params(c, &eq, fp32);
  eq.E0 = e0; eq.E = e1;
  if (fp32 && e1 > e0) {
    // entry classes of the chunk's cells (affine entries first in the fp32 tables), then the affine entries from the
    // memo tables and the general ones (the listed fifth or so) through the full computation
    CU(launch_classify(c->d_cell_ptr.p, c->d_ent_ptr.p, c->d_reads.p, ck.c0, ck.c1, c->d_perm.p, c->d_cell_na.p,
                       c->d_glist.p, c->stream));
    CU(launch_entq_aff(eq, c->stream));
    eq.glist = c->d_glist.p;
    c->st.gpu_launches += 2;
  }
  CU(launch_entq(eq, c->stream));
  if (e1 > e0) c->st.gpu_launches += 1;
  if (fp32 && nact) {
    // the fp32 scan leaves the sample-independent likelihoods (llk0s, llks00) to their own small kernel
    CellScalarParams cs;
    memset(&cs, 0, sizeof(cs));
    cs.nd = nd; cs.nA = nA; cs.sep0 = c->pg.sep0; cs.want_single = (c->cfg.flags & DMX_SKIP_SINGLE) ? 0 : 1;
    for (int n = 0; n < nA; ++n)
      if (c->dense_idx[n] >= 0) cs.plane_n[c->dense_idx[n]] = (int8_t)n;
    cs.ents = c->d_ents.p; cs.cell_ptr = c->d_cell_ptr.p; cs.cell_order = c->d_order.p + ck.o0;
    cs.p1_0 = c->d_p1_0.p; cs.l00 = c->d_l00.p; cs.lm = c->scan_gen == 2 ? c->d_lm.p : nullptr;
    CU(launch_cell_scalars(cs, nact, c->stream));
    c->st.gpu_launches += 1;
  }
  CU(cudaEventRecord(ev[1], c->stream));
  ScanParams sp = sp0;
  sp.cell_order = c->d_order.p + ck.o0;
  sp.parts = c->d_parts.p + ck.o0 * T;
  sp.ntasks = nact * T;
  CU(launch_pair_scan(sp, c->maps, fp32 ? c->scan_gen : 0, c->stream));
  CU(cudaEventRecord(ev[2], c->stream));
  MergeParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.nv = nv; mp.nA = nA; mp.T = T; mp.sep0 = c->pg.sep0; mp.eps_rel = sp.eps_rel; mp.eps_ent = sp.eps_ent; mp.cell_ptr = c->d_cell_ptr.p;
  mp.w_single = (1. - c->cfg.doublet_prior) / nv;
  for (int n = 0; n < nA; ++n)
    mp.w_double[n] = c->cfg.doublet_prior / nv / (nv - 1) / (nA - 1) / (c->alpha[n] == 0.5 ? 2.0 : 1.0);
  mp.cell_order = sp.cell_order; mp.tt = c->d_tt.p; mp.parts = sp.parts; mp.row0 = c->d_row0.p;
  mp.singA = c->d_singA.p; mp.lb = c->d_lb.p; mp.sing = c->d_sing.p; mp.out = c->d_out.p;
  if (dev_rescore) {
    mp.lists = c->d_lists.p;
    for (int n = 0; n < nA; ++n) mp.mode[n] = c->pg.mode[n];
    mp.any_long = c->d_any_long.p + ck.idx;
  }
  // read back on the third stream as soon as each array is final (all per-cell arrays are indexed by cell id: one
  // contiguous slice per chunk): the scan's outputs while merge and the re-scorer run, the rest after them
  const int64_t c0 = ck.c0, n = ck.c1 - ck.c0;
  if (n > 0) {
    CU(cudaStreamWaitEvent(c->stream3, ev[2], 0));
    CU(cudaMemcpyAsync(c->r_row0.p + c0 * nA * nv, c->d_row0.p + c0 * nA * nv, (size_t)n * nA * nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
    CU(cudaMemcpyAsync(c->r_l00.p + c0 * nA, c->d_l00.p + c0 * nA, (size_t)n * nA * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
    if (sp.want_single) {
      CU(cudaMemcpyAsync(c->r_p1.p + c0 * nv, c->d_p1.p + c0 * nv, (size_t)n * nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
      CU(cudaMemcpyAsync(c->r_p1_0.p + c0, c->d_p1_0.p + c0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
    }
  }
  CU(launch_merge(mp, nact, c->stream));
  if (nact) c->st.gpu_launches += 2;
  CU(cudaEventRecord(ev[4], c->stream));
  if (n > 0) {
    CU(cudaStreamWaitEvent(c->stream3, ev[4], 0));
    CU(cudaMemcpyAsync(c->r_sing.p + c0 * nv, c->d_sing.p + c0 * nv, (size_t)n * nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream3));
    CU(cudaMemcpyAsync(c->r_out.p + c0, c->d_out.p + c0, n * sizeof(CellOut), cudaMemcpyDeviceToHost, c->stream3));
  }
  if (dev_rescore) {
    RescoreParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.nv = nv; rp.gp_stride = c->nvp * 3 + 8; rp.nA = nA; rp.UQ = c->UQ; rp.sep0 = c->pg.sep0;
    for (int n2 = 0; n2 < nA; ++n2) rp.dense_idx[n2] = (int8_t)c->dense_idx[n2];
    rp.gp = c->d_gprow.p + 8; rp.ent_snp = c->d_ent_snp.p; rp.entq = c->d_entq.p; rp.cell_ptr = c->d_cell_ptr.p;
    rp.cell_order = sp.cell_order; rp.lists = c->d_lists.p;
    rp.ecode = fp32 ? c->d_ecode.p : nullptr; rp.memo_q = c->d_memo_q.p;
    rp.terms = c->d_terms.p;
    rp.any_long = c->d_any_long.p + ck.idx;
    CU(launch_rescore(rp, nact, c->stream));
    if (nact) c->st.gpu_launches += 2 * (RMAX / RESC_FAST);
    // the decision logic on the device; the cells it cannot decide (flag 1) go through the host finaliser below
    FinaliseParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.nv = nv; fp.nA = nA; fp.eps_rel = sp.eps_rel; fp.eps_ent = sp.eps_ent; fp.w_single = (1. - c->cfg.doublet_prior) / nv;
    for (int n2 = 0; n2 < nA; ++n2) fp.mode[n2] = c->pg.mode[n2];
    fp.cell_ptr = c->d_cell_ptr.p; fp.cell_order = sp.cell_order; fp.sing = c->d_sing.p; fp.row0 = c->d_row0.p;
    fp.l00 = c->d_l00.p; fp.out = c->d_out.p; fp.lists = c->d_lists.p; fp.best = c->d_best.p; fp.flag = c->d_flag.p;
    CU(launch_finalise(fp, nact, c->stream));
    if (nact) c->st.gpu_launches += 1;
  }
  CU(cudaEventRecord(ev[3], c->stream));
  if (n > 0 && dev_rescore) {
    CU(cudaStreamWaitEvent(c->stream3, ev[3], 0));
    CU(cudaMemcpyAsync(c->best.p + c0, c->d_best.p + c0, n * sizeof(dmx_best_rec), cudaMemcpyDeviceToHost, c->stream3));
    CU(cudaMemcpyAsync(c->r_flag.p + c0, c->d_flag.p + c0, n * sizeof(uint8_t), cudaMemcpyDeviceToHost, c->stream3));
    // (the re-score lists stay on the device: run_path fetches them for the cells the host finaliser takes)
  }
  return DMX_OK;
}

// The whole path.  `pipelined`: the host->device copies run on a second stream, chunk by chunk, under the kernels of
// the previous chunks; in both modes the records of a chunk are read back and finalised on the host threads while the
// device works on the next chunk.
int run_path(dmx_ctx* c, bool pipelined) {
  CU(cudaSetDevice(c->cfg.device));
  const double t0 = now_ms();
  const int nv = c->cfg.nv, nA = c->cfg.n_alpha;
  const std::vector<int32_t>& order = pipelined ? c->order_p : c->order;
  const std::vector<Chunk>& chunks = pipelined ? c->chunks_p : c->chunks;
  const int64_t nc = c->ncells, nact = (int64_t)order.size();
  const int T = (int)c->tt.size();
  const int K = (int)chunks.size();
  c->st.gpu_launches = 0;
  c->st.n_rescored = 0;
  c->st.n_fullcell = 0;

  const int nd = (c->UQ - QHEAD) / QPLANE;
  const bool fp32 = (c->cfg.flags & DMX_FP32_PAIRS) != 0;
  const int64_t pstride = (c->E + NB) * FPLANE;
  const bool exact = (c->cfg.flags & DMX_EXACT_TIES) != 0;
  const bool dev_rescore = exact && c->gpu_exact && !(c->cfg.flags & DMX_HOST_FINALISE);
  if (pipelined) {
    const int rc = check_inputs(c);
    if (rc != DMX_OK) return rc;
    RC(reserve_inputs(c));
  }
  CU(c->d_entq.reserve((size_t)c->E * c->UQ + (size_t)NB * c->UQ + 2));   // slack: a stage always reads NB records
  CU(c->d_ents.reserve((size_t)c->E * (2 + 2 * nd) + 2));
  if (fp32) {
    CU(c->d_entf_head.reserve((size_t)(c->E + NB) * FHEAD));
    CU(c->d_entf_plane.reserve((size_t)pstride * std::max(nd, 1)));
    CU(c->d_perm.reserve((size_t)c->E + 1)); CU(c->d_snp_perm.reserve((size_t)c->E + NB)); CU(c->d_cell_na.reserve((size_t)nc + 1));
    CU(c->d_ecode.reserve((size_t)c->E + 1)); CU(c->d_glist.reserve((size_t)c->E + 1));
  }
  CU(c->d_parts.reserve((size_t)nact * T + 1));
  CU(c->d_row0.reserve((size_t)nc * nA * nv + 1)); CU(c->d_singA.reserve((size_t)nc * nv + 1));
  CU(c->d_lb.reserve((size_t)nc * nv + 1)); CU(c->d_p1.reserve((size_t)nc * nv + 1));
  CU(c->d_p1_0.reserve(nc + 1)); CU(c->d_l00.reserve((size_t)nc * nA + 1));
  {
    const char* g = getenv("DMX_SCAN_GEN");
    c->scan_gen = (g && atoi(g) == 1) ? 1 : 2;
  }
  if (fp32 && c->scan_gen == 2) {
    CU(c->d_lm.reserve((size_t)nc * nA + 1));
    CU(c->d_drow.reserve((size_t)2 * c->nsnps * (c->nvp + 8) + 256));
    CU(c->d_srow.reserve((size_t)c->nsnps * (c->nvp + 8) + 256));
  }
  CU(c->d_sing.reserve((size_t)nc * nv + 1)); CU(c->d_out.reserve(nc + 1));
  CU(c->best.reserve(nc + 1));
  CU(c->r_out.reserve(nc + 1)); CU(c->r_sing.reserve((size_t)nc * nv + 1));
  CU(c->r_row0.reserve((size_t)nc * nA * nv + 1)); CU(c->r_l00.reserve((size_t)nc * nA + 1));
  CU(c->r_p1.reserve((size_t)nc * nv + 1)); CU(c->r_p1_0.reserve(nc + 1));
  if (dev_rescore) {
    CU(c->d_lists.reserve(nc + 1));
    CU(c->r_lists.reserve(nc + 1));
    CU(c->d_terms.reserve((size_t)c->E * RESC_FAST + 1));
    CU(c->d_best.reserve(nc + 1));
    CU(c->d_flag.reserve(nc + 1)); CU(c->r_flag.reserve(nc + 1));
    CU(c->d_any_long.reserve(64));
  }
  if ((int)c->cev.size() < 7 * K + 4) {
    const size_t old = c->cev.size();
    c->cev.resize(7 * K + 4, nullptr);
    for (size_t i = old; i < c->cev.size(); ++i) CU(cudaEventCreate(&c->cev[i]));
  }
  cudaEvent_t* ev_up = c->cev.data();            // [K + 2]: copies start, common inputs there, chunk k there
  cudaEvent_t* ev_k = c->cev.data() + K + 2;     // [K][5] kernel phases, then [K] chunk read back
  cudaEvent_t* ev_done = ev_k + 5 * K;
  cudaEvent_t ev_first = c->cev[7 * K + 2], ev_last = c->cev[7 * K + 3];

  int64_t h2d_bytes = 0;
  if (pipelined) {
    CU(cudaEventRecord(ev_up[0], c->stream2));
    RC(upload_common(c, c->stream2, &h2d_bytes, order));
    CU(cudaEventRecord(ev_up[1], c->stream2));
    // (the chunks' copies are enqueued one by one below, each right before its kernels: packing chunk k + 1 on the host
    // then runs under the copies and kernels of chunk k)
  }
  if (!pipelined && !order.empty())   // dmx_run may have left its own schedule there
    CU(cudaMemcpyAsync(c->d_order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(ev_first, c->stream));
  // outputs of cells that are never scheduled stay zero (the reference prints zeros for them in .single)
  CU(cudaMemsetAsync(c->d_p1.p, 0, (size_t)nc * nv * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_p1_0.p, 0, (size_t)nc * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_sing.p, 0, (size_t)nc * nv * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_l00.p, 0, (size_t)nc * nA * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_row0.p, 0, (size_t)nc * nA * nv * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_out.p, 0, (size_t)nc * sizeof(CellOut), c->stream));
  if (dev_rescore) {
    CU(cudaMemsetAsync(c->d_flag.p, 1, (size_t)nc, c->stream));
    CU(cudaMemsetAsync(c->d_any_long.p, 0, 64 * sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(c->d_best.p, 0, (size_t)nc * sizeof(dmx_best_rec), c->stream));
  }
  if (pipelined) CU(cudaStreamWaitEvent(c->stream, ev_up[1], 0));
  // tables derived from the GP matrix: once per matrix (a run with the inputs resident finds them in place)
  const int want_tab = (fp32 && c->scan_gen == 2) ? 2 : 1;
  if (pipelined) c->gptab_state = 0;
  if (c->gptab_state < 1) {
    CU(launch_gprow(c->d_gp.p, c->d_gprow.p, c->nsnps, nv, c->nvp, c->stream));
    if (c->nsnps) c->st.gpu_launches += 1;
  }
  if (want_tab == 2 && c->gptab_state < 2 && c->nsnps) {
    CU(launch_dtab(c->d_gprow.p, c->d_drow.p, c->d_srow.p, c->nsnps, c->nvp, c->stream));
    c->st.gpu_launches += 1;
  }
  c->gptab_state = std::max(c->gptab_state, want_tab);
  if (want_tab == 2) CU(cudaMemsetAsync(c->d_lm.p, 0, (size_t)nc * nA * sizeof(double), c->stream));
  c->entq_full = !fp32;
  if (fp32 && !c->memo_ready && c->nsnps > 0 && c->E > 0) RC(build_memo(c));
  ScanParams sp;
  fill_scan_params(c, &sp);
  {
    const int rc = encode_maps(c);
    if (rc != DMX_OK) return rc;
  }
  for (int k = 0; k < K; ++k) {
    if (pipelined) {
      RC(upload_chunk(c, c->chunks_p[k], c->stream2, &h2d_bytes, true));
      CU(cudaEventRecord(ev_up[2 + k], c->stream2));
      CU(cudaStreamWaitEvent(c->stream, ev_up[2 + k], 0));
    }
    const int rc = enqueue_chunk(c, chunks[k], sp, dev_rescore, ev_k + 5 * k);
    if (rc != DMX_OK) return rc;
    CU(cudaEventRecord(ev_done[k], c->stream3));
  }
  CU(cudaEventRecord(ev_last, c->stream));
  if (pipelined) {
    c->gp_dirty = false;
    c->plp_dirty = false;
    c->uploaded = true;
  }

  // ---- finalise, chunk by chunk as the records arrive ---------------------------------------------------------
  if (!dev_rescore && nc) memset(c->best.p, 0, (size_t)nc * sizeof(dmx_best_rec));   // else every record comes from the device
  const HostView hv = host_view(c);
  const ExactScorer scorer(hv, c->err256);
  const int nth = host_threads(c);
  std::vector<std::vector<int32_t>> overflow(nth);
  std::vector<int64_t> resc(nth, 0), resc_gpu(nth, 0);
  auto do_cell = [&](int64_t cell, const std::vector<Hyp>& cands, int tid) {
    dmx_best_rec* r = &c->best.p[cell];
    CellGpu g;
    g.co = &c->r_out.p[cell];
    g.sing = c->r_sing.p + (size_t)cell * nv;
    g.row0 = c->r_row0.p + (size_t)cell * nA * nv;
    g.l00 = c->r_l00.p + (size_t)cell * nA;
    int32_t from_gpu = 0;
    const double eps_abs = sp.eps_ent * (double)(c->h_cell_ptr[cell + 1] - c->h_cell_ptr[cell]);
    finalise_cell(hv, scorer, cell, g, cands, c->cfg.doublet_prior, sp.eps_rel, eps_abs, exact,
                  dev_rescore ? &c->r_lists.p[cell] : nullptr, r, &from_gpu);
    resc[tid] += r->n_rescored;
    resc_gpu[tid] += from_gpu;
  };
  double fin_ms = 0.0;
  const bool getenv_host_fin = getenv("DMX_HOST_DECIDE") != nullptr;   // debugging: every cell through the host logic
  for (int k = 0; k < K; ++k) {
    CU(cudaEventSynchronize(ev_done[k]));
    const double tf0 = now_ms();
    const int64_t c0 = chunks[k].c0, n = chunks[k].c1 - c0;
    auto host_cell = [&](int64_t cell, int tid) {
      dmx_best_rec* r = &c->best.p[cell];
      memset(r, 0, sizeof(*r));
      r->status = c->status[cell];
      r->n_snp = (int32_t)(c->h_cell_ptr[cell + 1] - c->h_cell_ptr[cell]);
      if (r->status != 0) return;
      const CellOut& co = c->r_out.p[cell];
      if (co.ncand > MAX_CAND) { overflow[tid].push_back((int32_t)cell); return; }
      thread_local std::vector<Hyp> cands;
      cands.clear();
      expand_cands(c, co, &cands);
      do_cell(cell, cands, tid);
    };
    if (dev_rescore && !getenv_host_fin) {
      // the records of this chunk came back complete; only the cells finalise_kernel left alone (flag 1: not
      // scheduled, or undecidable on the device) go through the host logic
      thread_local std::vector<int64_t> todo;
      todo.clear();
      for (int64_t i = 0; i < n; ++i) {
        if (c->r_flag.p[c0 + i]) todo.push_back(c0 + i);
        else { resc[0] += c->best.p[c0 + i].n_rescored; resc_gpu[0] += c->best.p[c0 + i].n_rescored; }
      }
      // their re-score lists: a few single copies, or the chunk's slice
      int64_t nlist = 0;
      for (const int64_t cell : todo) nlist += c->status[cell] == 0;
      // (synchronous copies: the chunk's lists are final, and the streams -- non-blocking ones -- carry later chunks)
      if (nlist > 0 && nlist < 32) {
        for (const int64_t cell : todo)
          if (c->status[cell] == 0)
            CU(cudaMemcpy(c->r_lists.p + cell, c->d_lists.p + cell, sizeof(RescList), cudaMemcpyDeviceToHost));
      } else if (nlist > 0) {
        CU(cudaMemcpy(c->r_lists.p + c0, c->d_lists.p + c0, n * sizeof(RescList), cudaMemcpyDeviceToHost));
      }
      if (todo.size() < 64) for (const int64_t cell : todo) host_cell(cell, 0);
      else c->pool->run((int64_t)todo.size(), nth, [&](int64_t i, int tid) { host_cell(todo[i], tid); });
    } else {
      if (dev_rescore && n > 0)
        CU(cudaMemcpy(c->r_lists.p + c0, c->d_lists.p + c0, n * sizeof(RescList), cudaMemcpyDeviceToHost));
      c->pool->run(n, nth, [&](int64_t i, int tid) { host_cell(c0 + i, tid); });
    }
    fin_ms += now_ms() - tf0;
  }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->stream3));
  if (pipelined) CU(cudaStreamSynchronize(c->stream2));
  c->st.d2h_bytes = nc * (int64_t)(sizeof(CellOut) + (dev_rescore ? sizeof(dmx_best_rec) + 1 : 0)) + (int64_t)nc * nv * 8 * (1 + nA + (sp.want_single ? 1 : 0)) +
                    nc * nA * 8 + (sp.want_single ? nc * 8 : 0);
  float ms = 0;
  c->st.prep_ms = 0; c->st.pair_kernel_ms = 0; c->st.d2h_ms = 0;
  for (int k = 0; k < K; ++k) {
    cudaEventElapsedTime(&ms, ev_k[5 * k], ev_k[5 * k + 1]); c->st.prep_ms += ms;
    cudaEventElapsedTime(&ms, ev_k[5 * k + 1], ev_k[5 * k + 2]); c->st.pair_kernel_ms += ms;
    cudaEventElapsedTime(&ms, ev_k[5 * k + 3], ev_done[k]); c->st.d2h_ms += ms;   // read back not hidden by kernels
  }
  cudaEventElapsedTime(&ms, ev_first, ev_last); c->st.kernel_ms = ms;   // first memset .. last kernel
  if (pipelined) {
    cudaEventElapsedTime(&ms, ev_up[0], ev_up[K + 1]); c->st.h2d_ms = ms;
    c->st.h2d_bytes = h2d_bytes;
  }
  const double tf1 = now_ms();
  std::vector<int32_t> ov;
  for (auto& v : overflow) ov.insert(ov.end(), v.begin(), v.end());
  if (!ov.empty()) {
    // candidate list overflow (massive exact ties, e.g. GT-field genotypes on a handful of SNPs): take the
    // candidates from a dense llksAB scan of just those cells
    const size_t per = (size_t)nv * nv * nA;
    const size_t batch = std::max<size_t>(1, ((size_t)256 << 20) / (per * sizeof(double)));
    std::vector<double> dense_host(std::min(batch, ov.size()) * per);
    for (size_t lo = 0; lo < ov.size(); lo += batch) {
      const size_t n = std::min(batch, ov.size() - lo);
      std::vector<int32_t> cells(ov.begin() + lo, ov.begin() + lo + n);
      int rc = dense_scan(c, cells, dense_host.data());
      if (rc != DMX_OK) return rc;
      c->pool->run((int64_t)n, nth, [&](int64_t i, int tid) {
        std::vector<Hyp> cands;
        const int64_t cl = cells[i];
        candidates_from_dense(dense_host.data() + (size_t)i * per, nv, nA, 1, sp.eps_rel,
                              sp.eps_ent * (double)(c->h_cell_ptr[cl + 1] - c->h_cell_ptr[cl]), &cands);
        do_cell(cells[i], cands, tid);
        c->best.p[cells[i]].exact_mask |= DMX_X_FULLCELL;
      });
    }
    c->st.n_fullcell = (int64_t)ov.size();
  }
  fin_ms += now_ms() - tf1;
  c->st.n_rescored_gpu = 0;
  for (int t = 0; t < nth; ++t) { c->st.n_rescored += resc[t]; c->st.n_rescored_gpu += resc_gpu[t]; }
  if (getenv("DMX_FIN_DEBUG") && dev_rescore) {
    int64_t nh = 0;
    for (int64_t i = 0; i < nc; ++i) nh += c->status[i] == 0 && c->r_flag.p[i] != 0;
    fprintf(stderr, "dmx: %lld of %lld cells finalised on the host, finaliser %.3f ms on %d threads\n", (long long)nh,
            (long long)nact, fin_ms, nth);
  }
  c->st.gpu_exact = c->gpu_exact ? 1 : 0;
  c->st.finalise_ms = fin_ms;
  c->st.n_entries = c->E;
  c->st.n_reads = c->R;
  c->st.n_tasks = nact * T;
  // SURVEY 8(d) gather model: per entry 12*nv (GP row) + 12 (gp0 row) + 4 (snp id) + 2*reads
  c->st.algo_bytes = 0;
  for (int64_t i = 0; i < nact; ++i) {
    const int32_t cell = order[i];
    const int64_t e0 = c->h_cell_ptr[cell], e1 = c->h_cell_ptr[cell + 1];
    c->st.algo_bytes += (double)(e1 - e0) * (12.0 * nv + 16.0) + 2.0 * (double)(c->h_ent_ptr[e1] - c->h_ent_ptr[e0]) +
                        8.0 + 200.0 + 8.0 * nv;
  }
  c->st.total_ms = now_ms() - t0;
  c->computed = true;
  return DMX_OK;
}
}  // namespace

int dmx_upload(dmx_ctx* c) {
  if (!c) return DMX_E_INVALID;
  if (!c->have_gp || !c->have_plp) return DMX_E_STATE;
  if (!c->kids.empty()) return sharded_run(c, 1);
  CU(cudaSetDevice(c->cfg.device));
  const double t0 = now_ms();
  int64_t bytes = 0;
  {
    const int rc = check_inputs(c);
    if (rc != DMX_OK) return rc;
  }
  RC(reserve_inputs(c));
  RC(upload_common(c, c->stream, &bytes, c->order));
  for (Chunk& ck : c->chunks) RC(upload_chunk(c, ck, c->stream, &bytes, false));
  c->gp_dirty = false;
  c->plp_dirty = false;
  CU(cudaStreamSynchronize(c->stream));
  c->st.h2d_ms = now_ms() - t0;
  c->st.h2d_bytes = bytes;
  c->uploaded = true;
  c->computed = false;
  return DMX_OK;
}

int dmx_compute(dmx_ctx* c) {
  if (!c) return DMX_E_INVALID;
  if (!c->uploaded) return DMX_E_STATE;
  if (!c->kids.empty()) return sharded_run(c, 2);
  return run_path(c, false);
}

// host buffers in, records out: copies, kernels and the finaliser overlap chunk by chunk
int dmx_run(dmx_ctx* c) {
  if (!c) return DMX_E_INVALID;
  if (!c->have_gp || !c->have_plp) return DMX_E_STATE;
  if (!c->kids.empty()) return sharded_run(c, 0);
  if (!c->gp_dirty && !c->plp_dirty && c->uploaded) return run_path(c, false);
  return run_path(c, true);
}

int dmx_get_best(dmx_ctx* c, dmx_best_rec* out) {
  if (!c || !out) return DMX_E_INVALID;
  if (!c->computed) return DMX_E_STATE;
  if (c->ncells) memcpy(out, c->best.p, c->ncells * sizeof(dmx_best_rec));
  return DMX_OK;
}

int dmx_get_single(dmx_ctx* c, double* llk, double* llk0) {
  if (!c || !llk || !llk0) return DMX_E_INVALID;
  if (!c->computed) return DMX_E_STATE;
  if (c->cfg.flags & DMX_SKIP_SINGLE) return DMX_E_STATE;
  memcpy(llk, c->r_p1.p, (size_t)c->ncells * c->cfg.nv * sizeof(double));
  memcpy(llk0, c->r_p1_0.p, (size_t)c->ncells * sizeof(double));
  return DMX_OK;
}

int dmx_get_sing2(dmx_ctx* c, double* llk_j00, double* llk00) {
  if (!c || !llk_j00 || !llk00) return DMX_E_INVALID;
  if (!c->computed) return DMX_E_STATE;
  memcpy(llk_j00, c->r_sing.p, (size_t)c->ncells * c->cfg.nv * sizeof(double));
  memcpy(llk00, c->r_l00.p, (size_t)c->ncells * c->cfg.n_alpha * sizeof(double));
  return DMX_OK;
}

int dmx_get_pair(dmx_ctx* c, int64_t cell_lo, int64_t cell_hi, double* llkAB) {
  if (!c || !llkAB || cell_lo < 0 || cell_hi > c->ncells || cell_lo > cell_hi) return DMX_E_INVALID;
  if (!c->computed) return DMX_E_STATE;
  const size_t per = (size_t)c->cfg.nv * c->cfg.nv * c->cfg.n_alpha;
  std::vector<int32_t> cells;
  std::vector<int64_t> dst;
  for (int64_t i = cell_lo; i < cell_hi; ++i) {
    if (c->status[i] == 0) {
      cells.push_back((int32_t)i);
      dst.push_back(i - cell_lo);
    } else {
      memset(llkAB + (size_t)(i - cell_lo) * per, 0, per * sizeof(double));   // no row in the output files
    }
  }
  return dense_scan(c, cells, llkAB, dst.data());
}

int dmx_rescore(dmx_ctx* c, int64_t cell, int32_t j, int32_t k, int32_t n, double* out) {
  if (!c || !out || !c->have_gp || !c->have_plp) return DMX_E_STATE;
  if (cell < 0 || cell >= c->ncells || j < 0 || j >= c->cfg.nv || k < 0 || k >= c->cfg.nv || n < 0 ||
      n >= c->cfg.n_alpha)
    return DMX_E_INVALID;
  if (!c->kids.empty()) {
    const int s = (int)(std::upper_bound(c->kid_lo.begin(), c->kid_lo.end(), cell) - c->kid_lo.begin()) - 1;
    return dmx_rescore(c->kids[(size_t)s], cell - c->kid_lo[(size_t)s], j, k, n, out);
  }
  const HostView hv = host_view(c);
  const ExactScorer scorer(hv, c->err256);
  Hyp h{j, k, n};
  scorer.eval(cell, &h, 1, out);
  return DMX_OK;
}

int dmx_exact_llk(int32_t nv, int32_t n_alpha, const double* alpha, const float* gp, const int64_t* cell_ptr,
                  const int32_t* ent_snp, const int64_t* ent_ptr, const uint16_t* reads, int64_t cell, int32_t j,
                  int32_t k, int32_t n, double* out) {
  if (!alpha || !gp || !cell_ptr || !ent_ptr || !out || nv < 1 || n_alpha < 1 || n_alpha > DMX_MAX_ALPHA || cell < 0 ||
      j < 0 || j >= nv || k < 0 || k >= nv || n < 0 || n >= n_alpha)
    return DMX_E_INVALID;
  HostView hv;
  hv.nv = nv; hv.nA = n_alpha; hv.alpha = alpha; hv.gp = gp;
  hv.cell_ptr = cell_ptr; hv.ent_snp = ent_snp; hv.ent_ptr = ent_ptr; hv.reads = reads;
  double err256[256];
  for (int i = 0; i <= 255; ++i) err256[i] = (i > 1) ? pow(0.1, i * 0.1) : 0.75;   // PhredHelper.cpp:30
  const ExactScorer scorer(hv, err256);
  Hyp h{j, k, n};
  scorer.eval(cell, &h, 1, out);
  return DMX_OK;
}

int dmx_get_stats(dmx_ctx* c, dmx_stats* out) {
  if (!c || !out) return DMX_E_INVALID;
  *out = c->st;
  return DMX_OK;
}

// ---- writers ---------------------------------------------------------------------------------------
namespace {

// printf into a growing text buffer (rows carry caller-supplied names of any length)
void appendf(std::string& out, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void appendf(std::string& out, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (n < 0) return;
  if ((size_t)n < sizeof(buf)) {
    out.append(buf, (size_t)n);
    return;
  }
  const size_t at = out.size();
  out.resize(at + (size_t)n + 1);
  va_start(ap, fmt);
  vsnprintf(&out[at], (size_t)n + 1, fmt, ap);
  va_end(ap);
  out.resize(at + (size_t)n);
}

bool write_all(FILE* f, const std::vector<std::string>& bufs) {
  for (const std::string& b : bufs)
    if (!b.empty() && fwrite(b.data(), 1, b.size(), f) != b.size()) return false;
  return true;
}

}  // namespace

// The rows are formatted by the context's host threads (blocks of cells in output order into per-block text buffers,
// the reference's printf formats) and written to the files in order by the calling thread.
int dmx_write_outputs(dmx_ctx* c, const char* const* barcodes, const char* const* samples, const char* prefix,
                      int write_pair) {
  if (!c || !barcodes || !samples || !prefix) return DMX_E_INVALID;
  if (!c->computed) return DMX_E_STATE;
  const int nv = c->cfg.nv, nA = c->cfg.n_alpha;
  const double dp = c->cfg.doublet_prior;
  const int64_t nc = c->ncells;
  const int nth = host_threads(c);
  // std::map<std::string,int32_t> iteration order (:472, :576)
  std::vector<int32_t> ord(nc);
  for (int64_t i = 0; i < nc; ++i) ord[i] = (int32_t)i;
  std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return strcmp(barcodes[a], barcodes[b]) < 0; });
  const std::string pre(prefix);
  const int64_t CB = 16;                                                  // cells per formatting block
  const int64_t slab = std::max<int64_t>(CB, (int64_t)nth * CB * 8);      // cells formatted between two writes
  std::vector<std::string> bufs, bufs2, bufsb;
  bool io_ok = true;
  if (!(c->cfg.flags & DMX_SKIP_SINGLE)) {
    FILE* ws = fopen((pre + ".single").c_str(), "w");
    if (!ws) return DMX_E_IO;                                           // :409
    fprintf(ws, "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n");   // :470
    for (int64_t o0 = 0; o0 < nc && io_ok; o0 += slab) {
      const int64_t o1 = std::min(nc, o0 + slab), nblk = (o1 - o0 + CB - 1) / CB;
      bufs.assign((size_t)nblk, std::string());
      c->pool->run(nblk, nth, [&](int64_t blk, int) {
        std::string& out = bufs[(size_t)blk];
        for (int64_t oi = o0 + blk * CB; oi < std::min(o1, o0 + (blk + 1) * CB); ++oi) {
          const int32_t i = ord[oi];
          if (c->status[i] == 1) continue;                              // :480
          const double* llks = c->r_p1.p + (size_t)i * nv;
          double sumLLK = -1e300;                                       // :478-489
          for (int j = 0; j < nv; ++j) {
            const double cur = llks[j];
            if (sumLLK > cur) sumLLK = sumLLK + log(1.0 + exp(cur - sumLLK));
            else sumLLK = cur + log(1.0 + exp(sumLLK - cur));
          }
          const int nsnp = (int)(c->h_cell_ptr[i + 1] - c->h_cell_ptr[i]);
          for (int j = 0; j < nv; ++j)                                  // :506
            appendf(out, "%s\t%s\t%d\t%d\t%d\t%d\t%.5lf\t%.5lf\t%.3lg\n", barcodes[i], samples[j], c->totl[i],
                    c->npass[i], c->uniq[i], nsnp, llks[j], c->r_p1_0.p[i], exp(llks[j] - sumLLK));
        }
      });
      io_ok = write_all(ws, bufs);
    }
    if (fclose(ws) != 0) io_ok = false;
    if (!io_ok) return DMX_E_IO;
  }
  FILE* w2 = fopen((pre + ".sing2").c_str(), "w");
  FILE* wb = fopen((pre + ".best").c_str(), "w");
  FILE* wp = write_pair ? fopen((pre + ".pair").c_str(), "w") : nullptr;
  if (!w2 || !wb || (write_pair && !wp)) {                              // :535
    if (w2) fclose(w2);
    if (wb) fclose(wb);
    if (wp) fclose(wp);
    return DMX_E_IO;
  }
  fprintf(w2, "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n");     // :533
  if (wp) fprintf(wp, "BARCODE\tSM1.ID\tSM2.ID\tLLK12\tPOSTPRB\n");                           // :570
  fprintf(wb,
          "BARCODE\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tBEST\tSNG.1ST\tSNG.LLK1\tSNG.2ND\tSNG.LLK2\tSNG.LLK0\tDBL.1ST\t"
          "DBL.2ND\tALPHA\tLLK12\tLLK1\tLLK2\tLLK10\tLLK20\tLLK00\tPRB.DBL\tPRB.SNG1\n");      // :571
  // cells with a row, in output order (:581, :592)
  std::vector<int32_t> live;
  live.reserve((size_t)nc);
  for (int64_t oi = 0; oi < nc; ++oi)
    if (c->status[ord[oi]] == 0) live.push_back(ord[oi]);
  const int64_t nl = (int64_t)live.size();
  const size_t per = (size_t)nv * nv * nA;
  std::vector<double> dense;
  // with --write-pair a slab is what one dense re-scan covers (64 MB of llksAB)
  const int64_t pair_batch = std::max<int64_t>(1, ((int64_t)64 << 20) / (int64_t)(per * sizeof(double)));
  const int64_t step = wp ? pair_batch : slab;
  const int64_t cb = wp ? 1 : CB;
  std::vector<std::string> bufsp;
  int rc = DMX_OK;
  for (int64_t l0 = 0; l0 < nl && rc == DMX_OK && io_ok; l0 += step) {
    const int64_t l1 = std::min(nl, l0 + step), nblk = (l1 - l0 + cb - 1) / cb;
    if (wp) {
      std::vector<int32_t> cells(live.begin() + l0, live.begin() + l1);
      dense.resize(cells.size() * per);
      rc = dense_scan(c, cells, dense.data());
      if (rc != DMX_OK) break;
      bufsp.assign((size_t)nblk, std::string());
    }
    bufs2.assign((size_t)nblk, std::string());
    bufsb.assign((size_t)nblk, std::string());
    c->pool->run(nblk, nth, [&](int64_t blk, int) {
      std::string& o2 = bufs2[(size_t)blk];
      std::string& ob = bufsb[(size_t)blk];
      for (int64_t li = l0 + blk * cb; li < std::min(l1, l0 + (blk + 1) * cb); ++li) {
        const int32_t i = live[(size_t)li];
        const dmx_best_rec& b = c->best.p[i];
        const double* sing = c->r_sing.p + (size_t)i * nv;
        const double l00_0 = c->r_l00.p[(size_t)i * nA];
        // exact values where the finaliser produced them
        auto sv = [&](int j) {
          if (j == b.sng1) return b.sng_llk1;
          if (j == b.sng2) return b.sng_llk2;
          if (j == b.dbl1) return b.llk1;
          if (j == b.dbl2) return b.llk2;
          return sing[j];
        };
        for (int j = 0; j < nv; ++j)                                      // :760-769
          appendf(o2, "%s\t%s\t%d\t%d\t%d\t%d\t%.4lf\t%.4lf\t%.3lg\n", barcodes[i], samples[j], c->totl[i], c->npass[i],
                  c->uniq[i], b.n_snp, sv(j), l00_0, exp(sv(j) - b.max_llk) * (1. - dp) / nv / b.sum_single);
        if (wp) {                                                         // :772-797
          std::string& op = bufsp[(size_t)blk];
          const double* AB = dense.data() + (size_t)(li - l0) * per;
          for (int j = 0; j < nv; ++j) {
            appendf(op, "%s\t%s\t%s\t%.3lf\t%.5lf\t%.5lg\n", barcodes[i], samples[j], samples[j], c->alpha[0], sv(j),
                    exp(sv(j) - b.max_llk) * (1. - dp) / nv / (b.sum_single + b.sum_double));
            for (int k = 0; k < nv; ++k)
              for (int n = 0; n < nA; ++n)
                if (n > 0 && j != k) {
                  if (j > k && c->alpha[n] == 0.5) continue;              // :785
                  double v = AB[((size_t)j * nv + k) * nA + n];
                  if (j == b.dbl1 && k == b.dbl2 && n == b.alpha_idx) v = b.llk12;
                  appendf(op, "%s\t%s\t%s\t%.3lf\t%.5lf\t%.5lg\n", barcodes[i], samples[j], samples[k], c->alpha[n], v,
                          exp(v - b.max_llk) * dp / nv / (nv - 1) / (nA - 1) / (b.sum_single + b.sum_double));
                }
          }
        }
        appendf(ob, "%s\t%d\t%d\t%d\t%d\t", barcodes[i], c->totl[i], c->npass[i], c->uniq[i], b.n_snp);   // :830
        if (b.call == 0)
          appendf(ob, "DBL-%s-%s-%.3lf", samples[b.dbl1], samples[b.dbl2], c->alpha[b.alpha_idx]);       // :838
        else if (b.call == 1)
          appendf(ob, "SNG-%s", samples[b.sng1]);                                                          // :845
        else
          appendf(ob, "AMB-%s-%s-%s/%s", samples[b.sng1], samples[b.sng2], samples[b.dbl1], samples[b.dbl2]);   // :851
        appendf(ob, "\t%s\t%.4lf", samples[b.sng1], b.sng_llk1);
        appendf(ob, "\t%s\t%.4lf\t%.4lf", samples[b.sng2], b.sng_llk2, b.sng_llk0);
        appendf(ob, "\t%s\t%s\t%.3lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.4lf\t%.3lg\t%.3lg\n", samples[b.dbl1],
                samples[b.dbl2], c->alpha[b.alpha_idx], b.llk12, b.llk1, b.llk2, b.llk10, b.llk20, b.llk00, b.prb_dbl,
                b.prb_sng1);
      }
    });
    io_ok = write_all(w2, bufs2) && write_all(wb, bufsb) && (!wp || write_all(wp, bufsp));
  }
  if (fclose(w2) != 0) io_ok = false;
  if (fclose(wb) != 0) io_ok = false;
  if (wp && fclose(wp) != 0) io_ok = false;
  if (rc != DMX_OK) return rc;
  return io_ok ? DMX_OK : DMX_E_IO;
}

// ---- VCF field -> genotype-probability matrix (kernels in dmx_field.cu) -------------------------------------------
namespace {

size_t field_in_bytes(int field, int nv) { return field == DMX_FIELD_GT ? (size_t)nv * 2 : (size_t)nv * 12; }

int field_device_ok(int device, int* sm_count) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_last_error = "no CUDA device";
    return DMX_E_NODEVICE;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) {
    g_last_error = "device is not compute capability 10.x";
    return DMX_E_NODEVICE;
  }
  CU(cudaSetDevice(device));
  CU(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
  return DMX_OK;
}

// phred2Prob (PhredHelper.cpp:32) with the host libm's pow(): the table the reference's EM multiplies by
void field_prob_table(double* t) {
  for (int i = 0; i <= 255; ++i) t[i] = pow(0.1, i * 0.1);
}

// scratch shared by the field entry points of one call: the phred table and the error flag
struct FieldScratch {
  double* d_prob = nullptr;
  int* d_flag = nullptr;
  ~FieldScratch() {
    if (d_prob) cudaFree(d_prob);
    if (d_flag) cudaFree(d_flag);
  }
  int init(cudaStream_t st) {
    double t[256];
    field_prob_table(t);
    CU(cudaMalloc((void**)&d_prob, sizeof(t)));
    CU(cudaMalloc((void**)&d_flag, sizeof(int)));
    CU(cudaMemcpyAsync(d_prob, t, sizeof(t), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    CU(cudaStreamSynchronize(st));   // t lives on this stack frame
    return DMX_OK;
  }
  int check(cudaStream_t st) {       // after the kernels: a GT allele index > 1 is not a biallelic record
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return flag ? DMX_E_INVALID : DMX_OK;
  }
};

// per-device cache for dmx_build_gp_device, which must not allocate or synchronise per call
struct FieldDevCache {
  double* d_prob = nullptr;
  int* d_flag = nullptr;
  int sm_count = 0;
};
FieldDevCache g_field_cache[64];
std::mutex g_field_mu;

}  // namespace

int dmx_build_gp_device(int32_t device, int32_t field, int32_t nv, int64_t nsnps, const void* d_in, double gt_error,
                        float* d_out, int64_t out_stride, double* d_af, void* stream) {
  if (field < DMX_FIELD_GT || field > DMX_FIELD_PL || nv < 1 || nv > 4095 || nsnps < 0 || out_stride < (int64_t)nv * 3)
    return DMX_E_INVALID;
  if (nsnps > 0 && (!d_in || !d_out)) return DMX_E_INVALID;
  if (device < 0 || device >= 64) return DMX_E_INVALID;
  FieldDevCache* fc = &g_field_cache[device];
  {
    std::lock_guard<std::mutex> lk(g_field_mu);
    if (!fc->d_prob) {
      int smc = 0;
      RC(field_device_ok(device, &smc));
      double t[256];
      field_prob_table(t);
      CU(cudaMalloc((void**)&fc->d_flag, sizeof(int)));
      CU(cudaMemset(fc->d_flag, 0, sizeof(int)));
      double* dp = nullptr;
      CU(cudaMalloc((void**)&dp, sizeof(t)));
      CU(cudaMemcpy(dp, t, sizeof(t), cudaMemcpyHostToDevice));
      fc->sm_count = smc;
      fc->d_prob = dp;
    }
  }
  CU(cudaSetDevice(device));
  FieldParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.field = field; fp.nv = nv; fp.nsnps = nsnps; fp.in = d_in; fp.out = d_out; fp.out_stride = out_stride;
  fp.af = d_af; fp.gt_error = gt_error; fp.prob = fc->d_prob; fp.err_flag = fc->d_flag;
  CU(launch_field(fp, fc->sm_count, stream));
  return DMX_OK;
}

int dmx_build_gp(int32_t device, int32_t field, int32_t nv, int64_t nsnps, const void* in, double gt_error,
                 float* gp_out, double* af_out, dmx_field_stats* stats) {
  if (field < DMX_FIELD_GT || field > DMX_FIELD_PL || nv < 1 || nv > 4095 || nsnps < 0) return DMX_E_INVALID;
  if (nsnps > 0 && (!in || !gp_out)) return DMX_E_INVALID;
  int sm_count = 0;
  RC(field_device_ok(device, &sm_count));
  const double t0 = now_ms();
  const size_t in_row = field_in_bytes(field, nv), out_row = (size_t)nv * 12;
  // chunks of ~16 MB of output: copies of one chunk overlap the kernel of the other (two streams, two buffer sets)
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(nsnps, 1), (16 << 20) / (int64_t)out_row));
  struct Side {
    cudaStream_t st = nullptr;
    void* d_in = nullptr;
    float* d_out = nullptr;
    double* d_af = nullptr;
    ~Side() {
      if (d_in) cudaFree(d_in);
      if (d_out) cudaFree(d_out);
      if (d_af) cudaFree(d_af);
      if (st) cudaStreamDestroy(st);
    }
  } side[2];
  std::vector<cudaEvent_t> evs;
  struct EvGuard {
    std::vector<cudaEvent_t>* v;
    ~EvGuard() { for (cudaEvent_t e : *v) cudaEventDestroy(e); }
  } evguard{&evs};
  const int nside = nsnps > chunk ? 2 : 1;
  for (int i = 0; i < nside; ++i) {
    CU(cudaStreamCreateWithFlags(&side[i].st, cudaStreamNonBlocking));
    CU(cudaMalloc(&side[i].d_in, (size_t)chunk * in_row));
    CU(cudaMalloc((void**)&side[i].d_out, (size_t)chunk * out_row));
    if (af_out) CU(cudaMalloc((void**)&side[i].d_af, (size_t)chunk * sizeof(double)));
  }
  FieldScratch sc;
  RC(sc.init(side[0].st));
  int64_t launches = 0;
  int k = 0;
  for (int64_t s0 = 0; s0 < nsnps; s0 += chunk, ++k) {
    Side& sd = side[k % nside];
    const int64_t n = std::min(chunk, nsnps - s0);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); evs.push_back(e0);
    CU(cudaEventCreate(&e1)); evs.push_back(e1);
    CU(cudaMemcpyAsync(sd.d_in, (const char*)in + (size_t)s0 * in_row, (size_t)n * in_row, cudaMemcpyHostToDevice, sd.st));
    FieldParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.field = field; fp.nv = nv; fp.nsnps = n; fp.in = sd.d_in; fp.out = sd.d_out; fp.out_stride = (int64_t)nv * 3;
    fp.af = af_out ? sd.d_af : nullptr; fp.gt_error = gt_error; fp.prob = sc.d_prob; fp.err_flag = sc.d_flag;
    CU(cudaEventRecord(e0, sd.st));
    CU(launch_field(fp, sm_count, sd.st));
    CU(cudaEventRecord(e1, sd.st));
    ++launches;
    CU(cudaMemcpyAsync((char*)gp_out + (size_t)s0 * out_row, sd.d_out, (size_t)n * out_row, cudaMemcpyDeviceToHost, sd.st));
    if (af_out && field != DMX_FIELD_GP)
      CU(cudaMemcpyAsync(af_out + s0, sd.d_af, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, sd.st));
  }
  for (int i = 0; i < nside; ++i) CU(cudaStreamSynchronize(side[i].st));
  const int rc = sc.check(side[0].st);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (size_t i = 0; i + 1 < evs.size(); i += 2) {
      float ms = 0;
      cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
      stats->kernel_ms += ms;
    }
    stats->total_ms = now_ms() - t0;
    stats->h2d_bytes = (int64_t)nsnps * (int64_t)in_row;
    stats->d2h_bytes = (int64_t)nsnps * (int64_t)(out_row + (af_out && field != DMX_FIELD_GP ? 8 : 0));
    stats->gpu_launches = launches;
    stats->algo_bytes = (double)nsnps * (double)(in_row + out_row);
  }
  if (rc != DMX_OK) g_last_error = "GT allele index > 1: not a biallelic record";
  return rc;
}

int dmx_site_filter(int32_t device, int32_t nv, int64_t nsnps, const int8_t* gt, int32_t min_mac, double min_callrate,
                    uint8_t* pass_out) {
  if (nv < 1 || nv > 4095 || nsnps < 0 || (nsnps > 0 && (!gt || !pass_out))) return DMX_E_INVALID;
  int sm_count = 0;
  RC(field_device_ok(device, &sm_count));
  if (nsnps == 0) return DMX_OK;
  cudaStream_t st = nullptr;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StGuard { cudaStream_t s; ~StGuard() { cudaStreamDestroy(s); } } stg{st};
  DevBuf<unsigned char> d_in, d_pass;
  struct Rel { DevBuf<unsigned char>*a, *b; ~Rel() { a->release(); b->release(); } } rel{&d_in, &d_pass};
  CU(d_in.reserve((size_t)nsnps * nv * 2));
  CU(d_pass.reserve((size_t)nsnps));
  FieldScratch sc;
  RC(sc.init(st));
  CU(cudaMemcpyAsync(d_in.p, gt, (size_t)nsnps * nv * 2, cudaMemcpyHostToDevice, st));
  FieldParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.field = DMX_FIELD_GT; fp.nv = nv; fp.nsnps = nsnps; fp.in = d_in.p; fp.out = nullptr; fp.out_stride = (int64_t)nv * 3;
  fp.prob = sc.d_prob; fp.err_flag = sc.d_flag; fp.pass = d_pass.p; fp.min_mac = min_mac; fp.min_callrate = min_callrate;
  CU(launch_field(fp, sm_count, st));
  CU(cudaMemcpyAsync(pass_out, d_pass.p, (size_t)nsnps, cudaMemcpyDeviceToHost, st));
  const int rc = sc.check(st);
  if (rc != DMX_OK) g_last_error = "GT allele index > 1: not a biallelic record";
  return rc;
}

int dmx_set_field(dmx_ctx* c, int32_t field, const void* in, int64_t nsnps, double gt_error) {
  if (!c || nsnps < 0 || (nsnps > 0 && !in) || field < DMX_FIELD_GT || field > DMX_FIELD_PL) return DMX_E_INVALID;
  if (!c->kids.empty()) {
    // every shard builds the matrix on its own device from the same FORMAT values
    RC(for_each_kid(c, [&](int k) { return dmx_set_field(c->kids[(size_t)k], field, in, nsnps, gt_error); }));
    c->nsnps = nsnps;
    c->h_gp = c->kids[0]->h_gp;
    c->have_gp = true;
    c->computed = false;
    return DMX_OK;
  }
  CU(cudaSetDevice(c->cfg.device));
  const int nv = c->cfg.nv, nvp = c->nvp;
  const size_t in_bytes = (size_t)nsnps * field_in_bytes(field, nv), n = (size_t)nsnps * nv * 3;
  int sm_count = 0;
  CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->cfg.device));
  CU(cudaStreamSynchronize(c->stream));
  c->nsnps = nsnps;
  CU(c->d_gp.reserve((size_t)nsnps * nvp * 3 + 4));
  CU(c->o_gp.reserve(n));
  DevBuf<unsigned char> d_in;
  struct Rel { DevBuf<unsigned char>* b; ~Rel() { b->release(); } } rel{&d_in};
  CU(d_in.reserve(in_bytes));
  FieldScratch sc;
  RC(sc.init(c->stream));
  CU(cudaMemcpyAsync(d_in.p, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
  if (nvp != nv) CU(cudaMemsetAsync(c->d_gp.p, 0, (size_t)nsnps * nvp * 3 * sizeof(float), c->stream));
  FieldParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.field = field; fp.nv = nv; fp.nsnps = nsnps; fp.in = d_in.p; fp.out = c->d_gp.p; fp.out_stride = (int64_t)nvp * 3;
  fp.gt_error = gt_error; fp.prob = sc.d_prob; fp.err_flag = sc.d_flag;
  CU(launch_field(fp, sm_count, c->stream));
  if (nsnps)
    CU(cudaMemcpy2DAsync(c->o_gp.p, (size_t)nv * 12, c->d_gp.p, (size_t)nvp * 12, (size_t)nv * 12, (size_t)nsnps,
                         cudaMemcpyDeviceToHost, c->stream));
  const int rc = sc.check(c->stream);
  c->have_gp = false;
  if (rc != DMX_OK) {
    g_last_error = "GT allele index > 1: not a biallelic record";
    return rc;
  }
  c->h_gp = c->o_gp.p;
  c->have_gp = true;
  c->gp_dirty = false;     // the device copy is already in place: dmx_upload / dmx_run skip the matrix
  c->gptab_state = 0;
  c->uploaded = false;
  c->computed = false;
  return DMX_OK;
}

}  // extern "C"

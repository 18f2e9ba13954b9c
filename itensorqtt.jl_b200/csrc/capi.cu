// extern "C" entry points of libqtt_b200.so (include/qtt_b200.h).  Plain pointers and sizes only; every C++ exception is
// caught here and turned into a status code + message.
#include "common.cuh"
#include <atomic>
#include <mutex>
#include <set>
#include <map>
#include <unordered_map>

namespace qtt {
void two_site_sweeps(Ctx* ctx, const Mpo& A, const Mps* b, Mps& x, const qtt_sweep_params& P, qtt_sweep_info* infos);

static std::string g_error;
static std::mutex g_mutex;

// ---- caching device allocator ---------------------------------------------------------------------------------
// cudaMalloc / cudaFree inside a sweep cost 20-100 ms spikes (measured: environment updates of the first round-1 build)
// and cudaFree synchronises the device.  Blocks are therefore recycled through per-device free lists; a block
// remembers the stream it was last used on, so re-use on the same stream is ordered by the stream itself and re-use
// on another stream first waits for the old one.  Everything is released by qtt_pool_trim() / at process exit.
namespace {
struct PoolBlock { void* p; size_t bytes; cudaStream_t st; };
struct DevicePool { std::multimap<size_t, PoolBlock> free_blocks; std::unordered_map<void*, size_t> live; size_t cached = 0; };
std::mutex g_pool_mutex;
std::set<cudaStream_t> g_live_streams;  // streams of live contexts (guarded by g_pool_mutex)
std::atomic<int> g_live_ctx[64];         // live contexts per device: kernels that fill the whole GPU size themselves down when > 1
std::map<int, DevicePool> g_pools;
thread_local Ctx* tl_ctx = nullptr;

size_t pool_round(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes <= (size_t(1) << 20)) {  // next power of two
    size_t r = 512;
    while (r < bytes) r <<= 1;
    return r;
  }
  const size_t q = size_t(1) << 20;
  return (bytes + q - 1) / q * q;
}
}  // namespace

int live_contexts(int device) { return (device >= 0 && device < 64) ? g_live_ctx[device].load() : 1; }

void set_current_ctx(Ctx* c) { tl_ctx = c; }

void* pool_alloc(size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t want = pool_round(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    DevicePool& P = g_pools[dev];
    auto it = P.free_blocks.find(want);
    if (it != P.free_blocks.end()) {
      PoolBlock b = it->second;
      P.free_blocks.erase(it);
      P.cached -= b.bytes;
      P.live[b.p] = b.bytes;
      if (b.st != st && b.st != nullptr) cudaStreamSynchronize(b.st);
      return b.p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    pool_trim();  // give the cached blocks back and retry once
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) fail(QTT_ERR_ALLOC, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  g_pools[dev].live[p] = want;
  return p;
}

void pool_free(void* p, cudaStream_t st) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  DevicePool& P = g_pools[dev];
  auto it = P.live.find(p);
  if (it == P.live.end()) {  // not ours (or another device is current): fall back to the driver
    cudaFree(p);
    return;
  }
  const size_t bytes = it->second;
  P.live.erase(it);
  // a chain may outlive the context that allocated it: a block tagged with a destroyed stream must never be synchronised on
  if (st != nullptr && g_live_streams.count(st) == 0) st = nullptr;
  P.free_blocks.emplace(bytes, PoolBlock{p, bytes, st});
  P.cached += bytes;
}

void pool_trim() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  DevicePool& P = g_pools[dev];
  cudaDeviceSynchronize();
  for (auto& kv : P.free_blocks) cudaFree(kv.second.p);
  P.free_blocks.clear();
  P.cached = 0;
}

// a stream is about to be destroyed: its cached blocks become quiescent
static void pool_forget_stream(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  g_live_streams.erase(st);
  for (auto& dp : g_pools)
    for (auto& kv : dp.second.free_blocks)
      if (kv.second.st == st) kv.second.st = nullptr;
}

void dbuf_reserve(DBuf& b, size_t n) {
  if (n <= b.cap && b.p) return;
  size_t cap = n < 16 ? 16 : n;
  cudaStream_t st = tl_ctx ? tl_ctx->stream : nullptr;
  double* np = static_cast<double*>(pool_alloc(cap * sizeof(double), st));
  if (b.p) pool_free(b.p, b.st);  // earlier readers of the old block are ahead of any later writer in stream order
  b.p = np;
  b.cap = cap;
  b.st = st;
}
void dbuf_free(DBuf& b) {
  if (b.p) pool_free(b.p, b.st);
  b.p = nullptr;
  b.cap = 0;
}

Mps::~Mps() { for (auto& c : core) dbuf_free(c); }
Mpo::~Mpo() { for (auto& c : core) dbuf_free(c); }

void Ctx::ensure_arena(size_t bytes) {
  if (bytes <= arena.cap) return;
  if (arena.top != 0) fail(QTT_ERR_ALLOC, "internal: arena growth requested while in use");
  if (arena.base) pool_free(arena.base, stream);
  size_t cap = bytes + bytes / 4 + (1 << 20);
  arena.base = nullptr;
  arena.cap = 0;
  arena.base = static_cast<char*>(pool_alloc(cap, stream));
  arena.cap = cap;
}

// RAII device scratch for the fine-grained entry points
struct Scratch {
  std::vector<double*> ptrs;
  cudaStream_t st = tl_ctx ? tl_ctx->stream : nullptr;
  double* get(size_t n) {
    double* p = static_cast<double*>(pool_alloc((n < 2 ? 2 : n) * sizeof(double), st));
    ptrs.push_back(p);
    return p;
  }
  double* upload(Ctx* ctx, const double* h, size_t n) {
    double* p = get(n);
    QTT_CUDA(cudaMemcpyAsync(p, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return p;
  }
  ~Scratch() { for (double* p : ptrs) pool_free(p, st); }
};

// host column-major (d0, d1, ...) == row-major [.., d1, d0]; bring to the row-major order `perm` of those reversed dims
double* upload_permuted(Ctx* ctx, Scratch& sc, const double* h, int nd, const int64_t* dims_colmajor, const int* perm_rowmajor) {
  size_t n = 1;
  int64_t rdims[6];
  for (int d = 0; d < nd; ++d) { n *= dims_colmajor[d]; rdims[d] = dims_colmajor[nd - 1 - d]; }
  double* raw = sc.upload(ctx, h, n);
  double* out = sc.get(n);
  permute(ctx, raw, out, nd, rdims, perm_rowmajor);
  return out;
}

template <typename F>
float timed_reps(Ctx* ctx, int reps, F&& body) {
  if (reps < 1) reps = 1;
  body();  // warm-up (also the run whose result is returned when reps == 1)
  if (reps == 1) return 0.f;
  QTT_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int r = 0; r < reps; ++r) body();
  QTT_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  QTT_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  QTT_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  return ms / reps;
}

}  // namespace qtt

using namespace qtt;

#define QTT_TRY(ctxp) try { qtt::set_current_ctx(ctxp);
#define QTT_CATCH(ctxp)                                                     \
  }                                                                         \
  catch (const qtt::Error& e) {                                             \
    if (ctxp) (ctxp)->err = e.what();                                       \
    { std::lock_guard<std::mutex> lk(g_mutex); g_error = e.what(); }        \
    return e.code;                                                          \
  }                                                                         \
  catch (const std::exception& e) {                                         \
    if (ctxp) (ctxp)->err = e.what();                                       \
    { std::lock_guard<std::mutex> lk(g_mutex); g_error = e.what(); }        \
    return QTT_ERR_ALLOC;                                                   \
  }                                                                         \
  return QTT_OK;

extern "C" {

int qtt_version(void) { return QTT_B200_VERSION; }
const char* qtt_global_error(void) { return g_error.c_str(); }
const char* qtt_last_error(const qtt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_error.c_str(); }

qtt_ctx* qtt_ctx_create(int device) {
  qtt_ctx* c = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(QTT_ERR_CUDA, "no CUDA device available (%s); libqtt_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) fail(QTT_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    QTT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QTT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) fail(QTT_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    c = new qtt_ctx();
    c->device = device;
    c->sms = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->clock_khz = prop.clockRate;
    QTT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
      std::lock_guard<std::mutex> lk(g_pool_mutex);
      g_live_streams.insert(c->stream);
    }
    if (device >= 0 && device < 64) { g_live_ctx[device].fetch_add(1); c->counted = true; }
    QTT_CUDA(cudaMalloc(&c->d_scal, 16384 * sizeof(double)));
    QTT_CUDA(cudaMemset(c->d_scal, 0, 16384 * sizeof(double)));
    QTT_CUDA(cudaMalloc(&c->d_sync, 256 * sizeof(unsigned)));
    QTT_CUDA(cudaMemset(c->d_sync, 0, 256 * sizeof(unsigned)));
    QTT_CUDA(cudaHostAlloc(&c->h_scal, 4096 * sizeof(double), cudaHostAllocMapped));
    memset(c->h_scal, 0, 4096 * sizeof(double));
    QTT_CUDA(cudaEventCreate(&c->ev0));
    QTT_CUDA(cudaEventCreate(&c->ev1));
    c->partial(1 << 17);
    return c;
  } catch (const std::exception& e) {
    std::lock_guard<std::mutex> lk(g_mutex);
    g_error = e.what();
    delete c;
    return nullptr;
  }
}

void qtt_ctx_destroy(qtt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->arena.base) pool_free(c->arena.base, c->stream);
  if (c->d_scal) cudaFree(c->d_scal);
  if (c->d_partial) cudaFree(c->d_partial);
  if (c->d_sync) cudaFree(c->d_sync);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) { pool_forget_stream(c->stream); cudaStreamDestroy(c->stream); }
  if (c->counted && c->device >= 0 && c->device < 64) g_live_ctx[c->device].fetch_sub(1);
  pool_trim();
  delete c;
}

int qtt_ctx_sync(qtt_ctx* ctx) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx, "null ctx");
  QTT_CUDA(cudaSetDevice(ctx->device));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int64_t qtt_ctx_launch_count(const qtt_ctx* ctx) { return ctx ? ctx->launches : 0; }

int qtt_ctx_device_info(const qtt_ctx* ctx, int64_t* out4) {
  if (!ctx || !out4) return QTT_ERR_ARG;
  out4[0] = ctx->sms; out4[1] = ctx->cc_major; out4[2] = ctx->cc_minor; out4[3] = ctx->clock_khz;
  return QTT_OK;
}

int qtt_set_profiling(qtt_ctx* ctx, int enabled) {
  if (!ctx) return QTT_ERR_ARG;
  ctx->profiling = enabled != 0;
  return QTT_OK;
}

// ---- chains ----------------------------------------------------------------------------------
static void upload_core(Ctx* ctx, DBuf& dst, const void* host, int nd, const int64_t* dims_colmajor, const int* perm, int dtype) {
  size_t n = 1;
  int64_t rdims[6];
  for (int d = 0; d < nd; ++d) { n *= dims_colmajor[d]; rdims[d] = dims_colmajor[nd - 1 - d]; }
  const int planes = dtype == QTT_C128 ? 2 : 1;
  dbuf_reserve(dst, n * planes);
  Scratch sc;
  if (dtype == QTT_F64) {
    double* raw = sc.upload(ctx, static_cast<const double*>(host), n);
    permute(ctx, raw, dst.p, nd, rdims, perm);
  } else {
    double* raw = sc.upload(ctx, static_cast<const double*>(host), 2 * n);
    double* re = sc.get(n);
    double* im = sc.get(n);
    deinterleave(ctx, raw, re, im, n);
    permute(ctx, re, dst.p, nd, rdims, perm);
    permute(ctx, im, dst.p + n, nd, rdims, perm);
  }
  // no synchronisation here: the staging blocks return to the stream-ordered pool, the caller syncs once per chain
}

int qtt_mps_upload(qtt_ctx* ctx, int n, const int64_t* chi, const void* const* cores, int dtype, qtt_mps** out) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && chi && cores && out && n >= 1, "qtt_mps_upload: bad arguments");
  QTT_REQUIRE(dtype == QTT_F64 || dtype == QTT_C128, "unsupported dtype %d", dtype);
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_mps* m = new qtt_mps();
  m->ctx = ctx; m->n = n; m->dtype = dtype;
  m->chi.assign(chi, chi + n + 1);
  m->core.resize(n);
  try {
    for (int j = 0; j < n; ++j) {
      QTT_REQUIRE(chi[j] >= 1 && chi[j + 1] >= 1 && cores[j], "bad core %d", j);
      // host col-major (a, s, c) == row-major [c][s][a]  ->  device [a][s][c]
      int64_t dims[3] = {chi[j], 2, chi[j + 1]};
      int perm[3] = {2, 1, 0};
      upload_core(ctx, m->core[j], cores[j], 3, dims, perm, dtype);
    }
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));  // host arrays are only borrowed for the duration of the call
  } catch (...) { delete m; throw; }
  *out = m;
  QTT_CATCH(ctx)
}

int qtt_mpo_upload(qtt_ctx* ctx, int n, const int64_t* w, const void* const* cores, int dtype, qtt_mpo** out) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && w && cores && out && n >= 1, "qtt_mpo_upload: bad arguments");
  QTT_REQUIRE(dtype == QTT_F64 || dtype == QTT_C128, "unsupported dtype %d", dtype);
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_mpo* m = new qtt_mpo();
  m->ctx = ctx; m->n = n; m->dtype = dtype;
  m->w.assign(w, w + n + 1);
  m->core.resize(n);
  try {
    for (int j = 0; j < n; ++j) {
      QTT_REQUIRE(w[j] >= 1 && w[j + 1] >= 1 && cores[j], "bad core %d", j);
      // host col-major (l, r, s, t) == row-major [t][s][r][l]  ->  device [l][r][s][t]
      int64_t dims[4] = {w[j], w[j + 1], 2, 2};
      int perm[4] = {3, 2, 1, 0};
      upload_core(ctx, m->core[j], cores[j], 4, dims, perm, dtype);
    }
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) { delete m; throw; }
  *out = m;
  QTT_CATCH(ctx)
}

int qtt_mps_length(const qtt_mps* x) { return x ? x->n : QTT_ERR_ARG; }
int qtt_mps_dtype(const qtt_mps* x) { return x ? x->dtype : QTT_ERR_ARG; }
int qtt_mps_dims(const qtt_mps* x, int64_t* chi_out) {
  if (!x || !chi_out) return QTT_ERR_ARG;
  for (int j = 0; j <= x->n; ++j) chi_out[j] = x->chi[j];
  return QTT_OK;
}
int qtt_mpo_length(const qtt_mpo* A) { return A ? A->n : QTT_ERR_ARG; }
int qtt_mpo_dims(const qtt_mpo* A, int64_t* w_out) {
  if (!A || !w_out) return QTT_ERR_ARG;
  for (int j = 0; j <= A->n; ++j) w_out[j] = A->w[j];
  return QTT_OK;
}

int qtt_mps_download(const qtt_mps* x, void* const* cores_out) {
  qtt_ctx* ctx = x ? static_cast<qtt_ctx*>(x->ctx) : nullptr;
  QTT_TRY(ctx)
  QTT_REQUIRE(x && cores_out, "qtt_mps_download: bad arguments");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;  // staging for all cores; one synchronisation at the end
  for (int j = 0; j < x->n; ++j) {
    const int64_t a = x->chi[j], c = x->chi[j + 1];
    const size_t n = (size_t)a * 2 * c;
    int64_t rdims[3] = {a, 2, c};  // device [a][s][c] -> row-major [c][s][a] == col-major (a, s, c)
    int perm[3] = {2, 1, 0};
    if (x->dtype == QTT_F64) {
      double* tmp = sc.get(n);
      permute(ctx, x->core[j].p, tmp, 3, rdims, perm);
      QTT_CUDA(cudaMemcpyAsync(cores_out[j], tmp, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      double* re = sc.get(n);
      double* im = sc.get(n);
      double* il = sc.get(2 * n);
      permute(ctx, x->core[j].p, re, 3, rdims, perm);
      permute(ctx, x->core[j].p + n, im, 3, rdims, perm);
      interleave(ctx, re, im, il, n);
      QTT_CUDA(cudaMemcpyAsync(cores_out[j], il, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int qtt_mps_clone(const qtt_mps* x, qtt_mps** out) {
  qtt_ctx* ctx = x ? static_cast<qtt_ctx*>(x->ctx) : nullptr;
  QTT_TRY(ctx)
  QTT_REQUIRE(x && out, "qtt_mps_clone: bad arguments");
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_mps* m = new qtt_mps();
  m->ctx = x->ctx; m->n = x->n; m->dtype = x->dtype; m->chi = x->chi;
  m->core.resize(x->n);
  try {
    for (int j = 0; j < x->n; ++j) {
      size_t n = (size_t)x->chi[j] * 2 * x->chi[j + 1] * (x->dtype == QTT_C128 ? 2 : 1);
      dbuf_reserve(m->core[j], n);
      QTT_CUDA(cudaMemcpyAsync(m->core[j].p, x->core[j].p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) { delete m; throw; }
  *out = m;
  QTT_CATCH(ctx)
}

void qtt_mps_free(qtt_mps* x) {
  if (!x) return;
  if (x->ctx) cudaSetDevice(x->ctx->device);
  delete x;
}
void qtt_mpo_free(qtt_mpo* A) {
  if (!A) return;
  if (A->ctx) cudaSetDevice(A->ctx->device);
  delete A;
}

// ---- sweeps ----------------------------------------------------------------------------------
int qtt_linsolve(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* b, qtt_mps* x, const qtt_sweep_params* p, qtt_sweep_info* per_sweep) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && A && b && x && p, "qtt_linsolve: null argument");
  QTT_REQUIRE(A->ctx == ctx && b->ctx == ctx && x->ctx == ctx, "chains belong to a different context");
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_sweep_params P = *p;
  if (P.solver != QTT_SOLVER_DENSE_QR && P.solver != QTT_SOLVER_PG_LSQ) P.solver = QTT_SOLVER_GMRES;
  two_site_sweeps(ctx, *A, b, *x, P, per_sweep);
  QTT_CATCH(ctx)
}

int qtt_dmrg_target(qtt_ctx* ctx, const qtt_mpo* A, qtt_mps* psi, const qtt_sweep_params* p, qtt_sweep_info* per_sweep) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && A && psi && p, "qtt_dmrg_target: null argument");
  QTT_REQUIRE(A->ctx == ctx && psi->ctx == ctx, "chains belong to a different context");
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_sweep_params P = *p;
  if (P.solver == QTT_SOLVER_GMRES) P.solver = QTT_SOLVER_LANCZOS;
  QTT_REQUIRE(P.solver != QTT_SOLVER_DENSE_QR && P.solver != QTT_SOLVER_PG_LSQ, "qtt_dmrg_target: solver %d is a linsolve solver", P.solver);
  two_site_sweeps(ctx, *A, nullptr, *psi, P, per_sweep);
  QTT_CATCH(ctx)
}

int qtt_apply(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* psi, double cutoff, int32_t maxdim, int32_t mindim, qtt_mps** out) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && A && psi && out, "qtt_apply: null argument");
  QTT_REQUIRE(A->ctx == ctx && psi->ctx == ctx, "chains belong to a different context");
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_mps* m = new qtt_mps();
  m->ctx = ctx;
  try {
    apply_mpo_mps(ctx, *A, *psi, cutoff, maxdim, mindim, *m);
  } catch (...) { delete m; throw; }
  *out = m;
  QTT_CATCH(ctx)
}

int qtt_vec_to_mps(qtt_ctx* ctx, const double* v, int32_t n, int32_t on_device, double cutoff, int32_t maxdim, int32_t mindim, qtt_mps** out,
                   double* maxtruncerr) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && v && out, "qtt_vec_to_mps: null argument");
  QTT_REQUIRE(n >= 1 && n <= 34, "qtt_vec_to_mps: n = %d out of range [1, 34]", n);
  QTT_CUDA(cudaSetDevice(ctx->device));
  const size_t total = (size_t)1 << n;
  {
    // the caller states where v lives; a mismatch would otherwise surface as an illegal address inside a kernel
    cudaPointerAttributes pa;
    const cudaError_t e = cudaPointerGetAttributes(&pa, v);
    cudaGetLastError();
    const bool is_dev = (e == cudaSuccess) && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    if (on_device) {
      QTT_REQUIRE(is_dev, "qtt_vec_to_mps: on_device = 1 but v is not a device pointer");
      QTT_REQUIRE(pa.type == cudaMemoryTypeManaged || pa.device == ctx->device, "qtt_vec_to_mps: v lives on device %d, the context on device %d",
                  pa.device, ctx->device);
    } else {
      QTT_REQUIRE(!(e == cudaSuccess && pa.type == cudaMemoryTypeDevice), "qtt_vec_to_mps: on_device = 0 but v is a device pointer");
    }
  }
  double* staged = nullptr;
  if (!on_device) {
    staged = static_cast<double*>(pool_alloc(total * sizeof(double), ctx->stream));
    cudaError_t e = cudaMemcpyAsync(staged, v, total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { pool_free(staged, ctx->stream); QTT_CUDA(e); }
  }
  qtt_mps* m = new qtt_mps();
  try {
    vec_to_mps(ctx, staged ? staged : v, n, cutoff, maxdim, mindim, *m, maxtruncerr);
  } catch (...) {
    delete m;
    if (staged) pool_free(staged, ctx->stream);
    throw;
  }
  if (staged) pool_free(staged, ctx->stream);
  if (total >= ((size_t)1 << 26)) pool_trim();  // do not keep multi-GB staging buffers cached
  *out = m;
  QTT_CATCH(ctx)
}

// ---- sum, dense expansion, point sampling ------------------------------------------------------
int qtt_mps_add(qtt_ctx* ctx, const qtt_mps* x, const qtt_mps* y, double alpha, double beta, double cutoff, int32_t maxdim, int32_t mindim,
                qtt_mps** out) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && x && y && out, "qtt_mps_add: null argument");
  QTT_REQUIRE(x->ctx == ctx && y->ctx == ctx, "chains belong to a different context");
  QTT_CUDA(cudaSetDevice(ctx->device));
  qtt_mps* m = new qtt_mps();
  m->ctx = ctx;
  try {
    mps_add(ctx, *x, *y, alpha, beta, cutoff, maxdim, mindim, *m);
  } catch (...) { delete m; throw; }
  *out = m;
  QTT_CATCH(ctx)
}

int qtt_mps_to_vec(qtt_ctx* ctx, const qtt_mps* x, void* out, int32_t on_device) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && x && out, "qtt_mps_to_vec: null argument");
  QTT_REQUIRE(x->ctx == ctx, "chain belongs to a different context");
  QTT_REQUIRE(x->n >= 1 && x->n <= 34, "qtt_mps_to_vec: n = %d out of range [1, 34]", x->n);
  QTT_CUDA(cudaSetDevice(ctx->device));
  {
    cudaPointerAttributes pa;
    const cudaError_t e = cudaPointerGetAttributes(&pa, out);
    cudaGetLastError();
    const bool is_dev = (e == cudaSuccess) && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    if (on_device) {
      QTT_REQUIRE(is_dev, "qtt_mps_to_vec: on_device = 1 but out is not a device pointer");
      QTT_REQUIRE(pa.type == cudaMemoryTypeManaged || pa.device == ctx->device, "qtt_mps_to_vec: out lives on device %d, the context on device %d",
                  pa.device, ctx->device);
    } else {
      QTT_REQUIRE(!(e == cudaSuccess && pa.type == cudaMemoryTypeDevice), "qtt_mps_to_vec: on_device = 0 but out is a device pointer");
    }
  }
  const size_t total = (size_t)1 << x->n;
  const bool cplx = x->dtype == QTT_C128;
  Scratch sc;
  if (!cplx) {
    double* dst = on_device ? static_cast<double*>(out) : sc.get(total);
    mps_to_vec(ctx, *x, dst, nullptr);
    if (!on_device) QTT_CUDA(cudaMemcpyAsync(out, dst, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    double* re = sc.get(total);
    double* im = sc.get(total);
    mps_to_vec(ctx, *x, re, im);
    double* il = on_device ? static_cast<double*>(out) : sc.get(2 * total);
    interleave(ctx, re, im, il, total);
    if (!on_device) QTT_CUDA(cudaMemcpyAsync(out, il, 2 * total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  // multi-GB staging buffers (host output, planar complex planes) do not stay cached; a real chain written straight into
  // the caller's device buffer staged nothing large
  if (total >= ((size_t)1 << 26) && (!on_device || cplx)) pool_trim();
  QTT_CATCH(ctx)
}

int qtt_mps_sample(qtt_ctx* ctx, const qtt_mps* x, const int64_t* idx, int64_t count, void* out) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && x && (count == 0 || (idx && out)), "qtt_mps_sample: null argument");
  QTT_REQUIRE(x->ctx == ctx, "chain belongs to a different context");
  QTT_REQUIRE(count >= 0, "qtt_mps_sample: negative count");
  QTT_CUDA(cudaSetDevice(ctx->device));
  if (count > 0) {
    const int64_t limit = x->n >= 63 ? INT64_MAX : ((int64_t)1 << x->n);
    for (int64_t i = 0; i < count; ++i)
      QTT_REQUIRE(idx[i] >= 0 && idx[i] < limit, "qtt_mps_sample: grid index %lld (entry %lld) outside [0, 2^%d)", (long long)idx[i],
                  (long long)i, x->n);
    const bool cplx = x->dtype == QTT_C128;
    Scratch sc;
    static_assert(sizeof(int64_t) == sizeof(double), "index staging reuses the double pool");
    int64_t* d_idx = reinterpret_cast<int64_t*>(sc.get((size_t)count));
    QTT_CUDA(cudaMemcpyAsync(d_idx, idx, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    double* d_out = sc.get((size_t)count * (cplx ? 2 : 1));
    mps_sample(ctx, *x, d_idx, count, d_out);
    if (!cplx) {
      QTT_CUDA(cudaMemcpyAsync(out, d_out, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      double* il = sc.get((size_t)2 * count);
      interleave(ctx, d_out, d_out + count, il, (size_t)count);
      QTT_CUDA(cudaMemcpyAsync(out, il, (size_t)2 * count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  QTT_CATCH(ctx)
}

int qtt_inner(qtt_ctx* ctx, const qtt_mps* x, const qtt_mps* y, double* out2) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && x && y && out2, "qtt_inner: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  mps_inner(ctx, *x, *y, out2);
  QTT_CATCH(ctx)
}

int qtt_residual(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* x, const qtt_mps* b, double* sq, double* sqn) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && A && x && b, "qtt_residual: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  mps_residual(ctx, *A, *x, *b, sq, sqn);
  QTT_CATCH(ctx)
}

// ---- fine-grained ------------------------------------------------------------------------------
int qtt_gemm(qtt_ctx* ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
             const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int32_t tile, int32_t reps, float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && A && B && C, "qtt_gemm: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  const size_t nA = (size_t)(transA ? K : M) * lda, nB = (size_t)(transB ? N : K) * ldb, nC = (size_t)M * ldc;
  double* dA = sc.upload(ctx, A, nA);
  double* dB = sc.upload(ctx, B, nB);
  double* dC0 = sc.upload(ctx, C, nC);
  double* dC = sc.get(nC);
  GemmDesc g;
  g.A = dA; g.B = dB; g.C = dC;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.transA = transA != 0; g.transB = transB != 0;
  g.alpha = alpha; g.beta = beta; g.tile = tile;
  float t = timed_reps(ctx, reps, [&] {
    if (beta != 0.0) QTT_CUDA(cudaMemcpyAsync(dC, dC0, nC * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    gemm(ctx, g);
  });
  if (ms) *ms = t;
  QTT_CUDA(cudaMemcpyAsync(C, dC, nC * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int qtt_matvec2(qtt_ctx* ctx, const double* L, const double* W1, const double* W2, const double* R, const double* v, double* out,
                int64_t chil, int64_t chir, int64_t wl, int64_t wm, int64_t wr, int32_t reps, float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && L && W1 && W2 && R && v && out, "qtt_matvec2: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  // L host (a, w, a') col-major == row-major [a'][w][a] -> device [w][a'][a]
  int64_t dL[3] = {chil, wl, chil}; int pL[3] = {1, 0, 2};
  double* dLp = upload_permuted(ctx, sc, L, 3, dL, pL);
  // R host (c, w, c') == row-major [c'][w][c] -> device [w][c][c']
  int64_t dR[3] = {chir, wr, chir}; int pR[3] = {1, 2, 0};
  double* dRp = upload_permuted(ctx, sc, R, 3, dR, pR);
  int64_t dW1[4] = {wl, wm, 2, 2}, dW2[4] = {wm, wr, 2, 2}; int pW[4] = {3, 2, 1, 0};
  double* dW1p = upload_permuted(ctx, sc, W1, 4, dW1, pW);
  double* dW2p = upload_permuted(ctx, sc, W2, 4, dW2, pW);
  // v host (a, s1, s2, c) == row-major [c][s2][s1][a] -> device [a][s1][s2][c]
  int64_t dV[4] = {chil, 2, 2, chir}; int pV[4] = {3, 2, 1, 0};
  double* dv = upload_permuted(ctx, sc, v, 4, dV, pV);
  const size_t len = (size_t)4 * chil * chir;
  double* dout = sc.get(len);
  double* W12 = sc.get((size_t)16 * wl * wr);
  double* scr = sc.get(matvec2_scratch((int)chil, (int)chir, (int)wl, (int)wr));
  build_w12(ctx, dW1p, dW2p, W12, (int)wl, (int)wm, (int)wr);
  float t = timed_reps(ctx, reps, [&] { matvec2(ctx, dLp, W12, dRp, dv, dout, (int)chil, (int)chir, (int)wl, (int)wr, scr); });
  if (ms) *ms = t;
  double* tmp = sc.get(len);
  int64_t rd[4] = {chil, 2, 2, chir}; int pb[4] = {3, 2, 1, 0};
  permute(ctx, dout, tmp, 4, rd, pb);
  QTT_CUDA(cudaMemcpyAsync(out, tmp, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int qtt_env_left(qtt_ctx* ctx, const double* L, const double* x, const double* W, double* Lnew, int64_t chia, int64_t chib, int64_t w,
                 int64_t w2, int32_t reps, float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && L && x && W && Lnew, "qtt_env_left: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  int64_t dL[3] = {chia, w, chia}; int pL[3] = {1, 0, 2};
  double* dLp = upload_permuted(ctx, sc, L, 3, dL, pL);
  int64_t dX[3] = {chia, 2, chib}; int pX[3] = {2, 1, 0};
  double* dx = upload_permuted(ctx, sc, x, 3, dX, pX);
  int64_t dW[4] = {w, w2, 2, 2}; int pW[4] = {3, 2, 1, 0};
  double* dWp = upload_permuted(ctx, sc, W, 4, dW, pW);
  const size_t nout = (size_t)w2 * chib * chib;
  double* dout = sc.get(nout);
  double* scr = sc.get(env_scratch((int)chia, (int)chib, (int)w, (int)w2));
  float t = timed_reps(ctx, reps, [&] { env_left(ctx, dLp, dx, dWp, dx, dout, (int)chia, (int)chib, (int)chia, (int)chib, (int)w, (int)w2, scr); });
  if (ms) *ms = t;
  // device [w2][b'][b] -> host (b, w2, b') col-major == row-major [b'][w2][b]
  double* tmp = sc.get(nout);
  int64_t rd[3] = {w2, chib, chib}; int pb[3] = {1, 0, 2};
  permute(ctx, dout, tmp, 3, rd, pb);
  QTT_CUDA(cudaMemcpyAsync(Lnew, tmp, nout * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int qtt_env_right(qtt_ctx* ctx, const double* R, const double* x, const double* W, double* Rnew, int64_t chia, int64_t chic, int64_t w,
                  int64_t w2, int32_t reps, float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && R && x && W && Rnew, "qtt_env_right: null argument");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  int64_t dR[3] = {chic, w2, chic}; int pR[3] = {1, 2, 0};
  double* dRp = upload_permuted(ctx, sc, R, 3, dR, pR);
  int64_t dX[3] = {chia, 2, chic}; int pX[3] = {2, 1, 0};
  double* dx = upload_permuted(ctx, sc, x, 3, dX, pX);
  int64_t dW[4] = {w, w2, 2, 2}; int pW[4] = {3, 2, 1, 0};
  double* dWp = upload_permuted(ctx, sc, W, 4, dW, pW);
  const size_t nout = (size_t)w * chia * chia;
  double* dout = sc.get(nout);
  double* scr = sc.get(env_scratch((int)chia, (int)chic, (int)w, (int)w2));
  float t = timed_reps(ctx, reps, [&] { env_right(ctx, dRp, dx, dWp, dx, dout, (int)chia, (int)chic, (int)chia, (int)chic, (int)w, (int)w2, scr); });
  if (ms) *ms = t;
  // device [w][a][a'] -> host (a, w, a') col-major == row-major [a'][w][a]
  double* tmp = sc.get(nout);
  int64_t rd[3] = {w, chia, chia}; int pb[3] = {2, 0, 1};
  permute(ctx, dout, tmp, 3, rd, pb);
  QTT_CUDA(cudaMemcpyAsync(Rnew, tmp, nout * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  QTT_CATCH(ctx)
}

int qtt_split2(qtt_ctx* ctx, const double* phi, int64_t chia, int64_t chic, int32_t ortho, double cutoff, int32_t maxdim, int32_t mindim,
               double* A_out, double* B_out, double* sigma_out, int32_t* info_out, double* truncerr_out, int32_t reps, float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && phi && A_out && B_out, "qtt_split2: null argument");
  QTT_REQUIRE(ortho == 0 || ortho == 1, "ortho must be 0 (left) or 1 (right)");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  int64_t dP[4] = {chia, 2, 2, chic}; int pP[4] = {3, 2, 1, 0};
  double* dphi = upload_permuted(ctx, sc, phi, 4, dP, pP);
  const int kmax = (int)(2 * chia < 2 * chic ? 2 * chia : 2 * chic);
  double* dA = sc.get((size_t)2 * chia * kmax);
  double* dB = sc.get((size_t)2 * chic * kmax);
  double* dsig = sc.get(kmax);
  ctx->ensure_arena(((size_t)16 * chia * chic * 4 + 65536) * sizeof(double));
  SplitResult sr;
  float t = timed_reps(ctx, reps, [&] { sr = split2(ctx, dphi, (int)chia, (int)chic, ortho, cutoff, maxdim, mindim, dA, dB, dsig, 0); });
  if (ms) *ms = t;
  const int k = sr.rank;
  // A device [a][s][k] -> host (a, s, k) col-major == row-major [k][s][a]
  double* tA = sc.get((size_t)2 * chia * k);
  int64_t rA[3] = {chia, 2, k}; int pr[3] = {2, 1, 0};
  permute(ctx, dA, tA, 3, rA, pr);
  double* tB = sc.get((size_t)2 * chic * k);
  int64_t rB[3] = {k, 2, chic};
  permute(ctx, dB, tB, 3, rB, pr);
  QTT_CUDA(cudaMemcpyAsync(A_out, tA, sizeof(double) * 2 * chia * k, cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaMemcpyAsync(B_out, tB, sizeof(double) * 2 * chic * k, cudaMemcpyDeviceToHost, ctx->stream));
  if (sigma_out) QTT_CUDA(cudaMemcpyAsync(sigma_out, dsig, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info_out) { info_out[0] = k; info_out[1] = sr.sweeps; }
  if (truncerr_out) *truncerr_out = sr.truncerr;
  QTT_CATCH(ctx)
}

int qtt_krylov_op(qtt_ctx* ctx, int32_t op, const double* V, int64_t nvec, int64_t len, double* w, double* h, double* out, int32_t reps,
                  float* ms) {
  QTT_TRY(ctx)
  QTT_REQUIRE(ctx && V && w && nvec >= 1 && len >= 1 && (op != 4 || (nvec <= 62 && kry_cgs2_fused_ok(ctx, (size_t)len))), "qtt_krylov_op: bad arguments");
  QTT_CUDA(cudaSetDevice(ctx->device));
  Scratch sc;
  double* dV = sc.upload(ctx, V, (size_t)nvec * len);
  double* dw0 = sc.upload(ctx, w, len);
  double* dw = sc.get(len);
  double* dh = sc.get(nvec + 2);
  double* dh2 = sc.get(nvec + 2);
  double* dn = sc.get(2);
  double* dS = sc.get(192);
  QTT_CUDA(cudaMemsetAsync(dS, 0, 192 * sizeof(double), ctx->stream));
  if (h && (op == 1)) QTT_CUDA(cudaMemcpyAsync(dh, h, sizeof(double) * nvec, cudaMemcpyHostToDevice, ctx->stream));
  float t = timed_reps(ctx, reps, [&] {
    QTT_CUDA(cudaMemcpyAsync(dw, dw0, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));
    switch (op) {
      case 0: kry_dots(ctx, dV, (int)nvec, len, dw, dh); break;
      case 1: kry_axpys(ctx, dV, (int)nvec, len, dh, dw, -1.0); break;
      case 2: kry_nrm2sq(ctx, dw, len, dn); break;
      case 3:
        kry_dots(ctx, dV, (int)nvec, len, dw, dh);
        kry_axpys(ctx, dV, (int)nvec, len, dh, dw, -1.0);
        kry_dots(ctx, dV, (int)nvec, len, dw, dh2);
        kry_axpys_nrm(ctx, dV, (int)nvec, len, dh2, dw, -1.0, dn);
        break;
      case 4:  // fused CGS2 step: h1 -> S[0..], h2 -> S[64..], |w|^2 -> S[128]
        kry_cgs2_fused(ctx, dV, (int)nvec, len, dw, nullptr, dS, 0, 64, 128, 129, 0, 0.0, 0.0, 0.0, nullptr);
        break;
      default: fail(QTT_ERR_ARG, "unknown krylov op %d", op);
    }
  });
  if (ms) *ms = t;
  QTT_CUDA(cudaMemcpyAsync(w, dw, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> hh(2 * nvec + 2, 0.0);
  double nn[2] = {0, 0};
  if (h && (op == 0 || op == 3)) QTT_CUDA(cudaMemcpyAsync(hh.data(), dh, sizeof(double) * nvec, cudaMemcpyDeviceToHost, ctx->stream));
  if (op == 3) QTT_CUDA(cudaMemcpyAsync(hh.data() + nvec, dh2, sizeof(double) * nvec, cudaMemcpyDeviceToHost, ctx->stream));
  if (op == 4) {
    QTT_CUDA(cudaMemcpyAsync(hh.data(), dS, sizeof(double) * nvec, cudaMemcpyDeviceToHost, ctx->stream));
    QTT_CUDA(cudaMemcpyAsync(hh.data() + nvec, dS + 64, sizeof(double) * nvec, cudaMemcpyDeviceToHost, ctx->stream));
    QTT_CUDA(cudaMemcpyAsync(nn, dS + 128, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (op == 2 || op == 3) QTT_CUDA(cudaMemcpyAsync(nn, dn, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h && (op == 0 || op == 3 || op == 4))
    for (int64_t i = 0; i < nvec; ++i) h[i] = hh[i] + (op >= 3 ? hh[nvec + i] : 0.0);
  if (out && (op == 2 || op == 3 || op == 4)) out[0] = sqrt(nn[0]);
  QTT_CATCH(ctx)
}

}  // extern "C"

// Internal definitions shared by the libqtt_b200 translation units (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <stdexcept>
#include <cstdlib>
#include <mutex>
#include "../../include/qtt_b200.h"

namespace qtt {

// ------------------------------------------------------------------ errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define QTT_CUDA(expr)                                                                               \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      ::qtt::fail(QTT_ERR_CUDA, "CUDA error '%s' at %s:%d: %s", cudaGetErrorString(e__), __FILE__, __LINE__, #expr); \
  } while (0)

#define QTT_REQUIRE(cond, ...)                           \
  do {                                                   \
    if (!(cond)) ::qtt::fail(QTT_ERR_ARG, __VA_ARGS__);  \
  } while (0)

// ------------------------------------------------------------------ device buffers
// Planar tensor storage: a real tensor has only `re`; a complex tensor stores the imaginary plane
// `plane` elements after the real one (planar layout keeps every contraction a real DMMA GEMM).
struct DBuf {
  double* p = nullptr;
  size_t cap = 0;  // elements
  cudaStream_t st = nullptr;  // stream of the context that allocated it (caching allocator, capi.cu)
};

struct Ctx;

// Simple stack arena for per-bond scratch (no cudaMalloc inside hot loops).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, top = 0, high = 0;
  double* alloc(size_t n_elems) {
    size_t bytes = (n_elems * sizeof(double) + 255) & ~size_t(255);
    if (top + bytes > cap) fail(QTT_ERR_ALLOC, "workspace arena exhausted: need %zu more bytes (cap %zu)", top + bytes - cap, cap);
    double* r = reinterpret_cast<double*>(base + top);
    top += bytes;
    if (top > high) high = top;
    return r;
  }
  size_t mark() const { return top; }
  void release(size_t m) { top = m; }
};

struct Ctx {
  int device = 0;
  int sms = 0;
  int cc_major = 0, cc_minor = 0, clock_khz = 0;
  cudaStream_t stream = nullptr;
  Arena arena;
  std::string err;
  int64_t launches = 0;
  unsigned dense_epoch = 0;  // value of the dense-QR panel barrier counter (d_sync + 240)
  unsigned bar_epoch = 0;  // value of the fused-Krylov barrier counter (d_sync + 4) after the launches so far
  unsigned gs_epoch = 0;   // value of the Gram-Schmidt select/re-orthogonalise barrier counter (d_sync + 232)
  bool counted = false;    // this context is included in the per-device live-context count
  int qr_last_rank = -1, qr_last_need = 0;  // numerical rank / mindim of the last split: picks the QR variant that goes first (svd_qr.cu)
  bool profiling = false;
  // small device scratch for scalars / flags / partial reductions
  double* d_scal = nullptr;      // 16384 doubles (GMRES scalar blocks at 0 and 8192, statistics at 8 ...)
  double* d_partial = nullptr;   // partial reduction buffer
  size_t partial_cap = 0;
  unsigned int* d_sync = nullptr;  // grid-barrier counters and flags (256 uints)
  double* h_scal = nullptr;        // pinned host mirror (4096 doubles)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  void ensure_arena(size_t bytes);
  double* partial(size_t n);
};

// cudaFuncSetAttribute is a per-DEVICE setting: a call site configures its kernel once per device, not once per process
// (one process may drive several GPUs through several contexts).  Thread-safe: batch.py drives one context per host thread
// and ctypes releases the GIL, so two threads can reach the same first use; the flag is set only AFTER the set-up has run,
// under the lock, and a thread that loses the race waits for it.
struct DevOnce {
  std::mutex mu;
  bool done[64] = {};
  template <typename F>
  void run(int dev, F&& setup) {
    if (dev < 0 || dev >= 64) { setup(); return; }
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return;
    setup();
    done[dev] = true;
  }
};

inline void c_launch_count(Ctx* c) { c->launches++; }

inline void check_launch(Ctx* c, const char* what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(QTT_ERR_CUDA, "kernel launch '%s' failed: %s", what, cudaGetErrorString(e));
}

// Launch with programmatic stream serialisation: the launch of a short kernel is processed while its predecessor in the
// stream drains (measured on the Arnoldi chain: -1.7 us per launch, 110.4 -> 107.3 ms per sweep).  The kernel MUST begin
// with dev::pdl_wait().  QTT_PDL=0 falls back to plain launches.
// live contexts on a device (capi.cu): with several of them (batch.py / bench.py --workload cfg5: one context = one stream per
// host thread) kernels that would fill the whole GPU by themselves use smaller grids so that the streams overlap
int live_contexts(int device);
// this context is the only one alive on its device: kernels with hand-rolled grid barriers may be launched non-cooperatively
// (all their CTAs become resident because nothing else holds SMs); QTT_COOP_ALWAYS=1 disables.  Why it is safe on ONE stream: a
// programmatically launched grid starts only after every CTA of its predecessor has exited or triggered
// (griddepcontrol.launch_dependents); the only kernels that trigger are matvec_a(_tma), whose successor is always matvec_b (no
// barrier), so a barrier kernel finds an empty GPU, and its own successor cannot start before all of its CTAs have exited.
inline bool sole_context(const Ctx* ctx) {
  static const bool always = getenv("QTT_COOP_ALWAYS") != nullptr && atoi(getenv("QTT_COOP_ALWAYS")) != 0;
  return !always && live_contexts(ctx->device) == 1;
}

template <typename... P, typename... A>
inline void launch_pdl(Ctx* ctx, const char* what, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, A... args) {
  static const bool on = !(getenv("QTT_PDL") && atoi(getenv("QTT_PDL")) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P(args)...);
  c_launch_count(ctx);
  if (e != cudaSuccess) fail(QTT_ERR_CUDA, "kernel launch '%s' failed: %s", what, cudaGetErrorString(e));
}

// ------------------------------------------------------------------ chains
struct Mps {
  Ctx* ctx = nullptr;
  int n = 0;
  int dtype = QTT_F64;
  std::vector<int64_t> chi;    // n+1
  std::vector<DBuf> core;      // device layout [a][s][c] row-major (c fastest); complex: planar, plane = a*2*c
  ~Mps();
};

struct Mpo {
  Ctx* ctx = nullptr;
  int n = 0;
  int dtype = QTT_F64;
  std::vector<int64_t> w;      // n+1
  std::vector<DBuf> core;      // device layout [wl][wr][s][t] row-major (t = s_out fastest); complex planar
  std::vector<std::vector<double>> host;  // host copy in the same layout (tiny; used to build fused W tables)
  ~Mpo();
};

void dbuf_reserve(DBuf& b, size_t n_elems);
void* pool_alloc(size_t bytes, cudaStream_t st);
void pool_free(void* p, cudaStream_t st);
void pool_trim();
void set_current_ctx(Ctx* c);
void dbuf_free(DBuf& b);

// ------------------------------------------------------------------ GEMM core (gemm_dmma.cu)
struct GemmDesc {
  const double* A = nullptr;
  const double* B = nullptr;
  double* C = nullptr;
  int M = 0, N = 0, K = 0;
  long lda = 0, ldb = 0, ldc = 0;  // row strides (elements) of the matrices AS STORED
  bool transA = false;             // A stored K x M (row-major) instead of M x K
  bool transB = false;             // B stored N x K (row-major) instead of K x N
  int batch = 1;                   // blockIdx.z = b1 + batch * b2
  long sA = 0, sB = 0, sC = 0;
  int batch2 = 1;
  long sA2 = 0, sB2 = 0, sC2 = 0;
  double alpha = 1.0, beta = 0.0;
  int tile = -1;                   // -1 = heuristic; otherwise index into the tile-config table
};
void gemm(Ctx* ctx, const GemmDesc& d);
const char* gemm_tile_name(int tile);
int gemm_num_tiles();

// ------------------------------------------------------------------ permutes (permute.cu)
// out[perm-ordered dims] = in; dims are ROW-MAJOR extents of `in`; out dim k = in dim perm[k].
void permute(Ctx* ctx, const double* in, double* out, int nd, const int64_t* dims, const int* perm);
// interleaved complex <-> planar
void deinterleave(Ctx* ctx, const double* in_interleaved, double* out_re, double* out_im, size_t n);
void interleave(Ctx* ctx, const double* in_re, const double* in_im, double* out_interleaved, size_t n);

// ------------------------------------------------------------------ hot-path contractions (contract.cu)
// Device layouts (all row-major, last index fastest):
//   L [w][a'][a]  (bra a', ket a)      R [w][c][c']  (ket c, bra c')
//   x [a][s][c]    phi / v [a][s1][s2][c]
void build_w12(Ctx* ctx, const double* W1, const double* W2, double* W12, int wl, int wm, int wr);
void matvec2(Ctx* ctx, const double* L, const double* W12, const double* R, const double* v, double* out,
             int chil, int chir, int wl, int wr, double* scratch /* >= (wl + wr) * 4 * chil * chir */);
size_t matvec2_scratch(int chil, int chir, int wl, int wr);
bool matvec2_fused_ok(int wl, int wr);
void matvec2_fused(Ctx* ctx, const double* L, const double* W12, const double* R, const double* v, double* out, int chil, int chir,
                   int wl, int wr, double* scratch /* >= 4 * chil * chir * wr */);
void env_left(Ctx* ctx, const double* L, const double* x, const double* W, const double* y, double* Lnew,
              int chia, int chib, int chia_y, int chib_y, int w, int w2, double* scratch);
void env_right(Ctx* ctx, const double* R, const double* x, const double* W, const double* y, double* Rnew,
               int chia, int chic, int chia_y, int chic_y, int w, int w2, double* scratch);
size_t env_scratch(int chia, int chib, int w, int w2);

// ------------------------------------------------------------------ Krylov vector ops (krylov.cu)
void kry_dots(Ctx* ctx, const double* V, int nvec, size_t len, const double* w, double* h /* device, nvec */);
void kry_dots_acc(Ctx* ctx, const double* V, int nvec, size_t len, const double* w, double* h /* h += V^T w */);
void kry_axpys(Ctx* ctx, const double* V, int nvec, size_t len, const double* h, double* w, double sign);
// w += sign * V h and, if nrm2sq_out != null, *nrm2sq_out = ||w_new||^2 (device scalar)
void kry_axpys_nrm(Ctx* ctx, const double* V, int nvec, size_t len, const double* h, double* w, double sign, double* nrm2sq_out);
bool kry_cgs2_fused_ok(const Ctx* ctx, size_t len);
void kry_cgs2_fused(Ctx* ctx, const double* V, int k, size_t len, double* w, double* vnext, double* S, int off_h1, int off_h2,
                    int off_nres2, int off_done, int post, double a0, double a1, double tol, volatile double* hstat);
// single-CTA local GMRES for small bonds (krylov_small.cu)
bool gmres_small_ok(int chil, int chir, int wl, int wr, int kd);
void gmres_small(Ctx* ctx, const double* L, const double* W12, const double* R, const double* beff, double* x, int chil, int chir, int wl,
                 int wr, double a0, double a1, double tol, int kd, int maxiter, bool fixed, double* V, double* stats_dev);
void kry_nrm2sq(Ctx* ctx, const double* w, size_t len, double* out /* device scalar */);
void kry_scale_to(Ctx* ctx, const double* w, size_t len, const double* denom_dev, bool sqrt_denom, double* out,
                  const double* done_dev = nullptr);
void kry_axpby(Ctx* ctx, size_t len, double a, const double* x, double b, const double* y, double c, const double* z, double* out);

// ------------------------------------------------------------------ dense local solvers (dense_solve.cu)
constexpr int DENSE_LEN_MAX = 8192;  // (4 chi_l chi_r) of the largest dense local problem (matrix 8192^2 = 512 MiB)
bool dense_local_ok(size_t len);
size_t dense_scratch(size_t len);  // doubles
// x = (a0 + a1 T)^{-1} beff by Householder QR (reference src/linsolve.jl:88-89); stats_dev[0..3] = (0, 0, zero pivots, max |R_ii|)
void dense_local_solve(Ctx* ctx, const double* L, const double* W12, const double* R, int chil, int chir, int wl, int wr, double a0,
                       double a1, const double* beff, double* x, double* scratch, double* stats_dev);
// x <- eigenvector of T whose eigenvalue is nearest `target` (reference src/dmrg_target.jl:3-7), unit norm;
// stats_dev[0..3] = (inverse-iteration steps, 1 - |<z_k, z_{k-1}>|, guarded pivots, max |R_ii|)
void dense_local_eigen(Ctx* ctx, const double* L, const double* W12, const double* R, int chil, int chir, int wl, int wr, double target,
                       double* x_inout, double* scratch, int maxit, double* stats_dev);
// `b_linsolve_exact_solver` (reference src/linsolve.jl:279-300): x = pinv(T) bT, T the (4 cbl cbr) x (4 chil chir) projected
// operator with the right-hand side chain as the bra (L [wl][cbl][chil], R [wr][chir][cbr]); minimum-norm least squares through
// the one-sided Jacobi SVD, singular values below eps max(m, N) sigma_1 dropped (LAPACK gelsd's default rcond).
bool pg_local_ok(size_t rows, size_t cols);
void pg_local_lstsq(Ctx* ctx, const double* L, const double* W12, const double* R, int cbl, int cbr, int chil, int chir, int wl, int wr,
                    const double* bT, double* x);

// ------------------------------------------------------------------ SVD split (svd_jacobi.cu)
struct SplitResult {
  int rank = 0;
  int sweeps = 0;
  double truncerr = 0.0;
};
// One-sided Jacobi on the rows of M (q x p row-major, device): on return Wn (allocated from the ctx arena, k x p) holds
// the k dominant right singular vectors as orthonormal rows (descending sigma); truncation per NDTensors.truncate!.
// The caller owns the arena mark taken before the call.
struct RowSvd {
  int k = 0;
  int sweeps = 0;
  double truncerr = 0.0;
  double* Wn = nullptr;
  double* Wni = nullptr;  // imaginary plane for complex input
};
// General form: optional imaginary plane Mi (planar complex rows) and eigen mode (M Hermitian PSD: the rows of the
// converged matrix are lambda_i q_i^H, the truncated spectrum is lambda_i = |row_i|).
RowSvd svd_rows_ex(Ctx* ctx, const double* M, const double* Mi, int q, int p, double cutoff, int maxdim, int mindim,
                   double* sigma_out, int max_sweeps, bool eigen_mode);
bool svd_rows_qr_supported(int q, int p);
RowSvd svd_rows_qr(Ctx* ctx, const double* M, int q, int p, double cutoff, int maxdim, int mindim, double* sigma_out, int max_sweeps,
                   bool eigen_mode = false);
RowSvd svd_rows(Ctx* ctx, const double* M, int q, int p, double cutoff, int maxdim, int mindim, double* sigma_out, int max_sweeps);
// phi [a][s1][s2][c] (device).  Writes A [a][s1][k] and B [k][s2][c] (device, capacity for k = min(2a, 2c)).
SplitResult split2(Ctx* ctx, const double* phi, int chia, int chic, int ortho, double cutoff, int maxdim, int mindim,
                   double* A_out, double* B_out, double* sigma_out /* device, may be null */, int max_sweeps);

// ------------------------------------------------------------------ TT-SVD of a sample vector (ttsvd.cu)
void vec_to_mps(Ctx* ctx, const double* v_dev, int n, double cutoff, int maxdim, int mindim, Mps& out, double* maxerr_out);

// ------------------------------------------------------------------ apply / metrics (apply.cu)
void apply_mpo_mps(Ctx* ctx, const Mpo& A, const Mps& psi, double cutoff, int maxdim, int mindim, Mps& out);
void mps_inner(Ctx* ctx, const Mps& x, const Mps& y, double* out2);
void mps_residual(Ctx* ctx, const Mpo& A, const Mps& x, const Mps& b, double* sq, double* sqn);

// ------------------------------------------------------------------ sum, dense expansion, point sampling (mps_algebra.cu)
void mps_add(Ctx* ctx, const Mps& x, const Mps& y, double alpha, double beta, double cutoff, int maxdim, int mindim, Mps& out);
void mps_to_vec(Ctx* ctx, const Mps& x, double* out_re_dev, double* out_im_dev /* null for a real chain */);
void mps_sample(Ctx* ctx, const Mps& x, const int64_t* idx_dev, int64_t count, double* out_dev /* [count] re (+ [count] im) */);

}  // namespace qtt

struct qtt_ctx : qtt::Ctx {};
struct qtt_mps : qtt::Mps {};
struct qtt_mpo : qtt::Mpo {};

/*
 * libqtt_b200.so -- C ABI of the B200-native (sm_100a) QTT sweep engine.
 *
 * Drop-in boundary for ONE hot path of ITensor/ITensorQTT.jl: the two-site DMRG-style sweep
 * (effective-operator matvec, environment updates, truncated-SVD site splitting) behind
 * `linsolve` / `dmrg_target` / `eigsolve_target`, and MPO.MPS `apply` for the laplacian / dft /
 * prolongation operators.  The reference has no FFI of its own (pure Julia); the seams replaced
 * here are Julia-level and each entry point cites the reference call it
 * stands in for.  Host Julia binds these with `ccall`; the test harness binds
 * them with Python `ctypes`.
 *
 * Conventions
 *   - every function returns 0 on success and a negative qtt_status otherwise; the text of the last
 *     error is available from qtt_last_error(ctx).  No C++ exception crosses this boundary.
 *   - host arrays are BORROWED for the duration of the call only (Julia: GC.@preserve).
 *   - host tensors are dense, COLUMN-MAJOR (Julia `Array`) with these index orders:
 *       MPS core j : (chi[j-1], 2, chi[j])                         [array(x[j], l_{j-1}, s_j, l_j)]
 *       MPO core j : (w[j-1], w[j], 2, 2) = (l, r, s_in, s_out)    [/root/reference/src/laplacian.jl:32]
 *     dtype 0 = Float64, 1 = ComplexF64 (interleaved re,im as in Julia).
 *   - site 1 is the most significant bit; index value 1 <-> bit 0 (reference src/function_to_mps.jl:170-173,
 *     src/project_bits.jl:7).  All Index identity / tags / prime levels stay in Julia.
 *   - a qtt_ctx owns one CUDA stream and its workspaces; it is NOT thread-safe: one ctx per host
 *     thread / task / GPU.  There is no CPU fallback: every entry point fails with QTT_ERR_CUDA if no
 *     sm_100-class device is usable.
 */
#ifndef QTT_B200_H
#define QTT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTT_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define QTT_API __attribute__((visibility("default")))
#else
#define QTT_API
#endif

typedef enum {
  QTT_OK = 0,
  QTT_ERR_ARG = -1,     /* bad argument (shape mismatch, null pointer, unsupported dtype ...) */
  QTT_ERR_CUDA = -2,    /* CUDA runtime / launch failure, or no usable device */
  QTT_ERR_ALLOC = -3,   /* out of device or host memory */
  QTT_ERR_NOCONV = -4,  /* an inner iteration (Jacobi SVD / eigen) hit its sweep limit */
  QTT_ERR_UNSUPPORTED = -5
} qtt_status;

enum { QTT_F64 = 0, QTT_C128 = 1 };

typedef struct qtt_ctx qtt_ctx;
typedef struct qtt_mps qtt_mps;
typedef struct qtt_mpo qtt_mpo;

/* ---- context ------------------------------------------------------------------------------ */
QTT_API int qtt_version(void);
/* Creates a context on CUDA device `device`; returns NULL on failure (qtt_global_error() has the text). */
QTT_API qtt_ctx* qtt_ctx_create(int device);
QTT_API void qtt_ctx_destroy(qtt_ctx* ctx);
QTT_API const char* qtt_last_error(const qtt_ctx* ctx);
QTT_API const char* qtt_global_error(void);
QTT_API int qtt_ctx_sync(qtt_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py `gpu_launches`) */
QTT_API int64_t qtt_ctx_launch_count(const qtt_ctx* ctx);
/* device properties: out[0]=SM count, out[1]=major, out[2]=minor, out[3]=clock kHz */
QTT_API int qtt_ctx_device_info(const qtt_ctx* ctx, int64_t* out4);

/* ---- chains: upload / download (replace `array(x[j], inds...)` / `itensor(array, inds...)` in the Julia shim) */
QTT_API int qtt_mps_upload(qtt_ctx* ctx, int n, const int64_t* chi /* n+1 */, const void* const* cores, int dtype, qtt_mps** out);
QTT_API int qtt_mpo_upload(qtt_ctx* ctx, int n, const int64_t* w /* n+1 */, const void* const* cores, int dtype, qtt_mpo** out);
QTT_API int qtt_mps_length(const qtt_mps* x);
QTT_API int qtt_mps_dtype(const qtt_mps* x);
QTT_API int qtt_mps_dims(const qtt_mps* x, int64_t* chi_out /* n+1 */);
QTT_API int qtt_mps_download(const qtt_mps* x, void* const* cores_out);
QTT_API int qtt_mps_clone(const qtt_mps* x, qtt_mps** out);
QTT_API int qtt_mpo_length(const qtt_mpo* A);
QTT_API int qtt_mpo_dims(const qtt_mpo* A, int64_t* w_out /* n+1 */);
QTT_API void qtt_mps_free(qtt_mps* x);
QTT_API void qtt_mpo_free(qtt_mpo* A);

/* ---- sweep engine --------------------------------------------------------------------------
 * Parameters of `tdvp(solver, P, t=Inf, x0; reverse_step=false, nsite=2, nsweeps, maxdim, mindim,
 * cutoff, normalize)` (upstream ITensorMPS driver reached from reference src/linsolve.jl:96,172,305
 * and src/dmrg_target.jl:11) plus the local-solver keywords the reference forwards to KrylovKit
 * (examples/airy_equation/src/linsolve.jl:24-34; src/linsolve.jl:266-273). */
typedef enum {
  QTT_SOLVER_GMRES = 0,        /* `linsolve`: KrylovKit GMRES on (a0 + a1 P) phi = proj(b) */
  QTT_SOLVER_LANCZOS = 1,      /* Krylov variant of `dmrg_target` (reference src/dmrg_target.jl:14-21): Ritz value nearest target */
  QTT_SOLVER_SHIFT_INVERT = 2, /* `eigsolve_target` (reference src/eigsolve_target.jl:2-9): eigsolve(x -> linsolve(P, x, x, -target)) */
  QTT_SOLVER_DENSE_QR = 3,     /* `linsolve_pseudoinverse` (reference src/linsolve.jl:54-97): dense (4 chi^2)^2 operator, `qr!(T) \ b`;
                                  solves (a0 + a1 T) x = b -- pass a0 = 0, a1 = 1 for the reference's behaviour (it ignores them) */
  QTT_SOLVER_DENSE_EIGEN = 4,  /* `dmrg_target` as written (reference src/dmrg_target.jl:1-12): dense Hermitian effective operator,
                                  eigenvector of the eigenvalue nearest `target`.  Dense kinds need 4 chi_l chi_r <= 8192. */
  QTT_SOLVER_PG_LSQ = 5        /* `b_linsolve` (reference src/linsolve.jl:175-307, live solver `b_linsolve_exact_solver` :279-300):
                                  Petrov-Galerkin sweep -- environments L.x.A.dag(b') with the right-hand side as the bra
                                  (`ProjMPOLinsolve`, b re-orthogonalised to every bond, :248-253), local right-hand side `bvec`
                                  (:255-261), dense SVD least squares `svd(A) \ b` on the (4 chi_b^2) x (4 chi_x^2) matrix.
                                  a0 / a1 are ignored as in the reference body. */
} qtt_solver_kind;

typedef struct {
  int32_t nsweeps;
  const int32_t* maxdim;  /* per-sweep arrays of length nsweeps (tdvp `process_sweeps` semantics); NULL = unlimited */
  const int32_t* mindim;  /* NULL = 1 */
  const double* cutoff;   /* NULL = 1e-16 (tdvp default) */
  int32_t solver;         /* qtt_solver_kind */
  double a0, a1;          /* shift / scale of (a0 + a1 A) x = b (real) */
  double tol;             /* KrylovKit `tol` (absolute residual norm) */
  int32_t krylovdim;      /* KrylovKit `krylovdim` */
  int32_t maxiter;        /* KrylovKit `maxiter` (restart cycles) */
  int32_t fixed_matvecs;  /* >0: benchmark "fixed-work" mode: exactly one cycle of this many Arnoldi steps per bond, tol ignored */
  int32_t normalize;      /* tdvp `normalize` */
  int32_t outputlevel;    /* >=1: print the "After sweep k: maxlinkdim=... maxerr=..." line to stdout */
  double target;          /* target eigenvalue for the eigen solvers */
  int32_t svd_max_sweeps; /* Jacobi sweep limit (0 = default 40) */
  int32_t x0_canonical;   /* 1: x0 is already right-orthonormal (orthogonality centre at site 1): skip `orthogonalize!(psi, 1)` */
} qtt_sweep_params;

typedef struct {
  int32_t maxlinkdim;
  int32_t svd_sweeps_max;  /* largest number of Jacobi sweeps any split needed */
  double maxtruncerr;
  double seconds;          /* wall time of the sweep (host clock, stream synchronised) */
  double solver_resid;     /* largest final local residual estimate over the bonds */
  int64_t matvecs;         /* effective-operator applications */
  double flops_matvec;     /* algorithmic FP64 flops of those matvecs (SURVEY.md 8d formula, actual bond dims) */
  double t_matvec, t_env, t_svd, t_krylov; /* CUDA-event milliseconds per phase (0 unless profiling enabled) */
  double energy;           /* eigen solvers: last local Ritz value; linsolve: 0 */
  double device_ms;        /* CUDA-event time of the sweep on the context's stream */
} qtt_sweep_info;

/* `linsolve(A::MPO, b::MPS, x0::MPS, a0, a1; nsweeps, maxdim, cutoff, tol, krylovdim, maxiter)`
 * (reference call sites examples/airy_equation/airy_equation.jl:429, src/eigsolve_target.jl:4;
 * body examples/airy_equation/src/linsolve.jl:24-34).  x is updated in place.  per_sweep: nsweeps entries or NULL.
 * p->solver = QTT_SOLVER_DENSE_QR selects the dense local solve of `linsolve_pseudoinverse` (src/linsolve.jl:54-97);
 * any other value means GMRES. */
QTT_API int qtt_linsolve(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* b, qtt_mps* x_inout, const qtt_sweep_params* p, qtt_sweep_info* per_sweep);
/* `dmrg_target(A, psi0; target_eigenvalue, nsweeps, maxdim, cutoff)` (reference src/dmrg_target.jl:1-12)
 * with the Krylov local solvers the reference sketches at :14-21: p->solver = QTT_SOLVER_LANCZOS (Ritz value nearest target) or
 * QTT_SOLVER_SHIFT_INVERT (`eigsolve_target`, src/eigsolve_target.jl:2-9; fixed_matvecs must be 0), or
 * QTT_SOLVER_DENSE_EIGEN (the dense `eigen` of :3-7 itself, small bonds only). */
QTT_API int qtt_dmrg_target(qtt_ctx* ctx, const qtt_mpo* A, qtt_mps* psi_inout, const qtt_sweep_params* p, qtt_sweep_info* per_sweep);
/* enable per-phase CUDA-event timing inside the sweeps (adds one event pair per phase per bond) */
QTT_API int qtt_set_profiling(qtt_ctx* ctx, int enabled);

/* ---- MPO.MPS apply -------------------------------------------------------------------------
 * `apply(A::MPO, psi::MPS; cutoff, maxdim, mindim)` (upstream density-matrix algorithm) as called by
 * reference src/dft.jl:124,131 and src/prolongation.jl:44,79.  Output cores carry the MPO's output
 * site index; chain reversal / output-bit relabelling stay in the caller (src/dft.jl:76-79,124). */
QTT_API int qtt_apply(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* psi, double cutoff, int32_t maxdim, int32_t mindim, qtt_mps** out);

/* ---- TT-SVD of a sampled function ------------------------------------------------------------
 * `vec_to_mps(v, s; cutoff, maxdim, mindim)` = `MPS(vec_to_tensor(v), s; ...)` (reference src/function_to_mps.jl:170-177):
 * the back end of `function_to_mps(f, s, xstart, xstop; alg="factorize", cutoff=1e-15)` (:41-52) once f is sampled on
 * `qtt_xrange` (:10-12).  v: 2^n doubles, v[j] = f(x_j), site 1 = most significant bit of j; host memory
 * (on_device = 0: uploaded once) or device memory on ctx's GPU (on_device = 1: read in place, not modified).
 * Left-to-right sweep of truncated SVDs with the NDTensors.truncate! rule (relative cutoff on the discarded weight);
 * the result is left-orthogonal with its orthogonality centre at site n, as upstream.  maxdim = 0: unlimited (this build
 * supports bond dimensions up to 512).  maxtruncerr (may be NULL) receives the largest per-bond discarded weight. */
QTT_API int qtt_vec_to_mps(qtt_ctx* ctx, const double* v, int32_t n, int32_t on_device, double cutoff, int32_t maxdim, int32_t mindim,
                           qtt_mps** out, double* maxtruncerr);

/* ---- sum, dense expansion, point sampling ------------------------------------------------------
 * out = alpha x + beta y, truncated: ITensorMPS `+(x, y; cutoff, maxdim)` as called at reference
 * src/fourier_interpolation.jl:8 (also src/function_to_mps.jl:123, src/laplacian.jl:106).  The direct sum is assembled on
 * the device and compressed by the density-matrix pass of `qtt_apply` (NDTensors.truncate! rule); x, y real or complex. */
QTT_API int qtt_mps_add(qtt_ctx* ctx, const qtt_mps* x, const qtt_mps* y, double alpha, double beta, double cutoff, int32_t maxdim,
                        int32_t mindim, qtt_mps** out);
/* `mps_to_discrete_function(psi)` (reference src/function_to_mps.jl:140-145, one dimension): all 2^n values, value j at the
 * bits of j with site 1 = most significant bit (the inverse of qtt_vec_to_mps).  out: 2^n doubles (f64 chain) or 2^n
 * interleaved complex (c128 chain); host memory (on_device = 0) or device memory on ctx's GPU (on_device = 1). */
QTT_API int qtt_mps_to_vec(qtt_ctx* ctx, const qtt_mps* x, void* out, int32_t on_device);
/* Values at `count` grid indices: `project_bits(psi, bits)` with every site projected (reference src/project_bits.jl:1-15,
 * scalar return :11-13), batched; 0 <= idx[i] < 2^n, bits of idx[i] with site 1 = most significant bit.  idx and out are
 * host arrays; out: count doubles or count interleaved complex. */
QTT_API int qtt_mps_sample(qtt_ctx* ctx, const qtt_mps* x, const int64_t* idx, int64_t count, void* out);

/* ---- metrics (reference src/mps_utils.jl:109-132) ------------------------------------------- */
/* inner(x, y) = <x|y>; out2 = {re, im} */
QTT_API int qtt_inner(qtt_ctx* ctx, const qtt_mps* x, const qtt_mps* y, double* out2);
/* sqeuclidean((A,x), b) and sqeuclidean_normalized((A,x), b) */
QTT_API int qtt_residual(qtt_ctx* ctx, const qtt_mpo* A, const qtt_mps* x, const qtt_mps* b, double* sqeuclid, double* sqeuclid_normalized);

/* ---- fine-grained kernels: parity tests, ncu captures, bench roofline (host, column-major arrays) ----
 * Index orders: L (a, w, a'), R (c, w, c'), W (wl, wr, s_in, s_out), v/out (a, s1, s2, c).
 * `reps` >= 1 repeats the device computation; *ms_per_rep (may be NULL) receives the CUDA-event average. */
/* `product(P::ProjMPO, v)` = noprime(v*L*W1*W2*R) (upstream; pattern at reference src/linsolve.jl:71-75,283-288) */
QTT_API int qtt_matvec2(qtt_ctx* ctx, const double* L, const double* W1, const double* W2, const double* R, const double* v, double* out,
                int64_t chil, int64_t chir, int64_t wl, int64_t wm, int64_t wr, int32_t reps, float* ms_per_rep);
/* `makeL!`: Lnew (b, w', b') = L * x * W * dag(x') (reference src/linsolve.jl:215; upstream ProjMPO) ; x (a, 2, b) */
QTT_API int qtt_env_left(qtt_ctx* ctx, const double* L, const double* x, const double* W, double* Lnew,
                 int64_t chia, int64_t chib, int64_t w, int64_t w2, int32_t reps, float* ms_per_rep);
/* `makeR!`: Rnew (a, w, a') = R * x * W * dag(x') (reference src/linsolve.jl:240); x (a, 2, c), R (c, w2, c') */
QTT_API int qtt_env_right(qtt_ctx* ctx, const double* R, const double* x, const double* W, double* Rnew,
                  int64_t chia, int64_t chic, int64_t w, int64_t w2, int32_t reps, float* ms_per_rep);
/* `replacebond!`: truncated SVD split of phi (a, 2, 2, c); ortho 0 = "left" (A = U, B = S V), 1 = "right" (A = U S, B = V).
 * A_out capacity (a, 2, min(2a,2c)), B_out capacity (min(2a,2c), 2, c), sigma_out capacity min(2a,2c).
 * info_out[0] = kept rank, [1] = Jacobi sweeps used; *truncerr_out = discarded weight fraction (truncation rule of
 * NDTensors.truncate!). */
QTT_API int qtt_split2(qtt_ctx* ctx, const double* phi, int64_t chia, int64_t chic, int32_t ortho, double cutoff, int32_t maxdim, int32_t mindim,
               double* A_out, double* B_out, double* sigma_out, int32_t* info_out, double* truncerr_out, int32_t reps, float* ms_per_rep);
/* C (M x N) = alpha op(A) op(B) + beta C, ROW-MAJOR host matrices; transA: A stored K x M; transB: B stored N x K. GEMM core test hook. */
QTT_API int qtt_gemm(qtt_ctx* ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
             const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int32_t tile_override, int32_t reps, float* ms_per_rep);
/* Krylov vector ops on nvec basis vectors of length len (row-major V[nvec][len]):
 * h = V^T w (op 0), w -= V h (op 1), out[0] = ||w||_2 (op 2), one CGS2 Arnoldi orthogonalisation step (op 3: h (nvec), out[0] = norm). */
QTT_API int qtt_krylov_op(qtt_ctx* ctx, int32_t op, const double* V, int64_t nvec, int64_t len, double* w_inout, double* h_inout, double* out,
                  int32_t reps, float* ms_per_rep);

#ifdef __cplusplus
}
#endif
#endif /* QTT_B200_H */

// Two-site sweep driver and device-resident local Krylov solvers.
//
// Replaces, for this path only:
//   ITensorMPS `tdvp(solver, P, t=Inf, x0; reverse_step=false, nsite=2, ...)`  -> two_site_sweeps()
//   `ProjMPO_MPS(A, [b])` environment caches + `position!`                      -> LA/RA/LB/RB arrays below
//   `proj(P::ProjMPS)` (reference src/linsolve.jl:13-22)                         -> project_rhs()
//   `KrylovKit.linsolve(P, b, x0, a0, a1; tol, krylovdim, maxiter)` (GMRES)      -> gmres_local()
//   `KrylovKit.eigsolve` (Lanczos / Krylov-Schur) for `dmrg_target`/`eigsolve_target` -> lanczos_local()
//   `replacebond!` (SVD + truncate!)                                            -> split2() (svd_jacobi.cu)
// All three hot loops (sweeps, bonds, Krylov iterations) run here; the only host<->device synchronisation per bond
// is the read-back of the kept rank after the split (fixed-work mode) plus one per GMRES restart cycle otherwise.
// GMRES scalars (Hessenberg column, Givens rotations, residual estimate, convergence flag) live in device memory and
// are advanced by a one-thread kernel; the host polls a pinned, device-written flag to stop enqueuing early.
#include "common.cuh"
#include "device_utils.cuh"
#include "krylov_scalars.cuh"
#include <chrono>
#include <functional>
#include <cstdlib>

namespace qtt {

void kry_dots_acc(Ctx*, const double*, int, size_t, const double*, double*);

namespace {

using namespace kscal;
int g_trace_sweeps = 0;  // sweeps completed in this process (QTT_DUMP addresses a sweep by this count)

__global__ void gmres_update_kernel(double* S, int k, double a0, double a1, double tol, int ld, volatile double* hstat) {
  dev::pdl_wait();
  gmres_step_scalars(S, k, a0, a1, tol, ld, hstat);
}

// Back substitution R yk = y, and the coefficients of the restart residual r = y[k] * [V, w/nres] (G_1^H ... G_k^H e_{k+1}).
// single-thread back substitution on a shared-memory copy of the scalar block (global-memory latency made it 49 us)
__global__ void __launch_bounds__(256) gmres_solve_kernel(double* S, int k_or_neg_kd, int ld) {
  dev::pdl_wait();
  extern __shared__ double Ss[];
  for (int i = threadIdx.x; i < OFF_END; i += blockDim.x) Ss[i] = S[i];
  __syncthreads();
  if (threadIdx.x == 0) gmres_solve_scalars(Ss, k_or_neg_kd, ld);
  __syncthreads();
  for (int i = threadIdx.x; i < OFF_R; i += blockDim.x) S[i] = Ss[i];  // only the vectors change (yk, coef, S_R0N2)
}

__global__ void set_scalars_kernel(double* p, int n, double v) {
  dev::pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void scale_inplace_kernel(double* w, size_t len, const double* nrm2sq) {
  dev::pdl_wait();
  double inv = rsqrt(*nrm2sq);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (size_t)gridDim.x * blockDim.x) w[e] *= inv;
}

struct LocalStats {
  int64_t matvecs = 0;
  double resid = 0.0;
  double theta = 0.0;
  int converged = 0;
};

using MatvecFn = std::function<void(const double*, double*)>;

struct KrylovWork {
  double* V = nullptr;  // (kd + 1) x len
  double* w = nullptr;
  double* r = nullptr;
};

void read_scalars(Ctx* ctx, const double* S, int off, int n, double* out) {
  QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + H_READ, S + off, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) out[i] = ctx->h_scal[H_READ + i];
}

// KrylovKit.linsolve(f, b, x0, GMRES(krylovdim, maxiter, tol), a0, a1)  [restated in oracle/sweep.py: gmres_krylovkit]
// Orthogonalisation is CGS2 (two classical Gram-Schmidt passes = 2 multi-dots + 2 multi-axpys), numerically equivalent
// to KrylovKit's ModifiedGramSchmidt2 but with one reduction per pass instead of one per basis vector.
void gmres_local(Ctx* ctx, const MatvecFn& mv, const double* beff, double* x, size_t len, double a0, double a1, double tol, int kd,
                 int maxiter, bool fixed, KrylovWork& ws, LocalStats& st, double* S = nullptr) {
  if (!S) S = ctx->d_scal;
  volatile double* hstat = ctx->h_scal;
  auto vec = [&](int i) { return ws.V + (size_t)i * len; };
  const bool fused = kry_cgs2_fused_ok(ctx, len);
  mv(x, ws.w);
  st.matvecs++;
  kry_axpby(ctx, len, 1.0, beff, -a0, x, -a1, ws.w, ws.r);
  kry_nrm2sq(ctx, ws.r, len, S + S_R0N2);
  double beta = 0.0;
  if (!fixed) {
    double v;
    read_scalars(ctx, S, S_R0N2, 1, &v);
    beta = sqrt(v);
    st.resid = beta;
    if (beta < tol) { st.converged = 1; return; }
  }
  int numiter = 0;
  while (numiter < maxiter) {
    numiter++;
    launch_pdl(ctx, "set_scalars", set_scalars_kernel, dim3(1), dim3(32), 0, S + S_DONE, 2, 0.0);
    kry_scale_to(ctx, ws.r, len, S + S_R0N2, true, vec(0));
    if (!fixed) hstat[H_DONE] = 0.0;
    int k;
    for (k = 1; k <= kd; ++k) {
      mv(vec(k - 1), ws.w);
      st.matvecs++;
      if (fused) {
        // one cooperative launch: CGS2, norm, Hessenberg/Givens update, v_k = w / |w|.  v_k is zeroed by the done flag of
        // the PREVIOUS step: at the step that sets the flag the vector is still finite (|w| != 0, entries <= |w|), its
        // coefficient is zero, and every later vector is written as zeros.
        kry_cgs2_fused(ctx, ws.V, k, len, ws.w, vec(k), S, OFF_H1, OFF_H2, S_NRES2, S_DONE, 1, a0, a1, fixed ? -1.0 : tol,
                       fixed ? nullptr : ctx->h_scal);
      } else {
        kry_dots(ctx, ws.V, k, len, ws.w, S + OFF_H1);
        kry_axpys(ctx, ws.V, k, len, S + OFF_H1, ws.w, -1.0);
        kry_dots(ctx, ws.V, k, len, ws.w, S + OFF_H2);
        kry_axpys_nrm(ctx, ws.V, k, len, S + OFF_H2, ws.w, -1.0, S + S_NRES2);
        launch_pdl(ctx, "gmres_update", gmres_update_kernel, dim3(1), dim3(1), 0, S, k, a0, a1, fixed ? -1.0 : tol, 64, fixed ? nullptr : ctx->h_scal);
        kry_scale_to(ctx, ws.w, len, S + S_NRES2, true, vec(k), S + S_DONE);
      }
      if (!fixed && hstat[H_DONE] != 0.0) break;
    }
    int kfin = kd;
    if (!fixed) {
      double v[8];
      read_scalars(ctx, S, 0, 8, v);
      kfin = (int)v[S_K];
      beta = v[S_BETA];
      st.resid = beta;
    }
    // fixed mode: the device-side step count (S_K < kd after an early breakdown) is used by the kernel itself; the
    // coefficients beyond it are zero and the corresponding basis vectors were written as zeros.
    launch_pdl(ctx, "gmres_solve", gmres_solve_kernel, dim3(1), dim3(256), OFF_END * sizeof(double), S, fixed ? -kd : kfin, 64);
    kry_axpys(ctx, ws.V, kfin, len, S + OFF_YK, x, 1.0);
    if (fixed) return;
    if (beta > tol) {
      // residual without another operator application (KrylovKit rotates the basis; here: explicit combination)
      QTT_CUDA(cudaMemsetAsync(ws.r, 0, sizeof(double) * len, ctx->stream));
      kry_axpys(ctx, ws.V, kfin + 1, len, S + OFF_COEF, ws.r, 1.0);
    } else {
      mv(x, ws.w);
      st.matvecs++;
      kry_axpby(ctx, len, 1.0, beff, -a0, x, -a1, ws.w, ws.r);
      kry_nrm2sq(ctx, ws.r, len, S + S_R0N2);
      double v;
      read_scalars(ctx, S, S_R0N2, 1, &v);
      beta = sqrt(v);
      st.resid = beta;
      if (beta < tol) { st.converged = 1; return; }
    }
  }
}

// ---- Lanczos with full (CGS2) orthogonalisation: local eigen solve of the Krylov `dmrg_target` variant ------------
// Projected matrix T (k x k, symmetric) is accumulated on the device; its eigen-decomposition (k <= 62) is done by a
// single-warp cyclic Jacobi kernel; the Ritz vector nearest `target` is assembled with one multi-axpy.
__global__ void lanczos_store_col_kernel(double* S, int k, int ld) { lanczos_step_scalars(S, k, ld); }

__global__ void __launch_bounds__(32) ritz_kernel(double* S, int kd, int ld, double target, int which, double* ritz_coef) {
  const int k = (int)S[S_K];
  // cyclic Jacobi eigen-decomposition of the k x k symmetric T in shared memory (k <= 62); lane-parallel row/col updates
  extern __shared__ double ritz_sm[];
  double(*A)[65] = reinterpret_cast<double(*)[65]>(ritz_sm);
  double(*Q)[65] = reinterpret_cast<double(*)[65]>(ritz_sm + 64 * 65);
  const int lane = threadIdx.x;
  double* T = S + OFF_R;
  for (int i = 0; i < k; ++i)
    for (int j = lane; j < k; j += 32) {
      A[i][j] = 0.5 * (T[i * ld + j] + T[j * ld + i]);
      Q[i][j] = (i == j) ? 1.0 : 0.0;
    }
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < k; ++i)
      for (int j = lane; j < k; j += 32)
        if (i != j) off += A[i][j] * A[i][j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
    double diag = 0.0;
    for (int j = lane; j < k; j += 32) diag += A[j][j] * A[j][j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diag += __shfl_xor_sync(0xffffffffu, diag, o);
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < k - 1; ++p)
      for (int q = p + 1; q < k; ++q) {
        double apq = A[p][q];
        if (fabs(apq) > 1e-300) {
          double app = A[p][p], aqq = A[q][q];
          double zeta = (aqq - app) / (2.0 * apq);
          double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = rsqrt(1.0 + t * t), s = c * t;
          __syncwarp();
          for (int j = lane; j < k; j += 32) {  // rows p, q
            double x1 = A[p][j], x2 = A[q][j];
            A[p][j] = c * x1 - s * x2;
            A[q][j] = s * x1 + c * x2;
          }
          __syncwarp();
          for (int j = lane; j < k; j += 32) {  // columns p, q (+ eigenvectors)
            double x1 = A[j][p], x2 = A[j][q];
            A[j][p] = c * x1 - s * x2;
            A[j][q] = s * x1 + c * x2;
            double q1 = Q[j][p], q2 = Q[j][q];
            Q[j][p] = c * q1 - s * q2;
            Q[j][q] = s * q1 + c * q2;
          }
          __syncwarp();
        }
      }
  }
  __syncwarp();
  if (lane == 0) {
    int best = 0;
    double bd = which ? -fabs(A[0][0]) : fabs(A[0][0] - target);
    for (int j = 1; j < k; ++j) {
      double d = which ? -fabs(A[j][j]) : fabs(A[j][j] - target);
      if (d < bd) { bd = d; best = j; }
    }
    S[S_THETA] = A[best][best];
    S[S_BETA] = fabs(S[S_NRES] * Q[k - 1][best]);  // Ritz residual estimate |beta_k u_k[last]|
    S[S_TMP] = (double)best;
  }
  __syncwarp();
  int best = (int)S[S_TMP];
  for (int j = lane; j < kd; j += 32) ritz_coef[j] = j < k ? Q[j][best] : 0.0;
}

// which = 0: Ritz value nearest `target`; which = 1: largest magnitude (KrylovKit :LM, used by the shift-invert solver)
void lanczos_local(Ctx* ctx, const MatvecFn& mv, double* x, size_t len, double target, double tol, int kd, int maxiter, bool fixed,
                   KrylovWork& ws, LocalStats& st, double* S = nullptr, int which = 0, bool check_each = false) {
  if (!S) S = ctx->d_scal;
  auto vec = [&](int i) { return ws.V + (size_t)i * len; };
  const bool fused = kry_cgs2_fused_ok(ctx, len);
  auto run_ritz = [&]() {
    constexpr int kRitzSmem = 2 * 64 * 65 * (int)sizeof(double);
    static DevOnce once;
    once.run(ctx->device, [&] {
      QTT_CUDA(cudaFuncSetAttribute(ritz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRitzSmem));
    });
    ritz_kernel<<<1, 32, kRitzSmem, ctx->stream>>>(S, kd, 64, target, which, S + OFF_YK);
    check_launch(ctx, "ritz");
  };
  for (int it = 0; it < maxiter; ++it) {
    int kused = kd;
    bool early = false;
    kry_nrm2sq(ctx, x, len, S + S_R0N2);
    launch_pdl(ctx, "set_scalars", set_scalars_kernel, dim3(1), dim3(32), 0, S + S_DONE, 2, 0.0);
    kry_scale_to(ctx, x, len, S + S_R0N2, true, vec(0));
    QTT_CUDA(cudaMemsetAsync(S + OFF_R, 0, sizeof(double) * 64 * 64, ctx->stream));
    for (int k = 1; k <= kd; ++k) {
      mv(vec(k - 1), ws.w);
      st.matvecs++;
      if (fused) {
        kry_cgs2_fused(ctx, ws.V, k, len, ws.w, k < kd ? vec(k) : nullptr, S, OFF_H1, OFF_H2, S_NRES2, S_DONE, 2, 0.0, 0.0, 0.0, nullptr);
      } else {
        kry_dots(ctx, ws.V, k, len, ws.w, S + OFF_H1);
        kry_axpys(ctx, ws.V, k, len, S + OFF_H1, ws.w, -1.0);
        kry_dots(ctx, ws.V, k, len, ws.w, S + OFF_H2);
        kry_axpys_nrm(ctx, ws.V, k, len, S + OFF_H2, ws.w, -1.0, S + S_NRES2);
        lanczos_store_col_kernel<<<1, 1, 0, ctx->stream>>>(S, k, 64);
        check_launch(ctx, "lanczos_store_col");
        if (k < kd) kry_scale_to(ctx, ws.w, len, S + S_NRES2, true, vec(k), S + S_DONE);
      }
      kused = k;
      if (check_each && k >= 2 && k < kd) {
        // expensive operator (shift-invert: one inner solve per application): test the Ritz pair after every expansion,
        // as KrylovKit's eigsolve does
        run_ritz();
        double v[8];
        read_scalars(ctx, S, 0, 8, v);
        if (v[S_BETA] < tol || v[S_DONE] != 0.0) { early = true; break; }
      }
    }
    if (!early) run_ritz();
    QTT_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
    kry_axpys(ctx, ws.V, kused, len, S + OFF_YK, x, 1.0);
    if (fixed) { st.resid = 0.0; return; }
    double v[8];
    read_scalars(ctx, S, 0, 8, v);
    st.resid = v[S_BETA];
    st.theta = v[S_THETA];
    if (v[S_BETA] < tol) { st.converged = 1; return; }
  }
}

// ------------------------------------------------------------------ overlap environments <x|b> (ProjMPS)
// Layouts: Lb [ab][ax], Rb [cb][cx]  (b link slowest).
void benv_left(Ctx* ctx, const double* Lb, const double* x, const double* bc, double* Lbn, int ax, int cx, int ab, int cb) {
  size_t mark = ctx->arena.mark();
  double* T = ctx->arena.alloc((size_t)ab * 2 * cx);
  GemmDesc g;
  g.A = Lb; g.B = x; g.C = T;
  g.M = ab; g.N = 2 * cx; g.K = ax; g.lda = ax; g.ldb = 2 * cx; g.ldc = 2 * cx;
  gemm(ctx, g);
  GemmDesc h;
  h.A = bc; h.B = T; h.C = Lbn; h.transA = true;
  h.M = cb; h.N = cx; h.K = 2 * ab; h.lda = cb; h.ldb = cx; h.ldc = cx;
  gemm(ctx, h);
  ctx->arena.release(mark);
}

void benv_right(Ctx* ctx, const double* Rb, const double* x, const double* bc, double* Rbn, int ax, int cx, int ab, int cb) {
  size_t mark = ctx->arena.mark();
  double* T = ctx->arena.alloc((size_t)ab * 2 * cx);
  GemmDesc g;
  g.A = bc; g.B = Rb; g.C = T;
  g.M = 2 * ab; g.N = cx; g.K = cb; g.lda = cb; g.ldb = cx; g.ldc = cx;
  gemm(ctx, g);
  GemmDesc h;
  h.A = T; h.B = x; h.C = Rbn; h.transB = true;
  h.M = ab; h.N = ax; h.K = 2 * cx; h.lda = 2 * cx; h.ldb = 2 * cx; h.ldc = ax;
  gemm(ctx, h);
  ctx->arena.release(mark);
}

// proj(P::ProjMPS), reference src/linsolve.jl:13-22: beff[a][s1 s2 c] = Lb[ab][a] b1[ab][s1][m] b2[m][s2][cb] Rb[cb][c]
void project_rhs(Ctx* ctx, const double* Lb, const double* b1, const double* b2, const double* Rb, double* beff, int ax, int cx,
                 int ab, int mb, int cb) {
  size_t mark = ctx->arena.mark();
  double* bb = ctx->arena.alloc((size_t)4 * ab * cb);
  double* T1 = ctx->arena.alloc((size_t)4 * ab * cx);
  GemmDesc g;
  g.A = b1; g.B = b2; g.C = bb;
  g.M = 2 * ab; g.N = 2 * cb; g.K = mb; g.lda = mb; g.ldb = 2 * cb; g.ldc = 2 * cb;
  gemm(ctx, g);
  GemmDesc h;
  h.A = bb; h.B = Rb; h.C = T1;
  h.M = 4 * ab; h.N = cx; h.K = cb; h.lda = cb; h.ldb = cx; h.ldc = cx;
  gemm(ctx, h);
  GemmDesc f;
  f.A = Lb; f.B = T1; f.C = beff; f.transA = true;
  f.M = ax; f.N = 4 * cx; f.K = ab; f.lda = ax; f.ldb = 4 * cx; f.ldc = 4 * cx;
  gemm(ctx, f);
  ctx->arena.release(mark);
}

void set_one(Ctx* ctx, DBuf& b) {
  dbuf_reserve(b, 2);
  launch_pdl(ctx, "set_one", set_scalars_kernel, dim3(1), dim3(32), 0, b.p, 1, 1.0);
}

double matvec_flops(double chil, double chir, double wl, double wm, double wr) {
  const double d = 2.0;
  return 2.0 * (chil * (wl * chil) * (d * d * chir) + (wl * d) * (chil * d * chir) * (wm * d) + (wm * d) * (chil * d * chir) * (wr * d) +
                (wr * chir) * (chil * d * d) * chir);
}

struct PhaseTimer {
  Ctx* ctx;
  double* acc;
  bool on;
  PhaseTimer(Ctx* c, double* a) : ctx(c), acc(a), on(c->profiling && a) {
    if (on) cudaEventRecord(ctx->ev0, ctx->stream);
  }
  ~PhaseTimer() {
    if (on) {
      cudaEventRecord(ctx->ev1, ctx->stream);
      cudaEventSynchronize(ctx->ev1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
      *acc += ms;
    }
  }
};

}  // namespace

// Bring x into right-orthonormal form (orthogonality centre at site 1): `orthogonalize!(psi, 1)` of the tdvp driver.
// Each step is an un-truncated row-space factorisation (one-sided Jacobi) of x[j] viewed as (chi_l) x (2 chi_r).
void right_orthonormalize(Ctx* ctx, Mps& x) {
  for (int j = x.n - 1; j >= 1; --j) {
    const int a = (int)x.chi[j], c = (int)x.chi[j + 1], ap = (int)x.chi[j - 1];
    size_t need = ((size_t)4 * a * 2 * c + (size_t)2 * ap * a + 4096) * sizeof(double) * 2;
    ctx->ensure_arena(need);
    size_t mark = ctx->arena.mark();
    RowSvd r = svd_rows(ctx, x.core[j].p, a, 2 * c, 0.0, 0, 1, nullptr, 0);
    const int k = r.k;
    double* Rf = ctx->arena.alloc((size_t)a * k);
    GemmDesc g;
    g.A = x.core[j].p; g.B = r.Wn; g.C = Rf; g.transB = true;
    g.M = a; g.N = k; g.K = 2 * c; g.lda = 2 * c; g.ldb = 2 * c; g.ldc = k;
    gemm(ctx, g);
    double* Xn = ctx->arena.alloc((size_t)2 * ap * k);
    GemmDesc h;
    h.A = x.core[j - 1].p; h.B = Rf; h.C = Xn;
    h.M = 2 * ap; h.N = k; h.K = a; h.lda = a; h.ldb = k; h.ldc = k;
    gemm(ctx, h);
    QTT_CUDA(cudaMemcpyAsync(x.core[j].p, r.Wn, sizeof(double) * (size_t)k * 2 * c, cudaMemcpyDeviceToDevice, ctx->stream));
    QTT_CUDA(cudaMemcpyAsync(x.core[j - 1].p, Xn, sizeof(double) * (size_t)2 * ap * k, cudaMemcpyDeviceToDevice, ctx->stream));
    x.chi[j] = k;
    ctx->arena.release(mark);
  }
}

// One step of `orthogonalize!(y, j)` on a device chain (un-truncated; bond dimensions as a QR would leave them).
// forward: centre j-1 -> j (core j-1 becomes left-orthonormal); backward: centre j+1 -> j (core j+1 right-orthonormal).
static void center_backward(Ctx* ctx, Mps& y, int j) {
  const int jj = j + 1;
  const int a = (int)y.chi[jj], c = (int)y.chi[jj + 1], ap = (int)y.chi[jj - 1];
  ctx->ensure_arena((((size_t)4 * a * 2 * c + (size_t)2 * ap * a + 4096) * 2 + 65536) * sizeof(double));
  size_t mark = ctx->arena.mark();
  const int full = a < 2 * c ? a : 2 * c;
  RowSvd r = svd_rows(ctx, y.core[jj].p, a, 2 * c, 0.0, 0, full, nullptr, 0);
  const int k = r.k;
  double* Rf = ctx->arena.alloc((size_t)a * k);
  GemmDesc g;
  g.A = y.core[jj].p; g.B = r.Wn; g.C = Rf; g.transB = true;
  g.M = a; g.N = k; g.K = 2 * c; g.lda = 2 * c; g.ldb = 2 * c; g.ldc = k;
  gemm(ctx, g);
  double* Xn = ctx->arena.alloc((size_t)2 * ap * k);
  GemmDesc h;
  h.A = y.core[jj - 1].p; h.B = Rf; h.C = Xn;
  h.M = 2 * ap; h.N = k; h.K = a; h.lda = a; h.ldb = k; h.ldc = k;
  gemm(ctx, h);
  dbuf_reserve(y.core[jj], (size_t)k * 2 * c);
  dbuf_reserve(y.core[jj - 1], (size_t)2 * ap * k);
  QTT_CUDA(cudaMemcpyAsync(y.core[jj].p, r.Wn, sizeof(double) * (size_t)k * 2 * c, cudaMemcpyDeviceToDevice, ctx->stream));
  QTT_CUDA(cudaMemcpyAsync(y.core[jj - 1].p, Xn, sizeof(double) * (size_t)2 * ap * k, cudaMemcpyDeviceToDevice, ctx->stream));
  y.chi[jj] = k;
  ctx->arena.release(mark);
}

static void center_forward(Ctx* ctx, Mps& y, int j) {
  const int jj = j - 1;
  const int a = (int)y.chi[jj], c = (int)y.chi[jj + 1], cn = (int)y.chi[jj + 2];
  ctx->ensure_arena((((size_t)6 * a * 2 * c + (size_t)2 * c * cn + 4096) * 2 + 65536) * sizeof(double));
  size_t mark = ctx->arena.mark();
  // M = core[jj] as (2a x c); its transpose's row space gives Q^T: M = Q R, Q = Wn^T (2a x k), R = Wn M (k x c)
  double* Mt = ctx->arena.alloc((size_t)c * 2 * a);
  int64_t dims[2] = {2 * a, c};
  int pm[2] = {1, 0};
  permute(ctx, y.core[jj].p, Mt, 2, dims, pm);
  const int full = c < 2 * a ? c : 2 * a;
  RowSvd r = svd_rows(ctx, Mt, c, 2 * a, 0.0, 0, full, nullptr, 0);
  const int k = r.k;
  double* Rf = ctx->arena.alloc((size_t)k * c);
  GemmDesc g;
  g.A = r.Wn; g.B = y.core[jj].p; g.C = Rf;
  g.M = k; g.N = c; g.K = 2 * a; g.lda = 2 * a; g.ldb = c; g.ldc = c;
  gemm(ctx, g);
  double* Xn = ctx->arena.alloc((size_t)k * 2 * cn);
  GemmDesc h;
  h.A = Rf; h.B = y.core[jj + 1].p; h.C = Xn;
  h.M = k; h.N = 2 * cn; h.K = c; h.lda = c; h.ldb = 2 * cn; h.ldc = 2 * cn;
  gemm(ctx, h);
  double* Q = ctx->arena.alloc((size_t)2 * a * k);
  int64_t dims2[2] = {k, 2 * a};
  permute(ctx, r.Wn, Q, 2, dims2, pm);
  dbuf_reserve(y.core[jj], (size_t)2 * a * k);
  dbuf_reserve(y.core[jj + 1], (size_t)k * 2 * cn);
  QTT_CUDA(cudaMemcpyAsync(y.core[jj].p, Q, sizeof(double) * (size_t)2 * a * k, cudaMemcpyDeviceToDevice, ctx->stream));
  QTT_CUDA(cudaMemcpyAsync(y.core[jj + 1].p, Xn, sizeof(double) * (size_t)k * 2 * cn, cudaMemcpyDeviceToDevice, ctx->stream));
  y.chi[jj + 1] = k;
  ctx->arena.release(mark);
}

// solver_kind: QTT_SOLVER_GMRES needs b; the eigen solvers ignore it.
void two_site_sweeps(Ctx* ctx, const Mpo& A, const Mps* b, Mps& x, const qtt_sweep_params& P, qtt_sweep_info* infos) {
  const int n = A.n;
  QTT_REQUIRE(n >= 2, "need at least 2 sites");
  QTT_REQUIRE(x.n == n, "MPS length %d != MPO length %d", x.n, n);
  QTT_REQUIRE(A.dtype == QTT_F64 && x.dtype == QTT_F64 && (!b || b->dtype == QTT_F64),
              "sweep engine: only Float64 chains are supported in this build");
  QTT_REQUIRE(A.w[0] == 1 && A.w[n] == 1 && x.chi[0] == 1 && x.chi[n] == 1, "boundary bond dimensions must be 1");
  if (b) QTT_REQUIRE(b->n == n && b->chi[0] == 1 && b->chi[n] == 1, "rhs MPS shape mismatch");
  QTT_REQUIRE(P.nsweeps >= 0, "nsweeps < 0");
  const bool dense_qr = (P.solver == QTT_SOLVER_DENSE_QR), dense_eig = (P.solver == QTT_SOLVER_DENSE_EIGEN);
  const bool pg = (P.solver == QTT_SOLVER_PG_LSQ);   // `b_linsolve`: Petrov-Galerkin environments, dense SVD least squares
  const bool dense = dense_qr || dense_eig || pg;
  const bool fixed = !dense && P.fixed_matvecs > 0;
  const int kd = dense ? 1 : (fixed ? P.fixed_matvecs : P.krylovdim);
  QTT_REQUIRE(kd >= 1 && kd <= KD_MAX, "krylovdim must be in [1, %d]", KD_MAX);
  const int maxiter = P.maxiter > 0 ? P.maxiter : 1;
  const bool is_lin = (P.solver == QTT_SOLVER_GMRES) || dense_qr || pg;
  QTT_REQUIRE(is_lin || dense_eig || P.solver == QTT_SOLVER_LANCZOS || P.solver == QTT_SOLVER_SHIFT_INVERT, "unknown solver kind %d",
              P.solver);
  const bool shift_invert = (P.solver == QTT_SOLVER_SHIFT_INVERT);
  QTT_REQUIRE(!(shift_invert && fixed), "fixed_matvecs is not defined for the shift-invert solver");
  QTT_REQUIRE(!is_lin || b, "linsolve needs a right-hand side");

  // QTT_TRACE=1 (with profiling enabled): per-bond phase timings on stderr; QTT_DUMP="sweep,dir,bond,path" saves that bond's phi
  const bool trace = ctx->profiling && getenv("QTT_TRACE") != nullptr;
  int dump_sw = -1, dump_dir = -1, dump_bond = -1;
  static char dump_path_buf[512];
  const char* dump_path = nullptr;
  double* dump_buf = nullptr;
  DBuf dump_store;
  if (trace && getenv("QTT_DUMP") && sscanf(getenv("QTT_DUMP"), "%d,%d,%d,%500s", &dump_sw, &dump_dir, &dump_bond, dump_path_buf) == 4)
    dump_path = dump_path_buf;
  // The per-call set-up (orthogonalize!, right-environment build) is part of the FIRST sweep's device time: the events
  // bracket everything a `linsolve(...; nsweeps)` call enqueues.
  cudaEvent_t evs0 = nullptr, evs1 = nullptr;
  QTT_CUDA(cudaEventCreate(&evs0));
  QTT_CUDA(cudaEventCreate(&evs1));
  std::vector<DBuf> LA(n + 1), RA(n + 1), LB(n + 1), RB(n + 1);
  Mps bra;   // b_linsolve only
  auto cleanup = [&]() {
    for (auto* v : {&LA, &RA, &LB, &RB})
      for (auto& d : *v) dbuf_free(d);
    if (evs0) cudaEventDestroy(evs0);
    if (evs1) cudaEventDestroy(evs1);
    evs0 = evs1 = nullptr;
  };
  try {
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    auto t_call = std::chrono::steady_clock::now();
    QTT_CUDA(cudaEventRecord(evs0, ctx->stream));
    if (!P.x0_canonical) right_orthonormalize(ctx, x);
    // `ProjMPOLinsolve` (reference src/linsolve.jl:175-253): a private copy of b, re-gauged to every bond, is the bra of the
    // operator environments
    if (pg) {
      bra.ctx = ctx; bra.n = n; bra.dtype = QTT_F64; bra.chi = b->chi; bra.core.resize(n);
      for (int j = 0; j < n; ++j) {
        const size_t ne = (size_t)b->chi[j] * 2 * b->chi[j + 1];
        dbuf_reserve(bra.core[j], ne);
        QTT_CUDA(cudaMemcpyAsync(bra.core[j].p, b->core[j].p, ne * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      }
      right_orthonormalize(ctx, bra);
    }
    set_one(ctx, LA[0]);
    set_one(ctx, RA[n]);
    if (b && !pg) { set_one(ctx, LB[0]); set_one(ctx, RB[n]); }
    int wmax = 1;
    for (int j = 0; j <= n; ++j) wmax = wmax > (int)A.w[j] ? wmax : (int)A.w[j];

    auto update_right = [&](int j) {  // RA[j] from RA[j+1] and site j
      const int a = (int)x.chi[j], c = (int)x.chi[j + 1], w = (int)A.w[j], w2 = (int)A.w[j + 1];
      const int ay = pg ? (int)bra.chi[j] : a, cy = pg ? (int)bra.chi[j + 1] : c;
      const int am = a > ay ? a : ay, cm = c > cy ? c : cy;
      ctx->ensure_arena((env_scratch(am, cm, w, w2) + 4096) * sizeof(double) + ((size_t)8 * am * cm + 4096) * sizeof(double));
      size_t mark = ctx->arena.mark();
      double* scr = ctx->arena.alloc(env_scratch(am, cm, w, w2));
      dbuf_reserve(RA[j], (size_t)w * a * ay);
      env_right(ctx, RA[j + 1].p, x.core[j].p, A.core[j].p, pg ? bra.core[j].p : x.core[j].p, RA[j].p, a, c, ay, cy, w, w2, scr);
      ctx->arena.release(mark);
      if (b && !pg) {
        const int ab = (int)b->chi[j], cb = (int)b->chi[j + 1];
        dbuf_reserve(RB[j], (size_t)ab * a);
        benv_right(ctx, RB[j + 1].p, x.core[j].p, b->core[j].p, RB[j].p, a, c, ab, cb);
      }
    };
    auto update_left = [&](int j) {  // LA[j+1] from LA[j] and site j
      const int a = (int)x.chi[j], c = (int)x.chi[j + 1], w = (int)A.w[j], w2 = (int)A.w[j + 1];
      const int ay = pg ? (int)bra.chi[j] : a, cy = pg ? (int)bra.chi[j + 1] : c;
      const int am = a > ay ? a : ay, cm = c > cy ? c : cy;
      ctx->ensure_arena((env_scratch(am, cm, w, w2) + 4096) * sizeof(double) + ((size_t)8 * am * cm + 4096) * sizeof(double));
      size_t mark = ctx->arena.mark();
      double* scr = ctx->arena.alloc(env_scratch(am, cm, w, w2));
      dbuf_reserve(LA[j + 1], (size_t)w2 * cy * c);
      env_left(ctx, LA[j].p, x.core[j].p, A.core[j].p, pg ? bra.core[j].p : x.core[j].p, LA[j + 1].p, a, c, ay, cy, w, w2, scr);
      ctx->arena.release(mark);
      if (b && !pg) {
        const int ab = (int)b->chi[j], cb = (int)b->chi[j + 1];
        dbuf_reserve(LB[j + 1], (size_t)cb * c);
        benv_left(ctx, LB[j].p, x.core[j].p, b->core[j].p, LB[j + 1].p, a, c, ab, cb);
      }
    };

    for (int j = n - 1; j >= 2; --j) update_right(j);

    for (int sw = 0; sw < P.nsweeps; ++sw) {
      auto t0 = t_call;
      if (sw > 0) {
        QTT_CUDA(cudaStreamSynchronize(ctx->stream));
        t0 = std::chrono::steady_clock::now();
        QTT_CUDA(cudaEventRecord(evs0, ctx->stream));
      }
      qtt_sweep_info info;
      memset(&info, 0, sizeof(info));
      const int maxdim = P.maxdim ? P.maxdim[sw] : 0;
      const int mindim = P.mindim ? P.mindim[sw] : 1;
      const double cutoff = P.cutoff ? P.cutoff[sw] : 1e-16;
      for (int dir = 0; dir < 2; ++dir) {
        for (int step = 0; step < n - 1; ++step) {
          const int j = dir == 0 ? step : (n - 2 - step);
          if (pg) {
            // `position!(P::ProjMPOLinsolve, x, pos)` (reference src/linsolve.jl:248-253): b's orthogonality centre follows the
            // bond; the left environment that contains the re-gauged core is rebuilt (the right ones only ever see
            // right-orthonormal cores of b, which do not change)
            if (dir == 0 && j > 0) {
              center_forward(ctx, bra, j);
              update_left(j - 1);
            } else if (dir == 1 && j < n - 2) {
              center_backward(ctx, bra, j);
            }
          }
          const int chil = (int)x.chi[j], chim = (int)x.chi[j + 1], chir = (int)x.chi[j + 2];
          const int wl = (int)A.w[j], wm = (int)A.w[j + 1], wr = (int)A.w[j + 2];
          const size_t len = (size_t)4 * chil * chir;
          size_t need = (len * (size_t)(2 * (kd + 1) + 16) + matvec2_scratch(chil, chir, wl, wr) + (size_t)16 * wl * wr + 65536) * sizeof(double);
          need += (size_t)256 * (kd + 24);
          if (pg) {
            const size_t mrows = (size_t)4 * bra.chi[j] * bra.chi[j + 2];
            QTT_REQUIRE(pg_local_ok(mrows, len), "b_linsolve: the dense local operator %zu x %zu exceeds the cap at bond %d", mrows, len, j);
            need += ((size_t)6 * mrows * len + (size_t)32 * (len + mrows) + 262144) * sizeof(double);
          } else if (dense) {
            QTT_REQUIRE(dense_local_ok(len), "dense local solver: 4 chi_l chi_r = %zu exceeds the cap (%d, QTT_DENSE_LEN_MAX) at bond %d; use a Krylov solver",
                        len, DENSE_LEN_MAX, j);
            need += (dense_scratch(len) + 64) * sizeof(double);
          }
          ctx->ensure_arena(need);
          size_t mark = ctx->arena.mark();
          double* phi = ctx->arena.alloc(len);
          double* W12 = ctx->arena.alloc((size_t)16 * wl * wr);
          double* mvscr = ctx->arena.alloc(matvec2_scratch(chil, chir, wl, wr));
          KrylovWork ws;
          ws.V = ctx->arena.alloc(len * (size_t)(kd + 1));
          ws.w = ctx->arena.alloc(len);
          ws.r = ctx->arena.alloc(len);
          {
            GemmDesc g;  // phi[(a s1)][(s2 c)] = x[j] . x[j+1]
            g.A = x.core[j].p; g.B = x.core[j + 1].p; g.C = phi;
            g.M = 2 * chil; g.N = 2 * chir; g.K = chim; g.lda = chim; g.ldb = 2 * chir; g.ldc = 2 * chir;
            gemm(ctx, g);
          }
          build_w12(ctx, A.core[j].p, A.core[j + 1].p, W12, wl, wm, wr);
          const double* Lp = LA[j].p;
          const double* Rp = RA[j + 2].p;
          MatvecFn mv = [&](const double* in, double* out) { matvec2(ctx, Lp, W12, Rp, in, out, chil, chir, wl, wr, mvscr); };
          LocalStats st;
          bool small_pending = false, dense_pending = false;
          const double tk0 = info.t_krylov, ts0 = info.t_svd, te0 = info.t_env;
          const int kdl = (size_t)kd < len ? kd : (int)len;  // a Krylov space cannot exceed the local dimension
          {
            PhaseTimer pt(ctx, &info.t_krylov);
            if (pg) {
              // `b_linsolve_exact_solver` (reference src/linsolve.jl:279-300): bT = `bvec(P)` (the two cores of b between the
              // environments, :255-261), x = svd(T) \ bT on the rectangular projected operator
              const int cbl = (int)bra.chi[j], cbm = (int)bra.chi[j + 1], cbr = (int)bra.chi[j + 2];
              double* bT = ctx->arena.alloc((size_t)4 * cbl * cbr);
              GemmDesc gb;
              gb.A = bra.core[j].p; gb.B = bra.core[j + 1].p; gb.C = bT;
              gb.M = 2 * cbl; gb.N = 2 * cbr; gb.K = cbm; gb.lda = cbm; gb.ldb = 2 * cbr; gb.ldc = 2 * cbr;
              gemm(ctx, gb);
              pg_local_lstsq(ctx, Lp, W12, Rp, cbl, cbr, chil, chir, wl, wr, bT, phi);
            } else if (is_lin) {
              double* beff = ctx->arena.alloc(len);
              project_rhs(ctx, LB[j].p, b->core[j].p, b->core[j + 1].p, RB[j + 2].p, beff, chil, chir, (int)b->chi[j], (int)b->chi[j + 1],
                          (int)b->chi[j + 2]);
              if (dense_qr) {
                // `linsolve_pseudoinverse` (reference src/linsolve.jl:67-92): dense operator, Householder QR, back substitution;
                // the true residual |b - (a0 + a1 P) x| is measured with one matvec for `solver_resid`
                double* dscr = ctx->arena.alloc(dense_scratch(len));
                dense_local_solve(ctx, Lp, W12, Rp, chil, chir, wl, wr, P.a0, P.a1, beff, phi, dscr, ctx->d_scal + 8);
                mv(phi, ws.w);
                kry_axpby(ctx, len, 1.0, beff, -P.a0, phi, -P.a1, ws.w, ws.r);
                kry_nrm2sq(ctx, ws.r, len, ctx->d_scal + 12);
                QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 72, ctx->d_scal + 8, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                dense_pending = true;
              } else if (gmres_small_ok(chil, chir, wl, wr, kdl)) {
                // whole local solve in one single-CTA kernel (krylov_small.cu); statistics are read after the split's sync
                gmres_small(ctx, Lp, W12, Rp, beff, phi, chil, chir, wl, wr, P.a0, P.a1, P.tol, kdl, maxiter, fixed, ws.V, ctx->d_scal + 8);
                QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 72, ctx->d_scal + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                small_pending = true;
              } else {
                gmres_local(ctx, mv, beff, phi, len, P.a0, P.a1, P.tol, kdl, maxiter, fixed, ws, st);
              }
            } else if (dense_eig) {
              // `dmrg_target` as written (reference src/dmrg_target.jl:3-7): dense operator, eigenvector nearest the target;
              // energy = Rayleigh quotient <x, P x> of the returned unit vector
              double* dscr = ctx->arena.alloc(dense_scratch(len));
              dense_local_eigen(ctx, Lp, W12, Rp, chil, chir, wl, wr, P.target, phi, dscr, 0, ctx->d_scal + 8);
              mv(phi, ws.w);
              kry_dots(ctx, phi, 1, len, ws.w, ctx->d_scal + 12);
              QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 72, ctx->d_scal + 8, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
              dense_pending = true;
            } else if (shift_invert) {
              // `eigsolve_target` (reference src/eigsolve_target.jl:2-9) as the local solver (src/dmrg_target.jl:14-21):
              // eigsolve(x -> linsolve(P, x, x, -target)), dominant eigenpair of the inverse.  The inner GMRES uses the default
              // scalar block, the outer Lanczos a second one.
              KrylovWork wi;
              wi.V = ctx->arena.alloc(len * (size_t)(kdl + 1));
              wi.w = ctx->arena.alloc(len);
              wi.r = ctx->arena.alloc(len);
              MatvecFn finv = [&](const double* in, double* out) {
                QTT_CUDA(cudaMemcpyAsync(out, in, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));  // x0 = b = in
                LocalStats si;
                gmres_local(ctx, mv, in, out, len, -P.target, 1.0, P.tol, kdl, maxiter, false, wi, si);
                st.matvecs += si.matvecs - 1;  // lanczos_local counts one application per call
              };
              lanczos_local(ctx, finv, phi, len, 0.0, P.tol, kdl, maxiter, false, ws, st, ctx->d_scal + 8192, 1, true);
              info.energy = (st.theta != 0.0) ? P.target + 1.0 / st.theta : P.target;
            } else {
              lanczos_local(ctx, mv, phi, len, P.target, P.tol, kdl, maxiter, fixed, ws, st);
              info.energy = st.theta;
            }
          }
          if (fixed && is_lin && !pg && !small_pending && !dense_pending)
            QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 70, ctx->d_scal + S_BETA, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
          if (!small_pending && !dense_pending) {
            info.matvecs += st.matvecs;
            info.flops_matvec += (double)st.matvecs * matvec_flops(chil, chir, wl, wm, wr);
            if (st.resid > info.solver_resid) info.solver_resid = st.resid;
          }
          if (P.normalize) {
            kry_nrm2sq(ctx, phi, len, ctx->d_scal + S_TMP);
            long nb = (long)((len + 255) / 256);
            if (nb > 4L * ctx->sms) nb = 4L * ctx->sms;
            launch_pdl(ctx, "scale_inplace", scale_inplace_kernel, dim3((unsigned)nb), dim3(256), 0, phi, len, ctx->d_scal + S_TMP);
          }
          int kcap = 2 * chil < 2 * chir ? 2 * chil : 2 * chir;
          if (maxdim > 0 && maxdim < kcap) kcap = maxdim;
          dbuf_reserve(x.core[j], (size_t)2 * chil * kcap);
          dbuf_reserve(x.core[j + 1], (size_t)2 * chir * kcap);
          if (trace && dump_path && g_trace_sweeps == dump_sw && dir == dump_dir && j == dump_bond) {
            dbuf_reserve(dump_store, len);
            dump_buf = dump_store.p;
            QTT_CUDA(cudaMemcpyAsync(dump_buf, phi, len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
          }
          SplitResult sr;
          {
            PhaseTimer pt(ctx, &info.t_svd);
            sr = split2(ctx, phi, chil, chir, dir, cutoff, maxdim, mindim, x.core[j].p, x.core[j + 1].p, nullptr, P.svd_max_sweeps);
          }
          if (dense_pending) {  // split2 synchronised the stream: the dense solver's statistics have arrived
            info.matvecs += 1;
            info.flops_matvec += matvec_flops(chil, chir, wl, wm, wr);
            if (dense_qr) {
              if (ctx->h_scal[74] > 0.0)
                fail(QTT_ERR_NOCONV, "dense QR local solve: exactly singular projected operator at bond %d (zero pivot)", j);
              const double rr = sqrt(ctx->h_scal[76]);
              if (rr > info.solver_resid) info.solver_resid = rr;
            } else {
              info.energy = ctx->h_scal[76];
              if (ctx->h_scal[73] > info.solver_resid) info.solver_resid = ctx->h_scal[73];
            }
          } else if (small_pending) {  // split2 synchronised the stream: the single-CTA solver's statistics have arrived
            const double mvs = ctx->h_scal[72];
            info.matvecs += (int64_t)mvs;
            info.flops_matvec += mvs * matvec_flops(chil, chir, wl, wm, wr);
            if (ctx->h_scal[73] > info.solver_resid) info.solver_resid = ctx->h_scal[73];
          } else if (fixed && is_lin && !pg && ctx->h_scal[70] > info.solver_resid) {
            info.solver_resid = ctx->h_scal[70];  // split2 synchronised
          }
          x.chi[j + 1] = sr.rank;
          if (sr.truncerr > info.maxtruncerr) info.maxtruncerr = sr.truncerr;
          if (sr.sweeps > info.svd_sweeps_max) info.svd_sweeps_max = sr.sweeps;
          ctx->arena.release(mark);
          {
            PhaseTimer pt(ctx, &info.t_env);
            // the environment past the turning point is never used (position! would not build it either)
            if (dir == 0 && j < n - 2) {
              if (!pg) update_left(j);   // b_linsolve: built at the next bond, after b's gauge has moved
            } else if (dir == 1 && j > 0) {
              update_right(j + 1);
            }
          }
          if (trace) {
            fprintf(stderr, "[qtt trace] sweep %d dir %d bond %d chi (%d,%d,%d)->%d krylov %.3f ms svd %.3f ms (%d jacobi sweeps) env %.3f ms\n", sw, dir,
                    j, chil, chim, chir, sr.rank, info.t_krylov - tk0, info.t_svd - ts0, sr.sweeps, info.t_env - te0);
            if (dump_path && g_trace_sweeps == dump_sw && dir == dump_dir && j == dump_bond) {
              // debugging aid: the solved two-site tensor of one bond, as raw doubles [a][s1][s2][c]
              std::vector<double> h(len);
              QTT_CUDA(cudaMemcpy(h.data(), dump_buf, len * sizeof(double), cudaMemcpyDeviceToHost));
              if (FILE* f = fopen(dump_path, "wb")) { fwrite(h.data(), sizeof(double), len, f); fclose(f); }
            }
          }
        }
      }
      QTT_CUDA(cudaEventRecord(evs1, ctx->stream));
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      ++g_trace_sweeps;
      info.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs0, evs1);
        info.device_ms = ms;
      }
      int mld = 1;
      for (int j = 1; j < n; ++j) mld = mld > (int)x.chi[j] ? mld : (int)x.chi[j];
      info.maxlinkdim = mld;
      if (P.outputlevel >= 1) {
        printf("After sweep %d: maxlinkdim=%d maxerr=%.2E time=%.3f matvecs=%lld\n", sw + 1, mld, info.maxtruncerr, info.seconds,
               (long long)info.matvecs);
        fflush(stdout);
      }
      if (infos) infos[sw] = info;
    }
  } catch (...) {
    cleanup();
    dbuf_free(dump_store);
    throw;
  }
  cleanup();
  dbuf_free(dump_store);
}

}  // namespace qtt

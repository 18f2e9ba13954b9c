// Device-resident scalar state of the local Krylov solvers (GMRES Hessenberg column / Givens rotations / residual
// estimate; Lanczos projected matrix), shared by the sweep driver (sweep.cu) and the fused CGS2 Arnoldi kernel
// (krylov.cu), which commits one step's scalars from its last CTA-0 thread.
#pragma once
#include "common.cuh"

namespace qtt {
namespace kscal {

constexpr int KD_MAX = 62;
// layout of the scalar block in ctx->d_scal (doubles)
enum {
  S_BETA = 0, S_NRES2 = 1, S_R0N2 = 2, S_DONE = 3, S_K = 4, S_NRES = 5, S_THETA = 6, S_TMP = 7,
  OFF_H1 = 16, OFF_H2 = 16 + 64, OFF_Y = 16 + 128, OFF_GC = 16 + 192, OFF_GS = 16 + 256, OFF_YK = 16 + 320,
  OFF_COEF = 16 + 384, OFF_R = 512, OFF_END = 512 + 64 * 64
};
// pinned host status block (ctx->h_scal + 64 ...)
enum { H_DONE = 64, H_K = 65, H_BETA = 66, H_READ = 80 };

__device__ __forceinline__ void givens(double f, double g, double& c, double& s, double& r) {
  if (g == 0.0) { c = 1.0; s = 0.0; r = f; return; }
  if (f == 0.0) { c = 0.0; s = g > 0 ? 1.0 : -1.0; r = fabs(g); return; }
  double h = hypot(f, g);
  double sg = f > 0 ? 1.0 : -1.0;
  c = fabs(f) / h;
  s = sg * g / h;
  r = sg * h;
}

// One Arnoldi step's scalar work (KrylovKit GMRES inner loop): fold the shift into the Hessenberg
// column, apply the stored Givens rotations, create the new one, rotate the rhs, update the residual estimate.
__device__ inline void gmres_step_scalars(double* S, int k, double a0, double a1, double tol, int ld, volatile double* hstat) {
  if (S[S_DONE] != 0.0) return;
  double nres = sqrt(S[S_NRES2]);
  S[S_NRES] = nres;
  double* Rm = S + OFF_R;
  double* y = S + OFF_Y;
  double* gc = S + OFF_GC;
  double* gs = S + OFF_GS;
  if (k == 1) y[0] = sqrt(S[S_R0N2]);
  double hn2 = nres * nres;
  for (int i = 0; i < k; ++i) {
    double h = S[OFF_H1 + i] + S[OFF_H2 + i];
    hn2 += h * h;
    Rm[i * ld + (k - 1)] = (i == k - 1) ? (a0 + a1 * h) : (a1 * h);
  }
  // Krylov space exhausted (invariant subspace): the next basis vector would be rounding noise
  const bool breakdown = !(nres > 1e-13 * sqrt(hn2));
  for (int i = 0; i < k - 1; ++i) {
    double t1 = Rm[i * ld + (k - 1)], t2 = Rm[(i + 1) * ld + (k - 1)];
    Rm[i * ld + (k - 1)] = gc[i] * t1 + gs[i] * t2;
    Rm[(i + 1) * ld + (k - 1)] = -gs[i] * t1 + gc[i] * t2;
  }
  double c, s, r;
  givens(Rm[(k - 1) * ld + (k - 1)], a1 * nres, c, s, r);
  gc[k - 1] = c;
  gs[k - 1] = s;
  Rm[(k - 1) * ld + (k - 1)] = r;
  double yk = y[k - 1];
  y[k - 1] = c * yk;
  y[k] = -s * yk;
  double beta = fabs(y[k]);
  S[S_BETA] = beta;
  S[S_K] = (double)k;
  if (!(beta > tol) || breakdown) S[S_DONE] = 1.0;
  if (hstat) {
    hstat[H_K] = (double)k;
    hstat[H_BETA] = beta;
    hstat[H_DONE] = S[S_DONE];
    __threadfence_system();
  }
}

// Same update with the inputs already on chip (the fused CGS2 kernel): h = h1 + h2 and the stored rotations gc/gs come
// from shared memory, the running Hessenberg column lives in registers, and global memory only receives fire-and-forget
// stores.  The plain version above pays a store -> load round trip per rotation (~0.25 us each) when S is in global memory.
__device__ inline void gmres_step_scalars_onchip(double* S, int k, double a0, double a1, double tol, int ld, volatile double* hstat,
                                                 const double* h1s, const double* h2s, const double* gcs, const double* gss, double y_prev,
                                                 double r0n2, double done_prev, double nres2) {
  if (done_prev != 0.0) return;
  const double nres = sqrt(nres2);
  S[S_NRES] = nres;
  double* Rm = S + OFF_R;
  double hn2 = nres * nres;
  double cur = 0.0;
  for (int i = 0; i < k; ++i) {
    const double h = h1s[i] + h2s[i];
    hn2 += h * h;
    const double col = (i == k - 1) ? (a0 + a1 * h) : (a1 * h);
    if (i == 0) {
      cur = col;
    } else {
      const double t1 = cur, t2 = col;
      Rm[(i - 1) * ld + (k - 1)] = gcs[i - 1] * t1 + gss[i - 1] * t2;
      cur = -gss[i - 1] * t1 + gcs[i - 1] * t2;
    }
  }
  const bool breakdown = !(nres > 1e-13 * sqrt(hn2));
  double c, s, r;
  givens(cur, a1 * nres, c, s, r);
  S[OFF_GC + k - 1] = c;
  S[OFF_GS + k - 1] = s;
  Rm[(k - 1) * ld + (k - 1)] = r;
  const double yk = (k == 1) ? sqrt(r0n2) : y_prev;
  S[OFF_Y + k - 1] = c * yk;
  S[OFF_Y + k] = -s * yk;
  const double beta = fabs(s * yk);
  S[S_BETA] = beta;
  S[S_K] = (double)k;
  const double done = (!(beta > tol) || breakdown) ? 1.0 : 0.0;
  if (done != 0.0) S[S_DONE] = 1.0;
  if (hstat) {
    hstat[H_K] = (double)k;
    hstat[H_BETA] = beta;
    hstat[H_DONE] = done;
    __threadfence_system();
  }
}

// Back substitution R yk = y, and the coefficients of the restart residual r = y[k] * [V, w/nres] (G_1^H ... G_k^H e_{k+1}).
// k >= 0: use k columns.  k < 0: use the device-side step count S_K and zero-fill yk up to kd = -k.
__device__ inline void gmres_solve_scalars(double* S, int k_or_neg_kd, int ld) {
  int k = k_or_neg_kd;
  int kd = 0;
  if (k < 0) { kd = -k; k = (int)S[S_K]; }
  double* Rm = S + OFF_R;
  double* y = S + OFF_Y;
  double* gc = S + OFF_GC;
  double* gs = S + OFF_GS;
  double* yk = S + OFF_YK;
  double* coef = S + OFF_COEF;
  for (int i = k - 1; i >= 0; --i) {
    double acc = y[i];
    for (int j = i + 1; j < k; ++j) acc -= Rm[i * ld + j] * yk[j];
    yk[i] = acc / Rm[i * ld + i];
  }
  for (int i = 0; i <= k; ++i) coef[i] = 0.0;
  coef[k] = 1.0;
  for (int i = k - 1; i >= 0; --i) {
    double t1 = coef[i], t2 = coef[i + 1];
    coef[i] = gc[i] * t1 - gs[i] * t2;
    coef[i + 1] = gs[i] * t1 + gc[i] * t2;
  }
  for (int i = 0; i <= k; ++i) coef[i] *= y[k];
  S[S_R0N2] = y[k] * y[k];
  for (int i = k; i < kd; ++i) yk[i] = 0.0;
}

// column k-1 of the Lanczos projected matrix from the CGS2 coefficients; subdiagonal = residual norm
__device__ inline void lanczos_step_scalars(double* S, int k, int ld) {
  if (S[S_DONE] != 0.0) return;
  double* T = S + OFF_R;
  double nres = sqrt(S[S_NRES2]);
  double hn2 = nres * nres;
  for (int i = 0; i < k; ++i) {
    double h = S[OFF_H1 + i] + S[OFF_H2 + i];
    hn2 += h * h;
    T[i * ld + (k - 1)] = h;
    T[(k - 1) * ld + i] = h;
  }
  S[S_NRES] = nres;
  S[S_K] = (double)k;
  if (!(nres > 1e-13 * sqrt(hn2))) S[S_DONE] = 1.0;  // invariant subspace: stop expanding (KrylovKit: normres <= tol)
}

}  // namespace kscal
}  // namespace qtt

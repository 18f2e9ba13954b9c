// Site splitting, fast path (real FP64): rank-revealing Householder QR followed by one-sided Jacobi on the rows of R.
// Replaces `replacebond!` -> `svd` (LAPACK gesdd) -> `truncate!`; this is the "QR followed by
// a one-sided Jacobi SVD kernel with cutoff/maxdim truncation" of the north star (Drmac-Veselic preconditioning).
//
// Problem: Z (q x p, rows z_i) -> the k dominant right singular vectors as orthonormal rows Wn (k x p) + sigma^2.
// Plain one-sided Jacobi on the rows of Z needs 14 sweeps on a flat 512 x 512 spectrum and 24-34 sweeps on the graded /
// numerically rank-deficient two-site tensors of a converging sweep (measured, profiles/r01_svd_notes.md); each sweep
// is q-1 latency-bound rounds.  Here instead
//   1. Z^T = Q R by Householder reflections applied from the RIGHT to the rows of the stacked matrix SE = [Z; I_p]
//      (so the identity block accumulates Q).  Panels of up to 32 rows are chosen by residual norm (largest first,
//      exact norms recomputed after every panel) and pivoted inside the panel; the factorisation STOPS as soon as
//      every remaining residual norm is below the zero threshold, so its cost is proportional to the numerical rank r.
//   2. Y = (SE[:, :r])^T = [R | Q^T] (r rows): one-sided Jacobi orthogonalises the rows of R (rotation angles from the
//      R part only) and carries the Q^T part along.  R R^T is far closer to diagonal than Z Z^T (9-11 sweeps instead
//      of up to 34, and only r-1 rounds per sweep).  On convergence row i is [sigma_i u_i^T | v_i^T]: the carried part
//      is already the orthonormal right singular vector, no division by sigma_i, so tiny sigma cost no accuracy.
//   3. descending sort of sigma^2, NDTensors.truncate! rule, gather of the kept rows.
// Everything is enqueued without a host round trip; only the kept rank is read back at the end.
#include "common.cuh"
#include "device_utils.cuh"
#include <cstdlib>
#include <string>
#include <cooperative_groups.h>

namespace qtt {
namespace {

constexpr int JB = 4;  // rows per block of the block-cyclic Jacobi ordering (8 was measured slower)
using dev::warp_sum;
using dev::ld_acquire;
using dev::grid_sync;

constexpr int NB = 32;          // panel width (reflectors per panel)
constexpr int PANEL_T = 512;    // threads of the panel / apply kernels (16 warps)
constexpr int PANEL_W = PANEL_T / 32;

// device-resident control block
struct QrCtl {
  int r;        // reflectors (= rows of R) so far
  int done;     // no row above the zero threshold is left, or r reached min(q, p)
  int nbp;      // reflectors produced by the most recent panel (0 = the panel kernel did nothing)
  int serial;   // panel serial number
  double F2;    // |Z|_F^2
  double thr;   // zero threshold on squared norms
  // block Gram-Schmidt variant (gs_* kernels)
  int r_next;   // r after the panel in flight
  int nsel;     // rows selected for the panel in flight
  int sel[32];  // their indices
  int rj;       // rows of Y with a non-zero R part (the rows Jacobi rotates; 0 = all r rows).  Completion rows follow them.
};

// SE rows [0,q) = Z (zero padded to ldp), rows [q, q+p) = identity; norms[i] = |z_i|^2
__global__ void qr_init_kernel(const double* __restrict__ M, int q, int p, int ldp, double* __restrict__ SE, double* __restrict__ norms,
                               int* __restrict__ rowstate, int* __restrict__ inpanel, QrCtl* ctl) {
  dev::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int GW = (gridDim.x * blockDim.x) >> 5;
  for (int i = gw; i < q + p; i += GW) {
    double* dst = SE + (size_t)i * ldp;
    if (i < q) {
      const double* src = M + (size_t)i * p;
      double acc = 0.0;
      for (int c = lane; c < ldp; c += 32) {
        double v = c < p ? src[c] : 0.0;
        dst[c] = v;
        acc += v * v;
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        norms[i] = acc;
        rowstate[i] = 0;
        inpanel[i] = -1;
      }
    } else {
      const int d = i - q;
      for (int c = lane; c < ldp; c += 32) dst[c] = (c == d) ? 1.0 : 0.0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->r = 0; ctl->done = 0; ctl->nbp = 0; ctl->serial = 0; ctl->F2 = -1.0; ctl->thr = 0.0; ctl->rj = 0;
  }
}

// One panel: select up to NB unprocessed rows by residual norm, Householder-factor them with in-panel pivoting.
// Vp (NB x ldp) receives the reflectors (v_j = 1, zeros left of j), taus their scalars.
__global__ void __launch_bounds__(PANEL_T) qr_panel_kernel(double* __restrict__ SE, int q, int p, int ldp, int rmax, int nb, double zero_tol2,
                                                           const double* __restrict__ norms, int* __restrict__ rowstate,
                                                           int* __restrict__ inpanel, double* __restrict__ Vp,
                                                           double* __restrict__ taus, QrCtl* ctl) {
  dev::pdl_wait();
  extern __shared__ double sm[];
  double* P = sm;                                      // [NB][ldp]
  double* nrm = sm + (size_t)nb * ldp;                 // [q] candidate norms (<0: not a candidate)
  __shared__ int sel[NB];
  __shared__ int proc[NB];      // in-panel: 1 once the row has become a reflector
  __shared__ int order[NB];     // order[t] = panel slot that became reflector t
  __shared__ double res[NB];
  __shared__ double beta_s[NB];
  __shared__ double tau_s[NB];
  __shared__ int s_nsel, s_piv, s_stop;
  __shared__ double s_tau;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) ctl->nbp = 0;
  if (ctl->done) return;
  const int r0 = ctl->r;
  const int serial = ctl->serial + 1;  // this panel's serial number (published below; qr_apply reads it)
  __syncthreads();
  if (tid == 0) ctl->serial = serial;
  // |Z|_F^2 once (the initial norms are the full row norms)
  if (ctl->F2 < 0.0) {
    __shared__ double red[PANEL_W];
    double acc = 0.0;
    for (int i = tid; i < q; i += PANEL_T) acc += norms[i];
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int k = 0; k < PANEL_W; ++k) s += red[k];
      ctl->F2 = s;
      ctl->thr = zero_tol2 * s;
    }
    __syncthreads();
  }
  const double thr = ctl->thr;
  int want = rmax - r0;
  if (want > nb) want = nb;
  // ---- selection: the `want` largest residual norms among the unprocessed rows, if above the threshold
  for (int i = tid; i < q; i += PANEL_T) nrm[i] = (rowstate[i] == 0) ? norms[i] : -1.0;
  if (tid < NB) { sel[tid] = -1; proc[tid] = 0; }
  if (tid == 0) { s_nsel = 0; s_stop = 0; }
  __syncthreads();
  for (int i = tid; i < q; i += PANEL_T) {
    const double ni = nrm[i];
    if (!(ni > thr)) continue;
    int rank = 0;
    for (int j = 0; j < q; ++j) {
      const double nj = nrm[j];
      rank += (nj > ni) || (nj == ni && j < i);
    }
    if (rank < want) {
      sel[rank] = i;
      atomicAdd(&s_nsel, 1);
    }
  }
  __syncthreads();
  const int nsel = s_nsel;  // ranks 0..nsel-1 are all filled (ranks are distinct and dense among rows above thr)
  if (nsel == 0) {
    if (tid == 0) ctl->done = 1;
    return;
  }
  // ---- load the panel rows
  for (int u = warp; u < nsel; u += PANEL_W) {
    const double* src = SE + (size_t)sel[u] * ldp;
    for (int c = lane; c < ldp; c += 32) P[(size_t)u * ldp + c] = src[c];
  }
  __syncthreads();
  int nfin = 0;
  // residual norms |row[r0:]|^2 of the panel rows; afterwards they are refreshed inside the update pass of every step
  for (int u = warp; u < nsel; u += PANEL_W) {
    const double* row = P + (size_t)u * ldp;
    double acc = 0.0;
    for (int c = r0 + lane; c < p; c += 32) acc += row[c] * row[c];
    acc = warp_sum(acc);
    if (lane == 0) res[u] = acc;
  }
  for (int t = 0; t < nsel; ++t) {
    const int j = r0 + t;
    __syncthreads();  // res[] of the previous update pass (or the initial norms) is complete
    if (warp == 0) {
      // in-panel pivot: the largest residual among the unprocessed rows (processed rows carry -1)
      double best = lane < nsel ? res[lane] : -1.0;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (!(best > thr)) {
        if (lane == 0) s_stop = 1;
      } else {
        // dlarfg: beta = -sign(x_j) |x[j:]|, tau = (beta - x_j) / beta, v = x[j+1:] / (x_j - beta), v_j = 1
        double* x = P + (size_t)bi * ldp;
        const double xj = x[j];
        const double nrmx = sqrt(best);
        const double beta = -copysign(nrmx, xj);
        const double tau = (beta - xj) / beta;
        const double scale = 1.0 / (xj - beta);
        for (int c = j + 1 + lane; c < p; c += 32) x[c] *= scale;
        __syncwarp();
        if (lane == 0) {
          x[j] = 1.0;
          beta_s[t] = beta;
          tau_s[t] = tau;
          s_tau = tau;
          s_piv = bi;
          order[t] = bi;
          proc[bi] = 1;
          res[bi] = -1.0;
        }
      }
    }
    __syncthreads();
    if (s_stop) break;
    const int piv = s_piv;
    const double* x = P + (size_t)piv * ldp;
    const double tau = s_tau;
    for (int u = warp; u < nsel; u += PANEL_W) {
      if (proc[u]) continue;
      double* row = P + (size_t)u * ldp;
      double d = 0.0;
      for (int c = j + lane; c < p; c += 32) d += row[c] * x[c];
      d = warp_sum(d) * tau;
      double acc = 0.0;  // residual for the next step: |row[j+1:]|^2 after the reflection
      for (int c = j + lane; c < p; c += 32) {
        const double v = row[c] - d * x[c];
        row[c] = v;
        if (c > j) acc += v * v;
      }
      acc = warp_sum(acc);
      if (lane == 0) res[u] = acc;
    }
    nfin = t + 1;
  }
  __syncthreads();
  // ---- write back: reflectors, finished rows of R^T, the partially updated leftovers
  for (int t = warp; t < nb; t += PANEL_W) {
    double* vrow = Vp + (size_t)t * ldp;
    if (t < nfin) {
      const double* x = P + (size_t)order[t] * ldp;
      const int j = r0 + t;
      for (int c = lane; c < ldp; c += 32) vrow[c] = (c >= j && c < p) ? x[c] : 0.0;
    }
  }
  for (int u = warp; u < nsel; u += PANEL_W) {
    double* dst = SE + (size_t)sel[u] * ldp;
    const double* row = P + (size_t)u * ldp;
    if (proc[u]) {
      int t = 0;
      for (int k = 0; k < nfin; ++k)
        if (order[k] == u) t = k;
      const int j = r0 + t;
      for (int c = r0 + lane; c < ldp; c += 32) dst[c] = c < j ? row[c] : (c == j ? beta_s[t] : 0.0);
      if (lane == 0) rowstate[sel[u]] = 1;
    } else {
      for (int c = r0 + lane; c < ldp; c += 32) dst[c] = row[c];
    }
    if (lane == 0) inpanel[sel[u]] = serial;
  }
  if (tid < nfin) taus[tid] = tau_s[tid];
  if (tid == 0) {
    ctl->r = r0 + nfin;
    ctl->nbp = nfin;
    if (r0 + nfin >= rmax) ctl->done = 1;
  }
}

// Applies the panel's reflectors (in order) from the right to every row of SE that was not in the panel, and refreshes
// the residual norms |z_i[r:]|^2 of all unprocessed rows.  NPL = row elements per lane.
template <int NPL>
__global__ void __launch_bounds__(PANEL_T) qr_apply_kernel(double* __restrict__ SE, int q, int p, int ldp, const int* __restrict__ rowstate,
                                                           const int* __restrict__ inpanel, const double* __restrict__ Vp,
                                                           const double* __restrict__ taus, double* __restrict__ norms, QrCtl* ctl) {
  dev::pdl_wait();
  extern __shared__ double sm[];
  const int nbp = ctl->nbp;
  if (nbp == 0) return;
  const int serial = ctl->serial;
  const int rnew = ctl->r;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* V = sm;  // [nbp][ldp]
  for (int i = tid; i < nbp * ldp; i += PANEL_T) V[i] = Vp[i];
  __shared__ double tau_s[NB];
  if (tid < nbp) tau_s[tid] = taus[tid];
  __syncthreads();
  const int gw = blockIdx.x * PANEL_W + warp;
  const int GW = gridDim.x * PANEL_W;
  for (int i = gw; i < q + p; i += GW) {
    if (i < q && rowstate[i] == 1) continue;
    const bool inpan = i < q && inpanel[i] == serial;
    double* row = SE + (size_t)i * ldp;
    double x[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int c = lane + 32 * k;
      x[k] = c < p ? row[c] : 0.0;
    }
    if (!inpan) {
      for (int t = 0; t < nbp; ++t) {
        const double* v = V + (size_t)t * ldp;
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          const int c = lane + 32 * k;
          if (c < p) d += x[k] * v[c];
        }
        d = warp_sum(d) * tau_s[t];
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
          const int c = lane + 32 * k;
          if (c < p) x[k] -= d * v[c];
        }
      }
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int c = lane + 32 * k;
        if (c < p) row[c] = x[k];
      }
    }
    if (i < q) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int c = lane + 32 * k;
        if (c >= rnew && c < p) acc += x[k] * x[k];
      }
      acc = warp_sum(acc);
      if (lane == 0) norms[i] = acc;
    }
  }
}

// generic row length: rows stay in global / L1
__global__ void __launch_bounds__(PANEL_T) qr_apply_generic_kernel(double* __restrict__ SE, int q, int p, int ldp,
                                                                   const int* __restrict__ rowstate, const int* __restrict__ inpanel,
                                                                   const double* __restrict__ Vp, const double* __restrict__ taus,
                                                                   double* __restrict__ norms, QrCtl* ctl) {
  dev::pdl_wait();
  const int nbp = ctl->nbp;
  if (nbp == 0) return;
  const int serial = ctl->serial;
  const int rnew = ctl->r;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PANEL_W + warp;
  const int GW = gridDim.x * PANEL_W;
  for (int i = gw; i < q + p; i += GW) {
    if (i < q && rowstate[i] == 1) continue;
    const bool inpan = i < q && inpanel[i] == serial;
    double* row = SE + (size_t)i * ldp;
    if (!inpan) {
      for (int t = 0; t < nbp; ++t) {
        const double* v = Vp + (size_t)t * ldp;
        double d = 0.0;
        for (int c = lane; c < p; c += 32) d += row[c] * v[c];
        d = warp_sum(d) * taus[t];
        for (int c = lane; c < p; c += 32) row[c] -= d * v[c];
        __syncwarp();
      }
    }
    if (i < q) {
      double acc = 0.0;
      for (int c = rnew + lane; c < p; c += 32) acc += row[c] * row[c];
      acc = warp_sum(acc);
      if (lane == 0) norms[i] = acc;
    }
  }
}


// Y[j][i] = SE[i][j] for j < r (tiled transpose); columns [q, qd) of Y are zero padding.
__global__ void build_y_kernel(const double* __restrict__ SE, int q, int p, int ldp, int qd, int ldy, double* __restrict__ Y, const QrCtl* ctl) {
  dev::pdl_wait();
  __shared__ double tile[32][33];
  const int r = ctl->r;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * 32;  // SE row block
  const int j0 = blockIdx.y * 32;  // SE column block (= Y row block)
  if (j0 >= r) return;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int i = i0 + ty + k, j = j0 + tx;
    tile[ty + k][tx] = (i < q + p && j < r) ? SE[(size_t)i * ldp + j] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int j = j0 + ty + k, i = i0 + tx;
    if (j < r && i < q + p) {
      const int col = i < q ? i : (qd + (i - q));
      Y[(size_t)j * ldy + col] = tile[tx][ty + k];
    }
  }
  if (qd != q && blockIdx.x == 0) {
    for (int j = j0 + ty * 32 + tx; j < j0 + 32 && j < r; j += 256) Y[(size_t)j * ldy + q] = 0.0;
  }
}

// ------------------------------------------------------------------ Jacobi on the rows of [R | Q^T]
struct JqArgs {
  double* Y;       // [rcap][ld]
  int ld, ndot;    // row stride / length; the rotation angles use columns [0, ndot)
  int kmax;        // min(q, p): length of the spectrum that is truncated (entries beyond r are zero)
  int max_sweeps;
  double tol;
  unsigned* sync;
  double* sig2;    // [kmax]
  int* perm;       // [kmax]
  double* norms;   // [rcap]
  int* info;       // [0] kept rank, [1] sweeps, [2] converged, [3] r
  double* derr;
  double cutoff;
  int maxdim, mindim;
  const QrCtl* ctl;
  long long* dbg;  // optional (QTT_TRACE): cycle counters of CTA 0 [barrier, load, steps, store]
  int eigen_mode;  // 1: the input is Hermitian PSD and the spectrum that is truncated is |row| (its eigenvalues), not |row|^2
};

// Jacobi rotation that orthogonalises two rows with |a|^2 = alpha, |b|^2 = beta, a.b = gamma: tan(theta) = t is the
// smaller root of t^2 + 2 zeta t - 1 = 0, zeta = (beta - alpha) / (2 gamma), written with ONE division and one
// (overflow-safe) hypot: t = 2 gamma / (d + sign(d) hypot(d, 2 gamma)), d = beta - alpha.
__device__ __forceinline__ void jrotation(double alpha, double beta, double gamma, double& c, double& s) {
  const double d = beta - alpha, g2 = 2.0 * gamma;
  const double h = hypot(d, g2);
  const double t = g2 / (d + copysign(h, d));
  c = rsqrt(1.0 + t * t);
  s = c * t;
}

template <int NCH>
__device__ __forceinline__ bool jq_rotate_reg(double* __restrict__ ra, double* __restrict__ rb, int ld, int ndot, double tol, double thr,
                                              int lane) {
  double2 za[NCH], zb[NCH];
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int col = 2 * lane + 64 * k;
    if (col < ld) {
      za[k] = __ldcg(reinterpret_cast<const double2*>(ra + col));
      zb[k] = __ldcg(reinterpret_cast<const double2*>(rb + col));
    } else {
      za[k] = make_double2(0.0, 0.0);
      zb[k] = make_double2(0.0, 0.0);
    }
    if (col < ndot) {
      alpha += za[k].x * za[k].x + za[k].y * za[k].y;
      beta += zb[k].x * zb[k].x + zb[k].y * zb[k].y;
      gamma += za[k].x * zb[k].x + za[k].y * zb[k].y;
    }
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gamma = warp_sum(gamma);
  if (!(alpha > thr) || !(beta > thr)) return false;
  const double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(fabs(gamma) > tol * lim)) return false;
  double c, s;
  jrotation(alpha, beta, gamma, c, s);
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int col = 2 * lane + 64 * k;
    if (col < ld) {
      __stcg(reinterpret_cast<double2*>(ra + col), make_double2(c * za[k].x - s * zb[k].x, c * za[k].y - s * zb[k].y));
      __stcg(reinterpret_cast<double2*>(rb + col), make_double2(s * za[k].x + c * zb[k].x, s * za[k].y + c * zb[k].y));
    }
  }
  return true;
}

__device__ __forceinline__ bool jq_rotate_mem(double* __restrict__ ra, double* __restrict__ rb, int ld, int ndot, double tol, double thr,
                                              int lane) {
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int col = 2 * lane; col < ndot; col += 64) {
    double2 a = __ldcg(reinterpret_cast<const double2*>(ra + col));
    double2 b = __ldcg(reinterpret_cast<const double2*>(rb + col));
    alpha += a.x * a.x + a.y * a.y;
    beta += b.x * b.x + b.y * b.y;
    gamma += a.x * b.x + a.y * b.y;
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gamma = warp_sum(gamma);
  if (!(alpha > thr) || !(beta > thr)) return false;
  const double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(fabs(gamma) > tol * lim)) return false;
  double c, s;
  jrotation(alpha, beta, gamma, c, s);
  for (int col = 2 * lane; col < ld; col += 64) {
    double2 a = __ldcg(reinterpret_cast<const double2*>(ra + col));
    double2 b = __ldcg(reinterpret_cast<const double2*>(rb + col));
    __stcg(reinterpret_cast<double2*>(ra + col), make_double2(c * a.x - s * b.x, c * a.y - s * b.y));
    __stcg(reinterpret_cast<double2*>(rb + col), make_double2(s * a.x + c * b.x, s * a.y + c * b.y));
  }
  return true;
}

template <int NCH>
__global__ void __launch_bounds__(64) jacobi_qr_kernel(JqArgs a) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int GW = gridDim.x * wpb;
  const int q = a.ctl->r;  // rows of R
  const double thr = a.ctl->thr;
  const int qe = (q + 1) & ~1;
  const int npairs = qe / 2, rounds = qe - 1;
  unsigned target = 0;
  int sweeps = 0;
  int converged = (q < 2) ? 1 : 0;
  for (int sw = 0; sw < a.max_sweeps && !converged; ++sw) {
    bool rotated = false;
    for (int r = 0; r < rounds; ++r) {
      for (int i = gw; i < npairs; i += GW) {
        int ia, ib;
        if (i == 0) {
          ia = r;
          ib = qe - 1;
        } else {
          ia = (r + i) % (qe - 1);
          ib = (r - i + (qe - 1)) % (qe - 1);
        }
        if (ia >= q || ib >= q) continue;
        if (ia > ib) {
          int tmp = ia;
          ia = ib;
          ib = tmp;
        }
        double* ra = a.Y + (size_t)ia * a.ld;
        double* rb = a.Y + (size_t)ib * a.ld;
        bool rot;
        if constexpr (NCH > 0) rot = jq_rotate_reg<NCH>(ra, rb, a.ld, a.ndot, a.tol, thr, lane);
        else rot = jq_rotate_mem(ra, rb, a.ld, a.ndot, a.tol, thr, lane);
        rotated |= rot;
      }
      grid_sync(a.sync, target);
    }
    if (rotated && lane == 0) atomicAdd(a.sync + 2 + sw, 1u);
    grid_sync(a.sync, target);
    sweeps = sw + 1;
    if (ld_acquire(a.sync + 2 + sw) == 0u) converged = 1;
  }
  // sigma_i^2 = |R part of row i|^2
  for (int i = gw; i < q; i += GW) {
    const double* ri = a.Y + (size_t)i * a.ld;
    double acc = 0.0;
    for (int col = 2 * lane; col < a.ndot; col += 64) {
      double2 v = __ldcg(reinterpret_cast<const double2*>(ri + col));
      acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if (!(acc > thr)) acc = 0.0;
    if (a.eigen_mode) acc = sqrt(acc);
    if (lane == 0) __stcg(a.norms + i, acc);
  }
  grid_sync(a.sync, target);
  dev::rank_sort_desc(a.norms, q, a.kmax, a.sig2, a.perm);
  grid_sync(a.sync, target);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err;
    a.info[0] = dev::truncate_spectrum(a.sig2, a.kmax, a.cutoff, a.maxdim, a.mindim, &err);
    a.info[1] = sweeps;
    a.info[2] = converged;
    a.info[3] = q;
    a.derr[0] = err;
  }
}

// ---- helpers of the block-cyclic kernel ---------------------------------------------------------------------------
// (Intermediate variants, measured and superseded -- one warp per pair with a grid barrier per round, two warps per pair,
// point-to-point neighbour flags instead of the barrier, 8-row blocks -- are recorded in profiles/r01_svd_notes.md.)
__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

// ---- block-cyclic variant: 4-row blocks, several rotation steps per grid barrier ---------------------------------------
// A round of the row-cyclic kernels costs ~3 us and is pure latency (barrier + L2 round trips), whatever the number of
// rows.  Here the tournament is played between BLOCKS of 4 rows: a CTA (4 warp pairs) takes two blocks P and M, keeps
// the P rows in registers and rotates them against all four M rows in 4 steps, the M rows circulating between the
// warp pairs through shared memory (one __syncthreads per step, ~0.4 us).  One grid barrier therefore covers 4
// rotation steps; the pairs inside a block are swept by three extra global rounds per sweep.  It is still a cyclic
// one-sided Jacobi ordering (every row pair exactly once per sweep).


// rows held as half-rows in registers (za: row A, zb: row B); role 0 computes the angle from the R part
template <int NC>
__device__ __forceinline__ bool rotate_halfrows(double2 (&za)[NC], double2 (&zb)[NC], int role, int unit_in_cta, double (*cs_s)[8], int& parity,
                                                double tol, double thr, int lane) {
  double* slot = cs_s[unit_in_cta] + 4 * parity;  // double-buffered by call parity: one pair barrier per rotation
  parity ^= 1;
  if (role == 0) {
    double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      alpha += za[k].x * za[k].x + za[k].y * za[k].y;
      beta += zb[k].x * zb[k].x + zb[k].y * zb[k].y;
      gamma += za[k].x * zb[k].x + za[k].y * zb[k].y;
    }
    alpha = warp_sum(alpha);
    beta = warp_sum(beta);
    gamma = warp_sum(gamma);
    double c = 1.0, s = 0.0, rot = 0.0;
    // |gamma| > tol sqrt(alpha beta), without the square roots
    if ((alpha > thr) && (beta > thr) && (gamma * gamma > (tol * tol) * (alpha * beta))) {
      jrotation(alpha, beta, gamma, c, s);
      rot = 1.0;
    }
    if (lane == 0) {
      slot[0] = c;
      slot[1] = s;
      slot[2] = rot;
    }
  }
  pair_bar(1 + unit_in_cta);
  const double c = slot[0], s = slot[1];
  const bool rot = slot[2] != 0.0;
  if (rot) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double2 x = za[k], y = zb[k];
      za[k] = make_double2(c * x.x - s * y.x, c * x.y - s * y.y);
      zb[k] = make_double2(s * x.x + c * y.x, s * x.y + c * y.y);
    }
  }
  return rot;
}

template <int NC>
__device__ __forceinline__ void load_half(double2 (&z)[NC], const double* row, int len, int lane, bool live) {
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int col = 2 * lane + 64 * k;
    z[k] = (live && col < len) ? __ldcg(reinterpret_cast<const double2*>(row + col)) : make_double2(0.0, 0.0);
  }
}
template <int NC>
__device__ __forceinline__ void store_half(const double2 (&z)[NC], double* row, int len, int lane, bool live) {
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int col = 2 * lane + 64 * k;
    if (live && col < len) __stcg(reinterpret_cast<double2*>(row + col), z[k]);
  }
}

template <int NC, int JBT>
__global__ void __launch_bounds__(64 * JBT) jacobi_qr_block_kernel(JqArgs a) {
  static_assert(JBT == JB, "the kernel body is written for the global block size");
  extern __shared__ double mbuf[];  // [2][JB][ld]: circulating M rows
  __shared__ double cs_s[JB][8];
  int parity = 0;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int unit = warp >> 1;   // 0..JB-1
  const int role = warp & 1;    // 0: R part (angles), 1: carried part
  const int q = a.ctl->r;
  const double thr = a.ctl->thr;
  const int nbk = (((q + JB - 1) / JB) + 1) & ~1;  // blocks, padded to even
  const int nslots = nbk / 2, orounds = nbk - 1;
  const int c_lo = role == 0 ? 0 : a.ndot;
  const int len = role == 0 ? a.ndot : a.ld - a.ndot;
  const int ld = a.ld;
  unsigned target = 0;
  int sweeps = 0;
  int converged = (q < 2) ? 1 : 0;
  for (int sw = 0; sw < a.max_sweeps && !converged; ++sw) {
    bool rotated = false;
    // ---- pairs inside a block: three rounds over all blocks, rows through global memory
    for (int st = 0; st < JB - 1; ++st) {
      const int gunit = blockIdx.x * JB + unit, gunits = gridDim.x * JB;
      for (int pidx = gunit; pidx < (JB / 2) * nbk; pidx += gunits) {
        const int blk = pidx / (JB / 2), i = pidx % (JB / 2);
        int u0, u1;  // round-robin tournament inside the block
        if (i == 0) { u0 = st; u1 = JB - 1; }
        else { u0 = (st + i) % (JB - 1); u1 = (st - i + (JB - 1)) % (JB - 1); }
        const int ia = blk * JB + u0, ib = blk * JB + u1;
        const bool live = ia < q && ib < q;
        double2 za[NC], zb[NC];
        double* ra = a.Y + (size_t)ia * ld + c_lo;
        double* rb = a.Y + (size_t)ib * ld + c_lo;
        load_half<NC>(za, ra, len, lane, live);
        load_half<NC>(zb, rb, len, lane, live);
        if (rotate_halfrows<NC>(za, zb, role, unit, cs_s, parity, a.tol, thr, lane)) {
          rotated = true;
          store_half<NC>(za, ra, len, lane, live);
          store_half<NC>(zb, rb, len, lane, live);
        }
      }
      grid_sync(a.sync, target);
    }
    // ---- block tournament
    for (int R = 0; R < orounds; ++R) {
      long long tc0 = clock64();
      for (int i = blockIdx.x; i < nslots; i += gridDim.x) {
        int bp, bm;
        if (i == 0) {
          bp = R;
          bm = nbk - 1;
        } else {
          bp = (R + i) % (nbk - 1);
          bm = (R - i + (nbk - 1)) % (nbk - 1);
        }
        const int ip = bp * JB + unit;
        const bool plive = ip < q;
        double2 zp[NC], zm[NC];
        double* rp = a.Y + (size_t)ip * ld + c_lo;
        load_half<NC>(zp, rp, len, lane, plive);
        int v = unit;  // M row currently held
        load_half<NC>(zm, a.Y + (size_t)(bm * JB + v) * ld + c_lo, len, lane, bm * JB + v < q);
        bool pdirty = false;
        long long tc1 = clock64();
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[1] += tc1 - tc0;
#pragma unroll 1
        for (int t = 0; t < JB; ++t) {
          const bool rot = rotate_halfrows<NC>(zp, zm, role, unit, cs_s, parity, a.tol, thr, lane);
          pdirty |= rot;
          rotated |= rot;
          if (t + 1 < JB) {
            // hand M[v] to the previous unit, take M[v + 1] from the next one
            double* wb = mbuf + ((size_t)(t & 1) * JB + v) * ld + c_lo;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
              const int col = 2 * lane + 64 * k;
              if (col < len) *reinterpret_cast<double2*>(wb + col) = zm[k];
            }
            __syncthreads();
            v = (v + 1) & (JB - 1);
            const double* rbuf = mbuf + ((size_t)(t & 1) * JB + v) * ld + c_lo;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
              const int col = 2 * lane + 64 * k;
              zm[k] = col < len ? *reinterpret_cast<const double2*>(rbuf + col) : make_double2(0.0, 0.0);
            }
          }
        }
        long long tc2 = clock64();
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[2] += tc2 - tc1;
        if (pdirty) store_half<NC>(zp, rp, len, lane, plive);
        // the M row may have been rotated by another unit: always written back
        store_half<NC>(zm, a.Y + (size_t)(bm * JB + v) * ld + c_lo, len, lane, bm * JB + v < q);
        if (nslots > (int)gridDim.x) __syncthreads();  // mbuf is reused by this CTA's next slot
        tc0 = clock64();
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[3] += tc0 - tc2;
      }
      long long tb0 = clock64();
      grid_sync(a.sync, target);
      if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) { a.dbg[0] += clock64() - tb0; a.dbg[4] += 1; }
    }
    if (rotated && lane == 0 && role == 0) atomicAdd(a.sync + 2 + sw, 1u);
    grid_sync(a.sync, target);
    sweeps = sw + 1;
    if (ld_acquire(a.sync + 2 + sw) == 0u) converged = 1;
  }
  const int gw = blockIdx.x * (2 * JB) + warp;
  const int GW = gridDim.x * (2 * JB);
  for (int i = gw; i < q; i += GW) {
    const double* ri = a.Y + (size_t)i * a.ld;
    double acc = 0.0;
    for (int col = 2 * lane; col < a.ndot; col += 64) {
      double2 v = __ldcg(reinterpret_cast<const double2*>(ri + col));
      acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if (!(acc > thr)) acc = 0.0;
    if (a.eigen_mode) acc = sqrt(acc);
    if (lane == 0) __stcg(a.norms + i, acc);
  }
  grid_sync(a.sync, target);
  dev::rank_sort_desc(a.norms, q, a.kmax, a.sig2, a.perm);
  grid_sync(a.sync, target);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err;
    a.info[0] = dev::truncate_spectrum(a.sig2, a.kmax, a.cutoff, a.maxdim, a.mindim, &err);
    a.info[1] = sweeps;
    a.info[2] = converged;
    a.info[3] = q;
    a.derr[0] = err;
  }
}

template <int NC, int JBT>
void launch_jq_block(Ctx* ctx, JqArgs& a, int grid, size_t smem) {
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(jacobi_qr_block_kernel<NC, JBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  });
  void* args[] = {&a};
  QTT_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_qr_block_kernel<NC, JBT>, dim3(grid), dim3(64 * JBT), args, smem, ctx->stream));
  ctx->launches++;
}

// ---- blocked variant: Gram matrix of a block pair on the DMMA core, inner two-sided Jacobi in shared memory -----------
// The row-cyclic kernels above pay one latency-bound step (barrier + L2 round trips, ~1.4-5.7 us) per 1-4 rotation rounds
// whatever the row length: 128 grid barriers per sweep at r = 512.  Here the tournament is played between blocks of B = 32
// rows.  One thread-block CLUSTER (8 CTAs) takes a block pair (64 rows); CTA c owns a column slab of the R part and of the
// carried part.  Per round:
//   1. partial Gram matrix (upper-triangular 8 x 8 tiles of the 64 x 64 matrix) of the R-part slab on the DMMA core; the
//      accumulators are pushed straight into the staging buffer of the CTA that owns the entry (distributed shared memory),
//      the owner adds the cs partial sums in rank order and pushes the result to every CTA, so all CTAs hold a bit-identical,
//      exactly symmetric G;
//   2. every CTA runs the same inner Jacobi on G in shared memory: the B^2 cross pairs (i in P, j in M) in B steps of B
//      disjoint rotations, G <- Rot G Rot^T by symmetric 2 x 2 quads from one buffer into the other, while warp 0 already
//      derives the rotations of the next step from the old buffer (one barrier per step).  Angles come from the updated
//      Gram entries, which two-sided Jacobi keeps accurate relative to sqrt(g_ii g_jj) (Demmel-Veselic), and G is rebuilt
//      from the rows themselves every round, so the accuracy of one-sided Jacobi is kept;
//   3. the 32 x 32 recorded rotations are applied to the slab columns held in REGISTERS (one column = 64 doubles per thread,
//      statically unrolled schedule, (c, s) broadcast from shared memory) and written straight back to global memory.
// One grid barrier per round (15 per sweep at r = 512) moves the blocks between clusters through L2; the pairs inside a
// block are swept once per sweep by a round with the in-block tournament (B - 1 steps, rotations accumulated into J and
// applied as a DMMA GEMM).  Every row pair is still rotated exactly once per sweep (a cyclic one-sided Jacobi ordering).
namespace cg = cooperative_groups;
// Threads per CTA and their roles in the inner loop.  The look-ahead warp (warp 0) is latency-critical: a dependent chain of
// ~45 FP64 operations per step.  The warps that share its scheduler / FP64 pipe (warps 4, 8, 12) therefore stay idle in the
// loop (measured: the look-ahead step 1200 -> 600 cycles at B = 32): Gram update on warps 1-3 and 5-7, slab columns in
// registers on warps 9-11 and 13-15.
constexpr int GNT = 512;
constexpr int GQT = 192;          // threads of the Gram-update warps
constexpr int GAT = 192;          // threads of the slab warps: two adjacent lanes share one slab column
constexpr int GNW = GNT / 32;

// compile-time geometry for blocks of B rows
template <int B>
struct GG {
  static constexpr int R = 2 * B;                 // rows of a block pair
  static constexpr int H = B / 2;                 // rows per lane of a lane pair
  static constexpr int JLD = R + 4;               // row stride of J (= 4 mod 16: conflict-free DMMA fragment loads)
  static constexpr int LD = R + 1;                // row stride of G (odd: column accesses are conflict-free)
  static constexpr int TRI = (B * (B + 1)) / 2;   // quads (row pair <= column pair) of the symmetric update
  static constexpr int QPT = (TRI + GQT - 1) / GQT;  // quads per Gram-update thread
  static constexpr int T8 = R / 8;                // 8 x 8 tiles per dimension
  static constexpr int UT = (T8 * (T8 + 1)) / 2;  // upper-triangular tiles
  static constexpr int TPW = (UT + GNW - 1) / GNW;   // tiles per warp
  static constexpr int UTE = UT * 64;             // entries of the upper-triangular tiles
  static constexpr int NC = B <= 16 ? 2 : 1;      // slab columns per lane pair of the register warps
};

struct GjArgs {
  JqArgs j;
  int skip;        // measurement only (QTT_GRAM_SKIP): 1 = no Gram update, 2 = no slab rotations, 4 = slab rotations after the inner loop
  int wR, wQ;      // slab widths (multiples of 8) of the R part and of the carried part per CTA
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}

// Rotation from Gram entries (alpha, beta, gamma) without division or square root on the critical path of the inner loop:
//   r = 1 / hypot(d, 2 gamma), cos(2 theta) = |d| r, c^2 = (1 + cos 2 theta) / 2, s = sign(d) gamma r / c,
// with both reciprocal square roots taken from the hardware approximation (MUFU.RSQ64H, ~2^-22): the ANGLE only has to
// be approximately right (the off-diagonal element shrinks by 1e-6 instead of vanishing; convergence stays quadratic
// down to the 2.5e-15 threshold), what must hold to rounding is c^2 + s^2 = 1, and that is restored by two Newton
// terms on e = c^2 + s^2 - 1 (|e| ~ 1e-6: e^3 is below double rounding).  Measured in isolation: 310 -> 150 cycles.
__device__ __forceinline__ double rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ void grot(double alpha, double beta, double gamma, double& c, double& s) {
  const double d = beta - alpha, g2 = 2.0 * gamma;
  const double r = rsqrt_approx(d * d + g2 * g2);
  const double c2 = 0.5 + 0.5 * (fabs(d) * r);
  const double rc = rsqrt_approx(c2);
  const double c0 = c2 * rc;
  const double s0 = copysign(0.5 * r * rc, d) * g2;
  const double e = fma(c0, c0, fma(s0, s0, -1.0));
  const double k = fma(e, fma(e, 0.375, -0.5), 1.0);   // (1 + e)^(-1/2) = 1 - e/2 + 3 e^2 / 8 - ...
  c = c0 * k;
  s = s0 * k;
}

__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 0;\n" ::: "memory"); }


// position of pair a of a step inside the (c, s) table: the two lanes of a lane pair (h = a / H) read adjacent entries
template <int B>
__device__ __forceinline__ int rpos(int a) { return ((a & (B / 2 - 1)) << 1) | (a / (B / 2)); }

// The recorded rotations of a cross round applied to one slab column shared by two adjacent lanes (h = lane & 1).
// Lane h keeps the P rows H h ... H h + H - 1 (xp) and a sliding window of H M rows (xm): in step t, P row a = H h + j meets
// M row (a + t) mod B, which sits in slot (j + t) mod H of the SAME lane; after the step the two lanes swap slot t mod H
// (M row t leaves lane 0 for lane 1, M row t + H comes back).  B doubles per lane instead of 2 B, all register indices static.
// SYNC: one CTA barrier after every step (the column follows the inner Jacobi step by step, see the kernel).
// NC slab columns per lane pair share every (c, s) load (blocks of 16 rows: 2 columns = 32 doubles per lane).
template <int B, int NC, bool SYNC>
__device__ __forceinline__ void apply_cross(double (&xp)[NC][B / 2], double (&xm)[NC][B / 2], int h, const double2* __restrict__ cs2) {
  constexpr int H = B / 2;
  const double2* tab = cs2 + h;
#pragma unroll
  for (int t = 0; t < B; ++t) {
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double2 r = tab[t * B + 2 * j];
#pragma unroll
      for (int n = 0; n < NC; ++n) {
        const double u = xp[n][j], v = xm[n][(j + t) & (H - 1)];
        xp[n][j] = r.x * u - r.y * v;
        xm[n][(j + t) & (H - 1)] = r.y * u + r.x * v;
      }
    }
#pragma unroll
    for (int n = 0; n < NC; ++n) xm[n][t & (H - 1)] = __shfl_xor_sync(0xffffffffu, xm[n][t & (H - 1)], 1);
    if (SYNC) cta_bar();
  }
}

template <int B>
__global__ void __launch_bounds__(GNT, 1) jacobi_gram_kernel(GjArgs ga) {
  dev::pdl_wait();   // no-op unless launched with programmatic stream serialisation (sole context)
  using K = GG<B>;
  constexpr int GR = K::R, GLD = K::LD, GJLD = K::JLD, H = K::H, NC = K::NC;
  constexpr int GAC = (GAT / 2) * NC;   // slab columns the register warps hold at a time
  const JqArgs& a = ga.j;
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wR = ga.wR, wQ = ga.wQ, W = wR + wQ, SLD = wR + 4;
  extern __shared__ __align__(16) double gsm[];
  double* S = gsm;                      // [GR][SLD] R-part slab (Gram operand)
  double* A = S + (size_t)GR * SLD;     // region A [GR * GJLD]: staging of the reduce; then (c, s) tables (cross) or J (inside-block round)
  double* G0 = A + GR * GJLD;           // [GR][GLD]
  double* G1 = G0 + GR * GLD;           // [GR][GLD]  (cross rounds: only the entries (i, j), i <= j, of G0 / G1 are maintained)
  double2* cs2 = reinterpret_cast<double2*>(A);  // [B steps][B pairs at rpos()]
  double* J = A;
  __shared__ unsigned char tri_r[K::TRI], tri_c[K::TRI];
  __shared__ int plo[2][B], phi[2][B];
  __shared__ double pc[2][B], ps[2][B];
  __shared__ int s_any;
  for (int k = tid; k < K::TRI; k += GNT) {
    int r = 0, rem = k;
    while (rem >= B - r) { rem -= B - r; ++r; }
    tri_r[k] = (unsigned char)r;
    tri_c[k] = (unsigned char)(r + rem);
  }
  __syncthreads();
  // static work assignment: Gram-update thread qt owns the quads qt, qt + 224, ... of the triangular enumeration
  const bool is_quad = warp >= 1 && warp < 8 && warp != 4;
  const bool is_apply = warp >= 9 && warp != 12;
  const int qt = (warp - 1 - (warp > 4)) * 32 + lane;                 // 0 .. 191 on the Gram-update warps
  const int at = (warp - 9 - (warp > 12)) * 32 + lane, acol = at >> 1, ah = at & 1;   // 0 .. 191 on the slab warps
  int qra[K::QPT], qcb[K::QPT];
  bool qok[K::QPT];
#pragma unroll
  for (int u = 0; u < K::QPT; ++u) {
    const int k = qt + u * GQT;
    qok[u] = is_quad && k < K::TRI;
    qra[u] = qok[u] ? tri_r[k] : 0;
    qcb[u] = qok[u] ? tri_c[k] : 0;
  }
  // Gram tiles of this warp: tl = warp, warp + 16, ... of the upper-triangular 8 x 8 tiles
  int gti[K::TPW], gtj[K::TPW];
#pragma unroll
  for (int u = 0; u < K::TPW; ++u) {
    int ti = 0, rem = warp + u * GNW;
    if (rem >= K::UT) rem = K::UT - 1;   // masked below
    while (rem >= K::T8 - ti) { rem -= K::T8 - ti; ++ti; }
    gti[u] = ti; gtj[u] = ti + rem;
  }

  const int qall = a.ctl->r;                               // rows of Y (sorted / gathered at the end)
  const int q = a.ctl->rj > 0 ? a.ctl->rj : qall;          // rows with a non-zero R part: the ones that are rotated
  const double thr = a.ctl->thr;
  const double tol2 = a.tol * a.tol;
  const int nbk = (((q + B - 1) / B) + 1) & ~1;
  const int nslots = nbk / 2, orounds = nbk - 1;
  const int ld = a.ld, ndot = a.ndot;
  const int cR0 = crank * wR, cQ0 = ndot + crank * wQ;   // this CTA's global column ranges
  const int per = K::UTE / cs;                             // entries of the upper-triangular tiles owned per CTA (cs divides UTE, per even)
  unsigned target = 0;
  int sweeps = 0;
  int converged = (q < 2) ? 1 : 0;
  long long pacc[8] = {};
  __syncthreads();
  for (int sw = 0; sw < a.max_sweeps && !converged; ++sw) {
    bool rotated = false;
    for (int R = -1; R < orounds; ++R) {   // R = -1: pairs inside the blocks
      for (int slot = cid; slot < nslots; slot += ncl) {
        int bp, bm;
        if (R < 0) { bp = 2 * slot; bm = 2 * slot + 1; }
        else if (slot == 0) { bp = R; bm = nbk - 1; }
        else { bp = (R + slot) % (nbk - 1); bm = (R - slot + (nbk - 1)) % (nbk - 1); }
        auto grow = [&](int i) { return i < B ? bp * B + i : bm * B + (i - B); };
        const bool dbg = a.dbg && blockIdx.x == 0 && tid == 0;
        const bool dbg_all = a.dbg && tid == 0;   // every CTA: per-phase maximum / sum over the CTAs (skew behind the grid barrier)
        long long tc = dbg_all ? clock64() : 0;
        auto lap = [&](int slot_) {
          if (dbg_all) { long long n = clock64(); if (dbg) a.dbg[slot_] += n - tc; pacc[slot_] += n - tc; tc = n; }
        };
        // slab column of the register warps (cross rounds): requested now, used after the Gram / reduce phases
        double xp[NC][H], xm[NC][H];
        // lane pair `acol` holds the slab columns c0 + acol + n * (GAT / 2), n < NC
        auto col_of = [&](int c, int& gc) { gc = c < wR ? cR0 + c : cQ0 + (c - wR); return c < W && gc < (c < wR ? ndot : ld); };
        auto col_load = [&](int c0) {
#pragma unroll
          for (int n = 0; n < NC; ++n) {
            int gc;
            const bool ok = col_of(c0 + acol + n * (GAT / 2), gc);
#pragma unroll
            for (int j = 0; j < H; ++j) {
              const int gp = grow(H * ah + j), gm = grow(B + H * ah + j);
              xp[n][j] = (ok && gp < q) ? __ldcg(a.Y + (size_t)gp * ld + gc) : 0.0;
              xm[n][j] = (ok && gm < q) ? __ldcg(a.Y + (size_t)gm * ld + gc) : 0.0;
            }
          }
        };
        auto col_store = [&](int c0) {
#pragma unroll
          for (int n = 0; n < NC; ++n) {
            int gc;
            const bool ok = col_of(c0 + acol + n * (GAT / 2), gc);
#pragma unroll
            for (int j = 0; j < H; ++j) {
              const int gp = grow(H * ah + j), gm = grow(B + H * ah + j);
              if (ok && gp < q) __stcg(a.Y + (size_t)gp * ld + gc, xp[n][j]);
              if (ok && gm < q) __stcg(a.Y + (size_t)gm * ld + gc, xm[n][j]);
            }
          }
        };
        if (R >= 0 && is_apply) col_load(0);
        // ---- 1. R-part slab load (cp.async, zero fill beyond the rank / the row length)
        {
          const int cpr = wR / 2;  // 16-byte chunks per slab row
          for (int id = tid; id < GR * cpr; id += GNT) {
            const int i = id / cpr, c = (id % cpr) * 2;
            const int gi = grow(i), gc = cR0 + c;
            const bool ok = gi < q && gc < ndot;
            cp_async16(S + (size_t)i * SLD + c, ok ? a.Y + (size_t)gi * ld + gc : a.Y, ok ? 16 : 0);
          }
          asm volatile("cp.async.commit_group;\n" ::);
          asm volatile("cp.async.wait_group 0;\n" ::);
          __syncthreads();
        }
        lap(1);
        // ---- 2. partial Gram, upper-triangular tiles only, pushed to the owners' staging buffers
        {
          double acc[K::TPW][2] = {};
          const double* sa[K::TPW];
          const double* sb[K::TPW];
#pragma unroll
          for (int u = 0; u < K::TPW; ++u) {
            sa[u] = S + (size_t)(gti[u] * 8 + g) * SLD + t4;
            sb[u] = S + (size_t)(gtj[u] * 8 + g) * SLD + t4;
          }
          if (warp < K::UT) {
            for (int k0 = 0; k0 < wR; k0 += 4) {
#pragma unroll
              for (int u = 0; u < K::TPW; ++u) dmma884(acc[u][0], acc[u][1], sa[u][k0], sb[u][k0]);
            }
          }
#pragma unroll
          for (int u = 0; u < K::TPW; ++u) {
            if (warp + u * GNW >= K::UT) continue;
            const int idx = (warp + u * GNW) * 64 + g * 8 + 2 * t4;   // (tile, row g, columns 2 t4, 2 t4 + 1)
            double* dst = cluster.map_shared_rank(A, idx / per) + crank * per + (idx % per);   // per is even: the pair stays together
            *reinterpret_cast<double2*>(dst) = make_double2(acc[u][0], acc[u][1]);
          }
        }
        lap(2);
        cluster.sync();
        lap(3);
        // ---- 3. owner: sum the cs partials in rank order, push the entry and its mirror to every CTA of the cluster
        for (int e = tid; e < per; e += GNT) {
          double sum = 0.0;
          for (int k = 0; k < cs; ++k) sum += A[k * per + e];
          const int idx = crank * per + e;
          const int tl = idx >> 6, within = idx & 63;
          int ti = 0, rem = tl;
          while (rem >= K::T8 - ti) { rem -= K::T8 - ti; ++ti; }
          const int tj = ti + rem;
          const int i = ti * 8 + (within >> 3), jx = tj * 8 + (within & 7);
          for (int k = 0; k < cs; ++k) {
            double* Gk = cluster.map_shared_rank(G0, k);
            Gk[i * GLD + jx] = sum;
            Gk[jx * GLD + i] = sum;
          }
        }
        cluster.sync();
        lap(4);
        // ---- 4. anything to rotate?  (identical decision in every CTA of the cluster: G is bit-identical)
        const bool intra = R < 0;
        {
          int cand = 0;
          for (int e = tid; e < GR * GR; e += GNT) {
            const int i = e / GR, jx = e % GR;
            const bool pairok = intra ? (i < jx && (i / B) == (jx / B)) : (i < B && jx >= B);
            if (pairok) {
              const double gii = G0[i * GLD + i], gjj = G0[jx * GLD + jx], gij = G0[i * GLD + jx];
              cand |= (gii > thr) && (gjj > thr) && (gij * gij > tol2 * (gii * gjj));
            }
          }
          if (tid == 0) s_any = 0;
          if (!__syncthreads_or(cand)) continue;
        }
        if (a.dbg && tid == 0 && crank == 0 && sw < 64) atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg) + 32 + sw, 1ull);
        lap(5);
        if (!intra) {
          // ---- 5a. cross round: B steps; step t pairs row a (P) with row B + (a + t) mod B (M).  Per step ONE barrier:
          //   warp 0       rotations of step t + 1 from the buffer that step t reads (look-ahead through the bilinear form)
          //   warps 1-3, 5-7    G <- Rot G Rot^T, symmetric quads (statically assigned), upper triangle, buffer to buffer
          //   warps 9-11, 13-15   the rotations of step t on the slab columns, in registers
          // Every role runs its own branch-free loop (a taken branch costs the lone look-ahead warp an instruction fetch:
          // measured ~50 cycles each, ten of them per step in the first version of this loop).
          int anyrot = 0;
          if (warp == 0) {
            const int lo = lane & (B - 1), hi = B + lo;
            const double al = G0[lo * GLD + lo], be = G0[hi * GLD + hi], gm = G0[lo * GLD + hi];
            double c, s;
            const bool rot = lane < B && (al > thr) && (be > thr) && (gm * gm > tol2 * (al * be));
            grot(al, be, gm, c, s);
            if (lane < B) cs2[rpos<B>(lane)] = rot ? make_double2(c, s) : make_double2(1.0, 0.0);
            anyrot |= rot;
          }
          __syncthreads();
          if (is_apply) {
            if (ga.skip & 6) {
              for (int st = 0; st < B; ++st) cta_bar();
              if (ga.skip & 4) apply_cross<B, NC, false>(xp, xm, ah, cs2);   // A/B: rotations applied after the loop
            } else {
              apply_cross<B, NC, true>(xp, xm, ah, cs2);
            }
          } else if (warp == 0) {
            // look-ahead: entries (lo, lo), (hi, hi), (lo, hi) of the NEXT matrix from the buffer step st reads,
            // G'[x][y] = sum_{x' in {x, px}} sum_{y' in {y, py}} k_x(x') k_y(y') G[x'][y']   (stored entries: row index <= column index).
            // The last pass (st = B - 1) writes a table row nobody reads.
            const int x_ = lane & (B - 1);
            long long tcomp = 0, tbar = 0, tprev = a.dbg ? clock64() : 0;
#pragma unroll 1
            for (int st = 0; st < B; ++st) {
              const double* __restrict__ Gc = (st & 1) ? G1 : G0;
              const double2* rt = cs2 + st * B;
              const int y_ = B + ((x_ + st + 1) & (B - 1));
              const int px = B + ((x_ + st) & (B - 1));          // partner of P row x in step st
              const int ay = (y_ - B - st) & (B - 1);             // pair index of M row y in step st; its partner is row ay
              const double2 rx = rt[rpos<B>(x_)], ry = rt[rpos<B>(ay)];
              const double kx0 = rx.x, kx1 = -rx.y;              // G'[x] = c G[x] - s G[px]
              const double ky0 = ry.x, ky1 = ry.y;               // G'[y] = c G[y] + s G[ay]
              const int xa0 = min(x_, ay), xa1 = max(x_, ay);
              const int py0 = min(px, y_), py1 = max(px, y_);
              const double gxx = Gc[x_ * GLD + x_], gxp = Gc[x_ * GLD + px], gpp = Gc[px * GLD + px];
              const double gyy = Gc[y_ * GLD + y_], gya = Gc[ay * GLD + y_], gaa = Gc[ay * GLD + ay];
              const double gxy = Gc[x_ * GLD + y_], gxa = Gc[xa0 * GLD + xa1], gpy = Gc[py0 * GLD + py1], gpa = Gc[ay * GLD + px];
              const double al = kx0 * (kx0 * gxx + kx1 * gxp) + kx1 * (kx0 * gxp + kx1 * gpp);
              const double be = ky0 * (ky0 * gyy + ky1 * gya) + ky1 * (ky0 * gya + ky1 * gaa);
              const double gm = kx0 * (ky0 * gxy + ky1 * gxa) + kx1 * (ky0 * gpy + ky1 * gpa);
              double c, s;
              const bool rot = lane < B && st + 1 < B && (al > thr) && (be > thr) && (gm * gm > tol2 * (al * be));
              grot(al, be, gm, c, s);
              if (lane < B) cs2[(st + 1) * B + rpos<B>(lane)] = rot ? make_double2(c, s) : make_double2(1.0, 0.0);
              anyrot |= rot;
              if (a.dbg) { const long long t1 = clock64(); tcomp += t1 - tprev; tprev = t1; }
              cta_bar();
              if (a.dbg) { const long long t1 = clock64(); tbar += t1 - tprev; tprev = t1; }
            }
            if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[9] += tcomp; a.dbg[10] += tbar; }
          } else if (is_quad && !(ga.skip & 1)) {
#pragma unroll 1
            for (int st = 0; st < B; ++st) {
              const double* __restrict__ Gc = (st & 1) ? G1 : G0;
              double* __restrict__ Gn = (st & 1) ? G0 : G1;
              const double2* rt = cs2 + st * B;
              double gq[K::QPT][4];
              double2 r1[K::QPT], r2[K::QPT];
              int o00[K::QPT], o01[K::QPT], o10[K::QPT], o11[K::QPT];
#pragma unroll
              for (int u = 0; u < K::QPT; ++u) {
                // rows {ia, ja} x columns {ib, jb}, ia <= ib < B <= ja, jb; stored positions have row index <= column index.
                // Threads without a quad (qok false) recompute quad 0 and do not store.
                const int ia = qra[u], ja = B + ((qra[u] + st) & (B - 1)), ib = qcb[u], jb = B + ((qcb[u] + st) & (B - 1));
                o00[u] = ia * GLD + ib; o01[u] = ia * GLD + jb; o10[u] = ib * GLD + ja;
                o11[u] = min(ja, jb) * GLD + max(ja, jb);
                r1[u] = rt[rpos<B>(qra[u])]; r2[u] = rt[rpos<B>(qcb[u])];
                gq[u][0] = Gc[o00[u]]; gq[u][1] = Gc[o01[u]]; gq[u][2] = Gc[o10[u]]; gq[u][3] = Gc[o11[u]];
              }
#pragma unroll
              for (int u = 0; u < K::QPT; ++u) {
                const double t00 = r1[u].x * gq[u][0] - r1[u].y * gq[u][2], t01 = r1[u].x * gq[u][1] - r1[u].y * gq[u][3];
                const double t10 = r1[u].y * gq[u][0] + r1[u].x * gq[u][2], t11 = r1[u].y * gq[u][1] + r1[u].x * gq[u][3];
                if (qok[u]) {
                  Gn[o00[u]] = r2[u].x * t00 - r2[u].y * t01;
                  Gn[o01[u]] = r2[u].y * t00 + r2[u].x * t01;
                  if (qra[u] != qcb[u]) Gn[o10[u]] = r2[u].x * t10 - r2[u].y * t11;   // diagonal quad: (ib, ja) is (ia, jb)
                  Gn[o11[u]] = r2[u].y * t10 + r2[u].x * t11;
                }
              }
              cta_bar();
            }
          } else {
#pragma unroll 1
            for (int st = 0; st < B; ++st) cta_bar();
          }
          if (warp == 0 && __any_sync(0xffffffffu, anyrot) && lane == 0) s_any = 1;
          __syncthreads();
          lap(6);
          if (s_any) {
            rotated = true;
            // ---- 6a. the register columns go straight back to global memory; slab columns beyond the GAC register columns
            //          are rotated afterwards from the complete table
            if (is_apply) {
              col_store(0);
              for (int c0 = GAC; c0 < W; c0 += GAC) {   // warp-uniform trip count: the lane pairs exchange rows by shuffle
                col_load(c0);
                apply_cross<B, NC, false>(xp, xm, ah, cs2);
                col_store(c0);
              }
            }
          }
          lap(7);
        } else {
          // ---- 5b. pairs inside the two blocks: B - 1 round-robin steps on G0 in place, rotations accumulated into J
          for (int e = tid; e < GR * GJLD; e += GNT) J[e] = ((e / GJLD) == (e % GJLD)) ? 1.0 : 0.0;
          auto make_rot = [&](int st, int par) {   // warp 0, lanes < B: B / 2 pairs per block
            const int l = lane & (B - 1);
            const int blk = l / H, i = l % H;
            int u0, u1;
            if (i == 0) { u0 = st; u1 = B - 1; }
            else { u0 = (st + i) % (B - 1); u1 = (st - i + (B - 1)) % (B - 1); }
            const int lo = blk * B + (u0 < u1 ? u0 : u1), hi = blk * B + (u0 < u1 ? u1 : u0);
            const double al = G0[lo * GLD + lo], be = G0[hi * GLD + hi], gm = G0[lo * GLD + hi];
            double c = 1.0, s = 0.0;
            const bool rot = lane < B && (al > thr) && (be > thr) && (gm * gm > tol2 * (al * be));
            if (rot) grot(al, be, gm, c, s);
            if (lane < B) { plo[par][lane] = lo; phi[par][lane] = hi; pc[par][lane] = c; ps[par][lane] = s; }
            if (__any_sync(0xffffffffu, rot) && lane == 0) s_any = 1;
          };
          __syncthreads();
          if (warp == 0) make_rot(0, 0);
          __syncthreads();
          for (int st = 0; st < B - 1; ++st) {
            const int par = st & 1;
            for (int qd = tid; qd < B * B; qd += GNT) {
              const int ra = qd / B, cb = qd % B;
              const int ia = plo[par][ra], ja = phi[par][ra], ib = plo[par][cb], jb = phi[par][cb];
              const double ca = pc[par][ra], sa = ps[par][ra], cc = pc[par][cb], sc = ps[par][cb];
              const double g00 = G0[ia * GLD + ib], g01 = G0[ia * GLD + jb], g10 = G0[ja * GLD + ib], g11 = G0[ja * GLD + jb];
              const double t00 = ca * g00 - sa * g10, t01 = ca * g01 - sa * g11;
              const double t10 = sa * g00 + ca * g10, t11 = sa * g01 + ca * g11;
              G0[ia * GLD + ib] = cc * t00 - sc * t01;
              G0[ia * GLD + jb] = sc * t00 + cc * t01;
              G0[ja * GLD + ib] = cc * t10 - sc * t11;
              G0[ja * GLD + jb] = sc * t10 + cc * t11;
            }
            __syncthreads();
            if (warp == 0) {
              if (st + 1 < B - 1) make_rot(st + 1, par ^ 1);
            } else {
              for (int e = tid - 32; e < GR * B; e += GNT - 32) {
                const int m = e / B, cb = e % B;
                const int ib = plo[par][cb], jb = phi[par][cb];
                const double cc = pc[par][cb], sc = ps[par][cb];
                const double x = J[m * GJLD + ib], y = J[m * GJLD + jb];
                J[m * GJLD + ib] = cc * x - sc * y;
                J[m * GJLD + jb] = sc * x + cc * y;
              }
            }
            __syncthreads();
          }
          lap(6);
          if (s_any) {
            rotated = true;
            // ---- 6b. slab <- J^T slab (DMMA), B operand read from global memory, one 8-column strip per warp pass
            for (int strip = warp; strip < W / 8; strip += GNW) {
              const int n0 = strip * 8;
              const int cb = n0 < wR ? cR0 + n0 : cQ0 + (n0 - wR);
              const int lim = n0 < wR ? ndot : ld;
              double acc[K::T8][2] = {};
              for (int k0 = 0; k0 < GR; k0 += 4) {
                const int gk = grow(k0 + t4);
                const double bv = (gk < q && cb + g < lim) ? __ldcg(a.Y + (size_t)gk * ld + cb + g) : 0.0;
#pragma unroll
                for (int mt = 0; mt < K::T8; ++mt) {
                  const double av = J[(k0 + t4) * GJLD + mt * 8 + g];
                  dmma884(acc[mt][0], acc[mt][1], av, bv);
                }
              }
              __syncwarp();
              const int gc = cb + 2 * t4;
              if (gc < lim) {
#pragma unroll
                for (int mt = 0; mt < K::T8; ++mt) {
                  const int gi = grow(mt * 8 + g);
                  if (gi < q) __stcg(reinterpret_cast<double2*>(a.Y + (size_t)gi * ld + gc), make_double2(acc[mt][0], acc[mt][1]));
                }
              }
            }
          }
          lap(7);
        }
        if (nslots > ncl) cluster.sync();   // region A / G0 are reused by this cluster's next slot
      }
      {
        long long tb0 = clock64();
        grid_sync(a.sync, target);
        if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[0] += clock64() - tb0; a.dbg[8] += 1; }
      }
    }
    if (rotated && tid == 0 && crank == 0) atomicAdd(a.sync + 2 + sw, 1u);
    grid_sync(a.sync, target);
    sweeps = sw + 1;
    if (ld_acquire(a.sync + 2 + sw) == 0u) converged = 1;
  }
  if (a.dbg && tid == 0 && cid < nslots) {
    long long busy = 0;
    for (int k = 1; k < 8; ++k) {
      busy += pacc[k];
      atomicMax(reinterpret_cast<unsigned long long*>(a.dbg) + 16 + k, (unsigned long long)pacc[k]);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg) + 24 + k, (unsigned long long)pacc[k]);
    }
    atomicMax(reinterpret_cast<unsigned long long*>(a.dbg) + 16, (unsigned long long)busy);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg) + 24, (unsigned long long)busy);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg) + 15, 1ull);
  }
  const int gw = blockIdx.x * GNW + warp;
  const int GW = gridDim.x * GNW;
  for (int i = gw; i < qall; i += GW) {
    const double* ri = a.Y + (size_t)i * a.ld;
    double acc = 0.0;
    for (int col = 2 * lane; col < a.ndot; col += 64) {
      double2 v = __ldcg(reinterpret_cast<const double2*>(ri + col));
      acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if (!(acc > thr) || i >= q) acc = 0.0;
    if (a.eigen_mode) acc = sqrt(acc);
    if (lane == 0) __stcg(a.norms + i, acc);
  }
  grid_sync(a.sync, target);
  dev::rank_sort_desc(a.norms, qall, a.kmax, a.sig2, a.perm);
  grid_sync(a.sync, target);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err;
    a.info[0] = dev::truncate_spectrum(a.sig2, a.kmax, a.cutoff, a.maxdim, a.mindim, &err);
    a.info[1] = sweeps;
    a.info[2] = converged;
    a.info[3] = qall;
    a.derr[0] = err;
  }
}

// Launch: one cluster per block pair.  Cluster size 8 when the device holds one cluster per pair (15 fit on a B200: measured),
// else 5 (24 or more fit; the Gram entries of the upper-triangular tiles must divide evenly: blocks of 16 rows only).
template <int B>
bool launch_jq_gram_b(Ctx* ctx, JqArgs& a, int r) {
  using K = GG<B>;
  const int nbk = (((r + B - 1) / B) + 1) & ~1;
  const int nslots = nbk / 2;
  const int lq = a.ld - a.ndot;
  static DevOnce once;
  static int maxcl[64][2] = {};
  const int dv = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
  const int css[2] = {8, 5};
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(jacobi_gram_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int k = 0; k < 2; ++k) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(css[k] * 8);
      cfg.blockDim = dim3(GNT);
      cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = css[k]; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, jacobi_gram_kernel<B>, &cfg) != cudaSuccess) n = 0;
      (void)cudaGetLastError();
      maxcl[dv][k] = n;
    }
  });
  static const int cs_env = getenv("QTT_JACOBI_CS") ? atoi(getenv("QTT_JACOBI_CS")) : 0;
  int ki = 0;
  if (K::UTE % 5 == 0 && (K::UTE / 5) % 2 == 0 && (cs_env == 5 || (cs_env == 0 && nslots > maxcl[dv][0] && nslots <= maxcl[dv][1]))) ki = 1;
  const int cs = css[ki];
  if (maxcl[dv][ki] < 1) return false;
  const int wR = (((a.ndot + cs - 1) / cs) + 7) & ~7;
  const int wQ = (((lq + cs - 1) / cs) + 7) & ~7;
  const size_t smem = ((size_t)K::R * (wR + 4) + (size_t)K::R * K::JLD + (size_t)2 * K::R * K::LD) * sizeof(double);
  if (smem > 220 * 1024) return false;
  const int ncl = nslots < maxcl[dv][ki] ? nslots : maxcl[dv][ki];
  GjArgs ga;
  ga.j = a; ga.wR = wR; ga.wQ = wQ;
  ga.skip = getenv("QTT_GRAM_SKIP") ? atoi(getenv("QTT_GRAM_SKIP")) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl * cs);
  cfg.blockDim = dim3(GNT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  if (sole_context(ctx)) {   // nothing else competes for the SMs: all clusters become resident (krylov.cu has the argument)
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
  } else {
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
  }
  cfg.attrs = at; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, jacobi_gram_kernel<B>, ga);
  if (e != cudaSuccess) fail(QTT_ERR_CUDA, "jacobi_gram launch failed (%d clusters x %d): %s", ncl, cs, cudaGetErrorString(e));
  if (a.dbg) fprintf(stderr, "[qtt trace]   gram jacobi B=%d: %d clusters x %d CTAs (max active %d / %d), slab %d + %d columns, %zu B smem\n", B, ncl, cs, maxcl[dv][0], maxcl[dv][1], wR, wQ, smem);
  ctx->launches++;
  return true;
}

// r = the numerical rank the QR found (read back by the caller): blocks of 16 rows while every block pair gets its own
// cluster (measured on B200: 15 co-resident clusters of 8, i.e. r <= 480), blocks of 32 beyond.  Measured, 512 x 512 splits:
// graded spectrum (r = 256) 3.9 ms with blocks of 32 / 2.7 ms with blocks of 16; flat spectrum (r = 512) 9.0 / 10.8 ms.
bool launch_jq_gram(Ctx* ctx, JqArgs& a, int r) {
  static const int bsel = getenv("QTT_JACOBI_B") ? atoi(getenv("QTT_JACOBI_B")) : 0;
  const int slots16 = ((((r + 15) / 16) + 1) & ~1) / 2;
  if (bsel == 16 || (bsel == 0 && slots16 <= 20)) {
    if (launch_jq_gram_b<16>(ctx, a, r)) return true;
  }
  return launch_jq_gram_b<32>(ctx, a, r);
}

// ------------------------------------------------------------------ rank-revealing QR, second generation: block Gram-Schmidt
// The Householder panels above are one CTA working through 32 dependent reflectors (4.4 us each: 140 us per panel,
// 2.8 ms for a full-rank 512 x 512 tensor).  Here the orthonormal basis Q of the row space is built explicitly, 32 rows
// per panel, and everything but a 32 x 32 Cholesky factor runs on the DMMA core:
//   select    the <= 32 unprocessed rows with the largest residual norm above the zero threshold (as before);
//   reorth    P <- P - (P Q^T) Q against ALL earlier basis rows (second Gram-Schmidt pass: two DMMA GEMMs);
//   panel     CholeskyQR of the diagonally scaled panel: Gram matrix (DMMA), Cholesky with deferral (a row whose pivot
//             falls below 1e-6 of its own squared norm is left for a later panel, so the panel stays well conditioned),
//             Q1 = L^-1 D^-1 P (DMMA), then one Newton-Schulz step Q1 <- (1.5 I - 0.5 Q1 Q1^T) Q1 (orthogonality ~1e-9 -> eps);
//   update    coefficients C = Q1 Zw^T of every remaining row (they ARE the rows of R) and Zw <- Zw - C^T Q1 (two DMMA GEMMs),
//             fresh residual norms.
// Y = [R | Q^T] is written directly (no build_y transpose).  The factorisation stops at the numerical rank like the
// Householder variant.  Basis rows beyond the rank (needed only when `mindim` exceeds it) are not produced: the caller
// falls back to the Householder variant, whose accumulated Q supplies the orthonormal completion.
constexpr int GSB = 32;           // panel height
constexpr int GST = 512;          // threads of the panel kernel
constexpr double GS_DEFER = 1e-3; // pivot threshold relative to the row's own squared norm: an accepted row keeps >= 3% of its norm, so normalising it amplifies the rounding of its re-orthogonalisation (eps) by <= 30 (at 1e-6: isometry of the factors 3e-13, measured)

__global__ void gs_init_kernel(const double* __restrict__ M, int q, int p, int ldp, double* __restrict__ Zw, double* __restrict__ norms,
                               int* __restrict__ rowstate, QrCtl* ctl) {
  dev::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int GW = (gridDim.x * blockDim.x) >> 5;
  for (int i = gw; i < q; i += GW) {
    const double* src = M + (size_t)i * p;
    double* dst = Zw + (size_t)i * ldp;
    double acc = 0.0;
    for (int c = lane; c < ldp; c += 32) {
      const double v = c < p ? src[c] : 0.0;
      dst[c] = v;
      acc += v * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) { norms[i] = acc; rowstate[i] = 0; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->r = 0; ctl->done = 0; ctl->nbp = 0; ctl->serial = 0; ctl->F2 = -1.0; ctl->thr = 0.0; ctl->r_next = 0; ctl->nsel = 0; ctl->rj = 0;
  }
}

// One CTA per panel slot: every CTA repeats the (cheap, deterministic) selection, CTA u copies row sel[u] into Pbuf[u].
// (`first`: the zero threshold is derived from the initial norms by every CTA itself; no CTA reads what another one writes here.)
// thr_fixed > 0 (completion phase): that threshold instead of the one derived from |Z|_F.
__global__ void __launch_bounds__(256) gs_select_kernel(const double* __restrict__ Zw, int q, int p, int ldp, int rmax, double zero_tol2,
                                                        int first, double thr_fixed, const double* __restrict__ norms,
                                                        const int* __restrict__ rowstate, double* __restrict__ Pbuf, QrCtl* ctl) {
  dev::pdl_wait();
  extern __shared__ double cand[];   // [q] residual norm of the unprocessed rows, -1 otherwise
  __shared__ int sel[GSB];
  __shared__ double red[8];
  __shared__ double s_thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = blockIdx.x;
  double* prow = Pbuf + (size_t)u * ldp;
  const int r0 = ctl->r_next;
  if (ctl->done) {
    for (int c = tid; c < ldp; c += 256) prow[c] = 0.0;
    return;
  }
  double thr = thr_fixed > 0.0 ? thr_fixed : (first ? 0.0 : ctl->thr);
  if (first && !(thr_fixed > 0.0)) {   // |Z|_F^2 once (the initial norms are the full row norms); same summation order in every CTA
    double acc = 0.0;
    for (int i = tid; i < q; i += 256) acc += norms[i];
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int k = 0; k < 8; ++k) sum += red[k];
      s_thr = zero_tol2 * sum;
      if (u == 0) { ctl->F2 = sum; ctl->thr = s_thr; }
    }
    __syncthreads();
    thr = s_thr;
  }
  int want = rmax - r0;
  if (want > GSB) want = GSB;
  if (tid < GSB) sel[tid] = -1;
  for (int i = tid; i < q; i += 256) cand[i] = (rowstate[i] == 0) ? norms[i] : -1.0;
  __syncthreads();
  for (int i = tid; i < q; i += 256) {
    const double ni = cand[i];
    if (!(ni > thr)) continue;
    int rank = 0;
    for (int j = 0; j < q; ++j) {
      const double nj = cand[j];
      rank += (nj > ni) || (nj == ni && j < i);
    }
    if (rank < want) sel[rank] = i;   // ranks are distinct and dense among the rows above the threshold
  }
  __syncthreads();
  const int mine = sel[u];
  for (int c = tid; c < ldp; c += 256) prow[c] = mine >= 0 ? Zw[(size_t)mine * ldp + c] : 0.0;
  if (u == 0 && tid == 0) {
    int nsel = 0;
    for (int k = 0; k < GSB; ++k) {
      ctl->sel[k] = sel[k];
      nsel += sel[k] >= 0;
    }
    ctl->r = r0;
    ctl->nsel = nsel;
    ctl->nbp = 0;
  }
}

// Selection AND the second Gram-Schmidt pass of the selected rows in one cooperative launch (replaces gs_select + two skinny
// GEMMs, 34 us per panel).  32 CTAs.  Every CTA repeats the selection and stages the <= 32 selected rows in shared memory; then
//   phase 1   CTA b owns the basis rows j = b, b + 32, ...: coefficients C[j][u] = q_j . z_u for all 32 panel rows (global scratch);
//   -- one grid barrier (32 CTAs) --
//   phase 2   CTA b owns the columns [b cw, (b + 1) cw): Pbuf[u][c] = z_u[c] - sum_j C[j][u] q_j[c]  (C staged in shared memory).
// Every coefficient comes from the un-updated row (classical Gram-Schmidt, as the GEMM pair did), sums run over j in order:
// deterministic.  All CTAs always arrive at the barrier counter, so the host can advance its epoch by 32 per launch.
__global__ void __launch_bounds__(256) gs_select_reorth_kernel(const double* __restrict__ Zw, int q, int p, int ldp, int rmax, double zero_tol2,
                                                               int first, double thr_fixed, const double* __restrict__ norms,
                                                               const int* __restrict__ rowstate, double* __restrict__ Pbuf, QrCtl* ctl,
                                                               const double* __restrict__ Yq, int ldy, double* __restrict__ Cbuf,
                                                               unsigned* bar, unsigned epoch, int pld) {
  dev::pdl_wait();   // no-op unless launched with programmatic stream serialisation (sole context, see the launch site)
  extern __shared__ __align__(16) double ssm[];
  double* cand = ssm;                                   // [q rounded up to even]
  double* Ps = ssm + ((q + 1) & ~1);                    // [GSB][pld] the selected rows; after the barrier: C^T [GSB][pld]
  __shared__ int sel[GSB];
  __shared__ double red[8];
  __shared__ double s_thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  auto arrive = [&]() {
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(bar) : "memory");
  };
  const int r0 = ctl->r_next;
  if (ctl->done) {
    for (int c = tid; c < ldp; c += 256) Pbuf[(size_t)b * ldp + c] = 0.0;
    arrive();
    return;
  }
  double thr = thr_fixed > 0.0 ? thr_fixed : (first ? 0.0 : ctl->thr);
  if (first && !(thr_fixed > 0.0)) {   // |Z|_F^2 once (the initial norms are the full row norms); same summation order in every CTA
    double acc = 0.0;
    for (int i = tid; i < q; i += 256) acc += norms[i];
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int k = 0; k < 8; ++k) sum += red[k];
      s_thr = zero_tol2 * sum;
      if (b == 0) { ctl->F2 = sum; ctl->thr = s_thr; }
    }
    __syncthreads();
    thr = s_thr;
  }
  int want = rmax - r0;
  if (want > GSB) want = GSB;
  if (tid < GSB) sel[tid] = -1;
  for (int i = tid; i < q; i += 256) cand[i] = (rowstate[i] == 0) ? norms[i] : -1.0;
  __syncthreads();
  for (int i = tid; i < q; i += 256) {
    const double ni = cand[i];
    if (!(ni > thr)) continue;
    int rank = 0;
    for (int j = 0; j < q; ++j) {
      const double nj = cand[j];
      rank += (nj > ni) || (nj == ni && j < i);
    }
    if (rank < want) sel[rank] = i;   // ranks are distinct and dense among the rows above the threshold
  }
  __syncthreads();
  if (b == 0 && tid == 0) {
    int nsel = 0;
    for (int k = 0; k < GSB; ++k) {
      ctl->sel[k] = sel[k];
      nsel += sel[k] >= 0;
    }
    ctl->r = r0;
    ctl->nsel = nsel;
    ctl->nbp = 0;
  }
  const int cpr = ldp / 2;   // double2 per row (ldp is even)
  if (r0 == 0) {             // first panel: nothing to orthogonalise against
    const int mine = sel[b];
    for (int c = tid; c < cpr; c += 256) {
      double2 v = make_double2(0.0, 0.0);
      if (mine >= 0) v = *reinterpret_cast<const double2*>(Zw + (size_t)mine * ldp + 2 * c);
      *reinterpret_cast<double2*>(Pbuf + (size_t)b * ldp + 2 * c) = v;
    }
    arrive();
    return;
  }
  // ---- stage the selected rows and this CTA's basis rows (slot k = basis row b + 32 k); padded row stride pld (= 4 mod 16:
  //      conflict-free DMMA fragment loads)
  double* Qs = Ps + (size_t)GSB * pld;   // phase 1: [<= 16][pld] basis rows; phase 2: [r0][2 cw2] column slice of all basis rows
  const int njb = r0 > b ? (r0 - b + GSB - 1) / GSB : 0;
  for (int e = tid; e < GSB * cpr; e += 256) {
    const int u = e / cpr, c = (e - u * cpr) * 2;
    const int i = sel[u];
    cp_async16(Ps + (size_t)u * pld + c, i >= 0 ? Zw + (size_t)i * ldp + c : Zw, i >= 0 ? 16 : 0);
  }
  for (int e = tid; e < njb * cpr; e += 256) {
    const int k = e / cpr, c = (e - k * cpr) * 2;
    cp_async16(Qs + (size_t)k * pld + c, Yq + (size_t)(b + GSB * k) * ldy + c, 16);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  const int g = lane >> 2, t4 = lane & 3;
  // ---- phase 1 on the DMMA core: C[slot][u] = sum_c Qs[slot][c] Ps[u][c]; 2 x 4 tiles of 8 x 8, one per warp, K = ldp.
  //      (A scalar version of the two phases was bound by shared-memory bandwidth: 43 us per launch against 13 us.)
  {
    const int mt = warp >> 2, nt = warp & 3;
    if (mt * 8 < njb) {
      const double* pa = Qs + (size_t)(mt * 8 + g) * pld + t4;
      const double* pb = Ps + (size_t)(nt * 8 + g) * pld + t4;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
      const bool arow = mt * 8 + g < njb;   // slots beyond njb hold stale shared memory
      int k0 = 0;
      for (; k0 + 16 <= ldp; k0 += 16) {
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma884(a0[v], a1[v], arow ? pa[k0 + 4 * v] : 0.0, pb[k0 + 4 * v]);
      }
      for (; k0 < ldp; k0 += 4) {
        const bool kin = k0 + t4 < ldp;
        dmma884(a0[0], a1[0], (arow && kin) ? pa[k0] : 0.0, kin ? pb[k0] : 0.0);
      }
      const double v0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), v1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      if (arow) {   // transposed scratch: Cbuf[u][j], row stride pld
        const int jj = b + GSB * (mt * 8 + g), u = nt * 8 + 2 * t4;
        __stcg(Cbuf + (size_t)u * pld + jj, v0);
        __stcg(Cbuf + (size_t)(u + 1) * pld + jj, v1);
      }
    }
  }
  // ---- this CTA's column slice of the selected rows: the C fragments of phase 2 (4 x NT tiles: rows u, columns of the slice)
  const int cw2 = (cpr + GSB - 1) / GSB;   // double2 columns per CTA
  const int cw = 2 * cw2, clo = b * cw;    // columns [clo, clo + cw) of the rows
  const int NT = (cw + 7) / 8;             // <= 2 for ldp <= 512
  const int mt2 = warp & 3, nt2 = warp >> 2;
  const int xu = mt2 * 8 + g, xc = nt2 * 8 + 2 * t4;   // this lane's two entries: row xu, slice columns xc, xc + 1
  const bool xok = nt2 < NT && xc < cw && clo + xc < ldp;
  double x0 = 0.0, x1 = 0.0;
  if (xok) {
    const double2 v = *reinterpret_cast<const double2*>(Ps + (size_t)xu * pld + clo + xc);
    x0 = v.x; x1 = v.y;
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(bar) : "memory");
    while ((int)(ld_acquire(bar) - (epoch + (unsigned)GSB)) < 0) {
    }
  }
  __syncthreads();
  // ---- phase 2: X[u][c] -= sum_j C[j][u] q_j[c]; coefficient table [u][j] (stride pld) and the column slice of every basis row
  //      into shared memory, then one 8 x 8 tile per warp, K = r0
  double* Cs = Ps;
  const int r0c = (r0 + 1) / 2;   // 16-byte chunks per coefficient row
  for (int e = tid; e < GSB * r0c; e += 256) {
    const int u = e / r0c, c = (e - u * r0c) * 2;
    cp_async16(Cs + (size_t)u * pld + c, Cbuf + (size_t)u * pld + c, 16);
  }
  for (int e = tid; e < r0 * cw2; e += 256) {
    const int jj = e / cw2, c2 = e % cw2;
    const bool ok = clo + 2 * c2 < ldp;
    cp_async16(Qs + (size_t)jj * cw + 2 * c2, ok ? Yq + (size_t)jj * ldy + clo + 2 * c2 : Yq, ok ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  if (nt2 < NT) {
    const double* pa = Cs + (size_t)(mt2 * 8 + g) * pld + t4;     // A[g][t4] = C^T[u][j]
    const int bc = nt2 * 8 + g;                                      // B[t4][g] = Qs[j][bc]
    const bool bok = bc < cw;
    const double* pb = Qs + (size_t)t4 * cw + (bok ? bc : 0);
    double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
    int k0 = 0;
    for (; k0 + 16 <= r0; k0 += 16) {
#pragma unroll
      for (int v = 0; v < 4; ++v) dmma884(a0[v], a1[v], pa[k0 + 4 * v], bok ? pb[(size_t)(k0 + 4 * v) * cw] : 0.0);
    }
    for (; k0 < r0; k0 += 4) {
      const bool kin = k0 + t4 < r0;
      dmma884(a0[0], a1[0], kin ? pa[k0] : 0.0, (kin && bok) ? pb[(size_t)k0 * cw] : 0.0);
    }
    if (xok) {
      const double v0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), v1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      *reinterpret_cast<double2*>(Pbuf + (size_t)xu * ldp + clo + xc) = make_double2(x0 - v0, x1 - v1);
    }
  }
}

// CholeskyQR + Newton-Schulz of the (re-orthogonalised) panel Pbuf: accepted rows become new orthonormal basis rows, written
// (compacted, zero rows behind them); the panel rows go back to Zw so that the update GEMMs see them.
__global__ void __launch_bounds__(GST) gs_panel_kernel(const double* __restrict__ Pbuf, int p, int ldp, int pld, int rmax,
                                                       double* __restrict__ Zw, int* __restrict__ rowstate, double* __restrict__ Yq, int ldy,
                                                       double thr_fixed, QrCtl* ctl, long long* dbg) {
  dev::pdl_wait();
  long long tph = dbg ? clock64() : 0;
  auto lap = [&](int k) { if (dbg && threadIdx.x == 0) { const long long n = clock64(); dbg[k] += n - tph; tph = n; } };
  extern __shared__ __align__(16) double psm[];
  double* P = psm;                          // [GSB][pld]
  double* Gs = P + (size_t)GSB * pld;       // [GSB][GSB + 1] scaled Gram matrix, then L
  double* T = Gs + GSB * (GSB + 1);         // [GSB][GSB + 4] compacted L^-1 D^-1, then the Newton-Schulz matrix
  __shared__ double dn[GSB], sc[GSB], invd[GSB];
  __shared__ int acc[GSB], slot[GSB];
  __shared__ int s_na;
  constexpr int TLD = GSB + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const bool done = ctl->done != 0;
  const int nsel = done ? 0 : ctl->nsel;
  if (nsel == 0) {   // nothing above the zero threshold is left (or the basis is complete): the update kernel becomes a no-op
    if (tid == 0) { ctl->done = 1; ctl->nbp = 0; }
    return;
  }
  const int r0 = ctl->r;
  const double thr = thr_fixed > 0.0 ? thr_fixed : ctl->thr;
  // ---- load; the re-orthogonalised rows replace the working rows
  {
    const int cpr = ldp / 2;
    for (int e = tid; e < GSB * cpr; e += GST) {
      const int u = e / cpr, c = (e - u * cpr) * 2;
      cp_async16(P + (size_t)u * pld + c, Pbuf + (size_t)u * ldp + c, 16);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    for (int e = tid; e < GSB * (pld - ldp); e += GST) P[(size_t)(e / (pld - ldp)) * pld + ldp + e % (pld - ldp)] = 0.0;
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (tid < GSB) slot[tid] = ctl->sel[tid];   // (slot[] is reused: the row indices of the panel)
    __syncthreads();
    for (int e = tid; e < GSB * cpr; e += GST) {   // the working rows take the re-orthogonalised values
      const int u = e / cpr, c = (e - u * cpr) * 2;
      const int i = slot[u];
      if (i >= 0) *reinterpret_cast<double2*>(Zw + (size_t)i * ldp + c) = *reinterpret_cast<const double2*>(P + (size_t)u * pld + c);
    }
  }
  lap(0);
  // ---- squared norms and scales
  for (int u = warp; u < GSB; u += GST / 32) {
    const double* row = P + (size_t)u * pld;
    double a = 0.0;
    for (int c = lane; c < p; c += 32) a += row[c] * row[c];
    a = warp_sum(a);
    if (lane == 0) {
      dn[u] = a;
      sc[u] = (u < nsel && a > thr) ? rsqrt(a) : 0.0;
    }
  }
  __syncthreads();
  // ---- Gram matrix on the DMMA core: warp w -> 8 x 8 tile (w / 4, w % 4), four interleaved accumulators along K
  auto gram = [&](bool scaled) {
    if (warp >= 10) return;   // the 10 tiles of the upper triangle (the FP64 rate of ONE SM is what this kernel is bound by)
    int ti = 0, rem = warp;
    while (rem >= 4 - ti) { rem -= 4 - ti; ++ti; }
    const int tj = ti + rem;
    double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
    const double* pa = P + (size_t)(ti * 8 + g) * pld + t4;
    const double* pb = P + (size_t)(tj * 8 + g) * pld + t4;
    int k0 = 0;
    for (; k0 + 16 <= pld; k0 += 16) {
#pragma unroll
      for (int v = 0; v < 4; ++v) dmma884(a0[v], a1[v], pa[k0 + 4 * v], pb[k0 + 4 * v]);
    }
    for (; k0 < pld; k0 += 4) dmma884(a0[0], a1[0], pa[k0], pb[k0]);
    const double v0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), v1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
    const int i = ti * 8 + g, j = tj * 8 + 2 * t4;
    const double si = scaled ? sc[i] : 1.0;
    const double w0 = v0 * si * (scaled ? sc[j] : 1.0), w1 = v1 * si * (scaled ? sc[j + 1] : 1.0);
    Gs[i * (GSB + 1) + j] = w0;
    Gs[i * (GSB + 1) + j + 1] = w1;
    Gs[j * (GSB + 1) + i] = w0;
    Gs[(j + 1) * (GSB + 1) + i] = w1;
  };
  lap(1);
  gram(true);
  __syncthreads();
  lap(2);
  // ---- Cholesky with deferral, all 512 threads: thread e owns the entries (i, j) = (e / 32, e % 32) and (i + 16, j) of the
  //      32 x 32 matrix.  Right-looking, ONE barrier per step: the trailing update uses the unscaled column,
  //      G[i][j] -= G[i][t] G[j][t] / piv, and the scaled column L[i][t] = G[i][t] / sqrt(piv) goes to a separate array.
  //      (A single warp working through the 496 dependent dot products took 41 000 cycles per panel: measured.)
  {
    constexpr int LD = GSB + 1;
    double* Lm = T + GSB * TLD;          // [32][33] the Cholesky factor (zero columns for deferred / null rows)
    const int j = tid & 31, i0 = tid >> 5, i1 = i0 + 16;
    for (int t = 0; t < GSB; ++t) {
      const double piv = Gs[t * LD + t];
      const bool ok = sc[t] > 0.0 && piv > GS_DEFER;
      const double inv = ok ? rsqrt(piv) : 0.0, inv2 = inv * inv;
      const double gjt = Gs[j * LD + t];
      const double g0 = Gs[i0 * LD + t], g1 = Gs[i1 * LD + t];
      if (j == t) {
        Lm[i0 * LD + t] = i0 >= t ? g0 * inv : 0.0;
        Lm[i1 * LD + t] = i1 >= t ? g1 * inv : 0.0;
        if (i0 == 0) { invd[t] = inv; acc[t] = ok ? 1 : 0; }
      }
      __syncthreads();   // every thread has read column t before anybody overwrites entries of later columns it may need
      if (j > t) {
        if (i0 > t) Gs[i0 * LD + j] -= g0 * gjt * inv2;
        if (i1 > t) Gs[i1 * LD + j] -= g1 * gjt * inv2;
      }
      __syncthreads();
    }
    // ---- X = L^-1 (forward substitution on all 32 columns at once): thread owns X[i0][j], X[i1][j] and the running sums
    double s0 = 0.0, s1 = 0.0;
    double* X = Gs;                      // Gs is free now
    for (int t = 0; t < GSB; ++t) {
      // row t of X is final once the sums of steps < t are in: x[t][j] = (delta_tj - s[t][j]) invd[t]
      if (i0 == t) X[t * LD + j] = (((t == j && acc[j]) ? 1.0 : 0.0) - s0) * invd[t];
      if (i1 == t) X[t * LD + j] = (((t == j && acc[j]) ? 1.0 : 0.0) - s1) * invd[t];
      __syncthreads();
      const double xt = X[t * LD + j];
      if (i0 > t) s0 += Lm[i0 * LD + t] * xt;
      if (i1 > t) s1 += Lm[i1 * LD + t] * xt;
    }
    __syncthreads();
    if (tid == 0) {
      int na = 0;
      for (int t = 0; t < GSB; ++t) { slot[t] = na; na += acc[t]; }
      s_na = na;
    }
    __syncthreads();
    // T = compacted rows of X, columns scaled: T[slot(t)][c] = X[t][c] sc[c]; rows beyond na are zero
    const double x0v = X[i0 * LD + j], x1v = X[i1 * LD + j];
    __syncthreads();   // X (= Gs) and Lm (behind T) are both dead after this point; T itself is written next
    for (int e = tid; e < GSB * TLD; e += GST) T[e] = 0.0;
    __syncthreads();
    if (acc[i0]) T[slot[i0] * TLD + j] = x0v * sc[j];
    if (acc[i1]) T[slot[i1] * TLD + j] = x1v * sc[j];
  }
  __syncthreads();
  lap(3);
  const int na = s_na;
  // ---- rows <- T rows (32 x 32 times 32 x p on the DMMA core; a warp owns 8-column strips, so P is updated in place)
  auto apply_t = [&]() {
    for (int strip = warp; strip < pld / 8; strip += GST / 32) {
      const int n0 = strip * 8;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k0 = 0; k0 < GSB; k0 += 4) {
        const double bv = P[(size_t)(k0 + t4) * pld + n0 + g];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) dmma884(a0[mt], a1[mt], T[(mt * 8 + g) * TLD + k0 + t4], bv);
      }
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) *reinterpret_cast<double2*>(P + (size_t)(mt * 8 + g) * pld + n0 + 2 * t4) = make_double2(a0[mt], a1[mt]);
    }
  };
  apply_t();
  __syncthreads();
  lap(4);
  // ---- one Newton-Schulz step: Q1 <- (1.5 I - 0.5 Q1 Q1^T) Q1
  gram(false);
  __syncthreads();
  lap(5);
  for (int e = tid; e < GSB * GSB; e += GST) {
    const int i = e / GSB, j = e % GSB;
    T[i * TLD + j] = (i < na && j < na) ? ((i == j ? 1.5 : 0.0) - 0.5 * Gs[i * (GSB + 1) + j]) : 0.0;
  }
  __syncthreads();
  apply_t();
  __syncthreads();
  lap(6);
  // ---- results: the new basis rows go straight into the carried part of Y = [R | Q^T] (rows r0 ... r0 + na - 1)
  for (int e = tid; e < na * (ldp / 2); e += GST) {
    const int u = e / (ldp / 2), c = (e % (ldp / 2)) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (c < p) v = *reinterpret_cast<const double2*>(P + (size_t)u * pld + c);
    *reinterpret_cast<double2*>(Yq + (size_t)(r0 + u) * ldy + c) = v;
  }
  if (tid < GSB && tid < nsel) {
    const int i = ctl->sel[tid];
    if (i >= 0 && (acc[tid] || !(dn[tid] > thr))) rowstate[i] = 1;   // accepted, or numerically in the span already
  }
  lap(7);
  if (tid == 0) {
    if (dbg) dbg[8] += 1;
    ctl->nbp = na;
    ctl->r_next = r0 + na;
    ctl->serial = ctl->serial + 1;
    if (r0 + na >= rmax) ctl->done = 1;
  }
}

// The panel factorisation spread over ONE CLUSTER of 8 CTAs (the single-CTA kernel above is bound by the FP64 rate of one SM:
// the two Gram matrices and the two 32 x 32 x p products are 33 k of its 75 k cycles).  CTA c owns the columns
// [c cwid, (c + 1) cwid) of the panel; partial Gram matrices are all-gathered through distributed shared memory and summed in rank
// order by every CTA (bit-identical copies), the 32 x 32 Cholesky / inverse / Newton-Schulz matrices are computed redundantly.
constexpr int GPC = 8;      // CTAs per cluster
__global__ void __launch_bounds__(GST) gs_panel_cluster_kernel(const double* __restrict__ Pbuf, int p, int ldp, int cwid, int rmax,
                                                               double* __restrict__ Zw, int* __restrict__ rowstate, double* __restrict__ Yq,
                                                               int ldy, double thr_fixed, QrCtl* ctl) {
  dev::pdl_wait();
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  extern __shared__ __align__(16) double psm[];
  const int sld = cwid + 4;                 // row stride of the slice (cwid is a multiple of 16: = 4 mod 16)
  double* P = psm;                          // [GSB][sld] this CTA's column slice of the panel
  double* Gs = P + (size_t)GSB * sld;       // [GSB][GSB + 1] Gram matrix, then L
  double* T = Gs + GSB * (GSB + 1);         // [GSB][GSB + 4] compacted L^-1 D^-1, then the Newton-Schulz matrix; Lm behind it
  double* STG = T + GSB * (GSB + 4) + GSB * (GSB + 1);   // [2][GPC][640] all-gather staging of the upper-triangular Gram tiles
  __shared__ double dn[GSB], sc[GSB], invd[GSB];
  __shared__ int acc[GSB], slot[GSB];
  __shared__ int s_na;
  constexpr int TLD = GSB + 4;
  constexpr int UTE = 10 * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const bool done = ctl->done != 0;
  const int nsel = done ? 0 : ctl->nsel;
  if (nsel == 0) {   // nothing above the zero threshold is left (or the basis is complete): the update kernel becomes a no-op
    if (crank == 0 && tid == 0) { ctl->done = 1; ctl->nbp = 0; }
    return;
  }
  cluster.sync();    // every CTA of the cluster runs before the first distributed-shared-memory store
  const int r0 = ctl->r;
  const double thr = thr_fixed > 0.0 ? thr_fixed : ctl->thr;
  const int c0 = crank * cwid;              // first global column of the slice
  const int wcols = max(0, min(cwid, ldp - c0));   // valid columns (even)
  // ---- load the slice; the re-orthogonalised rows replace the working rows
  {
    const int cpr = cwid / 2;
    for (int e = tid; e < GSB * cpr; e += GST) {
      const int u = e / cpr, c = (e - u * cpr) * 2;
      const bool ok = c < wcols;
      cp_async16(P + (size_t)u * sld + c, ok ? Pbuf + (size_t)u * ldp + c0 + c : Pbuf, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    if (tid < GSB) slot[tid] = ctl->sel[tid];   // (slot[] is reused: the row indices of the panel)
    __syncthreads();
    for (int e = tid; e < GSB * cpr; e += GST) {   // the working rows take the re-orthogonalised values
      const int u = e / cpr, c = (e - u * cpr) * 2;
      const int i = slot[u];
      if (i >= 0 && c < wcols) *reinterpret_cast<double2*>(Zw + (size_t)i * ldp + c0 + c) = *reinterpret_cast<const double2*>(P + (size_t)u * sld + c);
    }
  }
  // ---- Gram matrix of the panel: partial over the slice (DMMA, 10 upper-triangular tiles), all-gather, sum in rank order
  auto gram_allreduce = [&](int buf) {
    if (warp < 10) {
      int ti = 0, rem = warp;
      while (rem >= 4 - ti) { rem -= 4 - ti; ++ti; }
      const int tj = ti + rem;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
      const double* pa = P + (size_t)(ti * 8 + g) * sld + t4;
      const double* pb = P + (size_t)(tj * 8 + g) * sld + t4;
      for (int k0 = 0; k0 < cwid; k0 += 16) {
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma884(a0[v], a1[v], pa[k0 + 4 * v], pb[k0 + 4 * v]);
      }
      const double2 v = make_double2((a0[0] + a0[1]) + (a0[2] + a0[3]), (a1[0] + a1[1]) + (a1[2] + a1[3]));
      double* base = STG + (size_t)buf * GPC * UTE + (size_t)crank * UTE + warp * 64 + g * 8 + 2 * t4;
      for (int k = 0; k < GPC; ++k) *reinterpret_cast<double2*>(cluster.map_shared_rank(base, k)) = v;
    }
    cluster.sync();
    const double* stg = STG + (size_t)buf * GPC * UTE;
    for (int e = tid; e < UTE; e += GST) {
      double sum = 0.0;
      for (int k = 0; k < GPC; ++k) sum += stg[k * UTE + e];
      const int tl = e >> 6, within = e & 63;
      int ti = 0, rem = tl;
      while (rem >= 4 - ti) { rem -= 4 - ti; ++ti; }
      const int tj = ti + rem;
      const int i = ti * 8 + (within >> 3), j = tj * 8 + (within & 7);
      Gs[i * (GSB + 1) + j] = sum;
      Gs[j * (GSB + 1) + i] = sum;
    }
    __syncthreads();
  };
  gram_allreduce(0);
  // ---- squared norms (the diagonal), scales, scaled Gram matrix
  if (tid < GSB) {
    const double a = Gs[tid * (GSB + 1) + tid];
    dn[tid] = a;
    sc[tid] = (tid < nsel && a > thr) ? rsqrt(a) : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < GSB * GSB; e += GST) {
    const int i = e / GSB, j = e % GSB;
    Gs[i * (GSB + 1) + j] *= sc[i] * sc[j];
  }
  __syncthreads();
  // ---- Cholesky with deferral, all 512 threads (as in gs_panel_kernel)
  {
    constexpr int LD = GSB + 1;
    double* Lm = T + GSB * TLD;          // [32][33] the Cholesky factor (zero columns for deferred / null rows)
    const int j = tid & 31, i0 = tid >> 5, i1 = i0 + 16;
    for (int t = 0; t < GSB; ++t) {
      const double piv = Gs[t * LD + t];
      const bool ok = sc[t] > 0.0 && piv > GS_DEFER;
      const double inv = ok ? rsqrt(piv) : 0.0, inv2 = inv * inv;
      const double gjt = Gs[j * LD + t];
      const double g0 = Gs[i0 * LD + t], g1 = Gs[i1 * LD + t];
      if (j == t) {
        Lm[i0 * LD + t] = i0 >= t ? g0 * inv : 0.0;
        Lm[i1 * LD + t] = i1 >= t ? g1 * inv : 0.0;
        if (i0 == 0) { invd[t] = inv; acc[t] = ok ? 1 : 0; }
      }
      // (no barrier here: everything read above lies in column t, which the update below never writes)
      if (j > t) {
        if (i0 > t) Gs[i0 * LD + j] -= g0 * gjt * inv2;
        if (i1 > t) Gs[i1 * LD + j] -= g1 * gjt * inv2;
      }
      __syncthreads();
    }
    double s0 = 0.0, s1 = 0.0;
    double* X = Gs;
    for (int t = 0; t < GSB; ++t) {
      if (i0 == t) X[t * LD + j] = (((t == j && acc[j]) ? 1.0 : 0.0) - s0) * invd[t];
      if (i1 == t) X[t * LD + j] = (((t == j && acc[j]) ? 1.0 : 0.0) - s1) * invd[t];
      __syncthreads();
      const double xt = X[t * LD + j];
      if (i0 > t) s0 += Lm[i0 * LD + t] * xt;
      if (i1 > t) s1 += Lm[i1 * LD + t] * xt;
    }
    __syncthreads();
    if (tid == 0) {
      int na = 0;
      for (int t = 0; t < GSB; ++t) { slot[t] = na; na += acc[t]; }
      s_na = na;
    }
    __syncthreads();
    const double x0v = X[i0 * LD + j], x1v = X[i1 * LD + j];
    __syncthreads();
    for (int e = tid; e < GSB * TLD; e += GST) T[e] = 0.0;
    __syncthreads();
    if (acc[i0]) T[slot[i0] * TLD + j] = x0v * sc[j];
    if (acc[i1]) T[slot[i1] * TLD + j] = x1v * sc[j];
  }
  __syncthreads();
  const int na = s_na;
  // ---- slice <- T slice (32 x 32 times 32 x cwid on the DMMA core; a warp owns 8-column strips, so P is updated in place)
  auto apply_t = [&]() {
    for (int strip = warp; strip < cwid / 8; strip += GST / 32) {
      const int n0 = strip * 8;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k0 = 0; k0 < GSB; k0 += 4) {
        const double bv = P[(size_t)(k0 + t4) * sld + n0 + g];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) dmma884(a0[mt], a1[mt], T[(mt * 8 + g) * TLD + k0 + t4], bv);
      }
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) *reinterpret_cast<double2*>(P + (size_t)(mt * 8 + g) * sld + n0 + 2 * t4) = make_double2(a0[mt], a1[mt]);
    }
  };
  apply_t();
  __syncthreads();
  // ---- one Newton-Schulz step: Q1 <- (1.5 I - 0.5 Q1 Q1^T) Q1
  gram_allreduce(1);
  for (int e = tid; e < GSB * GSB; e += GST) {
    const int i = e / GSB, j = e % GSB;
    T[i * TLD + j] = (i < na && j < na) ? ((i == j ? 1.5 : 0.0) - 0.5 * Gs[i * (GSB + 1) + j]) : 0.0;
  }
  __syncthreads();
  apply_t();
  __syncthreads();
  // ---- results: the new basis rows go straight into the carried part of Y = [R | Q^T] (rows r0 ... r0 + na - 1)
  {
    const int cpr = cwid / 2;
    for (int e = tid; e < na * cpr; e += GST) {
      const int u = e / cpr, c = (e - u * cpr) * 2;
      if (c >= wcols) continue;
      double2 v = make_double2(0.0, 0.0);
      if (c0 + c < p) v = *reinterpret_cast<const double2*>(P + (size_t)u * sld + c);
      *reinterpret_cast<double2*>(Yq + (size_t)(r0 + u) * ldy + c0 + c) = v;
    }
  }
  if (crank == 0) {
    if (tid < GSB && tid < nsel) {
      const int i = ctl->sel[tid];
      if (i >= 0 && (acc[tid] || !(dn[tid] > thr))) rowstate[i] = 1;   // accepted, or numerically in the span already
    }
    if (tid == 0) {
      ctl->nbp = na;
      ctl->r_next = r0 + na;
      ctl->serial = ctl->serial + 1;
      if (r0 + na >= rmax) ctl->done = 1;
    }
  }
}

// Right-looking update of the working rows against the panel's basis rows Q1 (<= 32, in shared memory):
//   c_t = z_i . q_t (all t from the same z_i: the rows of Q1 are orthonormal to rounding), z_i <- z_i - sum_t c_t q_t,
// the new residual norm, and the coefficients straight into Y[r0 + t][i] (rows of R).  A warp owns TWO rows at a time so that every
// shared-memory load of Q1 feeds two FMAs.  Replaces two latency-bound skinny GEMMs (K = 512 on 16 tiles: 14 us each).
template <int NPL>
__global__ void __launch_bounds__(256) gs_update_kernel(double* __restrict__ Zw, int q, int p, int ldp, const int* __restrict__ rowstate,
                                                        double* __restrict__ norms, double* __restrict__ Y, int qd, int ldy, int write_r,
                                                        const QrCtl* ctl) {
  dev::pdl_wait();
  extern __shared__ __align__(16) double qsm[];   // [na rounded up to 4][ldp]
  const int na = ctl->nbp;
  if (na == 0) return;
  const int r0 = ctl->r;
  const int na4 = (na + 3) & ~3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {   // stage the basis rows: all copies in flight at once (a load loop through registers was latency-bound: 16 us of 50)
    const int cpr = ldp / 2;
    for (int e = tid; e < na4 * cpr; e += 256) {
      const int t = e / cpr, c = (e - t * cpr) * 2;
      const bool ok = t < na;
      cp_async16(qsm + (size_t)t * ldp + c, ok ? Y + (size_t)(r0 + t) * ldy + qd + c : Y, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
  }
  const int npairs = (q + 1) / 2;
  for (int pr = blockIdx.x * 8 + warp; pr < npairs; pr += gridDim.x * 8) {
    const int i0 = 2 * pr, i1 = 2 * pr + 1;
    const bool ok1 = i1 < q;
    double x0[NPL], x1[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int c = lane + 32 * k;
      x0[k] = c < p ? Zw[(size_t)i0 * ldp + c] : 0.0;
      x1[k] = (ok1 && c < p) ? Zw[(size_t)i1 * ldp + c] : 0.0;
    }
    double c0 = 0.0, c1 = 0.0;   // lane t keeps the coefficients of basis row t
    for (int t = 0; t < na4; t += 4) {   // four basis rows at a time: eight interleaved warp reductions
      double d0[4] = {0.0, 0.0, 0.0, 0.0}, d1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int c = lane + 32 * k;
        if (c < ldp) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const double w = qsm[(size_t)(t + v) * ldp + c];
            d0[v] += x0[k] * w;
            d1[v] += x1[k] * w;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          d0[v] += __shfl_xor_sync(0xffffffffu, d0[v], o);
          d1[v] += __shfl_xor_sync(0xffffffffu, d1[v], o);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (lane == t + v) { c0 = d0[v]; c1 = d1[v]; }
    }
    for (int t = 0; t < na4; t += 4) {
      double e0[4], e1[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        e0[v] = __shfl_sync(0xffffffffu, c0, t + v);
        e1[v] = __shfl_sync(0xffffffffu, c1, t + v);
      }
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const int c = lane + 32 * k;
        if (c < ldp) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const double w = qsm[(size_t)(t + v) * ldp + c];
            x0[k] -= e0[v] * w;
            x1[k] -= e1[v] * w;
          }
        }
      }
    }
    double n0 = 0.0, n1 = 0.0;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int c = lane + 32 * k;
      if (c < p) {
        Zw[(size_t)i0 * ldp + c] = x0[k];
        if (ok1) Zw[(size_t)i1 * ldp + c] = x1[k];
      }
      n0 += x0[k] * x0[k];
      n1 += x1[k] * x1[k];
    }
    n0 = warp_sum(n0);
    n1 = warp_sum(n1);
    if (write_r && lane < na) {   // rows of R: Y[r0 + t][i] = z_i . q_t
      Y[(size_t)(r0 + lane) * ldy + i0] = c0;
      if (ok1) Y[(size_t)(r0 + lane) * ldy + i1] = c1;
    }
    if (lane == 0) {
      if (rowstate[i0] == 0) norms[i0] = n0;
      if (ok1 && rowstate[i1] == 0) norms[i1] = n1;
    }
  }
}

// The same update on the DMMA core, 8 working rows per CTA (64 CTAs at q = 512 instead of 32, and no shared-memory bandwidth
// wall: the scalar version re-reads the 32 basis rows for every pair of working rows).  C = Z Q1^T (8 x 32, K = ldp: four
// 8 x 8 tiles, the K range halved between two warps each), then Z -= C Q1 (64 column tiles, K = 32), norms from the fragments.
// Fixed summation orders throughout.  Shared memory: Q1 [32][pld], Z [8][pld], C halves [2][8][36].
constexpr int GUR = 8;   // working rows per CTA
__global__ void __launch_bounds__(256) gs_update_mma_kernel(double* __restrict__ Zw, int q, int p, int ldp, int pld,
                                                            const int* __restrict__ rowstate, double* __restrict__ norms,
                                                            double* __restrict__ Y, int qd, int ldy, int write_r, const QrCtl* ctl) {
  dev::pdl_wait();
  extern __shared__ __align__(16) double usm[];
  const int na = ctl->nbp;
  if (na == 0) return;
  const int r0 = ctl->r;
  double* Qs = usm;                          // [GSB][pld]
  double* Zs = Qs + (size_t)GSB * pld;       // [GUR][pld]
  double* Ch = Zs + (size_t)GUR * pld;       // [2][GUR][36] partial coefficients of the two K halves; [0] becomes the sum
  __shared__ double nrm_s[8][GUR];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i0 = blockIdx.x * GUR;
  const int cpr = ldp / 2;
  for (int e = tid; e < GSB * cpr; e += 256) {
    const int t = e / cpr, c = (e - t * cpr) * 2;
    const bool ok = t < na;
    cp_async16(Qs + (size_t)t * pld + c, ok ? Y + (size_t)(r0 + t) * ldy + qd + c : Y, ok ? 16 : 0);
  }
  for (int e = tid; e < GUR * cpr; e += 256) {
    const int il = e / cpr, c = (e - il * cpr) * 2;
    const bool ok = i0 + il < q;
    cp_async16(Zs + (size_t)il * pld + c, ok ? Zw + (size_t)(i0 + il) * ldp + c : Zw, ok ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  // ---- C = Z Q1^T: warp -> (n-tile nt = warp & 3, K half kh = warp >> 2)
  {
    const int nt = warp & 3, kh = warp >> 2;
    const int kspan = (((ldp + 1) / 2) + 3) & ~3;   // per half, multiple of 4
    const int kb = kh * kspan, ke = min(ldp, kb + kspan);
    const double* pa = Zs + (size_t)g * pld + t4;
    const double* pb = Qs + (size_t)(nt * 8 + g) * pld + t4;
    double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
    int k0 = kb;
    for (; k0 + 16 <= ke; k0 += 16) {
#pragma unroll
      for (int v = 0; v < 4; ++v) dmma884(a0[v], a1[v], pa[k0 + 4 * v], pb[k0 + 4 * v]);
    }
    for (; k0 < ke; k0 += 4) {
      const bool kin = k0 + t4 < ke;
      dmma884(a0[0], a1[0], kin ? pa[k0] : 0.0, kin ? pb[k0] : 0.0);
    }
    double* dst = Ch + (size_t)kh * GUR * 36 + g * 36 + nt * 8 + 2 * t4;
    dst[0] = (a0[0] + a0[1]) + (a0[2] + a0[3]);
    dst[1] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
  }
  __syncthreads();
  {
    const int il = tid >> 5, t = tid & 31;   // 8 x 32 coefficients
    const double c = Ch[il * 36 + t] + Ch[GUR * 36 + il * 36 + t];
    __syncthreads();
    Ch[il * 36 + t] = c;
    if (write_r && t < na && i0 + il < q) Y[(size_t)(r0 + t) * ldy + i0 + il] = c;   // rows of R: Y[r0 + t][i] = z_i . q_t
  }
  __syncthreads();
  // ---- Z -= C Q1: warp takes the column tiles warp, warp + 8, ...; K = 32
  double nacc = 0.0;   // this lane's share of |z_g|^2
  double ca[GSB / 4];
#pragma unroll
  for (int ks = 0; ks < GSB / 4; ++ks) ca[ks] = Ch[g * 36 + ks * 4 + t4];   // A fragments: C[g][k]
  for (int tile = warp; tile * 8 < ldp; tile += 8) {
    const int n0 = tile * 8;
    double d0[2] = {0.0, 0.0}, d1[2] = {0.0, 0.0};
#pragma unroll
    for (int ks = 0; ks < GSB / 4; ++ks) {
      const double bv = (n0 + g < ldp) ? Qs[(size_t)(ks * 4 + t4) * pld + n0 + g] : 0.0;
      dmma884(d0[ks & 1], d1[ks & 1], ca[ks], bv);
    }
    const int c = n0 + 2 * t4;
    if (c < ldp) {
      const double2 z = *reinterpret_cast<const double2*>(Zs + (size_t)g * pld + c);
      const double v0 = z.x - (d0[0] + d0[1]), v1 = z.y - (d1[0] + d1[1]);
      if (i0 + g < q) *reinterpret_cast<double2*>(Zw + (size_t)(i0 + g) * ldp + c) = make_double2(v0, v1);
      nacc += v0 * v0 + v1 * v1;
    }
  }
  nacc += __shfl_xor_sync(0xffffffffu, nacc, 1);
  nacc += __shfl_xor_sync(0xffffffffu, nacc, 2);
  if (t4 == 0) nrm_s[warp][g] = nacc;
  __syncthreads();
  if (tid < GUR && i0 + tid < q) {
    double n = 0.0;
    for (int w = 0; w < 8; ++w) n += nrm_s[w][tid];
    if (rowstate[i0 + tid] == 0) norms[i0 + tid] = n;
  }
}

__global__ void gs_commit_kernel(QrCtl* ctl, int reopen) {
  dev::pdl_wait();
  if (threadIdx.x == 0) {
    ctl->r = ctl->r_next;
    if (reopen) { ctl->done = 0; ctl->nsel = 0; ctl->nbp = 0; ctl->rj = ctl->r_next; }
  }
}

// Completion phase (`mindim` asks for more basis rows than the numerical rank): the working matrix is the projector
// I - Q^T Q, whose rows are the coordinate vectors with the span of Q removed.  Zw2 arrives holding -Q^T Q (DMMA GEMM).
__global__ void gs_projector_kernel(double* __restrict__ Zw2, int p, int ldp, double* __restrict__ norms2, int* __restrict__ rowstate2) {
  dev::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int GW = (gridDim.x * blockDim.x) >> 5;
  for (int i = gw; i < p; i += GW) {
    double* row = Zw2 + (size_t)i * ldp;
    double a = 0.0;
    for (int c = lane; c < ldp; c += 32) {
      double v = c < p ? row[c] : 0.0;
      if (c == i) v += 1.0;
      row[c] = v;
      a += v * v;
    }
    a = warp_sum(a);
    if (lane == 0) { norms2[i] = a; rowstate2[i] = 0; }
  }
}

// Wn[i][:] = carried (Q^T) part of row perm[i].  Positions beyond the numerical rank r (sigma = 0, kept only because of
// `mindim`) receive the orthonormal completion LAPACK's SVD would return there: column i of the accumulated Q (the identity
// block of SE), which is orthogonal to the first r columns.  A zero row instead would make the next bond's projected
// operator exactly singular.
__global__ void gather_rows_kernel(const double* __restrict__ Y, int ld, int off, int p, const int* __restrict__ perm,
                                   const double* __restrict__ sig2, int k, double* __restrict__ Wn, double* __restrict__ sigma_out,
                                   const double* __restrict__ SE, int q, int ldp, int eigen_mode) {
  dev::pdl_wait();
  const int i = blockIdx.y;
  if (i >= k) return;
  const int src = perm[i];
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < p; col += gridDim.x * blockDim.x)
    Wn[(size_t)i * p + col] = src >= 0 ? Y[(size_t)src * ld + off + col] : SE[(size_t)(q + col) * ldp + i];
  if (sigma_out && blockIdx.x == 0 && threadIdx.x == 0) sigma_out[i] = eigen_mode ? sig2[i] : sqrt(sig2[i]);
}

template <int NCH>
void launch_jq(Ctx* ctx, JqArgs& a, int grid) {
  void* args[] = {&a};
  QTT_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_qr_kernel<NCH>, dim3(grid), dim3(64), args, 0, ctx->stream));
  ctx->launches++;
}

template <typename K>
void set_smem(K kern, size_t bytes) {
  QTT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

}  // namespace

static int panel_width(int p) { return p <= 768 ? NB : NB / 2; }

bool svd_rows_qr_supported(int q, int p) {
  // the panel kernel keeps nb rows of length p (+ q candidate norms) in shared memory
  const int ldp = (p + 1) & ~1;
  const size_t smem = ((size_t)panel_width(p) * ldp + q) * sizeof(double);
  return (p % 2 == 0) && q >= 1 && smem <= 200 * 1024;
}

RowSvd svd_rows_qr(Ctx* ctx, const double* M, int q, int p, double cutoff, int maxdim, int mindim, double* sigma_out, int max_sweeps,
                   bool eigen_mode) {
  QTT_REQUIRE(svd_rows_qr_supported(q, p), "svd_rows_qr: unsupported shape %d x %d", q, p);
  QTT_REQUIRE(!eigen_mode || q == p, "svd_rows_qr: eigen mode needs a square matrix");
  const int ldp = (p + 1) & ~1;
  const int kmax = q < p ? q : p;
  const int qd = (q + 1) & ~1;
  const int ldy = qd + ldp;
  if (max_sweeps <= 0) max_sweeps = 40;
  QTT_REQUIRE(max_sweeps <= 200, "svd_max_sweeps too large");
  double* SE = ctx->arena.alloc((size_t)(q + p) * ldp);
  double* norms = ctx->arena.alloc(q + kmax);
  int* rowstate = reinterpret_cast<int*>(ctx->arena.alloc(q / 2 + 1));
  int* inpanel = reinterpret_cast<int*>(ctx->arena.alloc(q / 2 + 1));
  double* Vp = ctx->arena.alloc((size_t)NB * ldp);
  double* taus = ctx->arena.alloc(NB);
  QrCtl* ctl = reinterpret_cast<QrCtl*>(ctx->arena.alloc(64));
  double* Y = ctx->arena.alloc((size_t)(kmax + GSB) * ldy);
  double* sig2 = ctx->arena.alloc(kmax);
  int* perm = reinterpret_cast<int*>(ctx->arena.alloc(kmax / 2 + 1));
  int* info = reinterpret_cast<int*>(ctx->arena.alloc(4));
  double* derr = ctx->arena.alloc(2);

  const double tol = sqrt((double)p) * 1.1102230246251565e-16;
  // rows below zt * |Z|_F are rounding noise: never rotated, sigma reported as 0.  In eigen mode the truncated spectrum is
  // |row| itself and the reference's cutoffs go down to 1e-15: the floor is kept as low as the rotation noise allows
  // (same rule as svd_rows_ex).
  const double zt = (eigen_mode ? 2.0 : 8.0) * tol;
  const bool trace = ctx->profiling && getenv("QTT_TRACE") != nullptr;
  cudaEvent_t evq0 = nullptr;
  if (trace) { cudaEventCreate(&evq0); cudaEventRecord(evq0, ctx->stream); }
  // ---- block Gram-Schmidt QR (DMMA); falls through to the Householder panels when it does not apply
  static const bool gs_off = getenv("QTT_QR") != nullptr && std::string(getenv("QTT_QR")) == "householder";
  const int pld = ldp + ((4 - ldp % 16) + 16) % 16;
  const size_t gs_smem = ((size_t)GSB * pld + 2 * GSB * (GSB + 1) + GSB * (GSB + 4)) * sizeof(double);
  int need = mindim < kmax ? mindim : kmax;   // basis rows the truncation rule may keep whatever the spectrum
  bool gs_done = false;
  int known_r = -1;   // rank known on the host without a read-back (Gram-Schmidt path, no completion phase)
  if (!gs_off && p % 2 == 0 && kmax >= 64 && gs_smem <= 220 * 1024) {
    static DevOnce gs_once;
    gs_once.run(ctx->device, [&] {
      set_smem(gs_panel_kernel, 220 * 1024);
      set_smem(gs_update_kernel<8>, 200 * 1024);
      set_smem(gs_update_kernel<16>, 200 * 1024);
      set_smem(gs_update_kernel<24>, 200 * 1024);
      set_smem(gs_select_reorth_kernel, 216 * 1024);
      set_smem(gs_update_mma_kernel, 216 * 1024);
      set_smem(gs_panel_cluster_kernel, 200 * 1024);
    });
    double* Zw = SE;   // the (q + p) x ldp Householder workspace is free on this path
    double* Pbuf = ctx->arena.alloc((size_t)GSB * ldp);
    double* C2 = ctx->arena.alloc((size_t)(kmax + GSB) * GSB);
    double* Ct = ctx->arena.alloc((size_t)GSB * pld);   // coefficients of the fused re-orthogonalisation, [u][j]
    QTT_CUDA(cudaMemsetAsync(Y, 0, sizeof(double) * (size_t)(kmax + GSB) * ldy, ctx->stream));
    {
      int grid = (q + 7) / 8;
      if (grid > 4 * ctx->sms) grid = 4 * ctx->sms;
      launch_pdl(ctx, "gs_init", gs_init_kernel, dim3(grid), dim3(256), 0, M, q, p, ldp, Zw, norms, rowstate, ctl);
    }
    long long* gs_dbg = nullptr;
    if (trace) {
      gs_dbg = reinterpret_cast<long long*>(ctx->arena.alloc(16));
      QTT_CUDA(cudaMemsetAsync(gs_dbg, 0, 128, ctx->stream));
    }
    int* h_done = reinterpret_cast<int*>(ctx->h_scal + 17);
    // panels over the working rows W (nrows x ldp) until `rmax` basis rows exist or nothing above the threshold is left
    auto run_panels = [&](double* W, int nrows, double* nrm, int* rst, int rmax, int r_have, bool main_phase, double thr_fixed) -> bool {
      const int nominal = (rmax - r_have + GSB - 1) / GSB;
      const int hint_panels = (main_phase && ctx->qr_last_need == need && ctx->qr_last_rank > 0) ? (ctx->qr_last_rank + 21) / 22 + 1 : 0;
      // fused selection + second Gram-Schmidt pass (one cooperative launch of 32 CTAs) when its staging fits in shared memory
      const int cw2 = (ldp / 2 + GSB - 1) / GSB;
      const size_t qs_d = std::max((size_t)((kmax + GSB - 1) / GSB) * pld, (size_t)kmax * 2 * cw2);
      const size_t fsm = sizeof(double) * ((size_t)((nrows + 1) & ~1) + (size_t)GSB * pld + qs_d);
      static const bool fused_off = getenv("QTT_GS_FUSED") != nullptr && atoi(getenv("QTT_GS_FUSED")) == 0;
      const bool fused = !fused_off && fsm <= 216 * 1024 && kmax <= ldp && ldp <= 512;
      for (int pn = 0; pn < nominal + 16; ++pn) {
        if (fused) {
          const double* Wc = W; const double* nc = nrm; const int* rc = rst; const double* yq = Y + qd;
          double zt2 = zt * zt;
          int first = (main_phase && pn == 0) ? 1 : 0;
          unsigned* bar = ctx->d_sync + 232;
          unsigned epoch = ctx->gs_epoch;
          int nr = nrows, pp = p, lp = ldp, rm = rmax, ly = ldy;
          double tf = thr_fixed;
          int pl = pld;
          if (sole_context(ctx)) {   // 32 CTAs: resident together whenever nothing else competes for the SMs (krylov.cu has the argument)
            launch_pdl(ctx, "gs_select_reorth", gs_select_reorth_kernel, dim3(GSB), dim3(256), fsm, Wc, nr, pp, lp, rm, zt2, first, tf, nc, rc, Pbuf, ctl,
                       yq, ly, Ct, bar, epoch, pl);
          } else {
            void* args[] = {&Wc, &nr, &pp, &lp, &rm, &zt2, &first, &tf, &nc, &rc, &Pbuf, &ctl, &yq, &ly, &Ct, &bar, &epoch, &pl};
            QTT_CUDA(cudaLaunchCooperativeKernel((void*)gs_select_reorth_kernel, dim3(GSB), dim3(256), args, fsm, ctx->stream));
            ctx->launches++;
          }
          ctx->gs_epoch += (unsigned)GSB;
        } else {
        launch_pdl(ctx, "gs_select", gs_select_kernel, dim3(GSB), dim3(256), sizeof(double) * nrows, (const double*)W, nrows, p, ldp, rmax, zt * zt,
                   (main_phase && pn == 0) ? 1 : 0, thr_fixed, (const double*)nrm, (const int*)rst, Pbuf, ctl);
        }
        if (!fused && (pn > 0 || r_have > 0)) {
          int J = r_have + pn * GSB;   // upper bound of the basis rows so far (rows beyond r are zero)
          if (J > kmax) J = kmax;
          GemmDesc g1;   // C2 (J x 32) = Q (J x p) . Pbuf^T
          g1.A = Y + qd; g1.B = Pbuf; g1.C = C2; g1.transB = true;
          g1.M = J; g1.N = GSB; g1.K = p; g1.lda = ldy; g1.ldb = ldp; g1.ldc = GSB;
          gemm(ctx, g1);
          GemmDesc g2;   // Pbuf (32 x p) -= C2^T . Q
          g2.A = C2; g2.B = Y + qd; g2.C = Pbuf; g2.transA = true;
          g2.M = GSB; g2.N = p; g2.K = J; g2.lda = GSB; g2.ldb = ldy; g2.ldc = ldp; g2.alpha = -1.0; g2.beta = 1.0;
          gemm(ctx, g2);
        }
        static const bool pc_off = getenv("QTT_GS_PANEL") != nullptr && std::string(getenv("QTT_GS_PANEL")) == "single";
        if (!pc_off && !gs_dbg) {
          const int cwid = (((ldp + GPC - 1) / GPC) + 15) & ~15;
          const size_t csm = sizeof(double) * ((size_t)GSB * (cwid + 4) + 2 * GSB * (GSB + 1) + GSB * (GSB + 4) + (size_t)2 * GPC * 640);
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(GPC);
          cfg.blockDim = dim3(GST);
          cfg.dynamicSmemBytes = csm;
          cfg.stream = ctx->stream;
          cudaLaunchAttribute at[2];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = GPC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[1].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = 2;
          cudaError_t e = cudaLaunchKernelEx(&cfg, gs_panel_cluster_kernel, (const double*)Pbuf, p, ldp, cwid, rmax, W, rst, Y + qd, ldy, thr_fixed, ctl);
          if (e != cudaSuccess) fail(QTT_ERR_CUDA, "gs_panel_cluster launch failed: %s", cudaGetErrorString(e));
          ctx->launches++;
        } else
        launch_pdl(ctx, "gs_panel", gs_panel_kernel, dim3(1), dim3(GST), gs_smem, (const double*)Pbuf, p, ldp, pld, rmax, W, rst, Y + qd, ldy, thr_fixed, ctl,
                   gs_dbg);
        {   // coefficients of every row on the new basis rows (= rows of R), W <- W - C^T Q1, fresh residual norms
          int ug = ((nrows + 1) / 2 + 7) / 8;
          if (ug > ctx->sms) ug = ctx->sms;
          const size_t usm = sizeof(double) * (size_t)GSB * ldp;
          const int wr = main_phase ? 1 : 0;
          const size_t msm = sizeof(double) * ((size_t)(GSB + GUR) * pld + 2 * GUR * 36);
          static const bool mma_off = getenv("QTT_GS_UPDATE") != nullptr && std::string(getenv("QTT_GS_UPDATE")) == "scalar";
          if (!mma_off && msm <= 216 * 1024)
            launch_pdl(ctx, "gs_update", gs_update_mma_kernel, dim3((nrows + GUR - 1) / GUR), dim3(256), msm, W, nrows, p, ldp, pld, (const int*)rst, nrm, Y, qd,
                       ldy, wr, (const QrCtl*)ctl);
          else if (ldp <= 256) launch_pdl(ctx, "gs_update", gs_update_kernel<8>, dim3(ug), dim3(256), usm, W, nrows, p, ldp, (const int*)rst, nrm, Y, qd, ldy, wr, (const QrCtl*)ctl);
          else if (ldp <= 512) launch_pdl(ctx, "gs_update", gs_update_kernel<16>, dim3(ug), dim3(256), usm, W, nrows, p, ldp, (const int*)rst, nrm, Y, qd, ldy, wr, (const QrCtl*)ctl);
          else launch_pdl(ctx, "gs_update", gs_update_kernel<24>, dim3(ug), dim3(256), usm, W, nrows, p, ldp, (const int*)rst, nrm, Y, qd, ldy, wr, (const QrCtl*)ctl);
        }
        const int np1 = pn + 1;
        // has the factorisation finished?  Every poll drains the stream, so the first one waits for the number of panels the
        // previous split of this shape needed (ranks vary smoothly along a chain: ~22 rows are accepted per panel), then every
        // second panel; without a hint: after 1, 2, 4, 8, ... panels.  Panels launched after `done` are no-ops (~3 us each).
        const bool poll = hint_panels > 0 ? (np1 >= hint_panels && ((np1 - hint_panels) % 2 == 0)) : ((np1 & (np1 - 1)) == 0);
        if (poll || np1 >= nominal) {
          QTT_CUDA(cudaMemcpyAsync(h_done, &ctl->done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
          QTT_CUDA(cudaMemcpyAsync(h_done + 1, &ctl->r_next, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));   // the rank, once done
          QTT_CUDA(cudaStreamSynchronize(ctx->stream));
          if (*h_done) return true;
        }
      }
      return false;
    };
    // Rank-deficient tensors under `mindim` (the first sweeps of the benchmark: rank 2 ... 200, mindim 256): the Householder
    // variant gets the orthonormal completion for free from its accumulated Q, so it is the better choice while the rank is
    // small; the last split's rank (ranks vary smoothly along a chain) decides which variant goes first.
    const bool hh_first = ctx->qr_last_need == need && ctx->qr_last_rank >= 0 && ctx->qr_last_rank < (int)(0.45 * need);
    if (!hh_first) gs_done = run_panels(Zw, q, norms, rowstate, kmax, 0, true, 0.0);
    if (gs_dbg) {
      long long h[16];
      cudaMemcpy(h, gs_dbg, 128, cudaMemcpyDeviceToHost);
      if (h[8] > 0)
        fprintf(stderr, "[qtt trace]   gs_panel cycles per panel: load %.0f norms %.0f gram %.0f cholesky+inverse %.0f apply %.0f gram2 %.0f apply2 %.0f store %.0f (%lld panels)\n",
                (double)h[0] / h[8], (double)h[1] / h[8], (double)h[2] / h[8], (double)h[3] / h[8], (double)h[4] / h[8], (double)h[5] / h[8],
                (double)h[6] / h[8], (double)h[7] / h[8], h[8]);
    }
    if (gs_done) {
      launch_pdl(ctx, "gs_commit", gs_commit_kernel, dim3(1), dim3(32), 0, ctl, 0);
      int* h_r = reinterpret_cast<int*>(ctx->h_scal + 16);
      const int r_num = h_done[1];   // ctl->r_next as read by the poll that saw `done` (= ctl->r after the commit): no extra round trip
      if (r_num >= need) { known_r = r_num; }   // no completion phase: the Jacobi launch below needs no read-back either
      ctx->qr_last_rank = r_num;
      ctx->qr_last_need = need;
      if (r_num < need) {
        if (r_num < (int)(0.45 * need)) {
          gs_done = false;   // far below `mindim`: Householder path (its accumulated Q is the orthonormal completion)
        } else {
          // completion: Gram-Schmidt continues on the rows of the projector I - Q^T Q until `need` basis rows exist;
          // their R part stays zero (sigma = 0: never rotated, sorted last, kept only because of `mindim`)
          double* Zw2 = SE + (size_t)q * ldp;
          double* norms2 = ctx->arena.alloc(p);
          int* rowstate2 = reinterpret_cast<int*>(ctx->arena.alloc(p / 2 + 1));
          GemmDesc gp;   // Zw2 (p x p) = -Q^T Q
          gp.A = Y + qd; gp.B = Y + qd; gp.C = Zw2; gp.transA = true;
          gp.M = p; gp.N = p; gp.K = r_num; gp.lda = ldy; gp.ldb = ldy; gp.ldc = ldp; gp.alpha = -1.0; gp.beta = 0.0;
          gemm(ctx, gp);
          int grid = (p + 7) / 8;
          if (grid > 4 * ctx->sms) grid = 4 * ctx->sms;
          launch_pdl(ctx, "gs_projector", gs_projector_kernel, dim3(grid), dim3(256), 0, Zw2, p, ldp, norms2, rowstate2);
          launch_pdl(ctx, "gs_commit", gs_commit_kernel, dim3(1), dim3(32), 0, ctl, 1);
          // rows of the projector with squared norm above 0.05 only (their sum is p - r >= need - r, so there are always enough):
          // normalising a short row would amplify the rounding of its orthogonalisation (measured: isometry 1e-11 at 1e-12)
          const bool ok = run_panels(Zw2, p, norms2, rowstate2, need, r_num, false, 0.05);
          launch_pdl(ctx, "gs_commit", gs_commit_kernel, dim3(1), dim3(32), 0, ctl, 0);
          QTT_CUDA(cudaMemcpyAsync(h_r, &ctl->r, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
          QTT_CUDA(cudaStreamSynchronize(ctx->stream));
          if (!ok || *h_r < need) gs_done = false;
        }
      }
    }
  }
  if (!gs_done) {
    int rows = q + p;
    int grid = (rows + 7) / 8;
    if (grid > 4 * ctx->sms) grid = 4 * ctx->sms;
    launch_pdl(ctx, "qr_init", qr_init_kernel, dim3(grid), dim3(256), 0, M, q, p, ldp, SE, norms, rowstate, inpanel, ctl);
  }
  const int nb = panel_width(p);
  const size_t smem_panel = ((size_t)nb * ldp + q) * sizeof(double);
  const size_t smem_apply = (size_t)nb * ldp * sizeof(double);
  static DevOnce once;
  once.run(ctx->device, [&] {
    set_smem(qr_panel_kernel, 200 * 1024);
    set_smem(qr_apply_kernel<8>, 200 * 1024);
    set_smem(qr_apply_kernel<16>, 200 * 1024);
    set_smem(qr_apply_kernel<32>, 200 * 1024);
  });
  const int npanels = gs_done ? 0 : (kmax + nb - 1) / nb;
  int agrid = (q + p + PANEL_W - 1) / PANEL_W;
  if (agrid > ctx->sms) agrid = ctx->sms;
  for (int pn = 0; pn < npanels; ++pn) {
    // (a register-resident variant of this kernel -- rows in registers, reflector broadcast through shared memory -- was
    // measured slower in situ, 35 / 74 / 128 ms of split time per sweep against 32 / 69 / 127 ms, and was removed)
    launch_pdl(ctx, "qr_panel", qr_panel_kernel, dim3(1), dim3(PANEL_T), smem_panel, SE, q, p, ldp, kmax, nb, zt * zt, norms, rowstate, inpanel,
               Vp, taus, ctl);
    if (p <= 256) launch_pdl(ctx, "qr_apply", qr_apply_kernel<8>, dim3(agrid), dim3(PANEL_T), smem_apply, SE, q, p, ldp, rowstate, inpanel, Vp, taus, norms, ctl);
    else if (p <= 512) launch_pdl(ctx, "qr_apply", qr_apply_kernel<16>, dim3(agrid), dim3(PANEL_T), smem_apply, SE, q, p, ldp, rowstate, inpanel, Vp, taus, norms, ctl);
    else if (p <= 1024) launch_pdl(ctx, "qr_apply", qr_apply_kernel<32>, dim3(agrid), dim3(PANEL_T), smem_apply, SE, q, p, ldp, rowstate, inpanel, Vp, taus, norms, ctl);
    else launch_pdl(ctx, "qr_apply", qr_apply_generic_kernel, dim3(agrid), dim3(PANEL_T), 0, SE, q, p, ldp, rowstate, inpanel, Vp, taus, norms, ctl);
  }
  cudaEvent_t evq = nullptr, evj = nullptr;
  if (trace) {
    cudaEventCreate(&evq);
    cudaEventCreate(&evj);
    cudaEventRecord(evq, ctx->stream);
  }
  if (!gs_done) {
    dim3 grid((q + p + 31) / 32, (kmax + 31) / 32);
    launch_pdl(ctx, "build_y", build_y_kernel, grid, dim3(32, 8), 0, SE, q, p, ldp, qd, ldy, Y, ctl);
  }
  QTT_CUDA(cudaMemsetAsync(ctx->d_sync + 8, 0, sizeof(unsigned) * (2 + 200), ctx->stream));
  JqArgs a;
  a.Y = Y; a.ld = ldy; a.ndot = qd; a.kmax = kmax;
  a.max_sweeps = max_sweeps;
  a.tol = tol;
  a.sync = ctx->d_sync + 8;
  a.sig2 = sig2; a.perm = perm; a.norms = norms + q; a.info = info; a.derr = derr;
  a.cutoff = cutoff;
  a.maxdim = maxdim > 0 ? maxdim : kmax;
  if (a.maxdim > kmax) a.maxdim = kmax;
  a.mindim = mindim;
  a.ctl = ctl;
  a.eigen_mode = eigen_mode ? 1 : 0;
  a.dbg = nullptr;
  if (ctx->profiling && getenv("QTT_TRACE") != nullptr) {
    a.dbg = reinterpret_cast<long long*>(ctx->arena.alloc(96));   // [0, 32): phase counters; [32, 96): active slots per sweep
    QTT_CUDA(cudaMemsetAsync(a.dbg, 0, 768, ctx->stream));
  }
  const int npairs = ((kmax + 1) & ~1) / 2;
  int grid = (npairs + 1) / 2;
  if (grid > ctx->sms) grid = ctx->sms;
  if (grid < 1) grid = 1;
  const int half = qd > ldp ? qd : ldp;  // the longer of the two row parts
  static const bool gram_off = getenv("QTT_JACOBI") != nullptr && std::string(getenv("QTT_JACOBI")) == "cyclic";
  bool launched = false;
  if (!gram_off && kmax >= 64 && kmax <= 2048) {
    // the grid and the block size follow the numerical rank: one 4-byte read-back (the split synchronises for the kept rank anyway)
    int* h_r = reinterpret_cast<int*>(ctx->h_scal + 16);
    int r_rot = known_r;
    if (r_rot < 0) {
      QTT_CUDA(cudaMemcpyAsync(h_r, &ctl->r, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      QTT_CUDA(cudaMemcpyAsync(h_r + 1, &ctl->rj, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      r_rot = h_r[1] > 0 ? h_r[1] : h_r[0];
    }
    if (r_rot >= 32) launched = launch_jq_gram(ctx, a, r_rot);
  }
  if (launched) {
  } else if (half <= 1024 && kmax >= 16) {
    // block-cyclic kernel: one CTA (4 warp pairs) per pair of 4-row blocks
    const int nb4 = (((kmax + JB - 1) / JB) + 1) & ~1;
    int g4 = nb4 / 2;
    if (g4 > ctx->sms) g4 = ctx->sms;
    const size_t smem4 = (size_t)2 * JB * ldy * sizeof(double);
    if (half <= 64) launch_jq_block<1, JB>(ctx, a, g4, smem4);
    else if (half <= 128) launch_jq_block<2, JB>(ctx, a, g4, smem4);
    else if (half <= 256) launch_jq_block<4, JB>(ctx, a, g4, smem4);
    else if (half <= 512) launch_jq_block<8, JB>(ctx, a, g4, smem4);
    else launch_jq_block<16, JB>(ctx, a, g4, smem4);  // chi = 512: rows of 1024 + 1024 doubles
  } else if (ldy <= 128) launch_jq<2>(ctx, a, grid);
  else if (ldy <= 256) launch_jq<4>(ctx, a, grid);
  else if (ldy <= 512) launch_jq<8>(ctx, a, grid);
  else if (ldy <= 1024) launch_jq<16>(ctx, a, grid);
  else launch_jq<0>(ctx, a, grid);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(QTT_ERR_CUDA, "jacobi_qr launch failed: %s", cudaGetErrorString(e));
  }
  int* h_info = reinterpret_cast<int*>(ctx->h_scal);
  double* h_err = ctx->h_scal + 8;
  QTT_CUDA(cudaMemcpyAsync(h_info, info, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaMemcpyAsync(h_err, derr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (trace) {
    cudaEventRecord(evj, ctx->stream);
    cudaEventSynchronize(evj);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, evq, evj);
    float msq = 0.f;
    cudaEventElapsedTime(&msq, evq0, evq);
    cudaEventDestroy(evq0);
    fprintf(stderr, "[qtt trace]   svd_rows_qr %d x %d: %s QR %.3f ms, numerical rank %d, jacobi %d sweeps %.3f ms\n", q, p,
            gs_done ? "Gram-Schmidt" : "Householder", msq, h_info[3], h_info[1], ms);
    if (a.dbg) {
      long long h[96];
      cudaMemcpy(h, a.dbg, 768, cudaMemcpyDeviceToHost);
      if (h[8] > 0 && h[15] > 0) {
        const char* nm[8] = {"busy", "load", "gram", "csync1", "reduce", "check", "inner", "apply"};
        fprintf(stderr, "[qtt trace]   per round over %lld CTAs, max / mean cycles:", h[15]);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %.0f / %.0f", nm[k], (double)h[16 + k] / h[8], (double)h[24 + k] / h[15] / h[8]);
        fprintf(stderr, "\n");
      }
      if (h[8] > 0)
        fprintf(stderr, "[qtt trace]   gram jacobi CTA0 cycles per round: barrier %.0f load %.0f gram %.0f csync1 %.0f reduce %.0f check %.0f inner %.0f apply %.0f (%lld rounds)\n",
                (double)h[0] / h[8], (double)h[1] / h[8], (double)h[2] / h[8], (double)h[3] / h[8], (double)h[4] / h[8], (double)h[5] / h[8],
                (double)h[6] / h[8], (double)h[7] / h[8], h[8]);
      if (h[8] > 0) {
        fprintf(stderr, "[qtt trace]   block pairs with rotations per sweep:");
        for (int k = 0; k < h_info[1] && k < 64; ++k) fprintf(stderr, " %lld", h[32 + k]);
        fprintf(stderr, "\n");
      }
      if (h[8] > 0) fprintf(stderr, "[qtt trace]   per round: look-ahead warp computing %.0f, waiting at the step barrier %.0f cycles\n", (double)h[9] / h[8], (double)h[10] / h[8]);
      else if (h[4] > 0)
        fprintf(stderr, "[qtt trace]   block jacobi CTA0 cycles per outer round: barrier %.0f load %.0f steps %.0f store %.0f (%lld rounds)\n",
                (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4], h[4]);
    }
    cudaEventDestroy(evq);
    cudaEventDestroy(evj);
  }
  RowSvd res;
  res.k = h_info[0];
  res.sweeps = h_info[1];
  res.truncerr = *h_err;
  if (!h_info[2]) fail(QTT_ERR_NOCONV, "Jacobi SVD did not converge in %d sweeps (q=%d, p=%d, rank %d)", max_sweeps, q, p, h_info[3]);
  if (!gs_done) { ctx->qr_last_rank = h_info[3]; ctx->qr_last_need = need; }
  if (gs_done && res.k > h_info[3]) fail(QTT_ERR_UNSUPPORTED, "svd_rows_qr: %d rows kept but only %d basis rows exist on the Gram-Schmidt path", res.k, h_info[3]);
  res.Wn = ctx->arena.alloc((size_t)res.k * p);
  res.Wni = nullptr;
  dim3 grid2((p + 255) / 256, res.k);
  launch_pdl(ctx, "gather_rows", gather_rows_kernel, grid2, dim3(256), 0, Y, ldy, qd, p, perm, sig2, res.k, res.Wn, sigma_out, SE, q, ldp,
             a.eigen_mode);
  return res;
}

}  // namespace qtt

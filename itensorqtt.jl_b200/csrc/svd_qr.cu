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
    ctl->r = 0; ctl->done = 0; ctl->nbp = 0; ctl->serial = 0; ctl->F2 = -1.0; ctl->thr = 0.0;
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
  if (once.first(ctx->device)) {
    QTT_CUDA(cudaFuncSetAttribute(jacobi_qr_block_kernel<NC, JBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  void* args[] = {&a};
  QTT_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_qr_block_kernel<NC, JBT>, dim3(grid), dim3(64 * JBT), args, smem, ctx->stream));
  ctx->launches++;
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
  QrCtl* ctl = reinterpret_cast<QrCtl*>(ctx->arena.alloc(8));
  double* Y = ctx->arena.alloc((size_t)kmax * ldy);
  double* sig2 = ctx->arena.alloc(kmax);
  int* perm = reinterpret_cast<int*>(ctx->arena.alloc(kmax / 2 + 1));
  int* info = reinterpret_cast<int*>(ctx->arena.alloc(4));
  double* derr = ctx->arena.alloc(2);

  const double tol = sqrt((double)p) * 1.1102230246251565e-16;
  // rows below zt * |Z|_F are rounding noise: never rotated, sigma reported as 0.  In eigen mode the truncated spectrum is
  // |row| itself and the reference's cutoffs go down to 1e-15: the floor is kept as low as the rotation noise allows
  // (same rule as svd_rows_ex).
  const double zt = (eigen_mode ? 2.0 : 8.0) * tol;
  {
    int rows = q + p;
    int grid = (rows + 7) / 8;
    if (grid > 4 * ctx->sms) grid = 4 * ctx->sms;
    launch_pdl(ctx, "qr_init", qr_init_kernel, dim3(grid), dim3(256), 0, M, q, p, ldp, SE, norms, rowstate, inpanel, ctl);
  }
  const int nb = panel_width(p);
  const size_t smem_panel = ((size_t)nb * ldp + q) * sizeof(double);
  const size_t smem_apply = (size_t)nb * ldp * sizeof(double);
  static DevOnce once;
  if (once.first(ctx->device)) {
    set_smem(qr_panel_kernel, 200 * 1024);
    set_smem(qr_apply_kernel<8>, 200 * 1024);
    set_smem(qr_apply_kernel<16>, 200 * 1024);
    set_smem(qr_apply_kernel<32>, 200 * 1024);
  }
  const int npanels = (kmax + nb - 1) / nb;
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
  const bool trace = ctx->profiling && getenv("QTT_TRACE") != nullptr;
  cudaEvent_t evq = nullptr, evj = nullptr;
  if (trace) {
    cudaEventCreate(&evq);
    cudaEventCreate(&evj);
    cudaEventRecord(evq, ctx->stream);
  }
  {
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
    a.dbg = reinterpret_cast<long long*>(ctx->arena.alloc(8));
    QTT_CUDA(cudaMemsetAsync(a.dbg, 0, 64, ctx->stream));
  }
  const int npairs = ((kmax + 1) & ~1) / 2;
  int grid = (npairs + 1) / 2;
  if (grid > ctx->sms) grid = ctx->sms;
  if (grid < 1) grid = 1;
  const int half = qd > ldp ? qd : ldp;  // the longer of the two row parts
  if (half <= 1024 && kmax >= 16) {
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
    fprintf(stderr, "[qtt trace]   svd_rows_qr %d x %d: numerical rank %d, jacobi %d sweeps %.3f ms (after QR)\n", q, p, h_info[3], h_info[1], ms);
    if (a.dbg) {
      long long h[8];
      cudaMemcpy(h, a.dbg, 64, cudaMemcpyDeviceToHost);
      if (h[4] > 0)
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
  res.Wn = ctx->arena.alloc((size_t)res.k * p);
  res.Wni = nullptr;
  dim3 grid2((p + 255) / 256, res.k);
  launch_pdl(ctx, "gather_rows", gather_rows_kernel, grid2, dim3(256), 0, Y, ldy, qd, p, perm, sig2, res.k, res.Wn, sigma_out, SE, q, ldp,
             a.eigen_mode);
  return res;
}

}  // namespace qtt

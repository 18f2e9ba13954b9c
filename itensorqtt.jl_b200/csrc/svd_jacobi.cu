// Site splitting: truncated SVD of the two-site tensor by one-sided (Hestenes) Jacobi + the ITensors truncation rule.
// Replaces `replacebond!` -> `svd` (LAPACK gesdd) -> `truncate!`.
//
// Only the factor that becomes the orthonormal core is needed (U for ortho="left", V for ortho="right"); the other
// core is one GEMM with phi.  So the rows of Z = phi (ortho right) or Z = phi^T (ortho left) are orthogonalised by plane
// rotations; on convergence row i of Z is sigma_i v_i^T.  One-sided Jacobi computes small singular values to high
// RELATIVE accuracy, which the sigma^2-relative cutoffs of the reference (1e-12 ... 1e-15) require -- no Gram matrix.
//
// Parallel schedule: round-robin tournament, q/2 disjoint row pairs per round, one warp per pair with both rows held
// in registers (double2 per lane, coalesced), three warp-shuffle reductions (|zi|^2, |zj|^2, zi.zj) and an in-register
// rotation.  The whole decomposition (all rounds and sweeps, the convergence test, the descending sort of sigma^2 and
// the truncation scan) is ONE cooperative persistent kernel with a monotonic-counter grid barrier between rounds, so
// there is no host round trip until the kept rank is read back.
#include "common.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace qtt {

void permute(Ctx*, const double*, double*, int, const int64_t*, const int*);

namespace {

using dev::warp_sum;
using dev::ld_acquire;
using dev::grid_sync;

struct JacobiArgs {
  double* Zi;      // imaginary plane (same layout) or nullptr for real data
  int eigen_mode;  // 1: the spectrum that is truncated is |row| (eigenvalues of a Hermitian PSD matrix), not |row|^2
  double* Z;       // [q][ld] row-major, ld even, columns >= p are zero
  int q, p, ld;
  int max_sweeps;
  double tol;
  double zero_tol2;  // rows with |z|^2 <= zero_tol2 * |Z|_F^2 are numerically zero: never rotated, sigma^2 reported as 0
  unsigned* sync;  // [0] barrier counter, [1] spare, [2 + s] rotation count of sweep s
  double* sig2;    // [q] out: squared singular values, descending
  int* perm;       // [q] out: row index of the i-th largest
  double* norms;   // [q] scratch
  int* info;       // out: [0] kept rank, [1] sweeps used, [2] converged flag
  double* derr;    // out: [0] truncerr
  double cutoff;
  int maxdim, mindim;
};

__device__ __forceinline__ void rotation(double alpha, double beta, double gamma, double& c, double& s) {
  double zeta = (beta - alpha) / (2.0 * gamma);
  double t;
  if (fabs(zeta) > 1e150) t = 0.5 / zeta;
  else t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  c = rsqrt(1.0 + t * t);
  s = c * t;
}

// register-resident pair rotation: NCH double2 chunks per lane per row (covers p <= 128 * NCH / 2 ... i.e. 64*NCH columns)
template <int NCH>
__device__ __forceinline__ bool rotate_pair_reg(double* __restrict__ ra, double* __restrict__ rb, int ld, double tol, double thr,
                                                int lane) {
  double2 za[NCH], zb[NCH];
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    int col = 2 * lane + 64 * k;
    if (col < ld) {
      za[k] = __ldcg(reinterpret_cast<const double2*>(ra + col));
      zb[k] = __ldcg(reinterpret_cast<const double2*>(rb + col));
    } else {
      za[k] = make_double2(0.0, 0.0);
      zb[k] = make_double2(0.0, 0.0);
    }
    alpha += za[k].x * za[k].x + za[k].y * za[k].y;
    beta += zb[k].x * zb[k].x + zb[k].y * zb[k].y;
    gamma += za[k].x * zb[k].x + za[k].y * zb[k].y;
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gamma = warp_sum(gamma);
  if (!(alpha > thr) || !(beta > thr)) return false;
  double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(fabs(gamma) > tol * lim)) return false;
  double c, s;
  rotation(alpha, beta, gamma, c, s);
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    int col = 2 * lane + 64 * k;
    if (col < ld) {
      double2 na = make_double2(c * za[k].x - s * zb[k].x, c * za[k].y - s * zb[k].y);
      double2 nb = make_double2(s * za[k].x + c * zb[k].x, s * za[k].y + c * zb[k].y);
      __stcg(reinterpret_cast<double2*>(ra + col), na);
      __stcg(reinterpret_cast<double2*>(rb + col), nb);
    }
  }
  return true;
}

// generic two-pass version for long rows
__device__ __forceinline__ bool rotate_pair_mem(double* __restrict__ ra, double* __restrict__ rb, int ld, double tol, double thr,
                                                int lane) {
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int col = 2 * lane; col < ld; col += 64) {
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
  double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(fabs(gamma) > tol * lim)) return false;
  double c, s;
  rotation(alpha, beta, gamma, c, s);
  for (int col = 2 * lane; col < ld; col += 64) {
    double2 a = __ldcg(reinterpret_cast<const double2*>(ra + col));
    double2 b = __ldcg(reinterpret_cast<const double2*>(rb + col));
    __stcg(reinterpret_cast<double2*>(ra + col), make_double2(c * a.x - s * b.x, c * a.y - s * b.y));
    __stcg(reinterpret_cast<double2*>(rb + col), make_double2(s * a.x + c * b.x, s * a.y + c * b.y));
  }
  return true;
}

// complex rows (planar): gamma = <z_b, z_a> = sum z_a conj(z_b); z_b is first multiplied by the phase gamma/|gamma|, which
// makes the inner product real and positive, then the real rotation is applied.
__device__ __forceinline__ bool rotate_pair_cplx(double* __restrict__ ar, double* __restrict__ ai, double* __restrict__ br,
                                                 double* __restrict__ bi, int ld, double tol, double thr, int lane) {
  double alpha = 0.0, beta = 0.0, gr = 0.0, gi = 0.0;
  for (int col = 2 * lane; col < ld; col += 64) {
    double2 xr = __ldcg(reinterpret_cast<const double2*>(ar + col)), xi = __ldcg(reinterpret_cast<const double2*>(ai + col));
    double2 yr = __ldcg(reinterpret_cast<const double2*>(br + col)), yi = __ldcg(reinterpret_cast<const double2*>(bi + col));
    alpha += xr.x * xr.x + xi.x * xi.x + xr.y * xr.y + xi.y * xi.y;
    beta += yr.x * yr.x + yi.x * yi.x + yr.y * yr.y + yi.y * yi.y;
    gr += xr.x * yr.x + xi.x * yi.x + xr.y * yr.y + xi.y * yi.y;  // Re(x conj(y))
    gi += xi.x * yr.x - xr.x * yi.x + xi.y * yr.y - xr.y * yi.y;  // Im(x conj(y))
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gr = warp_sum(gr);
  gi = warp_sum(gi);
  if (!(alpha > thr) || !(beta > thr)) return false;
  double gabs = hypot(gr, gi);
  double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(gabs > tol * lim)) return false;
  double c, s;
  rotation(alpha, beta, gabs, c, s);
  const double wr = gr / gabs, wi = gi / gabs;  // phase omega
  for (int col = 2 * lane; col < ld; col += 64) {
    double2 xr = __ldcg(reinterpret_cast<const double2*>(ar + col)), xi = __ldcg(reinterpret_cast<const double2*>(ai + col));
    double2 yr = __ldcg(reinterpret_cast<const double2*>(br + col)), yi = __ldcg(reinterpret_cast<const double2*>(bi + col));
    // y~ = omega * y
    double2 tr = make_double2(wr * yr.x - wi * yi.x, wr * yr.y - wi * yi.y);
    double2 ti = make_double2(wr * yi.x + wi * yr.x, wr * yi.y + wi * yr.y);
    __stcg(reinterpret_cast<double2*>(ar + col), make_double2(c * xr.x - s * tr.x, c * xr.y - s * tr.y));
    __stcg(reinterpret_cast<double2*>(ai + col), make_double2(c * xi.x - s * ti.x, c * xi.y - s * ti.y));
    __stcg(reinterpret_cast<double2*>(br + col), make_double2(s * xr.x + c * tr.x, s * xr.y + c * tr.y));
    __stcg(reinterpret_cast<double2*>(bi + col), make_double2(s * xi.x + c * ti.x, s * xi.y + c * ti.y));
  }
  return true;
}

// the same with both rows held in registers between the inner products and the update (ld <= 64 NC): one pass over L2 instead of
// two (the rounds of the complex density matrices of `apply` are latency-bound: a round is a barrier plus these passes)
template <int NC>
__device__ __forceinline__ bool rotate_pair_cplx_reg(double* __restrict__ ar, double* __restrict__ ai, double* __restrict__ br,
                                                     double* __restrict__ bi, int ld, double tol, double thr, int lane) {
  double2 xr[NC], xi[NC], yr[NC], yi[NC];
  double alpha = 0.0, beta = 0.0, gr = 0.0, gi = 0.0;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int col = 2 * lane + 64 * k;
    if (col < ld) {
      xr[k] = __ldcg(reinterpret_cast<const double2*>(ar + col)); xi[k] = __ldcg(reinterpret_cast<const double2*>(ai + col));
      yr[k] = __ldcg(reinterpret_cast<const double2*>(br + col)); yi[k] = __ldcg(reinterpret_cast<const double2*>(bi + col));
    } else {
      xr[k] = xi[k] = yr[k] = yi[k] = make_double2(0.0, 0.0);
    }
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    alpha += xr[k].x * xr[k].x + xi[k].x * xi[k].x + xr[k].y * xr[k].y + xi[k].y * xi[k].y;
    beta += yr[k].x * yr[k].x + yi[k].x * yi[k].x + yr[k].y * yr[k].y + yi[k].y * yi[k].y;
    gr += xr[k].x * yr[k].x + xi[k].x * yi[k].x + xr[k].y * yr[k].y + xi[k].y * yi[k].y;  // Re(x conj(y))
    gi += xi[k].x * yr[k].x - xr[k].x * yi[k].x + xi[k].y * yr[k].y - xr[k].y * yi[k].y;  // Im(x conj(y))
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gr = warp_sum(gr);
  gi = warp_sum(gi);
  if (!(alpha > thr) || !(beta > thr)) return false;
  const double gabs = hypot(gr, gi);
  const double lim = sqrt(alpha) * sqrt(beta);
  if (!(lim > 0.0) || !(gabs > tol * lim)) return false;
  double c, s;
  rotation(alpha, beta, gabs, c, s);
  const double wr = gr / gabs, wi = gi / gabs;  // phase omega
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int col = 2 * lane + 64 * k;
    if (col < ld) {
      const double2 tr = make_double2(wr * yr[k].x - wi * yi[k].x, wr * yr[k].y - wi * yi[k].y);   // y~ = omega * y
      const double2 ti = make_double2(wr * yi[k].x + wi * yr[k].x, wr * yi[k].y + wi * yr[k].y);
      __stcg(reinterpret_cast<double2*>(ar + col), make_double2(c * xr[k].x - s * tr.x, c * xr[k].y - s * tr.y));
      __stcg(reinterpret_cast<double2*>(ai + col), make_double2(c * xi[k].x - s * ti.x, c * xi[k].y - s * ti.y));
      __stcg(reinterpret_cast<double2*>(br + col), make_double2(s * xr[k].x + c * tr.x, s * xr[k].y + c * tr.y));
      __stcg(reinterpret_cast<double2*>(bi + col), make_double2(s * xi[k].x + c * ti.x, s * xi[k].y + c * ti.y));
    }
  }
  return true;
}

__device__ __forceinline__ double row_norm2(const JacobiArgs& a, int i, int lane) {
  const double* ri = a.Z + (size_t)i * a.ld;
  double acc = 0.0;
  for (int col = 2 * lane; col < a.ld; col += 64) {
    double2 v = __ldcg(reinterpret_cast<const double2*>(ri + col));
    acc += v.x * v.x + v.y * v.y;
  }
  if (a.Zi) {
    const double* ii = a.Zi + (size_t)i * a.ld;
    for (int col = 2 * lane; col < a.ld; col += 64) {
      double2 v = __ldcg(reinterpret_cast<const double2*>(ii + col));
      acc += v.x * v.x + v.y * v.y;
    }
  }
  return warp_sum(acc);
}

// NCH > 0: real rows held in registers; NCH == 0: real rows, two passes; NCH == -1: complex rows (planar), two passes;
// NCH < -1: complex rows held in registers (-NCH chunks of 64 columns)
// CL: the whole grid is ONE thread-block cluster (<= 8 CTAs of 8 warps) and the per-round barrier is the hardware cluster
// barrier (release / acquire at cluster scope; the rows are accessed with .cg loads / stores, i.e. at L2) instead of the
// software grid barrier: the rounds of a small 2k x 2k density matrix are barrier-bound (profiles/r01_launches_apply_dft.md).
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int NCH, bool CL>
__global__ void __launch_bounds__(CL ? 256 : 64) jacobi_rows_kernel(JacobiArgs a) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int GW = gridDim.x * wpb;
  const int qe = (a.q + 1) & ~1;
  const int npairs = qe / 2, rounds = qe - 1;
  unsigned target = 0;
  auto bar = [&]() {
    if constexpr (CL) cluster_barrier();
    else grid_sync(a.sync, target);
  };
  int sweeps = 0;
  int converged = (a.q < 2) ? 1 : 0;
  // |Z|_F^2 (invariant under the rotations) sets the scale below which a row is numerically zero.  Without this a
  // rank-deficient Z (more rows than the row space has dimensions) never converges: the surplus rows are rounding
  // noise that keeps being re-orthogonalised at ever smaller scales.
  for (int i = gw; i < a.q; i += GW) {
    double acc = row_norm2(a, i, lane);
    if (lane == 0) __stcg(a.norms + i, acc);
  }
  bar();
  double F2 = 0.0;
  for (int j = lane; j < a.q; j += 32) F2 += __ldcg(a.norms + j);
  F2 = warp_sum(F2);
  const double thr = a.zero_tol2 * F2;
  bar();  // norms[] is rewritten after the sweeps
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
        if (ia >= a.q || ib >= a.q) continue;
        if (ia > ib) {
          int tmp = ia;
          ia = ib;
          ib = tmp;
        }
        double* ra = a.Z + (size_t)ia * a.ld;
        double* rb = a.Z + (size_t)ib * a.ld;
        bool rot;
        if constexpr (NCH > 0) rot = rotate_pair_reg<NCH>(ra, rb, a.ld, a.tol, thr, lane);
        else if constexpr (NCH == 0) rot = rotate_pair_mem(ra, rb, a.ld, a.tol, thr, lane);
        else if constexpr (NCH == -1) rot = rotate_pair_cplx(ra, a.Zi + (size_t)ia * a.ld, rb, a.Zi + (size_t)ib * a.ld, a.ld, a.tol, thr, lane);
        else rot = rotate_pair_cplx_reg<-NCH>(ra, a.Zi + (size_t)ia * a.ld, rb, a.Zi + (size_t)ib * a.ld, a.ld, a.tol, thr, lane);
        rotated |= rot;
      }
      bar();
    }
    if (rotated && lane == 0) atomicAdd(a.sync + 2 + sw, 1u);
    bar();
    sweeps = sw + 1;
    if (ld_acquire(a.sync + 2 + sw) == 0u) converged = 1;
  }
  // squared norms of the (now mutually orthogonal) rows
  for (int i = gw; i < a.q; i += GW) {
    double acc = row_norm2(a, i, lane);
    if (!(acc > thr)) acc = 0.0;
    if (a.eigen_mode) acc = sqrt(acc);
    if (lane == 0) __stcg(a.norms + i, acc);
  }
  bar();
  dev::rank_sort_desc(a.norms, a.q, a.q, a.sig2, a.perm);
  bar();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err;
    a.info[0] = dev::truncate_spectrum(a.sig2, a.q, a.cutoff, a.maxdim, a.mindim, &err);
    a.info[1] = sweeps;
    a.info[2] = converged;
    a.derr[0] = err;
  }
}

// Wn[i][:] = Z[perm[i]][:] / sqrt(sig2[i])  for i < k   (zero rows stay zero)
__global__ void gather_normalize_kernel(const double* __restrict__ Z, const double* __restrict__ Zi, int ld, int p,
                                        const int* __restrict__ perm, const double* __restrict__ sig2, int eigen_mode, int k,
                                        double* __restrict__ Wn, double* __restrict__ Wni, double* __restrict__ sigma_out) {
  int i = blockIdx.y;
  if (i >= k) return;
  double s = eigen_mode ? sig2[i] : sqrt(sig2[i]);  // row norm
  double inv = s > 0.0 ? 1.0 / s : 0.0;
  const size_t off = (size_t)perm[i] * ld;
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < p; col += gridDim.x * blockDim.x) {
    Wn[(size_t)i * p + col] = Z[off + col] * inv;
    if (Zi) Wni[(size_t)i * p + col] = Zi[off + col] * inv;
  }
  if (sigma_out && blockIdx.x == 0 && threadIdx.x == 0) sigma_out[i] = s;
}

// Z[r][0..ld) = src[r][0..p) zero padded
__global__ void pad_rows_kernel(const double* __restrict__ src, int q, int p, int ld, double* __restrict__ Z) {
  long total = (long)q * ld;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    int r = (int)(o / ld), c = (int)(o % ld);
    Z[o] = c < p ? src[(long)r * p + c] : 0.0;
  }
}

template <int NCH>
void launch_jacobi(Ctx* ctx, JacobiArgs& a, int grid, int cluster_ctas) {
  if (cluster_ctas > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_ctas);
    cfg.blockDim = dim3(256);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QTT_CUDA(cudaLaunchKernelEx(&cfg, jacobi_rows_kernel<NCH, true>, a));
    ctx->launches++;
    return;
  }
  void* args[] = {&a};
  QTT_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_rows_kernel<NCH, false>, dim3(grid), dim3(64), args, 0, ctx->stream));
  ctx->launches++;
}

}  // namespace

RowSvd svd_rows_ex(Ctx* ctx, const double* M, const double* Mi, int q, int p, double cutoff, int maxdim, int mindim,
                   double* sigma_out, int max_sweeps, bool eigen_mode) {
  // real Hermitian-PSD input (density matrices of a real `apply`: prolongation, retraction, real sums) of at least 32 rows:
  // rank-revealing QR + block-cyclic Jacobi (svd_qr.cu) with the eigenvalue spectrum; QTT_EIGEN_QR=0 keeps the plain kernel
  {
    static const bool eig_qr = !(getenv("QTT_EIGEN_QR") && atoi(getenv("QTT_EIGEN_QR")) == 0);
    if (eig_qr && eigen_mode && !Mi && q == p && q >= 32 && svd_rows_qr_supported(q, p))
      return svd_rows_qr(ctx, M, q, p, cutoff, maxdim, mindim, sigma_out, max_sweeps, true);
  }
  const int ld = (p + 1) & ~1;
  const int kmax = q < p ? q : p;
  const bool cplx = Mi != nullptr;
  if (max_sweeps <= 0) max_sweeps = 40;
  QTT_REQUIRE(max_sweeps <= 200, "svd_max_sweeps too large");
  double* Z = ctx->arena.alloc((size_t)q * ld);
  double* Zi = cplx ? ctx->arena.alloc((size_t)q * ld) : nullptr;
  double* sig2 = ctx->arena.alloc(q);
  double* norms = ctx->arena.alloc(q);
  int* perm = reinterpret_cast<int*>(ctx->arena.alloc((q + 1) / 2 + 1));
  int* info = reinterpret_cast<int*>(ctx->arena.alloc(4));
  double* derr = ctx->arena.alloc(2);
  auto stage = [&](const double* src, double* dst) {
    if (ld == p) {
      QTT_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)q * p, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      pad_rows_kernel<<<ctx->sms * 2, 256, 0, ctx->stream>>>(src, q, p, ld, dst);
      check_launch(ctx, "pad_rows");
    }
  };
  stage(M, Z);
  if (cplx) stage(Mi, Zi);
  QTT_CUDA(cudaMemsetAsync(ctx->d_sync + 8, 0, sizeof(unsigned) * (2 + 200), ctx->stream));
  JacobiArgs a;
  a.Z = Z; a.Zi = Zi; a.eigen_mode = eigen_mode ? 1 : 0;
  a.q = q; a.p = p; a.ld = ld;
  a.max_sweeps = max_sweeps;
  a.tol = sqrt((double)p) * 1.1102230246251565e-16;
  // rows below zero_tol * |Z|_F are rounding noise of the rotations.  In eigen mode the truncated spectrum is |row| itself
  // and the reference's cutoffs go down to 1e-15, so the floor is kept as low as the rotation noise allows.
  const double zt = (eigen_mode ? 2.0 : 8.0) * a.tol;
  a.zero_tol2 = zt * zt;
  a.sync = ctx->d_sync + 8;
  a.sig2 = sig2; a.perm = perm; a.norms = norms; a.info = info; a.derr = derr;
  a.cutoff = cutoff;
  a.maxdim = maxdim > 0 ? maxdim : kmax;
  if (a.maxdim > kmax) a.maxdim = kmax;
  a.mindim = mindim;
  const int npairs = ((q + 1) & ~1) / 2;
  int grid = (npairs + 1) / 2;
  if (grid > ctx->sms) grid = ctx->sms;
  if (grid < 1) grid = 1;
  // up to 64 row pairs (q <= 128): one cluster of <= 8 CTAs x 8 warps, one pair per warp and round, hardware cluster barrier per
  // round (QTT_JACOBI_CLUSTER=0 restores the cooperative grid).  Measured on the DFT apply: n = 20, chi = 64 (2k <= 124)
  // 26.4 -> 23.6 ms; with two pairs per warp (n = 24, chi = 128, 2k = 252) the 64 warps are work-bound and the grid of
  // up to 148 two-warp CTAs wins (91 vs 117 ms), so larger problems keep the grid.
  static const int cl_env = getenv("QTT_JACOBI_CLUSTER") ? atoi(getenv("QTT_JACOBI_CLUSTER")) : 1;
  int cl = 0;
  if (cl_env && npairs >= 2 && npairs <= 64) {
    cl = (npairs + 7) / 8;
    if (cl > 8) cl = 8;
  }
  if (cplx) {   // complex rows: in registers up to 512 columns (NCH = -(chunks of 64 columns)), two passes beyond
    if (ld <= 128) launch_jacobi<-2>(ctx, a, grid, cl);
    else if (ld <= 256) launch_jacobi<-4>(ctx, a, grid, cl);
    else if (ld <= 512) launch_jacobi<-8>(ctx, a, grid, cl);
    else launch_jacobi<-1>(ctx, a, grid, cl);
  }
  else if (ld <= 128) launch_jacobi<2>(ctx, a, grid, cl);
  else if (ld <= 256) launch_jacobi<4>(ctx, a, grid, cl);
  else if (ld <= 512) launch_jacobi<8>(ctx, a, grid, cl);
  else if (ld <= 1024) launch_jacobi<16>(ctx, a, grid, cl);
  else launch_jacobi<0>(ctx, a, grid, cl);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(QTT_ERR_CUDA, "jacobi launch failed: %s", cudaGetErrorString(e));
  }
  // the kept rank decides every later shape: one small read-back per split
  int* h_info = reinterpret_cast<int*>(ctx->h_scal);
  double* h_err = ctx->h_scal + 8;
  QTT_CUDA(cudaMemcpyAsync(h_info, info, sizeof(int) * 3, cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaMemcpyAsync(h_err, derr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  RowSvd res;
  res.k = h_info[0];
  res.sweeps = h_info[1];
  res.truncerr = *h_err;
  if (!h_info[2]) fail(QTT_ERR_NOCONV, "Jacobi SVD did not converge in %d sweeps (q=%d, p=%d)", max_sweeps, q, p);
  res.Wn = ctx->arena.alloc((size_t)res.k * p);
  res.Wni = cplx ? ctx->arena.alloc((size_t)res.k * p) : nullptr;
  dim3 grid2((p + 255) / 256, res.k);
  gather_normalize_kernel<<<grid2, 256, 0, ctx->stream>>>(Z, Zi, ld, p, perm, sig2, a.eigen_mode, res.k, res.Wn, res.Wni, sigma_out);
  check_launch(ctx, "gather_normalize");
  return res;
}

RowSvd svd_rows(Ctx* ctx, const double* M, int q, int p, double cutoff, int maxdim, int mindim, double* sigma_out, int max_sweeps) {
  // fast path: rank-revealing QR + Jacobi on the rows of R (svd_qr.cu); QTT_SVD=jacobi forces the plain kernel (A/B runs)
  static const bool plain = getenv("QTT_SVD") != nullptr && std::string(getenv("QTT_SVD")) == "jacobi";
  if (!plain && svd_rows_qr_supported(q, p)) return svd_rows_qr(ctx, M, q, p, cutoff, maxdim, mindim, sigma_out, max_sweeps);
  return svd_rows_ex(ctx, M, nullptr, q, p, cutoff, maxdim, mindim, sigma_out, max_sweeps, false);
}

SplitResult split2(Ctx* ctx, const double* phi, int chia, int chic, int ortho, double cutoff, int maxdim, int mindim, double* A_out,
                   double* B_out, double* sigma_out, int max_sweeps) {
  const int m = 2 * chia, n = 2 * chic;
  size_t mark = ctx->arena.mark();
  SplitResult res;
  try {
    if (ortho == 0) {
      // rows of phi^T (coalesced shared-memory transpose): A[(a s1)][k] = Wn^T ; B[k][(s2 c)] = Wn . phi
      double* phiT = ctx->arena.alloc((size_t)m * n);
      int64_t dims[2] = {m, n};
      int pm[2] = {1, 0};
      permute(ctx, phi, phiT, 2, dims, pm);
      RowSvd r = svd_rows(ctx, phiT, n, m, cutoff, maxdim, mindim, sigma_out, max_sweeps);
      int64_t dims2[2] = {r.k, m};
      permute(ctx, r.Wn, A_out, 2, dims2, pm);
      GemmDesc g;
      g.A = r.Wn; g.B = phi; g.C = B_out;
      g.M = r.k; g.N = n; g.K = m;
      g.lda = m; g.ldb = n; g.ldc = n;
      gemm(ctx, g);
      res.rank = r.k; res.sweeps = r.sweeps; res.truncerr = r.truncerr;
    } else {
      // B[k][(s2 c)] = Wn ; A[(a s1)][k] = phi . Wn^T
      RowSvd r = svd_rows(ctx, phi, m, n, cutoff, maxdim, mindim, sigma_out, max_sweeps);
      QTT_CUDA(cudaMemcpyAsync(B_out, r.Wn, sizeof(double) * (size_t)r.k * n, cudaMemcpyDeviceToDevice, ctx->stream));
      GemmDesc g;
      g.A = phi; g.B = r.Wn; g.C = A_out;
      g.transB = true;
      g.M = m; g.N = r.k; g.K = n;
      g.lda = n; g.ldb = n; g.ldc = r.k;
      gemm(ctx, g);
      res.rank = r.k; res.sweeps = r.sweeps; res.truncerr = r.truncerr;
    }
  } catch (...) {
    ctx->arena.release(mark);
    throw;
  }
  // Scratch lives in the arena: the consumer kernels are already enqueued on the same stream and the arena is only
  // re-used by later work on that stream, so releasing here is safe.
  ctx->arena.release(mark);
  return res;
}

}  // namespace qtt

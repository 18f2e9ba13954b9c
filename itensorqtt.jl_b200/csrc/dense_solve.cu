// Dense local solvers of the two-site sweep (small bonds only: the matrix is (4 chi_l chi_r)^2).
//
//   reference src/linsolve.jl:54-97  `linsolve_pseudoinverse`: T = lproj * W_j * W_{j+1} * rproj as a dense matrix,
//                                     `F = qr!(T_mat); x_vec = F \ b_vec`                         -> dense_local_solve()
//   reference src/dmrg_target.jl:1-12 `dmrg_target`: H = contract(PH, 1), `eigen(H; ishermitian=true)`, eigenvector of the
//                                     eigenvalue nearest `target_eigenvalue`                      -> dense_local_eigen()
//
// Both use one blocked Householder QR (compact-WY panels of 32 reflectors, trailing updates on the DMMA GEMM core).  The
// operator is stored TRANSPOSED (Tt[col][row]) so that a column of T -- the thing a reflector is built from -- is a
// contiguous row; the right-hand side is appended as one more row and receives Q^T b for free.
//   solve : R x = (Q^T b)[0:n] by blocked back substitution (single CTA, y in shared memory).
//   eigen : T - target = Q R  =>  (T - target)^T (T - target) = R^T R; inverse iteration z <- (R^T R)^{-1} z (two
//           triangular solves per step, no Q needed) converges to the right singular vector of the smallest singular value
//           of T - target, which for the symmetric effective operator is the eigenvector of the eigenvalue nearest the
//           target, at rate (|l1 - target| / |l2 - target|)^2 per step.  The whole iteration runs inside one kernel.
#include "common.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace qtt {
namespace {

using dev::warp_sum;

constexpr int DNB = 32;    // reflectors per panel
constexpr int DPT = 1024;  // threads of the single-CTA triangular-solve kernel
constexpr int PANEL_G_MAX = 32;  // CTAs of the cooperative panel kernel

// Tt[col][row] = a0 [row == col] + a1 sum_{wl, wr} L[wl][a'][a] W12[(wl s12)][(t12 wr)] R[wr][c][c']
// row = (a' t12 c'), col = (a s12 c).  Optional extra row `len` <- rhs.
__global__ void dense_op_t_kernel(const double* __restrict__ L, const double* __restrict__ W12, const double* __restrict__ R,
                                  double* __restrict__ Tt, long ld, int chil, int chir, int wl, int wr, double a0, double a1,
                                  const double* __restrict__ rhs) {
  const int len = 4 * chil * chir;
  const int col = blockIdx.y;
  if (col == len) {
    if (rhs)
      for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < len; row += gridDim.x * blockDim.x) Tt[(long)len * ld + row] = rhs[row];
    return;
  }
  const int a = col / (4 * chir), s12 = (col / chir) & 3, c = col % chir;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < len; row += gridDim.x * blockDim.x) {
    const int ap = row / (4 * chir), t12 = (row / chir) & 3, cp = row % chir;
    double acc = 0.0;
    for (int l = 0; l < wl; ++l) {
      const double lv = L[((long)l * chil + ap) * chil + a];
      if (lv == 0.0) continue;
      const double* w = W12 + (long)(l * 4 + s12) * (4 * wr) + t12 * wr;
      double inner = 0.0;
      for (int r = 0; r < wr; ++r) inner += w[r] * R[((long)r * chir + c) * chir + cp];
      acc += lv * inner;
    }
    Tt[(long)col * ld + row] = a1 * acc + (row == col ? a0 : 0.0);
  }
}

// One panel of the row-oriented Householder QR: rows j0 .. j0+nb-1 of Tt (length m), reflector jj annihilates
// row[j0+jj][j0+jj+1 : m].  Vp (nb x ldv, index i - j0) receives the reflectors with explicit unit diagonal and zeros left
// of it, Tm (DNB x DNB row-major) the upper-triangular compact-WY factor:  H_0 H_1 ... H_{nb-1} = I - Vp^T Tm Vp.
//
// Cooperative: CTA g keeps the column slice [j0 + g*slice, j0 + (g+1)*slice) of the whole panel in shared memory.  Per
// reflector ONE grid barrier: every CTA publishes its partial dots <r_jj, r_c> over the columns right of j (CTA 0 also the
// column-j entries), then all CTAs reduce the partials in the same fixed order -- bit-identical scalars everywhere -- and
// update their slice.  (A first single-CTA version that streamed the rows from L2, one warp per row, took 0.7-2.6 ms per
// panel at m = 4096; 90 % of a dense bond update.)
struct PanelArgs {
  double* Tt; long ld; int m, j0, nb;
  double* Vp; long ldv; double* Tm;
  double* part;      // [2][G][64] partial dots (double-buffered by step parity) + [2][64] column-j entries + [G][1024] Gram partials
  unsigned* bar;     // monotonic barrier counter
  unsigned epoch;    // its value when this launch starts
  int slice;         // columns per CTA (multiple of 8, >= DNB unless the grid is one CTA)
};
constexpr int PANEL_THREADS = 256;

__global__ void __launch_bounds__(PANEL_THREADS) dqr_panel_kernel(PanelArgs a) {
  extern __shared__ double P[];  // [DNB][slice]
  __shared__ double tau_s[DNB];
  __shared__ double dsum[DNB];
  __shared__ double Gs[DNB][DNB + 1];
  __shared__ double Ts[DNB][DNB + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = PANEL_THREADS / 32;
  const int G = gridDim.x, g = blockIdx.x, slice = a.slice, nb = a.nb;
  const int c0 = a.j0 + g * slice;                       // first global column of this slice
  int ncol = a.m - c0; if (ncol > slice) ncol = slice; if (ncol < 0) ncol = 0;
  unsigned target = a.epoch;
  double* colv_base = a.part + (size_t)2 * G * 64;
  double* gram_part = colv_base + 2 * 64;
  for (int r = warp; r < nb; r += NW)
    for (int i = lane; i < ncol; i += 32) P[r * slice + i] = a.Tt[(long)(a.j0 + r) * a.ld + c0 + i];
  __syncthreads();
  for (int jj = 0; jj < nb; ++jj) {
    const int j = a.j0 + jj;
    int lo = j + 1 - c0; if (lo < 0) lo = 0;             // local columns strictly right of j
    double* part = a.part + (size_t)(jj & 1) * G * 64;
    double* colv = colv_base + (jj & 1) * 64;
    const double* rj = P + jj * slice;
    for (int c = jj + warp; c < nb; c += NW) {
      const double* rc = P + c * slice;
      double acc = 0.0;
      for (int i = lo + lane; i < ncol; i += 32) acc += rj[i] * rc[i];
      acc = warp_sum(acc);
      if (lane == 0) {
        part[g * 64 + c] = acc;
        if (g == 0) colv[c] = rc[jj];                    // slice >= DNB when G > 1: column j lives in CTA 0 at local index jj
      }
    }
    target += G;
    if (G > 1) dev::grid_arrive_wait(a.bar, target);
    else __syncthreads();
    // reduce (same order in every CTA)
    for (int c = jj + warp; c < nb; c += NW) {
      double v = 0.0;
      for (int gg = lane; gg < G; gg += 32) v += __ldcg(part + gg * 64 + c);
      v = warp_sum(v);
      if (lane == 0) dsum[c] = v;
    }
    __syncthreads();
    const double alpha = __ldcg(colv + jj), s2 = dsum[jj];
    double tau = 0.0, beta = alpha, scale = 0.0;
    if (s2 != 0.0) {
      beta = -copysign(sqrt(alpha * alpha + s2), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (tau != 0.0) {
      for (int c = jj + 1 + warp; c < nb; c += NW) {
        double* rc = P + c * slice;
        const double f = tau * (__ldcg(colv + c) + scale * dsum[c]);
        const double fs = f * scale;
        for (int i = lo + lane; i < ncol; i += 32) rc[i] -= fs * rj[i];
        if (g == 0 && lane == 0) rc[jj] -= f;
      }
    }
    __syncthreads();
    {
      double* rw = P + jj * slice;
      for (int i = lo + tid; i < ncol; i += PANEL_THREADS) rw[i] *= scale;
      if (g == 0 && tid == 0) rw[jj] = beta;
      if (tid == 0) tau_s[jj] = tau;
    }
    __syncthreads();
  }
  // write back the factored slice and the explicit reflectors
  for (int r = warp; r < nb; r += NW) {
    const int jr = a.j0 + r;
    for (int i = lane; i < ncol; i += 32) {
      const double v = P[r * slice + i];
      const int col = c0 + i;
      a.Tt[(long)jr * a.ld + col] = v;
      a.Vp[(long)r * a.ldv + (col - a.j0)] = col > jr ? v : (col == jr ? 1.0 : 0.0);
    }
  }
  // Gram matrix of the reflectors G[x][y] = v_x . v_y (x < y), partial over this slice
  for (int e = warp; e < nb * nb; e += NW) {
    const int x = e / nb, y = e % nb;
    if (x >= y) continue;
    const int jy = a.j0 + y;                             // v_y vanishes left of jy, equals 1 at jy
    double acc = 0.0;
    for (int i = lane; i < ncol; i += 32) {
      const int col = c0 + i;
      if (col < jy) continue;
      const double vy = col == jy ? 1.0 : P[y * slice + i];
      acc += vy * P[x * slice + i];                      // col >= jy > jx: stored value is v_x
    }
    acc = warp_sum(acc);
    if (lane == 0) gram_part[(size_t)g * 1024 + x * DNB + y] = acc;
  }
  target += G;
  if (G > 1) dev::grid_arrive_wait(a.bar, target);
  else __syncthreads();
  if (g != 0) return;
  for (int e = tid; e < nb * nb; e += PANEL_THREADS) {
    const int x = e / nb, y = e % nb;
    if (x >= y) continue;
    double v = 0.0;
    for (int gg = 0; gg < G; ++gg) v += __ldcg(gram_part + (size_t)gg * 1024 + x * DNB + y);
    Gs[x][y] = v;
  }
  __syncthreads();
  // T(0:j, j) = -tau_j T(0:j, 0:j) G(0:j, j),  T(j, j) = tau_j   (LAPACK dlarft, forward / columnwise)
  if (warp == 0) {
    const int i = lane;
    for (int j = 0; j < DNB; ++j) Ts[i][j] = 0.0;
    __syncwarp();
    for (int j = 0; j < nb; ++j) {
      const double tj = tau_s[j];
      if (i < j) {
        double acc = 0.0;
        for (int k = i; k < j; ++k) acc += Ts[i][k] * Gs[k][j];
        Ts[i][j] = -tj * acc;
      } else if (i == j) {
        Ts[i][j] = tj;
      }
      __syncwarp();
    }
    for (int j = 0; j < DNB; ++j) a.Tm[i * DNB + j] = Ts[i][j];
  }
}

__device__ __forceinline__ double block_reduce(double v, double* red, bool take_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (take_max) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  } else {
    v = warp_sum(v);
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = take_max ? fmax(r, red[w]) : r + red[w];
  return r;
}

struct TriArgs {
  const double* Tt;  // R^T in the lower triangle: R(i, k) = Tt[k][i], k >= i
  long ld;
  int n;
  double* y;         // solve: in (Q^T b)[0:n], out x.  eigen: in start vector, out normalised eigenvector
  int eigen;         // 0: one backward solve; 1: inverse iteration on R^T R
  int maxit;
  double* stats;     // [0] iterations, [1] 1 - |<z_new, z_old>| at exit, [2] tiny / zero pivots met, [3] max |R_ii|
};

// y_s <- R^{-1} y_s (blocked back substitution)
__device__ void tri_backward(const TriArgs& a, double* ys, double (*blk)[DNB + 1], double pivmin, int* zero_piv) {
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c0 = ((n - 1) / DNB) * DNB; c0 >= 0; c0 -= DNB) {
    const int bs = n - c0 < DNB ? n - c0 : DNB;
    {
      const int k = warp, i = lane;  // blk[k][i] = R(c0 + i, c0 + k)
      blk[k][i] = (k < bs && i < bs) ? a.Tt[(long)(c0 + k) * a.ld + c0 + i] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double yi = lane < bs ? ys[c0 + lane] : 0.0;
      for (int k = bs - 1; k >= 0; --k) {
        double piv = blk[k][k];
        if (fabs(piv) <= pivmin) {
          if (lane == 0) ++*zero_piv;
          piv = (piv < 0.0) ? -pivmin : pivmin;
        }
        const double xk = __shfl_sync(0xffffffffu, yi, k) / piv;
        if (lane == k) yi = xk;
        else if (lane < k) yi -= xk * blk[k][lane];
      }
      if (lane < bs) ys[c0 + lane] = yi;
    }
    __syncthreads();
    for (int i = tid; i < c0; i += DPT) {
      double acc = 0.0;
      for (int k = 0; k < bs; ++k) acc += ys[c0 + k] * a.Tt[(long)(c0 + k) * a.ld + i];
      ys[i] -= acc;
    }
    __syncthreads();
  }
}

// y_s <- R^{-T} y_s (blocked forward substitution)
__device__ void tri_forward(const TriArgs& a, double* ys, double (*blk)[DNB + 1], double pivmin, int* zero_piv) {
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c0 = 0; c0 < n; c0 += DNB) {
    const int bs = n - c0 < DNB ? n - c0 : DNB;
    if (warp < bs) {
      const double* r = a.Tt + (long)(c0 + warp) * a.ld;
      double acc = 0.0;
      for (int k = lane; k < c0; k += 32) acc += r[k] * ys[k];
      acc = warp_sum(acc);
      if (lane == 0) ys[c0 + warp] -= acc;
    }
    {
      const int c = warp, k = lane;  // blk[c][k] = R(c0 + k, c0 + c), k <= c
      blk[c][k] = (c < bs && k < bs) ? a.Tt[(long)(c0 + c) * a.ld + c0 + k] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double zc = lane < bs ? ys[c0 + lane] : 0.0;
      for (int k = 0; k < bs; ++k) {
        double piv = blk[k][k];
        if (fabs(piv) <= pivmin) {
          if (lane == 0) ++*zero_piv;
          piv = (piv < 0.0) ? -pivmin : pivmin;
        }
        const double wk = __shfl_sync(0xffffffffu, zc, k) / piv;
        if (lane == k) zc = wk;
        else if (lane > k) zc -= wk * blk[lane][k];
      }
      if (lane < bs) ys[c0 + lane] = zc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(DPT) dtri_kernel(TriArgs a) {
  extern __shared__ double sm[];
  __shared__ double blk[DNB][DNB + 1];
  __shared__ double red[DPT / 32];
  __shared__ int zero_piv;
  const int n = a.n, tid = threadIdx.x;
  double* ys = sm;
  double* zs = sm + n;  // eigen mode only
  if (tid == 0) zero_piv = 0;
  double dmax = 0.0;
  for (int i = tid; i < n; i += DPT) {
    ys[i] = a.y[i];
    dmax = fmax(dmax, fabs(a.Tt[(long)i * a.ld + i]));
  }
  dmax = block_reduce(dmax, red, true);
  if (!a.eigen) {
    // `F \ b` for a square QR: plain triangular solve; an exactly zero pivot is reported (Julia: SingularException)
    tri_backward(a, ys, blk, 0.0, &zero_piv);
    for (int i = tid; i < n; i += DPT) a.y[i] = ys[i];
    if (tid == 0) { a.stats[0] = 0.0; a.stats[1] = 0.0; a.stats[2] = zero_piv; a.stats[3] = dmax; }
    return;
  }
  // inverse iteration; pivots below eps * max|R_ii| are replaced by that value (the usual safeguard when the shift is an
  // eigenvalue to working precision)
  const double pivmin = 2.220446049250313e-16 * dmax;
  double nrm = 0.0;
  for (int i = tid; i < n; i += DPT) nrm += ys[i] * ys[i];
  nrm = block_reduce(nrm, red, false);
  const double inv0 = nrm > 0.0 ? rsqrt(nrm) : 0.0;
  for (int i = tid; i < n; i += DPT) ys[i] = nrm > 0.0 ? ys[i] * inv0 : (i == 0 ? 1.0 : 0.0);
  __syncthreads();
  int it = 0;
  double conv = 1.0;
  int settled = 0;
  for (; it < a.maxit;) {
    for (int i = tid; i < n; i += DPT) zs[i] = ys[i];
    __syncthreads();
    tri_forward(a, ys, blk, pivmin, &zero_piv);
    tri_backward(a, ys, blk, pivmin, &zero_piv);
    double s = 0.0, d = 0.0;
    for (int i = tid; i < n; i += DPT) { s += ys[i] * ys[i]; d += ys[i] * zs[i]; }
    s = block_reduce(s, red, false);
    d = block_reduce(d, red, false);
    const double inv = 1.0 / sqrt(s);
    for (int i = tid; i < n; i += DPT) ys[i] *= inv;
    __syncthreads();
    ++it;
    conv = 1.0 - fabs(d) * inv;
    // two consecutive steps at rounding level: the direction no longer moves
    if (conv < 1e-15) { if (++settled >= 2) break; } else settled = 0;
  }
  for (int i = tid; i < n; i += DPT) a.y[i] = ys[i];
  if (tid == 0) { a.stats[0] = it; a.stats[1] = conv; a.stats[2] = zero_piv; a.stats[3] = dmax; }
}

size_t round8(size_t v) { return (v + 7) & ~size_t(7); }

struct DenseLayout {
  size_t ld, tt, vp, wk, wk2, tm, part, total;
};
DenseLayout dense_layout(size_t len) {
  DenseLayout d;
  d.ld = round8(len);
  d.tt = 0;
  size_t off = (len + 1) * d.ld;
  d.vp = off; off += (size_t)DNB * d.ld;
  d.wk = off; off += round8((len + 1) * DNB);
  d.wk2 = off; off += round8((len + 1) * DNB);
  d.tm = off; off += DNB * DNB;
  d.part = off; off += (size_t)2 * PANEL_G_MAX * 64 + 2 * 64 + (size_t)PANEL_G_MAX * 1024;
  d.total = off + 64;
  return d;
}

// Householder QR of the first n rows of Tt (each of length m = n) applied to all `nrow` rows (nrow - n appended rhs rows).
void dense_qr_rows(Ctx* ctx, double* base, const DenseLayout& d, int n, int nrow) {
  double* Tt = base + d.tt;
  double* Vp = base + d.vp;
  double* Wk = base + d.wk;
  double* Wk2 = base + d.wk2;
  double* Tm = base + d.tm;
  const int m = n;
  for (int j0 = 0; j0 < n; j0 += DNB) {
    const int nb = n - j0 < DNB ? n - j0 : DNB;
    {
      const int K = m - j0;
      int G = K <= 512 ? 1 : K / 128;  // small panels: one CTA, the whole panel in shared memory, no grid barrier
      if (G < 1) G = 1;
      if (G > PANEL_G_MAX) G = PANEL_G_MAX;
      if (G > ctx->sms) G = ctx->sms;
      int slice = (int)round8((size_t)(K + G - 1) / G);
      if (G > 1 && slice < DNB) slice = DNB;
      PanelArgs pa;
      pa.Tt = Tt; pa.ld = (long)d.ld; pa.m = m; pa.j0 = j0; pa.nb = nb; pa.Vp = Vp; pa.ldv = (long)d.ld; pa.Tm = Tm;
      pa.part = base + d.part; pa.bar = ctx->d_sync + 240; pa.epoch = ctx->dense_epoch; pa.slice = slice;
      const size_t smem = sizeof(double) * (size_t)DNB * slice;
      static DevOnce once;
      once.run(ctx->device, [&] {
        QTT_CUDA(cudaFuncSetAttribute(dqr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      });
      void* args[] = {&pa};
      QTT_CUDA(cudaLaunchCooperativeKernel((void*)dqr_panel_kernel, dim3(G), dim3(PANEL_THREADS), args, smem, ctx->stream));
      ctx->launches++;
      if (G > 1) ctx->dense_epoch += (unsigned)G * (unsigned)(nb + 1);
    }
    const int nrem = nrow - (j0 + nb);
    if (nrem <= 0) continue;
    double* C = Tt + (size_t)(j0 + nb) * d.ld + j0;
    const int K = m - j0;
    GemmDesc g1;  // Wk = C . Vp^T
    g1.A = C; g1.B = Vp; g1.C = Wk; g1.M = nrem; g1.N = nb; g1.K = K; g1.lda = (long)d.ld; g1.ldb = (long)d.ld; g1.ldc = DNB; g1.transB = true;
    gemm(ctx, g1);
    GemmDesc g2;  // Wk2 = Wk . Tm
    g2.A = Wk; g2.B = Tm; g2.C = Wk2; g2.M = nrem; g2.N = nb; g2.K = nb; g2.lda = DNB; g2.ldb = DNB; g2.ldc = DNB;
    gemm(ctx, g2);
    GemmDesc g3;  // C -= Wk2 . Vp
    g3.A = Wk2; g3.B = Vp; g3.C = C; g3.M = nrem; g3.N = K; g3.K = nb; g3.lda = DNB; g3.ldb = (long)d.ld; g3.ldc = (long)d.ld;
    g3.alpha = -1.0; g3.beta = 1.0;
    gemm(ctx, g3);
  }
}

void launch_tri(Ctx* ctx, TriArgs& a) {
  const size_t smem = sizeof(double) * (size_t)a.n * (a.eigen ? 2 : 1);
  static DevOnce once;  // (the kernel's 8.7 KB of static shared memory count against the 48 KB default)
  once.run(ctx->device, [&] { QTT_CUDA(cudaFuncSetAttribute(dtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024))); });
  dtri_kernel<<<1, DPT, smem, ctx->stream>>>(a);
  check_launch(ctx, "dtri");
}

void build_op(Ctx* ctx, const double* L, const double* W12, const double* R, double* Tt, size_t ld, int chil, int chir, int wl, int wr,
              double a0, double a1, const double* rhs) {
  const int len = 4 * chil * chir;
  int gx = (len + 255) / 256;
  if (gx > 8) gx = 8;
  dim3 grid(gx, len + (rhs ? 1 : 0));
  dense_op_t_kernel<<<grid, 256, 0, ctx->stream>>>(L, W12, R, Tt, (long)ld, chil, chir, wl, wr, a0, a1, rhs);
  check_launch(ctx, "dense_op_t");
}

}  // namespace

// ---- `b_linsolve` local solve: rectangular projected operator + SVD least squares ------------------------------------------
namespace {
// T[row][col], row = (a' t12 c') in the bra (= right-hand side) frame, col = (a s12 c) in the ket frame
__global__ void dense_op_rect_kernel(const double* __restrict__ L, const double* __restrict__ W12, const double* __restrict__ R,
                                     double* __restrict__ T, int cbl, int cbr, int chil, int chir, int wl, int wr) {
  const long N = 4L * chil * chir;
  const int row = blockIdx.y;
  const int ap = row / (4 * cbr), t12 = (row / cbr) & 3, cp = row % cbr;
  for (long col = blockIdx.x * (long)blockDim.x + threadIdx.x; col < N; col += (long)gridDim.x * blockDim.x) {
    const int a = (int)(col / (4 * chir)), s12 = (int)((col / chir) & 3), c = (int)(col % chir);
    double acc = 0.0;
    for (int l = 0; l < wl; ++l) {
      const double lv = L[((long)l * cbl + ap) * chil + a];
      if (lv == 0.0) continue;
      const double* w = W12 + (long)(l * 4 + s12) * (4 * wr) + t12 * wr;
      double inner = 0.0;
      for (int r = 0; r < wr; ++r) inner += w[r] * R[((long)r * chir + c) * cbr + cp];
      acc += lv * inner;
    }
    T[(long)row * N + col] = acc;
  }
}
// z_i <- z_i / sigma_i^2 for the singular values above rcond * sigma_1, 0 otherwise
__global__ void pinv_scale_kernel(double* __restrict__ z, const double* __restrict__ sigma, int k, double rcond) {
  const double smax = sigma[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    const double s = sigma[i];
    z[i] = (s > rcond * smax && s > 0.0) ? z[i] / (s * s) : 0.0;
  }
}
}  // namespace

bool pg_local_ok(size_t rows, size_t cols) { return rows >= 1 && cols >= 1 && rows * cols <= (size_t)DENSE_LEN_MAX * DENSE_LEN_MAX / 4; }

void pg_local_lstsq(Ctx* ctx, const double* L, const double* W12, const double* R, int cbl, int cbr, int chil, int chir, int wl, int wr,
                    const double* bT, double* x) {
  const int m = 4 * cbl * cbr;
  const long N = 4L * chil * chir;
  QTT_REQUIRE(pg_local_ok((size_t)m, (size_t)N), "b_linsolve: the dense local operator %d x %ld exceeds the cap", m, N);
  double* T = ctx->arena.alloc((size_t)m * N);
  const int kmax = m < N ? m : (int)N;
  double* sigma = ctx->arena.alloc(kmax);
  double* y = ctx->arena.alloc(N);
  double* z = ctx->arena.alloc(kmax);
  {
    int gx = (int)((N + 255) / 256);
    if (gx > 64) gx = 64;
    dense_op_rect_kernel<<<dim3(gx, m), 256, 0, ctx->stream>>>(L, W12, R, T, cbl, cbr, chil, chir, wl, wr);
    check_launch(ctx, "dense_op_rect");
  }
  // T = U S V^T: the rows of Wn are the right singular vectors of the non-zero singular values
  RowSvd sv = svd_rows(ctx, T, m, (int)N, 0.0, 0, 1, sigma, 0);
  GemmDesc g1;   // y = T^T b
  g1.A = T; g1.B = bT; g1.C = y; g1.transA = true;
  g1.M = (int)N; g1.N = 1; g1.K = m; g1.lda = N; g1.ldb = 1; g1.ldc = 1;
  gemm(ctx, g1);
  GemmDesc g2;   // z = V^T y
  g2.A = sv.Wn; g2.B = y; g2.C = z;
  g2.M = sv.k; g2.N = 1; g2.K = (int)N; g2.lda = N; g2.ldb = 1; g2.ldc = 1;
  gemm(ctx, g2);
  const double rcond = 2.220446049250313e-16 * (double)(m > N ? m : N);
  pinv_scale_kernel<<<1, 256, 0, ctx->stream>>>(z, sigma, sv.k, rcond);
  check_launch(ctx, "pinv_scale");
  GemmDesc g3;   // x = V z
  g3.A = sv.Wn; g3.B = z; g3.C = x; g3.transA = true;
  g3.M = (int)N; g3.N = 1; g3.K = sv.k; g3.lda = N; g3.ldb = 1; g3.ldc = 1;
  gemm(ctx, g3);
}

// QTT_DENSE_LEN_MAX (environment) lowers the cap, e.g. to bound the chi^4 memory of a batch of contexts
bool dense_local_ok(size_t len) {
  size_t cap = DENSE_LEN_MAX;
  if (const char* e = getenv("QTT_DENSE_LEN_MAX")) {
    const long v = atol(e);
    if (v >= 1 && (size_t)v < cap) cap = (size_t)v;
  }
  return len >= 1 && len <= cap;
}

size_t dense_scratch(size_t len) { return dense_layout(len).total; }

void dense_local_solve(Ctx* ctx, const double* L, const double* W12, const double* R, int chil, int chir, int wl, int wr, double a0,
                       double a1, const double* beff, double* x, double* scratch, double* stats_dev) {
  const size_t len = (size_t)4 * chil * chir;
  QTT_REQUIRE(dense_local_ok(len), "dense local solver: local dimension %zu exceeds the cap %d (the matrix is (4 chi^2)^2)", len, DENSE_LEN_MAX);
  const DenseLayout d = dense_layout(len);
  build_op(ctx, L, W12, R, scratch + d.tt, d.ld, chil, chir, wl, wr, a0, a1, beff);
  dense_qr_rows(ctx, scratch, d, (int)len, (int)len + 1);
  TriArgs t;
  t.Tt = scratch + d.tt; t.ld = (long)d.ld; t.n = (int)len; t.y = scratch + d.tt + len * d.ld; t.eigen = 0; t.maxit = 0; t.stats = stats_dev;
  launch_tri(ctx, t);
  QTT_CUDA(cudaMemcpyAsync(x, t.y, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));
}

void dense_local_eigen(Ctx* ctx, const double* L, const double* W12, const double* R, int chil, int chir, int wl, int wr, double target,
                       double* x_inout, double* scratch, int maxit, double* stats_dev) {
  const size_t len = (size_t)4 * chil * chir;
  QTT_REQUIRE(dense_local_ok(len), "dense local solver: local dimension %zu exceeds the cap %d (the matrix is (4 chi^2)^2)", len, DENSE_LEN_MAX);
  const DenseLayout d = dense_layout(len);
  build_op(ctx, L, W12, R, scratch + d.tt, d.ld, chil, chir, wl, wr, -target, 1.0, nullptr);
  dense_qr_rows(ctx, scratch, d, (int)len, (int)len);
  TriArgs t;
  t.Tt = scratch + d.tt; t.ld = (long)d.ld; t.n = (int)len; t.y = x_inout; t.eigen = 1; t.maxit = maxit > 0 ? maxit : 60; t.stats = stats_dev;
  launch_tri(ctx, t);
}

}  // namespace qtt

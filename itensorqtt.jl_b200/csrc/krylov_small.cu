// Local GMRES for SMALL bonds as ONE single-CTA kernel (whole solve: matvecs, CGS2, Givens, restarts).
//
// On the edge bonds of a chain (chi_l * chi_r <= 256: len = 4 chi_l chi_r <= 1024) and on the reference's own
// CPU-runnable configuration (Airy, n = 20, maxdim = 32: the solution has bond dimension <= 6) every kernel of the
// multi-launch path is launch-latency bound: ~60 us per Arnoldi step for microseconds of work (measured: 1.9 ms per edge
// bond, 16 of the 58 bond updates of the n = 30 benchmark sweep).  Here the operator (L, W12, R), the current vector and
// the matvec intermediates live in shared memory, the Krylov basis stays in L2, and the step costs a few microseconds.
// Semantics are those of gmres_local() in sweep.cu (KrylovKit GMRES restated, fixed-work mode included); sums are
// accumulated in a different order than on the DMMA path, so results agree to rounding, not bitwise.
#include "common.cuh"
#include "krylov_scalars.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace qtt {
namespace {

using namespace kscal;

constexpr int ST = 512;  // threads
constexpr int SLD = 32;  // leading dimension of the Hessenberg/R matrix (krylovdim <= 31)
constexpr int S_DOUBLES = OFF_R + SLD * SLD;

struct SmallArgs {
  int vstash;          // basis vectors kept in shared memory
  const double* L;     // [wl][chil][chil]
  const double* W12;   // [4 wl][4 wr]
  const double* R;     // [wr][chir][chir]
  const double* beff;  // [len]
  double* x;           // [len] in/out
  double* V;           // [(kd + 1)][len] workspace (global, L2 resident)
  int chil, chir, wl, wr;
  double a0, a1, tol;
  int kd, maxiter, fixed;
  double* stats;       // out: [0] matvecs, [1] resid, [2] converged, [3] last S_BETA
};

using dev::warp_sum;

struct Smem {
  double* L; double* R; double* W; double* v; double* w; double* T1; double* T2; double* S; double* h; double* red;
};

// w = P v = L . W12 . R . v   (all operands in shared memory)
__device__ void small_matvec(const Smem& m, int chil, int chir, int wl, int wr) {
  const int tid = threadIdx.x;
  const int len = 4 * chil * chir;
  const int c4 = 4 * chir;
  // stage 1: T1[w][a'][(s c)] = sum_a L[w][a'][a] v[a][(s c)]
  for (int o = tid; o < wl * len; o += ST) {
    const int sc = o % c4, wa = o / c4;  // wa = w * chil + a'
    const double* lrow = m.L + (size_t)wa * chil;
    double acc = 0.0;
    for (int a = 0; a < chil; ++a) acc += lrow[a] * m.v[a * c4 + sc];
    m.T1[o] = acc;
  }
  __syncthreads();
  // stage 2: T2[a'][o][c] = sum_{(w s)} W12[(w s)][o] T1[w][a'][s][c],  o = (t12, wr)
  const int nout = 4 * wr;
  for (int idx = tid; idx < chil * nout * chir; idx += ST) {
    const int c = idx % chir, rest = idx / chir;
    const int o = rest % nout, ap = rest / nout;
    double acc = 0.0;
    for (int w = 0; w < wl; ++w)
#pragma unroll
      for (int s = 0; s < 4; ++s) acc += m.W[(w * 4 + s) * nout + o] * m.T1[((w * chil + ap) * 4 + s) * chir + c];
    m.T2[idx] = acc;
  }
  __syncthreads();
  // stage 3: w[a'][t12][c'] = sum_{wr, c} T2[a'][(t12 wr)][c] R[wr][c][c']
  for (int o = tid; o < len; o += ST) {
    const int cp = o % chir, at = o / chir;  // at = a' * 4 + t12
    double acc = 0.0;
    for (int r = 0; r < wr; ++r) {
      const double* t2 = m.T2 + ((size_t)at * wr + r) * chir;
      const double* rr = m.R + (size_t)r * chir * chir + cp;
      for (int c = 0; c < chir; ++c) acc += t2[c] * rr[c * chir];
    }
    m.w[o] = acc;
  }
  __syncthreads();
}

__device__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ST / 32; ++k) s += red[k];
  __syncthreads();
  return s;
}

// The Krylov basis: the first `ns` vectors live in shared memory (as many as fit next to the operator), the others in
// global memory (L2).  A single CTA is latency bound on the L2 part, so its loads are issued in batches of 16.
struct Basis {
  double* g;   // global [(kd + 1)][len]; slots < ns unused
  double* s;   // shared [ns][len]
  int ns, len;
  __device__ __forceinline__ double ld(int i, int e) const { return i < ns ? s[(size_t)i * len + e] : __ldcg(g + (size_t)i * len + e); }
  __device__ __forceinline__ void st(int i, int e, double v) const {
    if (i < ns) s[(size_t)i * len + e] = v;
    else g[(size_t)i * len + e] = v;
  }
};

// sum_{i<k} coef[i] V[i][e]
__device__ __forceinline__ double comb(const Basis& B, int k, int e, const double* coef, double sign) {
  double x = 0.0;
  const int ks = k < B.ns ? k : B.ns;
  for (int i = 0; i < ks; ++i) x += coef[i] * B.s[(size_t)i * B.len + e];
  for (int i0 = ks; i0 < k; i0 += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (i0 + u < k) ? __ldcg(B.g + (size_t)(i0 + u) * B.len + e) : 0.0;
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (i0 + u < k) x += coef[i0 + u] * v[u];
  }
  return sign * x;
}

// h[i] = V[i] . w for i < k; then w -= V h   (one classical Gram-Schmidt pass)
__device__ void cgs_pass(const Basis& B, int k, double* w, double* h) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = B.len;
  for (int i = warp; i < k; i += ST / 32) {
    double acc = 0.0;
    if (i < B.ns) {
      const double* vi = B.s + (size_t)i * len;
      for (int e = lane; e < len; e += 32) acc += vi[e] * w[e];
    } else {
      const double* vi = B.g + (size_t)i * len;
      for (int e0 = 0; e0 < len; e0 += 512) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int e = e0 + lane + 32 * u;
          v[u] = e < len ? __ldcg(vi + e) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int e = e0 + lane + 32 * u;
          if (e < len) acc += v[u] * w[e];
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) h[i] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < len; e += ST) w[e] += comb(B, k, e, h, -1.0);
  __syncthreads();
}

__global__ void __launch_bounds__(ST) gmres_small_kernel(SmallArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const int chil = a.chil, chir = a.chir, wl = a.wl, wr = a.wr;
  const int len = 4 * chil * chir;
  Smem m;
  double* p = sm;
  m.L = p; p += wl * chil * chil;
  m.R = p; p += wr * chir * chir;
  m.W = p; p += 16 * wl * wr;
  m.v = p; p += len;
  m.w = p; p += len;
  m.T1 = p; p += (size_t)wl * len;
  m.T2 = p; p += (size_t)wr * len;
  m.S = p; p += S_DOUBLES;
  m.h = p; p += 128;
  m.red = p; p += 32;
  Basis B;
  B.g = a.V; B.s = p; B.ns = a.vstash; B.len = len;
  for (int i = tid; i < wl * chil * chil; i += ST) m.L[i] = a.L[i];
  for (int i = tid; i < wr * chir * chir; i += ST) m.R[i] = a.R[i];
  for (int i = tid; i < 16 * wl * wr; i += ST) m.W[i] = a.W12[i];
  for (int i = tid; i < S_DOUBLES; i += ST) m.S[i] = 0.0;
  for (int i = tid; i < len; i += ST) m.v[i] = a.x[i];
  __syncthreads();
  double* S = m.S;
  const int r0 = a.kd;  // the slot of v_kd doubles as the residual buffer between cycles
  const bool fixed = a.fixed != 0;
  const double tol = fixed ? -1.0 : a.tol;
  double matvecs = 0.0, resid = 0.0, converged = 0.0;

  // r = beff - a0 x - a1 P x
  small_matvec(m, chil, chir, wl, wr);
  matvecs += 1.0;
  double acc = 0.0;
  for (int e = tid; e < len; e += ST) {
    const double r = a.beff[e] - a.a0 * m.v[e] - a.a1 * m.w[e];
    B.st(r0, e, r);
    acc += r * r;
  }
  double r0n2 = block_sum(acc, m.red);
  if (tid == 0) S[S_R0N2] = r0n2;
  __syncthreads();
  bool finished = false;
  if (!fixed) {
    resid = sqrt(r0n2);
    if (resid < a.tol) { converged = 1.0; finished = true; }
  }
  for (int it = 0; it < a.maxiter && !finished; ++it) {
    if (tid == 0) { S[S_DONE] = 0.0; S[S_K] = 0.0; }
    {
      const double d = sqrt(S[S_R0N2]);
      const double inv = d != 0.0 ? 1.0 / d : 0.0;
      __syncthreads();
      for (int e = tid; e < len; e += ST) {
        const double v0 = B.ld(r0, e) * inv;
        B.st(0, e, v0);
        m.v[e] = v0;
      }
    }
    __syncthreads();
    int ksteps = 0;
    for (int k = 1; k <= a.kd; ++k) {
      small_matvec(m, chil, chir, wl, wr);
      matvecs += 1.0;
      cgs_pass(B, k, m.w, S + OFF_H1);
      cgs_pass(B, k, m.w, S + OFF_H2);
      double na = 0.0;
      for (int e = tid; e < len; e += ST) na += m.w[e] * m.w[e];
      const double n2 = block_sum(na, m.red);
      if (tid == 0) {
        S[S_NRES2] = n2;
        gmres_step_scalars(S, k, a.a0, a.a1, tol, SLD, nullptr);
      }
      __syncthreads();
      const bool dead = S[S_DONE] != 0.0;
      const double nrm = sqrt(n2);
      const double inv = (nrm != 0.0 && !dead) ? 1.0 / nrm : 0.0;
      for (int e = tid; e < len; e += ST) {
        const double vv = m.w[e] * inv;
        B.st(k, e, vv);
        m.v[e] = vv;
      }
      ksteps = k;
      __syncthreads();
      if (!fixed && dead) break;
    }
    // fixed mode runs all kd steps; S_K holds the number of useful ones (yk is zero-filled beyond it)
    const int kfin = fixed ? a.kd : (int)S[S_K];
    if (tid == 0) gmres_solve_scalars(S, fixed ? -a.kd : kfin, SLD);
    __syncthreads();
    for (int e = tid; e < len; e += ST) a.x[e] += comb(B, kfin, e, S + OFF_YK, 1.0);
    __syncthreads();
    resid = S[S_BETA];
    if (fixed) break;
    if (resid > a.tol) {
      // restart residual as the explicit combination of the basis (no extra operator application)
      for (int e = tid; e < len; e += ST) m.w[e] = comb(B, kfin + 1, e, S + OFF_COEF, 1.0);
      __syncthreads();
      for (int e = tid; e < len; e += ST) B.st(r0, e, m.w[e]);
      __syncthreads();
    } else {
      for (int e = tid; e < len; e += ST) m.v[e] = a.x[e];
      __syncthreads();
      small_matvec(m, chil, chir, wl, wr);
      matvecs += 1.0;
      double ra = 0.0;
      for (int e = tid; e < len; e += ST) {
        const double r = a.beff[e] - a.a0 * m.v[e] - a.a1 * m.w[e];
        B.st(r0, e, r);
        ra += r * r;
      }
      const double n2 = block_sum(ra, m.red);
      if (tid == 0) S[S_R0N2] = n2;
      __syncthreads();
      resid = sqrt(n2);
      if (resid < a.tol) { converged = 1.0; finished = true; }
    }
    (void)ksteps;
  }
  if (tid == 0) {
    a.stats[0] = matvecs;
    a.stats[1] = resid;
    a.stats[2] = converged;
    a.stats[3] = S[S_BETA];
  }
}

constexpr size_t SMALL_SMEM_BUDGET = 224 * 1024;  // dynamic shared memory of the single CTA (227 KB minus static)

size_t small_smem_doubles(int chil, int chir, int wl, int wr) {
  const size_t len = (size_t)4 * chil * chir;
  return (size_t)wl * chil * chil + (size_t)wr * chir * chir + (size_t)16 * wl * wr + 2 * len + (size_t)(wl + wr) * len + S_DOUBLES + 128 + 32;
}

}  // namespace

bool gmres_small_ok(int chil, int chir, int wl, int wr, int kd) {
  static const bool off = getenv("QTT_SMALL") != nullptr && atoi(getenv("QTT_SMALL")) == 0;
  if (off) return false;
  const size_t len = (size_t)4 * chil * chir;
  return kd >= 1 && kd < SLD && len <= 1024 && small_smem_doubles(chil, chir, wl, wr) * sizeof(double) <= 200 * 1024;
}

// V: workspace of (kd + 1) * len doubles; stats_dev: 4 doubles (matvecs, resid, converged, beta)
void gmres_small(Ctx* ctx, const double* L, const double* W12, const double* R, const double* beff, double* x, int chil, int chir, int wl,
                 int wr, double a0, double a1, double tol, int kd, int maxiter, bool fixed, double* V, double* stats_dev) {
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(gmres_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_BUDGET));
  });
  SmallArgs a;
  a.L = L; a.W12 = W12; a.R = R; a.beff = beff; a.x = x; a.V = V;
  a.chil = chil; a.chir = chir; a.wl = wl; a.wr = wr;
  a.a0 = a0; a.a1 = a1; a.tol = tol; a.kd = kd; a.maxiter = maxiter; a.fixed = fixed ? 1 : 0;
  a.stats = stats_dev;
  // as many basis vectors as fit stay in shared memory (all kd + 1 of them for len <= 512 at krylovdim 30)
  const size_t base = small_smem_doubles(chil, chir, wl, wr);
  const size_t len = (size_t)4 * chil * chir;
  size_t ns = (SMALL_SMEM_BUDGET / sizeof(double) - base) / len;
  if (ns > (size_t)kd + 1) ns = (size_t)kd + 1;
  a.vstash = (int)ns;
  const size_t smem = (base + ns * len) * sizeof(double);
  gmres_small_kernel<<<1, ST, smem, ctx->stream>>>(a);
  check_launch(ctx, "gmres_small");
}

}  // namespace qtt

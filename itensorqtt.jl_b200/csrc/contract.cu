// Hot-path contractions of the two-site sweep: effective-operator matvec and environment updates.
//
// Layouts are chosen so that every large contraction is a plain row-major GEMM on the DMMA core with NO
// index permute (the reference pays a permutedims + allocation per ITensor `*`):
//   L [w][a'][a]   R [w][c][c']   x [a][s][c]   v [a][s1][s2][c]      (last index fastest)
// and the skinny contractions over the MPO indices (K = 2w..4w, not tensor-core shaped) are fused into
// one bandwidth-bound kernel each, which also performs the layout change the next GEMM needs.
#include "common.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace qtt {
namespace {

// W12[(wl s1 s2)][(t1 t2 wr)] = sum_wm W1[wl][wm][s1][t1] W2[wm][wr][s2][t2]
__global__ void build_w12_kernel(const double* __restrict__ W1, const double* __restrict__ W2, double* __restrict__ W12, int wl,
                                 int wm, int wr) {
  dev::pdl_wait();
  int total = 16 * wl * wr;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
    int col = o % (4 * wr), row = o / (4 * wr);
    int r = col % wr, t2 = (col / wr) & 1, t1 = col / (2 * wr);
    int s2 = row & 1, s1 = (row >> 1) & 1, l = row >> 2;
    double acc = 0.0;
    for (int m = 0; m < wm; ++m) acc += W1[((l * wm + m) * 2 + s1) * 2 + t1] * W2[((m * wr + r) * 2 + s2) * 2 + t2];
    W12[o] = acc;
  }
}

// T2[a'][t1 t2][wr][c] = sum_{wl, s1 s2} W12[(wl s12)][(t12 wr)] T1[wl][a'][s12][c]
// one thread per (a', c); c is the coalesced index on both sides.
template <int WLMAX>
__global__ void apply_w12_kernel(const double* __restrict__ T1, const double* __restrict__ W12, double* __restrict__ T2, int chil,
                                 int chir, int wl, int wr) {
  extern __shared__ double w12s[];
  const int nin = 4 * wl, nout = 4 * wr;
  for (int i = threadIdx.x; i < nin * nout; i += blockDim.x) w12s[i] = W12[i];
  __syncthreads();
  const int ap = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= chir) return;
  double in[4 * WLMAX];
#pragma unroll
  for (int l = 0; l < WLMAX; ++l) {
    if (l < wl) {
#pragma unroll
      for (int s = 0; s < 4; ++s) in[l * 4 + s] = T1[(((long)l * chil + ap) * 4 + s) * chir + c];
    }
  }
  for (int o = 0; o < nout; ++o) {  // o = (t12, wr)
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4 * WLMAX; ++i)
      if (i < nin) acc += w12s[i * nout + o] * in[i];
    T2[((long)ap * nout + o) * chir + c] = acc;
  }
}

// generic (any wl): no register caching
__global__ void apply_w12_generic_kernel(const double* __restrict__ T1, const double* __restrict__ W12, double* __restrict__ T2,
                                         int chil, int chir, int wl, int wr) {
  dev::pdl_wait();
  const int nin = 4 * wl, nout = 4 * wr;
  const int ap = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= chir) return;
  for (int o = 0; o < nout; ++o) {
    double acc = 0.0;
    for (int i = 0; i < nin; ++i) {
      int l = i >> 2, s = i & 3;
      acc += W12[i * nout + o] * T1[(((long)l * chil + ap) * 4 + s) * chir + c];
    }
    T2[((long)ap * nout + o) * chir + c] = acc;
  }
}

// out[i] = sum_b P[b][i]
__global__ void sum_partials_kernel(const double* __restrict__ P, double* __restrict__ out, long n, int nb) {
  dev::pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double acc = P[i];
    for (int b = 1; b < nb; ++b) acc += P[(long)b * n + i];
    out[i] = acc;
  }
}

// U[w2][a'][t][b] = sum_{w, s} W[w][w2][s][t] T[w][a'][s][b]      (left environment, middle step)
__global__ void env_left_mid_kernel(const double* __restrict__ T, const double* __restrict__ W, double* __restrict__ U, int chia,
                                    int chib, int w, int w2) {
  dev::pdl_wait();
  extern __shared__ double ws[];
  for (int i = threadIdx.x; i < w * w2 * 4; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const int ap = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= chib) return;
  for (int o = 0; o < w2; ++o) {
    double acc0 = 0.0, acc1 = 0.0;
    for (int l = 0; l < w; ++l) {
      double t0 = T[(((long)l * chia + ap) * 2 + 0) * chib + b];
      double t1 = T[(((long)l * chia + ap) * 2 + 1) * chib + b];
      const double* wp = ws + ((l * w2 + o) * 2) * 2;  // [s][t]
      acc0 += wp[0] * t0 + wp[2] * t1;
      acc1 += wp[1] * t0 + wp[3] * t1;
    }
    U[(((long)o * chia + ap) * 2 + 0) * chib + b] = acc0;
    U[(((long)o * chia + ap) * 2 + 1) * chib + b] = acc1;
  }
}

// U[w][s][c][a'] = sum_{w2, t} W[w][w2][s][t] T[w2][c][a'][t]      (right environment, middle step)
__global__ void env_right_mid_kernel(const double* __restrict__ T, const double* __restrict__ W, double* __restrict__ U, int chic,
                                     int chia, int w, int w2) {
  dev::pdl_wait();
  extern __shared__ double ws[];
  for (int i = threadIdx.x; i < w * w2 * 4; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const int c = blockIdx.y;
  const int ap = blockIdx.x * blockDim.x + threadIdx.x;
  if (ap >= chia) return;
  for (int o = 0; o < w; ++o) {
    double acc0 = 0.0, acc1 = 0.0;
    for (int l = 0; l < w2; ++l) {
      const double* tp = T + (((long)l * chic + c) * chia + ap) * 2;
      double t0 = tp[0], t1 = tp[1];
      const double* wp = ws + ((o * w2 + l) * 2) * 2;  // [s][t]
      acc0 += wp[0] * t0 + wp[1] * t1;
      acc1 += wp[2] * t0 + wp[3] * t1;
    }
    U[(((long)o * 2 + 0) * chic + c) * chia + ap] = acc0;
    U[(((long)o * 2 + 1) * chic + c) * chia + ap] = acc1;
  }
}

}  // namespace

void build_w12(Ctx* ctx, const double* W1, const double* W2, double* W12, int wl, int wm, int wr) {
  int total = 16 * wl * wr;
  launch_pdl(ctx, "build_w12", build_w12_kernel, dim3((total + 127) / 128), dim3(128), 0, W1, W2, W12, wl, wm, wr);
}

size_t matvec2_scratch(int chil, int chir, int wl, int wr) {
  size_t base = (size_t)4 * chil * chir;
  size_t wmax = wl > wr ? wl : wr;
  return base * wmax + base * wr + 64;
}

void matvec2(Ctx* ctx, const double* L, const double* W12, const double* R, const double* v, double* out, int chil, int chir,
             int wl, int wr, double* scratch) {
  // fused path (matvec_fused.cu): two kernels, no T1 intermediate; QTT_MATVEC=unfused keeps the 4-launch path (A/B runs)
  static const bool unfused = getenv("QTT_MATVEC") != nullptr && std::string(getenv("QTT_MATVEC")) == "unfused";
  if (!unfused && matvec2_fused_ok(wl, wr)) {
    matvec2_fused(ctx, L, W12, R, v, out, chil, chir, wl, wr, scratch);
    return;
  }
  const size_t base = (size_t)4 * chil * chir;
  const size_t wmax = wl > wr ? wl : wr;
  double* T1 = scratch;                // [wl][a'][s1 s2][c]; later reused for the split-K partials
  double* T2 = scratch + base * wmax;  // [a'][t1 t2][wr][c]
  // stage 1: T1[(wl a')][(s12 c)] = L[(wl a')][a] . v[a][(s12 c)]
  GemmDesc g1;
  g1.A = L; g1.B = v; g1.C = T1;
  g1.M = wl * chil; g1.N = 4 * chir; g1.K = chil;
  g1.lda = chil; g1.ldb = 4 * chir; g1.ldc = 4 * chir;
  gemm(ctx, g1);
  // stages 2+3 fused: apply W12 = W1.W2 over (wl, s1, s2) -> (t1, t2, wr)
  {
    const int threads = chir >= 128 ? 128 : (chir >= 64 ? 64 : 32);
    dim3 grid((chir + threads - 1) / threads, chil);
    size_t smem = (size_t)16 * wl * wr * sizeof(double);
    if (wl <= 4) apply_w12_kernel<4><<<grid, threads, smem, ctx->stream>>>(T1, W12, T2, chil, chir, wl, wr);
    else if (wl <= 8 && smem <= 48 * 1024) apply_w12_kernel<8><<<grid, threads, smem, ctx->stream>>>(T1, W12, T2, chil, chir, wl, wr);
    else launch_pdl(ctx, "apply_w12", apply_w12_generic_kernel, dim3(grid), dim3(threads), 0, T1, W12, T2, chil, chir, wl, wr);
  }
  // stage 4: out[(a' t12)][c'] = T2[(a' t12)][(wr c)] . R[(wr c)][c'].  The output is small (4 chil x chir)
  // and K is long (wr chir): when it cannot fill the SMs it is split over wr into partial products.
  const long tiles64 = (long)((4 * chil + 63) / 64) * ((chir + 63) / 64);
  const bool split = wr > 1 && tiles64 < 2L * ctx->sms;
  if (!split) {
    GemmDesc g4;
    g4.A = T2; g4.B = R; g4.C = out;
    g4.M = 4 * chil; g4.N = chir; g4.K = wr * chir;
    g4.lda = (long)wr * chir; g4.ldb = chir; g4.ldc = chir;
    gemm(ctx, g4);
  } else {
    double* P = T1;
    GemmDesc g4;
    g4.A = T2; g4.B = R; g4.C = P;
    g4.M = 4 * chil; g4.N = chir; g4.K = chir;
    g4.lda = (long)wr * chir; g4.ldb = chir; g4.ldc = chir;
    g4.batch = wr;
    g4.sA = chir; g4.sB = (long)chir * chir; g4.sC = (long)base;
    gemm(ctx, g4);
    long nblk = ((long)base + 255) / 256;
    if (nblk > 8L * ctx->sms) nblk = 8L * ctx->sms;
    launch_pdl(ctx, "sum_partials", sum_partials_kernel, dim3((unsigned)nblk), dim3(256), 0, P, out, (long)base, wr);
  }
}

size_t env_scratch(int chia, int chib, int w, int w2) {
  size_t m = (size_t)2 * chia * chib;
  return m * w + m * w2 + 64;
}

// Lnew[w2][b'][b] = sum L[w][a'][a] x[a][s][b] W[w][w2][s][t] y[a'][t][b']   (y = bra chain; y = x for Galerkin)
void env_left(Ctx* ctx, const double* L, const double* x, const double* W, const double* y, double* Lnew, int chia, int chib,
              int chia_y, int chib_y, int w, int w2, double* scratch) {
  double* T = scratch;                                  // [w][a'][s][b]
  double* U = scratch + (size_t)2 * w * chia_y * chib;  // [w2][a'][t][b]
  GemmDesc g1;
  g1.A = L; g1.B = x; g1.C = T;
  g1.M = w * chia_y; g1.N = 2 * chib; g1.K = chia;
  g1.lda = chia; g1.ldb = 2 * chib; g1.ldc = 2 * chib;
  gemm(ctx, g1);
  {
    const int threads = chib >= 128 ? 128 : (chib >= 64 ? 64 : 32);
    dim3 grid((chib + threads - 1) / threads, chia_y);
    launch_pdl(ctx, "env_left_mid", env_left_mid_kernel, dim3(grid), dim3(threads), (size_t)w * w2 * 4 * sizeof(double), T, W, U, chia_y, chib, w, w2);
  }
  GemmDesc g3;
  g3.A = y; g3.B = U; g3.C = Lnew;
  g3.transA = true;
  g3.M = chib_y; g3.N = chib; g3.K = 2 * chia_y;
  g3.lda = chib_y; g3.ldb = chib; g3.ldc = chib;
  g3.batch = w2;
  g3.sA = 0; g3.sB = (long)2 * chia_y * chib; g3.sC = (long)chib_y * chib;
  gemm(ctx, g3);
}

// Rnew[w][a][a'] = sum x[a][s][c] W[w][w2][s][t] y[a'][t][c'] R[w2][c][c']
void env_right(Ctx* ctx, const double* R, const double* x, const double* W, const double* y, double* Rnew, int chia, int chic,
               int chia_y, int chic_y, int w, int w2, double* scratch) {
  double* T = scratch;                                  // [w2][c][a'][t]
  double* U = scratch + (size_t)2 * w2 * chic * chia_y;  // [w][s][c][a']
  GemmDesc g1;
  g1.A = R; g1.B = y; g1.C = T;
  g1.transB = true;
  g1.M = w2 * chic; g1.N = 2 * chia_y; g1.K = chic_y;
  g1.lda = chic_y; g1.ldb = chic_y; g1.ldc = 2 * chia_y;
  gemm(ctx, g1);
  {
    const int threads = chia_y >= 128 ? 128 : (chia_y >= 64 ? 64 : 32);
    dim3 grid((chia_y + threads - 1) / threads, chic);
    launch_pdl(ctx, "env_right_mid", env_right_mid_kernel, dim3(grid), dim3(threads), (size_t)w * w2 * 4 * sizeof(double), T, W, U, chic, chia_y, w, w2);
  }
  GemmDesc g3;
  g3.A = x; g3.B = U; g3.C = Rnew;
  g3.M = chia; g3.N = chia_y; g3.K = 2 * chic;
  g3.lda = 2 * chic; g3.ldb = chia_y; g3.ldc = chia_y;
  g3.batch = w;
  g3.sA = 0; g3.sB = (long)2 * chic * chia_y; g3.sC = (long)chia * chia_y;
  gemm(ctx, g3);
}

}  // namespace qtt

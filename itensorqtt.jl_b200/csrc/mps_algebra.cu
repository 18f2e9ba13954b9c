// Device-resident MPS algebra around the apply path: truncated sum, dense expansion and point sampling.
//
//   mps_add      ITensorMPS `+(x, y; cutoff, maxdim)` as called at reference src/fourier_interpolation.jl:8 (and the
//                `+(...; cutoff)` of src/function_to_mps.jl:123, src/laplacian.jl:106): the block-diagonal direct sum is
//                assembled on the device and compressed by the density-matrix pass of apply.cu with an identity operator
//                (the density matrices of the direct sum ARE those of the sum state, so this is the sum truncated at
//                `cutoff` in the NDTensors.truncate! sense).
//   mps_to_vec   `mps_to_discrete_function(psi)` (reference src/function_to_mps.jl:140-145, 1-D): all 2^n values, site 1 =
//                most significant bit.  The chain is cut at a bond m near the middle: left part [2^m][chi_m] and right
//                part [chi_m][2^(n-m)] are built by GEMM chains on cores as they lie in memory (no permute), the
//                HBM-bound final product writes the 2^n values once.
//   mps_sample   `project_bits(psi, bits)` with every site projected (reference src/project_bits.jl:1-15, scalar return
//                :11-13), batched over many grid indices: one CTA walks one index through the chain with the running
//                row vector in shared memory; reads of the selected core slices are coalesced along the right bond.
#include "common.cuh"
#include <algorithm>

namespace qtt {
namespace {

// dst[r * R + col] (+)= scale * src[r * c + col], r < rows, col < c
__global__ void place_block_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t rows, int64_t c, int64_t R,
                                   double scale, int accumulate) {
  const size_t total = (size_t)rows * (size_t)c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (size_t)c, col = i % (size_t)c;
    const double v = scale * src[i];
    double* d = dst + r * (size_t)R + col;
    *d = accumulate ? *d + v : v;
  }
}

void place_block(Ctx* ctx, const double* src, double* dst, int64_t rows, int64_t c, int64_t R, double scale, bool accumulate) {
  const size_t total = (size_t)rows * (size_t)c;
  unsigned g = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->sms * 8);
  place_block_kernel<<<g ? g : 1, 256, 0, ctx->stream>>>(src, dst, rows, c, R, scale, accumulate ? 1 : 0);
  check_launch(ctx, "place_block");
}

struct PoolBuf {  // stream-ordered scratch that returns to the caching pool
  double* p = nullptr;
  cudaStream_t st = nullptr;
  PoolBuf() = default;
  PoolBuf(size_t n_elems, cudaStream_t s) : st(s) { p = static_cast<double*>(pool_alloc((n_elems < 2 ? 2 : n_elems) * sizeof(double), s)); }
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
  PoolBuf(PoolBuf&& o) noexcept : p(o.p), st(o.st) { o.p = nullptr; }
  PoolBuf& operator=(PoolBuf&& o) noexcept {
    if (this != &o) {
      reset();
      p = o.p; st = o.st; o.p = nullptr;
    }
    return *this;
  }
  void reset() {
    if (p) pool_free(p, st);
    p = nullptr;
  }
  ~PoolBuf() { reset(); }
};

void gemm_rm(Ctx* ctx, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, double alpha, double beta) {
  QTT_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "mps_to_vec: GEMM extent out of range");
  GemmDesc g;
  g.A = A; g.B = B; g.C = C;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.lda = (long)K; g.ldb = (long)N; g.ldc = (long)N;
  g.alpha = alpha; g.beta = beta;
  gemm(ctx, g);
}

// planar complex (or real when the imaginary planes are null) C = A . B, all row-major
void gemm_planar(Ctx* ctx, const double* Ar, const double* Ai, const double* Br, const double* Bi, double* Cr, double* Ci, int64_t M,
                 int64_t N, int64_t K) {
  gemm_rm(ctx, Ar, Br, Cr, M, N, K, 1.0, 0.0);
  if (!Ci) return;
  if (Ai && Bi) gemm_rm(ctx, Ai, Bi, Cr, M, N, K, -1.0, 1.0);
  bool first = true;
  if (Bi) { gemm_rm(ctx, Ar, Bi, Ci, M, N, K, 1.0, 0.0); first = false; }
  if (Ai) { gemm_rm(ctx, Ai, Br, Ci, M, N, K, 1.0, first ? 0.0 : 1.0); first = false; }
  if (first) QTT_CUDA(cudaMemsetAsync(Ci, 0, (size_t)M * (size_t)N * sizeof(double), ctx->stream));
}

constexpr int SAMPLE_MAX_SITES = 64;
struct SampleArgs {
  const double* core[SAMPLE_MAX_SITES];
  int chi[SAMPLE_MAX_SITES + 1];
  int n;
  int chimax;
  const int64_t* idx;
  int64_t count;
  double* out;  // [count] real plane, then [count] imaginary plane when complex
};

template <bool CPLX>
__global__ void __launch_bounds__(256) sample_kernel(SampleArgs a) {
  extern __shared__ double sv[];  // two row vectors of chimax (x2 planes when complex)
  const int planes = CPLX ? 2 : 1;
  double* buf0 = sv;
  double* buf1 = sv + (size_t)planes * a.chimax;
  for (int64_t s = blockIdx.x; s < a.count; s += gridDim.x) {
    const int64_t gi = a.idx[s];
    double* cur = buf0;
    double* nxt = buf1;
    __syncthreads();
    if (threadIdx.x == 0) {
      cur[0] = 1.0;
      if (CPLX) cur[a.chimax] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < a.n; ++j) {
      const int ca = a.chi[j], cc = a.chi[j + 1];
      const int bit = (int)((gi >> (a.n - 1 - j)) & 1);  // site 1 = most significant bit
      const double* re = a.core[j] + (size_t)bit * cc;   // [a][s][c]: element (aa, bit, c) at (aa * 2 + bit) * cc + c
      const double* im = CPLX ? re + (size_t)ca * 2 * cc : nullptr;
      for (int c = threadIdx.x; c < cc; c += blockDim.x) {
        double o_r = 0.0, o_i = 0.0;
        for (int aa = 0; aa < ca; ++aa) {
          const double xr = re[(size_t)aa * 2 * cc + c];
          const double vr = cur[aa];
          if (CPLX) {
            const double xi = im[(size_t)aa * 2 * cc + c];
            const double vi = cur[a.chimax + aa];
            o_r += vr * xr - vi * xi;
            o_i += vr * xi + vi * xr;
          } else {
            o_r += vr * xr;
          }
        }
        nxt[c] = o_r;
        if (CPLX) nxt[a.chimax + c] = o_i;
      }
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0) {
      a.out[s] = cur[0];
      if (CPLX) a.out[a.count + s] = cur[a.chimax];
    }
  }
}

}  // namespace

// out = alpha x + beta y, truncated (cutoff / maxdim / mindim as in apply)
void mps_add(Ctx* ctx, const Mps& x, const Mps& y, double alpha, double beta, double cutoff, int maxdim, int mindim, Mps& out) {
  const int n = x.n;
  QTT_REQUIRE(y.n == n && n >= 1, "mps_add: chain lengths differ (%d vs %d)", x.n, y.n);
  QTT_REQUIRE(x.chi[0] == 1 && x.chi[n] == 1 && y.chi[0] == 1 && y.chi[n] == 1, "mps_add: boundary bond dimensions must be 1");
  const bool cplx = x.dtype == QTT_C128 || y.dtype == QTT_C128;
  const Mps* in[2] = {&x, &y};
  const double coef[2] = {alpha, beta};
  Mps sum;
  sum.ctx = ctx; sum.n = n; sum.dtype = cplx ? QTT_C128 : QTT_F64;
  sum.chi.assign(n + 1, 1);
  for (int j = 1; j < n; ++j) sum.chi[j] = x.chi[j] + y.chi[j];
  sum.core.resize(n);
  for (int j = 0; j < n; ++j) {
    const int64_t Lj = sum.chi[j], Rj = sum.chi[j + 1];
    const size_t plane = (size_t)Lj * 2 * Rj;
    dbuf_reserve(sum.core[j], plane * (cplx ? 2 : 1));
    QTT_CUDA(cudaMemsetAsync(sum.core[j].p, 0, plane * (cplx ? 2 : 1) * sizeof(double), ctx->stream));
    int64_t l0 = 0, r0 = 0;
    for (int i = 0; i < 2; ++i) {
      const Mps& p = *in[i];
      const int64_t a = p.chi[j], c = p.chi[j + 1];
      const size_t sp = (size_t)a * 2 * c;
      const int64_t lo = (j == 0) ? 0 : l0, ro = (j == n - 1) ? 0 : r0;
      const double scale = (j == 0) ? coef[i] : 1.0;
      double* dst = sum.core[j].p + (size_t)lo * 2 * Rj + ro;
      // n == 1 (and only then) both blocks land on the same entries: accumulate the second
      const bool acc = (n == 1 && i == 1);
      place_block(ctx, p.core[j].p, dst, a * 2, c, Rj, scale, acc);
      if (p.dtype == QTT_C128) place_block(ctx, p.core[j].p + sp, dst + plane, a * 2, c, Rj, scale, acc);
      l0 += a; r0 += c;
    }
  }
  // identity operator chain (w = 1)
  Mpo I;
  I.ctx = ctx; I.n = n; I.dtype = QTT_F64;
  I.w.assign(n + 1, 1);
  I.core.resize(n);
  I.host.assign(n, std::vector<double>{1.0, 0.0, 0.0, 1.0});
  for (int j = 0; j < n; ++j) {
    dbuf_reserve(I.core[j], 4);
    QTT_CUDA(cudaMemcpyAsync(I.core[j].p, I.host[j].data(), 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  apply_mpo_mps(ctx, I, sum, cutoff, maxdim, mindim, out);
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));  // `sum` and `I` are released on return
}

// all 2^n values; out_re (and out_im for a complex chain) are device buffers of 2^n doubles
void mps_to_vec(Ctx* ctx, const Mps& x, double* out_re, double* out_im) {
  const int n = x.n;
  QTT_REQUIRE(n >= 1 && n <= 34, "mps_to_vec: n = %d out of range [1, 34]", n);
  QTT_REQUIRE(x.chi[0] == 1 && x.chi[n] == 1, "mps_to_vec: boundary bond dimensions must be 1");
  const bool cplx = x.dtype == QTT_C128;
  QTT_REQUIRE(!cplx || out_im, "mps_to_vec: complex chain needs an imaginary output plane");
  auto re = [&](int j) { return x.core[j].p; };
  auto im = [&](int j) { return cplx ? x.core[j].p + (size_t)x.chi[j] * 2 * x.chi[j + 1] : nullptr; };
  if (n == 1) {
    QTT_CUDA(cudaMemcpyAsync(out_re, re(0), 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (cplx) QTT_CUDA(cudaMemcpyAsync(out_im, im(0), 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    else if (out_im) QTT_CUDA(cudaMemsetAsync(out_im, 0, 2 * sizeof(double), ctx->stream));
    return;
  }
  // cut: the smallest bond within 3 sites of the middle (both factors stay far below 2^31 rows / columns)
  int m = n / 2;
  if (m < 1) m = 1;
  for (int c = std::max(1, n / 2 - 3); c <= std::min(n - 1, n / 2 + 3); ++c)
    if (x.chi[c] < x.chi[m]) m = c;
  const int planes = cplx ? 2 : 1;
  // left part T [2^m][chi_m]
  const double *Lr = re(0), *Li = im(0);
  PoolBuf lbuf;
  for (int j = 1; j < m; ++j) {
    const int64_t rows = 1LL << j, k = x.chi[j], cols = 2 * x.chi[j + 1];
    PoolBuf nb((size_t)rows * cols * planes, ctx->stream);
    double* nr = nb.p;
    double* ni = cplx ? nb.p + (size_t)rows * cols : nullptr;
    gemm_planar(ctx, Lr, Li, re(j), im(j), nr, ni, rows, cols, k);
    lbuf = std::move(nb);  // the previous buffer returns to the stream-ordered pool after its last use was enqueued
    Lr = nr; Li = ni;
  }
  // right part R [chi_m][2^(n-m)]
  const double *Rr = re(n - 1), *Ri = im(n - 1);
  PoolBuf rbuf;
  for (int j = n - 2; j >= m; --j) {
    const int64_t rows = x.chi[j] * 2, k = x.chi[j + 1], cols = 1LL << (n - 1 - j);
    PoolBuf nb((size_t)rows * cols * planes, ctx->stream);
    double* nr = nb.p;
    double* ni = cplx ? nb.p + (size_t)rows * cols : nullptr;
    gemm_planar(ctx, re(j), im(j), Rr, Ri, nr, ni, rows, cols, k);
    rbuf = std::move(nb);
    Rr = nr; Ri = ni;
  }
  gemm_planar(ctx, Lr, Li, Rr, Ri, out_re, cplx ? out_im : nullptr, 1LL << m, 1LL << (n - m), x.chi[m]);
  if (!cplx && out_im) QTT_CUDA(cudaMemsetAsync(out_im, 0, ((size_t)1 << n) * sizeof(double), ctx->stream));
}

// values at `count` grid indices (device array idx, 0 <= idx < 2^n); out: [count] real (+ [count] imaginary) on the device
void mps_sample(Ctx* ctx, const Mps& x, const int64_t* idx_dev, int64_t count, double* out_dev) {
  const int n = x.n;
  QTT_REQUIRE(n >= 1 && n <= SAMPLE_MAX_SITES - 1, "mps_sample: n = %d out of range [1, %d]", n, SAMPLE_MAX_SITES - 1);
  QTT_REQUIRE(x.chi[0] == 1 && x.chi[n] == 1, "mps_sample: boundary bond dimensions must be 1");
  if (count <= 0) return;
  const bool cplx = x.dtype == QTT_C128;
  SampleArgs a;
  a.n = n;
  a.chimax = 1;
  for (int j = 0; j < n; ++j) a.core[j] = x.core[j].p;
  for (int j = 0; j <= n; ++j) {
    QTT_REQUIRE(x.chi[j] < (1LL << 30), "mps_sample: bond dimension out of range");
    a.chi[j] = (int)x.chi[j];
    a.chimax = std::max(a.chimax, a.chi[j]);
  }
  a.idx = idx_dev; a.count = count; a.out = out_dev;
  const size_t smem = (size_t)2 * (cplx ? 2 : 1) * a.chimax * sizeof(double);
  QTT_REQUIRE(smem <= 200 * 1024, "mps_sample: bond dimension %d exceeds the shared-memory row-vector budget", a.chimax);
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QTT_CUDA(cudaFuncSetAttribute(sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  });
  const unsigned grid = (unsigned)std::min<int64_t>(count, (int64_t)ctx->sms * 8);
  if (cplx) sample_kernel<true><<<grid, 256, smem, ctx->stream>>>(a);
  else sample_kernel<false><<<grid, 256, smem, ctx->stream>>>(a);
  check_launch(ctx, "mps_sample");
}

}  // namespace qtt

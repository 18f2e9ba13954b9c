// TT-SVD of a 2^n-point sample vector on the device: `vec_to_mps(v, s; cutoff, maxdim, mindim)` = `MPS(vec_to_tensor(v), s; ...)`
// (reference src/function_to_mps.jl:170-177), the back end of `function_to_mps(f, s, xstart, xstop; alg="factorize")`
// (:41-52) once f is sampled on `qtt_xrange` (:10-12).  SURVEY.md 8(f) row 1.
//
// Site 1 is the most significant bit (vec_to_tensor reverses the column-major dims, :170-173), so the row-major view of v
// as (2 x 2^{n-1}) already has the site-1 index as its row.  Step j holds the remainder C as a short-fat (q x N) matrix,
// q = 2 r_{j-1}, N = 2^{n-j}, and needs the dominant left singular vectors U (q x r_j) and C' = U^T C -- whose row-major
// storage IS the next step's (2 r_j x N/2) matrix, no data movement.
//
// HBM-bound streaming formulation (C is up to 8.6 GB at n = 30 while q <= 2 maxdim is tiny):
//   1. G = C C^T (q x q): one read of C.  q <= 16: register-accumulating streaming kernel; larger q: split-K batches of the
//      DMMA GEMM + a fixed-order reduction.
//   2. W = full orthonormal eigenbasis of G, descending (rank-revealing QR + Jacobi of svd_qr.cu, with the orthonormal
//      completion for the null space): q x q, exactly orthogonal whatever the rounding noise in G.
//   3. B = W C (one read, one write) and the squared row norms of B.  Because W is orthogonal, the sum of the squared
//      norms of the rows dropped is EXACTLY the discarded weight of the projection, so the NDTensors.truncate! rule is
//      applied to those norms (not to the eigenvalues of G, whose small end is rounding noise of the Gram matrix):
//      the truncation error honoured is the true one even where sigma_k / sigma_1 < sqrt(eps).
//   4. core_j = W[:k]^T, C' = B[:k] (a prefix of the buffer: zero copy).
// Deterministic: all reductions have a fixed order.
#include "common.cuh"
#include "device_utils.cuh"
#include <chrono>
#include <cstdlib>

namespace qtt {
namespace {

using dev::warp_sum;
constexpr int TS_T = 256;

template <int Q>
__global__ void __launch_bounds__(TS_T) gram_small_kernel(const double* __restrict__ C, int q, long N, double* __restrict__ part) {
  constexpr int NP = Q * (Q + 1) / 2;
  double acc[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) acc[i] = 0.0;
  for (long col = (long)blockIdx.x * TS_T + threadIdx.x; col < N; col += (long)gridDim.x * TS_T) {
    double v[Q];
#pragma unroll
    for (int a = 0; a < Q; ++a) v[a] = a < q ? __ldg(C + (long)a * N + col) : 0.0;
    int idx = 0;
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
      for (int b = a; b < Q; ++b) acc[idx++] += v[a] * v[b];
  }
  __shared__ double red[TS_T / 32][NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const double t = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NP; i += TS_T) {
    double t = 0.0;
    for (int w = 0; w < TS_T / 32; ++w) t += red[w][i];
    part[(size_t)blockIdx.x * NP + i] = t;
  }
}

// G (q x q, symmetric) from the per-block packed upper triangles of gram_small_kernel<Q>
__global__ void gram_small_reduce_kernel(const double* __restrict__ part, int nblocks, int Q, int q, double* __restrict__ G) {
  const int NP = Q * (Q + 1) / 2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < q * q; e += gridDim.x * blockDim.x) {
    int a = e / q, b = e % q;
    if (a > b) { const int t = a; a = b; b = t; }
    const int idx = a * Q - a * (a - 1) / 2 + (b - a);
    double t = 0.0;
    for (int s = 0; s < nblocks; ++s) t += part[(size_t)s * NP + idx];
    G[e] = t;
  }
}

// out[e] = sum_s part[s][e]
__global__ void sum_slices_kernel(const double* __restrict__ part, int nslices, long n, double* __restrict__ out) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int s = 0; s < nslices; ++s) t += part[(size_t)s * n + e];
    out[e] = t;
  }
}

// B = W C for q <= Q rows, plus per-block partial squared row norms of B
template <int Q>
__global__ void __launch_bounds__(TS_T) project_small_kernel(const double* __restrict__ W, const double* __restrict__ C, double* __restrict__ B,
                                                             int q, long N, double* __restrict__ npart) {
  __shared__ double Ws[Q * Q];
  __shared__ double red[TS_T / 32][Q];
  for (int i = threadIdx.x; i < Q * Q; i += TS_T) {
    const int r = i / Q, c = i % Q;
    Ws[i] = (r < q && c < q) ? W[r * q + c] : 0.0;
  }
  __syncthreads();
  double nacc[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) nacc[i] = 0.0;
  for (long col = (long)blockIdx.x * TS_T + threadIdx.x; col < N; col += (long)gridDim.x * TS_T) {
    double v[Q];
#pragma unroll
    for (int a = 0; a < Q; ++a) v[a] = a < q ? __ldg(C + (long)a * N + col) : 0.0;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      double o = 0.0;
#pragma unroll
      for (int a = 0; a < Q; ++a) o += Ws[i * Q + a] * v[a];
      if (i < q) B[(long)i * N + col] = o;
      nacc[i] += o * o;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const double t = warp_sum(nacc[i]);
    if (lane == 0) red[warp][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < Q) {
    double t = 0.0;
    for (int w = 0; w < TS_T / 32; ++w) t += red[w][threadIdx.x];
    npart[(size_t)blockIdx.x * Q + threadIdx.x] = t;
  }
}

// partial squared norms of the rows of B (q x N): block (chunk, row)
__global__ void __launch_bounds__(TS_T) rownorm_kernel(const double* __restrict__ B, long N, long chunk, double* __restrict__ npart, int q) {
  const int row = blockIdx.y;
  const long c0 = (long)blockIdx.x * chunk;
  long c1 = c0 + chunk; if (c1 > N) c1 = N;
  const double* r = B + (long)row * N;
  double acc = 0.0;
  for (long c = c0 + threadIdx.x; c < c1; c += TS_T) { const double v = r[c]; acc += v * v; }
  __shared__ double red[TS_T / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < TS_T / 32; ++w) t += red[w];
    npart[(size_t)blockIdx.x * q + row] = t;   // slice-major: sum_slices_kernel reduces it
  }
}

int small_q(int q) { return q <= 2 ? 2 : q <= 4 ? 4 : q <= 8 ? 8 : q <= 16 ? 16 : 0; }

template <int Q>
void gram_small(Ctx* ctx, const double* C, int q, long N, double* part, int nblocks, double* G) {
  gram_small_kernel<Q><<<nblocks, TS_T, 0, ctx->stream>>>(C, q, N, part);
  check_launch(ctx, "ttsvd_gram_small");
  gram_small_reduce_kernel<<<1, 256, 0, ctx->stream>>>(part, nblocks, Q, q, G);
  check_launch(ctx, "ttsvd_gram_reduce");
}

template <int Q>
void project_small(Ctx* ctx, const double* W, const double* C, double* B, int q, long N, double* npart, int nblocks, double* norms) {
  project_small_kernel<Q><<<nblocks, TS_T, 0, ctx->stream>>>(W, C, B, q, N, npart);
  check_launch(ctx, "ttsvd_project_small");
  // npart is [block][Q]; rows >= q are zeros
  sum_slices_kernel<<<1, 64, 0, ctx->stream>>>(npart, nblocks, Q, norms);
  check_launch(ctx, "ttsvd_norm_reduce");
}

// scratch (doubles) one Gram + projection of a (q x N) matrix needs from the arena, beyond what svd_rows takes
struct StepPlan {
  int Qs;            // template size of the small-q kernels, 0 = GEMM path
  long nsplit, chunk;  // split-K slices of the Gram GEMM
  long nchunk_norm;  // column chunks of the row-norm kernel
  size_t part_elems;
};
StepPlan plan_step(const Ctx* ctx, int q, long N) {
  StepPlan p;
  p.Qs = small_q(q);
  p.nsplit = 1; p.chunk = N; p.nchunk_norm = 0;
  const int nblocks = ctx->sms * 4;
  if (!p.Qs) {
    long ns = N / 4096;
    const long cap = (1L << 24) / ((long)q * q);
    if (ns > cap) ns = cap;
    if (ns > 4096) ns = 4096;
    if (ns < 1) ns = 1;
    long p2 = 1;
    while (p2 * 2 <= ns) p2 *= 2;  // N is a power of two: keep the slices equal
    p.nsplit = p2; p.chunk = N / p2;
    p.nchunk_norm = N >= (1L << 16) ? ((N >> 14) > 1024 ? 1024 : (N >> 14)) : 1;
    p.part_elems = (size_t)p.nsplit * q * q + (size_t)p.nchunk_norm * q;
  } else {
    p.part_elems = (size_t)nblocks * (p.Qs * (p.Qs + 1) / 2 + p.Qs);
  }
  return p;
}
size_t step_scratch(const StepPlan& p, int q) { return p.part_elems + (size_t)32 * q * q + (size_t)64 * q + 65536; }

// G = C C^T
void gram(Ctx* ctx, const StepPlan& p, const double* C, int q, long N, double* part, double* G) {
  const int nblocks = ctx->sms * 4;
  if (p.Qs == 2) gram_small<2>(ctx, C, q, N, part, nblocks, G);
  else if (p.Qs == 4) gram_small<4>(ctx, C, q, N, part, nblocks, G);
  else if (p.Qs == 8) gram_small<8>(ctx, C, q, N, part, nblocks, G);
  else if (p.Qs == 16) gram_small<16>(ctx, C, q, N, part, nblocks, G);
  else {
    GemmDesc g;
    g.A = C; g.B = C; g.C = part; g.M = q; g.N = q; g.K = (int)p.chunk; g.lda = N; g.ldb = N; g.ldc = q; g.transB = true;
    g.batch = (int)p.nsplit; g.sA = p.chunk; g.sB = p.chunk; g.sC = (long)q * q;
    gemm(ctx, g);
    sum_slices_kernel<<<(q * q + 255) / 256, 256, 0, ctx->stream>>>(part, (int)p.nsplit, (long)q * q, G);
    check_launch(ctx, "ttsvd_gram_sum");
  }
}

// B = W C (W q x q row-major) and norms[i] = |B[i, :]|^2
void project(Ctx* ctx, const StepPlan& p, const double* W, const double* C, double* B, int q, long N, double* part, double* norms) {
  const int nblocks = ctx->sms * 4;
  if (p.Qs == 2) project_small<2>(ctx, W, C, B, q, N, part, nblocks, norms);
  else if (p.Qs == 4) project_small<4>(ctx, W, C, B, q, N, part, nblocks, norms);
  else if (p.Qs == 8) project_small<8>(ctx, W, C, B, q, N, part, nblocks, norms);
  else if (p.Qs == 16) project_small<16>(ctx, W, C, B, q, N, part, nblocks, norms);
  else {
    const long cn = N > (1L << 20) ? (1L << 20) : N;   // columns per batch entry (grid.y limit)
    GemmDesc g;
    g.A = W; g.B = C; g.C = B; g.M = q; g.N = (int)cn; g.K = q; g.lda = q; g.ldb = N; g.ldc = N;
    g.batch = (int)(N / cn); g.sA = 0; g.sB = cn; g.sC = cn;
    gemm(ctx, g);
    double* npart = part + (size_t)p.nsplit * q * q;
    const long nchunk = p.nchunk_norm;
    const long csz = (N + nchunk - 1) / nchunk;
    rownorm_kernel<<<dim3((unsigned)nchunk, (unsigned)q), TS_T, 0, ctx->stream>>>(B, N, csz, npart, q);
    check_launch(ctx, "ttsvd_rownorm");
    sum_slices_kernel<<<(q + 63) / 64, 64, 0, ctx->stream>>>(npart, (int)nchunk, (long)q, norms);
    check_launch(ctx, "ttsvd_norm_sum");
  }
}

// NDTensors.truncate!(P; cutoff, maxdim, mindim) on the host copy of the squared row norms (same rule as
// dev::truncate_spectrum; the norms are in descending order up to rounding noise)
int host_truncate(const double* p, int len, double cutoff, int maxdim, int mindim, double* truncerr) {
  int n = len;
  double err = 0.0;
  if (maxdim > len) maxdim = len;
  if (mindim < 1) mindim = 1;
  if (len > 1) {
    while (n > maxdim) err += p[--n];
    double scale = 0.0;
    for (int i = 0; i < len; ++i) scale += p[i];
    if (scale == 0.0) scale = 1.0;
    while (n > mindim && err + p[n - 1] <= cutoff * scale) err += p[--n];
    err /= scale;
  }
  if (n < 1) n = 1;
  *truncerr = err;
  return n;
}

}  // namespace

// v_dev: 2^n doubles on the device (not modified).  `out` receives n cores [r_{j-1}][2][r_j], orthogonality centre at site n.
void vec_to_mps(Ctx* ctx, const double* v_dev, int n, double cutoff, int maxdim, int mindim, Mps& out, double* maxerr_out) {
  QTT_REQUIRE(n >= 1 && n <= 34, "vec_to_mps: n = %d out of range [1, 34]", n);
  if (mindim < 1) mindim = 1;
  const long total = 1L << n;
  out.ctx = ctx; out.n = n; out.dtype = QTT_F64;
  out.chi.assign(n + 1, 1);
  out.core.resize(n);
  double maxerr = 0.0;
  if (n == 1) {
    dbuf_reserve(out.core[0], 2);
    QTT_CUDA(cudaMemcpyAsync(out.core[0].p, v_dev, 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (maxerr_out) *maxerr_out = 0.0;
    return;
  }
  // two ping-pong work buffers of 2^n doubles (B of step j has q_j N_j = 2 r_{j-1} 2^{n-j} <= 2^n elements)
  double* work[2] = {nullptr, nullptr};
  work[0] = static_cast<double*>(pool_alloc(sizeof(double) * (size_t)total, ctx->stream));
  double* h_norms = ctx->h_scal + 2048;  // pinned, 1024 doubles
  try {
    work[1] = static_cast<double*>(pool_alloc(sizeof(double) * (size_t)total, ctx->stream));
    const double* C = v_dev;
    int rprev = 1;
    const bool trace = getenv("QTT_TRACE") != nullptr;
    for (int j = 1; j <= n - 1; ++j) {
      auto tstep = std::chrono::steady_clock::now();
      int levels = 0;
      const int q = 2 * rprev;
      const long N = 1L << (n - j);
      QTT_REQUIRE(q <= 1024, "vec_to_mps: bond dimension %d at site %d is beyond this build's limit (512); pass maxdim", rprev, j);
      double* B = work[(j - 1) & 1];
      double* other = work[j & 1];  // free once B exists, unless it still holds nothing we need (C lives there for j >= 2)
      const StepPlan plan = plan_step(ctx, q, N);
      ctx->ensure_arena(3 * step_scratch(plan, q) * sizeof(double));
      const size_t mark = ctx->arena.mark();
      double* part = ctx->arena.alloc(plan.part_elems);
      double* G = ctx->arena.alloc((size_t)q * q);
      double* W = ctx->arena.alloc((size_t)q * q);
      double* norms = ctx->arena.alloc(q + 16);
      // 1.-3. Gram matrix, full orthonormal eigenbasis (descending), B = W C with its squared row norms
      gram(ctx, plan, C, q, N, part, G);
      {
        const size_t m2 = ctx->arena.mark();
        RowSvd ev = svd_rows(ctx, G, q, q, 0.0, q, q, nullptr, 0);
        QTT_REQUIRE(ev.k == q, "vec_to_mps: internal: eigenbasis of rank %d for a %d x %d Gram matrix", ev.k, q, q);
        QTT_CUDA(cudaMemcpyAsync(W, ev.Wn, sizeof(double) * (size_t)q * q, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->arena.release(m2);  // stream-ordered: later arena users are enqueued after the copy
      }
      project(ctx, plan, W, C, B, q, N, part, norms);
      int md = maxdim > 0 ? maxdim : q;
      if (md > q) md = q;
      if ((long)md > N) md = (int)N;  // rank <= number of columns
      QTT_CUDA(cudaMemcpyAsync(h_norms, norms, sizeof(double) * q, cudaMemcpyDeviceToHost, ctx->stream));
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      double err = 0.0;
      int k = host_truncate(h_norms, q, cutoff, md, mindim, &err);
      // 3b. The Gram matrix resolves singular values only down to ~sqrt(eps) sigma_1: rows whose weight is below 1e-12 of
      // the leading one (including the orthonormal completion of a numerically rank-deficient G) are a correct but
      // unsorted basis of the small end of the spectrum.  When the rule wants to keep some of them, the same procedure is
      // repeated on that block alone -- its own leading singular value is now the reference scale -- which resolves
      // another six orders of magnitude per level.
      int s0 = 0;
      for (int level = 0; level < 3 && j >= 2; ++level) {
        int s = s0;
        while (s < q && h_norms[s] > 1e-12 * h_norms[s0]) ++s;
        if ((q - s) & 1) --s;            // even block (svd_qr.cu)
        if (s <= s0 || k <= s || q - s < 2) break;
        const int mb = q - s;
        double* Bb = B + (long)s * N;
        const StepPlan pb = plan_step(ctx, mb, N);
        const size_t m2 = ctx->arena.mark();
        double* partb = ctx->arena.alloc(pb.part_elems);
        double* Gb = ctx->arena.alloc((size_t)mb * mb);
        double* Wb = ctx->arena.alloc((size_t)mb * mb);
        double* Wnew = ctx->arena.alloc((size_t)mb * q);
        double* normsb = ctx->arena.alloc(mb + 16);
        gram(ctx, pb, Bb, mb, N, partb, Gb);
        {
          const size_t m3 = ctx->arena.mark();
          RowSvd ev = svd_rows(ctx, Gb, mb, mb, 0.0, mb, mb, nullptr, 0);
          QTT_REQUIRE(ev.k == mb, "vec_to_mps: internal: block eigenbasis of rank %d for %d rows", ev.k, mb);
          QTT_CUDA(cudaMemcpyAsync(Wb, ev.Wn, sizeof(double) * (size_t)mb * mb, cudaMemcpyDeviceToDevice, ctx->stream));
          ctx->arena.release(m3);
        }
        project(ctx, pb, Wb, Bb, other, mb, N, partb, normsb);   // rotated block -> the free work buffer, then back in place
        QTT_CUDA(cudaMemcpyAsync(Bb, other, sizeof(double) * (size_t)mb * N, cudaMemcpyDeviceToDevice, ctx->stream));
        {
          GemmDesc g;  // W[s:q] <- Wb . W[s:q]
          g.A = Wb; g.B = W + (size_t)s * q; g.C = Wnew; g.M = mb; g.N = q; g.K = mb; g.lda = mb; g.ldb = q; g.ldc = q;
          gemm(ctx, g);
          QTT_CUDA(cudaMemcpyAsync(W + (size_t)s * q, Wnew, sizeof(double) * (size_t)mb * q, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        QTT_CUDA(cudaMemcpyAsync(h_norms + s, normsb, sizeof(double) * mb, cudaMemcpyDeviceToHost, ctx->stream));
        QTT_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->arena.release(m2);
        k = host_truncate(h_norms, q, cutoff, md, mindim, &err);
        s0 = s;
        ++levels;
      }
      if (err > maxerr) maxerr = err;
      // 4. core_j [(a s)][k] = W[:k]^T
      dbuf_reserve(out.core[j - 1], (size_t)q * k);
      {
        int64_t dims[2] = {k, q};
        int pm[2] = {1, 0};
        permute(ctx, W, out.core[j - 1].p, 2, dims, pm);
      }
      out.chi[j] = k;
      rprev = k;
      C = B;  // C' = first k rows of B, viewed as (2k x N/2)
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));  // the arena scratch (W) is released next
      ctx->arena.release(mark);
      if (trace)
        fprintf(stderr, "[qtt trace] ttsvd step %d: q %d x N %ld -> k %d (%d refinement levels) %.3f ms\n", j, q, N, k, levels,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tstep).count());
    }
    dbuf_reserve(out.core[n - 1], (size_t)rprev * 2);
    QTT_CUDA(cudaMemcpyAsync(out.core[n - 1].p, C, sizeof(double) * (size_t)rprev * 2, cudaMemcpyDeviceToDevice, ctx->stream));
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    if (work[0]) pool_free(work[0], ctx->stream);
    if (work[1]) pool_free(work[1], ctx->stream);
    throw;
  }
  pool_free(work[0], ctx->stream);
  pool_free(work[1], ctx->stream);
  if (maxerr_out) *maxerr_out = maxerr;
}

}  // namespace qtt

// Krylov vector kernels (replace KrylovKit's orthogonaliser / vector interface on ITensors):
// multi-dot h = V^T w, multi-axpy w -= V h, nrm2, scale, axpby.  All are HBM/L2-bound streaming kernels:
// 16-byte vector loads, warp-shuffle reductions, deterministic two-level reduction (per-CTA partials written
// to a buffer, reduced in a fixed order by the last CTA to arrive -- no floating-point atomics, so results
// are bit-reproducible run to run).
#include "common.cuh"

namespace qtt {
namespace {

constexpr int KT = 256;          // threads per CTA
constexpr int EPT = 4;           // elements per thread per chunk
constexpr int CHUNK = KT * EPT;  // 1024 elements per CTA iteration

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[cta][i] = sum over the CTA's elements of V[i][e] * w[e] for up to NV vectors at once (accumulators in
// registers, ONE block-level reduction for all of them); the last CTA to arrive reduces the partials in a fixed order.
constexpr int NV = 32;
__global__ void __launch_bounds__(KT) dots_kernel(const double* __restrict__ V, int nvec, size_t len, const double* __restrict__ w,
                                                  double* __restrict__ partial, double* __restrict__ h, unsigned int* counter,
                                                  int accumulate) {
  __shared__ double red[KT / 32][NV + 1];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  const bool vec_ok = ((len & 1) == 0);
  for (size_t base = (size_t)blockIdx.x * CHUNK; base < len; base += (size_t)gridDim.x * CHUNK) {
    size_t e = base + (size_t)tid * EPT;
    if (e + EPT <= len && vec_ok) {
      double2 b0 = *reinterpret_cast<const double2*>(w + e), b1 = *reinterpret_cast<const double2*>(w + e + 2);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i < nvec) {
          const double* vi = V + (size_t)i * len + e;
          double2 a0 = *reinterpret_cast<const double2*>(vi), a1 = *reinterpret_cast<const double2*>(vi + 2);
          acc[i] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
        }
      }
    } else {
      for (int k = 0; k < EPT; ++k) {
        if (e + k < len) {
          double wv = w[e + k];
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (i < nvec) acc[i] += V[(size_t)i * len + e + k] * wv;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nvec) {
      double v = warp_sum(acc[i]);
      if (lane == 0) red[warp][i] = v;
    }
  }
  __syncthreads();
  if (tid < nvec) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KT / 32; ++k) s += red[k][tid];
    partial[(size_t)blockIdx.x * nvec + tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // nvec * 8 threads: 8 lanes cooperate on each vector's column of partials
    const int grp = tid >> 3, sub = tid & 7;
    for (int i = grp; i < nvec; i += KT / 8) {
      double s = 0.0;
      for (unsigned int b = sub; b < gridDim.x; b += 8) s += __ldcg(partial + (size_t)b * nvec + i);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (sub == 0) h[i] = accumulate ? (h[i] + s) : s;
    }
    if (tid == 0) *counter = 0;
  }
}

// w[e] += sign * sum_i h[i] V[i][e]; optionally accumulates ||w_new||^2 into nrm2sq_out (deterministic, via partials)
__global__ void __launch_bounds__(KT) axpys_kernel(const double* __restrict__ V, int nvec, size_t len, const double* __restrict__ h,
                                                   double* __restrict__ w, double sign, double* __restrict__ partial,
                                                   double* __restrict__ nrm2sq_out, unsigned int* counter) {
  extern __shared__ double hs[];
  __shared__ double red[KT / 32];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < nvec; i += KT) hs[i] = sign * h[i];
  __syncthreads();
  double nacc = 0.0;
  for (size_t base = (size_t)blockIdx.x * CHUNK; base < len; base += (size_t)gridDim.x * CHUNK) {
    size_t e = base + (size_t)tid * EPT;
    if (e + EPT <= len && ((len & 1) == 0)) {
      double2 w0 = *reinterpret_cast<const double2*>(w + e), w1 = *reinterpret_cast<const double2*>(w + e + 2);
      for (int i = 0; i < nvec; ++i) {
        const double* vi = V + (size_t)i * len + e;
        double2 a0 = *reinterpret_cast<const double2*>(vi), a1 = *reinterpret_cast<const double2*>(vi + 2);
        double c = hs[i];
        w0.x += c * a0.x; w0.y += c * a0.y; w1.x += c * a1.x; w1.y += c * a1.y;
      }
      *reinterpret_cast<double2*>(w + e) = w0;
      *reinterpret_cast<double2*>(w + e + 2) = w1;
      nacc += w0.x * w0.x + w0.y * w0.y + w1.x * w1.x + w1.y * w1.y;
    } else {
      for (int k = 0; k < EPT; ++k) {
        if (e + k < len) {
          double x = w[e + k];
          for (int i = 0; i < nvec; ++i) x += hs[i] * V[(size_t)i * len + e + k];
          w[e + k] = x;
          nacc += x * x;
        }
      }
    }
  }
  if (nrm2sq_out == nullptr) return;
  nacc = warp_sum(nacc);
  if (lane == 0) red[warp] = nacc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < KT / 32; ++k) s += red[k];
    partial[blockIdx.x] = s;
    __threadfence();
    unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(partial + b);
    *nrm2sq_out = s;
    *counter = 0;
  }
}

// out[e] = w[e] / d  with d = *denom (or sqrt(*denom))
__global__ void __launch_bounds__(KT) scale_to_kernel(const double* __restrict__ w, size_t len, const double* __restrict__ denom,
                                                      int sqrt_denom, double* __restrict__ out, const double* __restrict__ done) {
  double d = *denom;
  if (sqrt_denom) d = sqrt(d);
  // once the Krylov space is exhausted / converged (device flag), later basis vectors are written as zeros so that
  // run-ahead iterations stay finite and contribute nothing
  const bool dead = (done != nullptr && *done != 0.0);
  const double inv = (d != 0.0 && !dead) ? 1.0 / d : 0.0;
  for (size_t e = (size_t)blockIdx.x * KT + threadIdx.x; e < len; e += (size_t)gridDim.x * KT) out[e] = w[e] * inv;
}

// out = a x + b y + c z  (null pointers are skipped)
__global__ void __launch_bounds__(KT) axpby_kernel(size_t len, double a, const double* __restrict__ x, double b,
                                                   const double* __restrict__ y, double c, const double* __restrict__ z,
                                                   double* __restrict__ out) {
  for (size_t e = (size_t)blockIdx.x * KT + threadIdx.x; e < len; e += (size_t)gridDim.x * KT) {
    double v = 0.0;
    if (x) v += a * x[e];
    if (y) v += b * y[e];
    if (z) v += c * z[e];
    out[e] = v;
  }
}

inline unsigned grid_for(const Ctx* ctx, size_t len, int per_cta) {
  size_t g = (len + per_cta - 1) / per_cta;
  size_t cap = (size_t)ctx->sms * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

double* Ctx::partial(size_t n) {
  if (n > partial_cap) {
    if (d_partial) cudaFree(d_partial);
    size_t cap = n < 65536 ? 65536 : n;
    QTT_CUDA(cudaMalloc(&d_partial, cap * sizeof(double)));
    partial_cap = cap;
  }
  return d_partial;
}

static void dots_impl(Ctx* ctx, const double* V, int nvec, size_t len, const double* w, double* h, int accumulate) {
  unsigned g = grid_for(ctx, len, CHUNK);
  for (int i0 = 0; i0 < nvec; i0 += NV) {
    int nv = nvec - i0 < NV ? nvec - i0 : NV;
    double* part = ctx->partial((size_t)g * nv);
    dots_kernel<<<g, KT, 0, ctx->stream>>>(V + (size_t)i0 * len, nv, len, w, part, h + i0, ctx->d_sync + 0, accumulate);
    check_launch(ctx, "kry_dots");
  }
}

void kry_dots(Ctx* ctx, const double* V, int nvec, size_t len, const double* w, double* h) {
  if (nvec <= 0) return;
  dots_impl(ctx, V, nvec, len, w, h, 0);
}

void kry_dots_acc(Ctx* ctx, const double* V, int nvec, size_t len, const double* w, double* h) {
  if (nvec <= 0) return;
  dots_impl(ctx, V, nvec, len, w, h, 1);
}

void kry_axpys_nrm(Ctx* ctx, const double* V, int nvec, size_t len, const double* h, double* w, double sign, double* nrm2sq_out) {
  unsigned g = grid_for(ctx, len, CHUNK);
  double* part = ctx->partial(g);
  axpys_kernel<<<g, KT, (size_t)(nvec > 0 ? nvec : 1) * sizeof(double), ctx->stream>>>(V, nvec, len, h, w, sign, part, nrm2sq_out,
                                                                                      ctx->d_sync + 1);
  check_launch(ctx, "kry_axpys");
}

void kry_axpys(Ctx* ctx, const double* V, int nvec, size_t len, const double* h, double* w, double sign) {
  if (nvec <= 0) return;
  kry_axpys_nrm(ctx, V, nvec, len, h, w, sign, nullptr);
}

void kry_nrm2sq(Ctx* ctx, const double* w, size_t len, double* out) {
  // ||w||^2 = w . w through the dot kernel (one "basis vector" = w itself)
  unsigned g = grid_for(ctx, len, CHUNK);
  double* part = ctx->partial(g);
  dots_kernel<<<g, KT, 0, ctx->stream>>>(w, 1, len, w, part, out, ctx->d_sync + 0, 0);
  check_launch(ctx, "kry_nrm2sq");
}

void kry_scale_to(Ctx* ctx, const double* w, size_t len, const double* denom_dev, bool sqrt_denom, double* out, const double* done_dev) {
  unsigned g = grid_for(ctx, len, KT);
  scale_to_kernel<<<g, KT, 0, ctx->stream>>>(w, len, denom_dev, sqrt_denom ? 1 : 0, out, done_dev);
  check_launch(ctx, "kry_scale_to");
}

void kry_axpby(Ctx* ctx, size_t len, double a, const double* x, double b, const double* y, double c, const double* z, double* out) {
  unsigned g = grid_for(ctx, len, KT);
  axpby_kernel<<<g, KT, 0, ctx->stream>>>(len, a, x, b, y, c, z, out);
  check_launch(ctx, "kry_axpby");
}

}  // namespace qtt

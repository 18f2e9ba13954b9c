// Krylov vector kernels (replace KrylovKit's orthogonaliser / vector interface on ITensors):
// multi-dot h = V^T w, multi-axpy w -= V h, nrm2, scale, axpby.  All are HBM/L2-bound streaming kernels:
// 16-byte vector loads, warp-shuffle reductions, deterministic two-level reduction (per-CTA partials written
// to a buffer, reduced in a fixed order by the last CTA to arrive -- no floating-point atomics, so results
// are bit-reproducible run to run).
#include "common.cuh"
#include "krylov_scalars.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace qtt {
namespace {

constexpr int KT = 256;          // threads per CTA
constexpr int EPT = 4;           // elements per thread per chunk
constexpr int CHUNK = KT * EPT;  // 1024 elements per CTA iteration

using dev::warp_sum;

// partial[cta][i] = sum over the CTA's elements of V[i][e] * w[e] for up to NV vectors at once (accumulators in
// registers, ONE block-level reduction for all of them); the last CTA to arrive reduces the partials in a fixed order.
constexpr int NV = 32;
__global__ void __launch_bounds__(KT) dots_kernel(const double* __restrict__ V, int nvec, size_t len, const double* __restrict__ w,
                                                  double* __restrict__ partial, double* __restrict__ h, unsigned int* counter,
                                                  int accumulate) {
  dev::pdl_wait();
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
  dev::pdl_wait();
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
  dev::pdl_wait();
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
  dev::pdl_wait();
  for (size_t e = (size_t)blockIdx.x * KT + threadIdx.x; e < len; e += (size_t)gridDim.x * KT) {
    double v = 0.0;
    if (x) v += a * x[e];
    if (y) v += b * y[e];
    if (z) v += c * z[e];
    out[e] = v;
  }
}


// ------------------------------------------------------------------ fused CGS2 Arnoldi step (one cooperative launch)
// w <- (I - V V^T)^2 w with both coefficient sets kept, ||w||, and v_next = w / ||w||: the six launches of the unfused
// path (dots, axpys, dots, axpys+nrm, scalar update, scale) become ONE persistent kernel with three grid barriers.
// Each CTA owns a contiguous slice of the vector, which stays in shared memory for the whole step; V is streamed from
// L2 (the Krylov basis of a chi=256 bond is 31 x 2 MiB and stays L2-resident).  Reductions are per-CTA partials summed
// in a fixed order by every CTA redundantly, so the coefficients are bit-reproducible and identical in all CTAs.
struct Cgs2Args {
  const double* V;
  int k;
  size_t len;
  double* w;        // in: the vector to orthogonalise; out: the orthogonalised, unnormalised vector
  double* vnext;    // out (may be null): w / ||w||, or zeros when dead
  double* S;        // scalar block (sweep.cu layout) or null
  double* part;     // [3][G][64] partials
  unsigned* bar;    // monotonic barrier counter
  unsigned epoch;   // counter value at which this launch starts
  int per;          // slice length per CTA (even)
  double eta2;      // DGKS threshold eta^2 (0: always reorthogonalise); see cgs2_fused_kernel
  int* counters;    // optional [2]: steps, steps that needed the second pass (trace)
  int dbgmode;      // development aid (QTT_CGS2_DBGMODE): 1 = return at entry, 2 = return after loading the w slice
  long long* dbg;   // optional [16]: globaltimer stamps of CTA 0 at the stage boundaries (QTT_CGS2_DBG)
  int kstash;       // the CTA's slices of V[0..kstash) are kept in shared memory after the first pass (read from L2 once, not 4x)
  int post;         // 0: store h1, h2, nrm2 only; 1: GMRES update; 2: Lanczos column
  double a0, a1, tol;
  volatile double* hstat;
  int off_h1, off_h2, off_nres2, off_done;
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) { dev::grid_arrive_wait(counter, target); }
__device__ __forceinline__ void stamp(const Cgs2Args& a, int slot) {
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[slot] = t;
  }
}

// MODE 0: V from L2; the slices of the first kstash vectors are also written to shared memory.
// MODE 1: the stashed slices come from shared memory, the rest from L2.  Vectors are processed in groups of 8 whose
// eight loads are issued back to back before the FMAs; the L2 groups of MODE 1 start at kstash.
template <int MODE>
__device__ __forceinline__ void slice_dots(const Cgs2Args& a, const double* ws, double* vs, size_t s0, int n2, double* red /*[nwarps][64]*/,
                                           double* part_b) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  if (MODE == 1) {
    for (int i0 = 0; i0 < a.kstash; i0 += 8) {
      double acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.0;
      for (int t2 = tid; t2 < n2; t2 += nthr) {
        const double2 wv = *reinterpret_cast<const double2*>(ws + 2 * t2);
        const double* vp = vs + (size_t)i0 * a.per + 2 * t2;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u < a.kstash) {
            const double2 v = *reinterpret_cast<const double2*>(vp + (size_t)u * a.per);
            acc[u] += v.x * wv.x + v.y * wv.y;
          }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        double v = warp_sum(acc[u]);
        if (lane == 0 && i0 + u < a.kstash) red[warp * 64 + i0 + u] = v;
      }
    }
  }
  for (int i0 = (MODE == 1 ? a.kstash : 0); i0 < a.k; i0 += 8) {
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    for (int t2 = tid; t2 < n2; t2 += nthr) {
      const double2 wv = *reinterpret_cast<const double2*>(ws + 2 * t2);
      const double* vp = a.V + s0 + 2 * (size_t)t2 + (size_t)i0 * a.len;
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (i0 + u < a.k) ? __ldcg(reinterpret_cast<const double2*>(vp + (size_t)u * a.len)) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += v[u].x * wv.x + v[u].y * wv.y;
      if (MODE == 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u < a.kstash) *reinterpret_cast<double2*>(vs + (size_t)(i0 + u) * a.per + 2 * t2) = v[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double v = warp_sum(acc[u]);
      if (lane == 0 && i0 + u < a.k) red[warp * 64 + i0 + u] = v;
    }
  }
  __syncthreads();
  if (tid < a.k) {
    double s = 0.0;
    for (int wdx = 0; wdx < nwarps; ++wdx) s += red[wdx * 64 + tid];
    __stcg(part_b + tid, s);
  }
}

// h[i] = sum_b part[b][i] in a fixed order, result in shared memory.  8 lanes per coefficient; every lane first issues
// ALL its loads (independent, one L2 latency in total) and only then adds them up -- a dependent load-add chain over
// 148 partials costs ~5 us per reduction, which was most of the fixed cost of the step.
__device__ __forceinline__ void reduce_partials(const double* part, int k, int G, double* h_s) {
  const int tid = threadIdx.x;
  for (int c0 = 0; c0 < k; c0 += (int)(blockDim.x >> 3)) {
    const int i = c0 + (tid >> 3), sub = tid & 7;
    double v[20];
#pragma unroll
    for (int u = 0; u < 20; ++u) {
      const int b = sub + 8 * u;
      v[u] = (i < k && b < G) ? __ldcg(part + (size_t)b * 64 + i) : 0.0;
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 20; ++u) s += v[u];
    for (int b = sub + 160; b < G; b += 8) s += (i < k) ? __ldcg(part + (size_t)b * 64 + i) : 0.0;  // G > 160 (not on B200)
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (i < k && sub == 0) h_s[i] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void slice_axpys(const Cgs2Args& a, double* ws, const double* vs, size_t s0, int n2, const double* h_s) {
  const int nthr = blockDim.x;
  for (int t2 = threadIdx.x; t2 < n2; t2 += nthr) {
    double2 x = *reinterpret_cast<const double2*>(ws + 2 * t2);
    const double* vp = a.V + s0 + 2 * (size_t)t2;
    for (int i = 0; i < a.kstash; ++i) {
      const double2 v = *reinterpret_cast<const double2*>(vs + (size_t)i * a.per + 2 * t2);
      const double c = h_s[i];
      x.x -= c * v.x;
      x.y -= c * v.y;
    }
    for (int i0 = a.kstash; i0 < a.k; i0 += 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (i0 + u < a.k) ? __ldcg(reinterpret_cast<const double2*>(vp + (size_t)(i0 + u) * a.len)) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u < a.k) {
          const double c = h_s[i0 + u];
          x.x -= c * v[u].x;
          x.y -= c * v[u].y;
        }
    }
    *reinterpret_cast<double2*>(ws + 2 * t2) = x;
  }
}

constexpr int FT_MAX = 512;
__global__ void __launch_bounds__(FT_MAX) cgs2_fused_kernel(Cgs2Args a) {
  dev::pdl_wait();   // no-op unless launched with programmatic stream serialisation
  extern __shared__ double dyn[];
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[13] = a.dbg[11];  // end stamp of the previous launch
  stamp(a, 12);
  if (a.dbgmode == 1) return;
  double* ws = dyn;                // [per]
  double* red = dyn + a.per;       // [nwarps][64]
  double* h1 = red + (blockDim.x >> 5) * 64;  // [64]
  double* h2 = h1 + 64;            // [64]
  double* vs = h2 + 64;            // [kstash][per] stashed slices of V
  __shared__ double nred[FT_MAX / 32];
  __shared__ double s_n2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  const size_t s0 = (size_t)b * a.per;
  size_t s1 = s0 + a.per;
  if (s1 > a.len) s1 = a.len;
  const int n2 = s0 < a.len ? (int)((s1 - s0) / 2) : 0;
  // the done flag of the PREVIOUS step decides whether v_next is written as zeros (the flag of this step is only
  // committed by CTA 0 after the last barrier; see sweep.cu for why lagging by one step is safe)
  const bool dead = a.S != nullptr && a.S[a.off_done] != 0.0;
  // CTA 0 stages the Givens rotations / rhs element of the earlier steps for the on-chip scalar update at the end
  __shared__ double gcs[64], gss[64];
  __shared__ double s_yprev, s_r0n2;
  if (b == 0 && a.S != nullptr && a.post == 1) {
    if (tid < 64) {
      gcs[tid] = a.S[kscal::OFF_GC + tid];
      gss[tid] = a.S[kscal::OFF_GS + tid];
    }
    if (tid == 0) {
      s_yprev = a.k >= 2 ? a.S[kscal::OFF_Y + a.k - 1] : 0.0;
      s_r0n2 = a.S[kscal::S_R0N2];
    }
  }
  stamp(a, 0);
  for (int t2 = tid; t2 < n2; t2 += nthr)
    *reinterpret_cast<double2*>(ws + 2 * t2) = __ldcg(reinterpret_cast<const double2*>(a.w + s0 + 2 * (size_t)t2));
  __syncthreads();
  stamp(a, 1);
  if (a.dbgmode == 2) return;
  double* P1 = a.part;
  double* P2 = a.part + (size_t)G * 64;
  double* P3 = a.part + (size_t)2 * G * 64;
  // |w|^2 of the incoming vector (column 63 of the partial row), for the reorthogonalisation test
  auto block_norm2 = [&](double* dst) {
    double nacc = 0.0;
    for (int t2 = tid; t2 < n2; t2 += nthr) {
      const double2 x = *reinterpret_cast<const double2*>(ws + 2 * t2);
      nacc += x.x * x.x + x.y * x.y;
    }
    nacc = warp_sum(nacc);
    if (lane == 0) nred[warp] = nacc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wdx = 0; wdx < nwarps; ++wdx) s += nred[wdx];
      __stcg(dst, s);
    }
  };
  if (a.eta2 > 0.0) block_norm2(P1 + (size_t)b * 64 + 63);
  slice_dots<0>(a, ws, vs, s0, n2, red, P1 + (size_t)b * 64);
  stamp(a, 2);
  grid_barrier(a.bar, a.epoch + (unsigned)G);
  stamp(a, 3);
  reduce_partials(P1, a.k, G, h1);
  stamp(a, 4);
  // DGKS test (Daniel-Gragg-Kaufman-Stewart): the second Gram-Schmidt pass is only needed when the first one cancelled
  // most of w.  |w'|^2 is ESTIMATED by Pythagoras, |w|^2 - |h1|^2, for the decision only (identical in every CTA: the
  // inputs are the fixed-order reductions); the norm that is used afterwards is computed from the actual vector.
  bool second = true;
  if (a.eta2 > 0.0) {
    if (warp == 0) {
      double sw = 0.0;
      for (int bb = lane; bb < G; bb += 32) sw += __ldcg(P1 + (size_t)bb * 64 + 63);
      sw = warp_sum(sw);
      double sh = 0.0;
      for (int i = lane; i < a.k; i += 32) sh += h1[i] * h1[i];
      sh = warp_sum(sh);
      if (lane == 0) s_n2 = ((sw - sh) >= a.eta2 * sw) ? 0.0 : 1.0;
    }
    __syncthreads();
    second = (s_n2 != 0.0);
    __syncthreads();
  }
  slice_axpys(a, ws, vs, s0, n2, h1);
  __syncthreads();
  stamp(a, 5);
  if (second) {
    slice_dots<1>(a, ws, vs, s0, n2, red, P2 + (size_t)b * 64);
    stamp(a, 6);
    grid_barrier(a.bar, a.epoch + 2u * (unsigned)G);
    stamp(a, 7);
    reduce_partials(P2, a.k, G, h2);
    slice_axpys(a, ws, vs, s0, n2, h2);
    __syncthreads();
    stamp(a, 8);
  } else {
    if (tid < 64) h2[tid] = 0.0;
    // keep the barrier counter in step with the host's bookkeeping (3 arrivals per CTA per launch)
    if (tid == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;\n" ::"l"(a.bar) : "memory");
  }
  if (a.counters && b == 0 && tid == 0) {
    atomicAdd(a.counters, 1);
    if (second) atomicAdd(a.counters + 1, 1);
  }
  // (|w''|^2 = |w'|^2 - |h2|^2 reduced together with the second dots would save this pass and barrier; measured: the
  // extra reduction work costs most of what the barrier saves -- 1 us of a 25 us step -- so the norm stays exact.)
  block_norm2(P3 + b);
  stamp(a, 9);
  grid_barrier(a.bar, a.epoch + 3u * (unsigned)G);
  stamp(a, 10);
  if (warp == 0) {
    double s = 0.0;
    for (int bb = lane; bb < G; bb += 32) s += __ldcg(P3 + bb);
    s = warp_sum(s);  // xor-butterfly: same order in every CTA
    if (lane == 0) s_n2 = s;
  }
  __syncthreads();
  const double n2sq = s_n2;
  const double nrm = sqrt(n2sq);
  const double inv = (nrm != 0.0 && !dead) ? 1.0 / nrm : 0.0;
  for (int t2 = tid; t2 < n2; t2 += nthr) {
    const double2 x = *reinterpret_cast<const double2*>(ws + 2 * t2);
    __stcg(reinterpret_cast<double2*>(a.w + s0 + 2 * (size_t)t2), x);
    if (a.vnext) __stcg(reinterpret_cast<double2*>(a.vnext + s0 + 2 * (size_t)t2), make_double2(x.x * inv, x.y * inv));
  }
  stamp(a, 11);
  if (b == 0 && a.S != nullptr) {
    if (tid < a.k) {
      a.S[a.off_h1 + tid] = h1[tid];
      a.S[a.off_h2 + tid] = h2[tid];
    }
    if (tid == 0) a.S[a.off_nres2] = n2sq;
    if (a.post == 1) {
      if (tid == 0)
        kscal::gmres_step_scalars_onchip(a.S, a.k, a.a0, a.a1, a.tol, 64, a.hstat, h1, h2, gcs, gss, s_yprev, s_r0n2, dead ? 1.0 : 0.0, n2sq);
    } else {
      __syncthreads();
      if (tid == 0 && a.post == 2) kscal::lanczos_step_scalars(a.S, a.k, 64);
    }
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
    launch_pdl(ctx, "kry_dots", dots_kernel, dim3(g), dim3(KT), 0, V + (size_t)i0 * len, nv, len, w, part, h + i0, ctx->d_sync + 0, accumulate);
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
  launch_pdl(ctx, "kry_axpys", axpys_kernel, dim3(g), dim3(KT), (size_t)(nvec > 0 ? nvec : 1) * sizeof(double), V, nvec, len, h, w, sign, part, nrm2sq_out,
                                                                                      ctx->d_sync + 1);
}

void kry_axpys(Ctx* ctx, const double* V, int nvec, size_t len, const double* h, double* w, double sign) {
  if (nvec <= 0) return;
  kry_axpys_nrm(ctx, V, nvec, len, h, w, sign, nullptr);
}

void kry_nrm2sq(Ctx* ctx, const double* w, size_t len, double* out) {
  // ||w||^2 = w . w through the dot kernel (one "basis vector" = w itself)
  unsigned g = grid_for(ctx, len, CHUNK);
  double* part = ctx->partial(g);
  launch_pdl(ctx, "kry_nrm2sq", dots_kernel, dim3(g), dim3(KT), 0, w, 1, len, w, part, out, ctx->d_sync + 0, 0);
}

void kry_scale_to(Ctx* ctx, const double* w, size_t len, const double* denom_dev, bool sqrt_denom, double* out, const double* done_dev) {
  unsigned g = grid_for(ctx, len, KT);
  launch_pdl(ctx, "kry_scale_to", scale_to_kernel, dim3(g), dim3(KT), 0, w, len, denom_dev, sqrt_denom ? 1 : 0, out, done_dev);
}

void kry_axpby(Ctx* ctx, size_t len, double a, const double* x, double b, const double* y, double c, const double* z, double* out) {
  unsigned g = grid_for(ctx, len, KT);
  launch_pdl(ctx, "kry_axpby", axpby_kernel, dim3(g), dim3(KT), 0, len, a, x, b, y, c, z, out);
}

}  // namespace qtt

namespace qtt {

bool kry_cgs2_fused_ok(const Ctx* ctx, size_t len) {
  if (len % 2 != 0) return false;
  const int G = ctx->sms;
  size_t per = (len + G - 1) / G;
  per = (per + 1) & ~size_t(1);
  return (per + 32 * 64 + 128) * sizeof(double) <= 200 * 1024;
}

// One fused CGS2 step.  S/off_*: scalar block receiving h1, h2, ||w||^2 (and the solver-specific update `post`).
void kry_cgs2_fused(Ctx* ctx, const double* V, int k, size_t len, double* w, double* vnext, double* S, int off_h1, int off_h2,
                    int off_nres2, int off_done, int post, double a0, double a1, double tol, volatile double* hstat) {
  QTT_REQUIRE(k >= 1 && k <= 64, "kry_cgs2_fused: k out of range");
  // one CTA per >= 512 elements: short vectors (edge / mid bonds) use few CTAs, whose grid barriers are cheaper (measured
  // at len = 16384 / 65536, k = 30: 25.2 / 27.7 us with 512, 30.3 / 31.8 us with 1024, no gain below 512)
  // With several live contexts on the device (independent instances on their own streams) a grid of up to 148 CTAs with a full
  // SM of shared memory each serialises the streams: one CTA per >= 2048 elements lets four of these kernels run side by side
  // (measured, 8 streams of n = 28, chi = 128 instances: 19.3 -> 31.3 instances/s; 4096: 29.3).
  static const int min_per_env = getenv("QTT_CGS2_MINPER") ? atoi(getenv("QTT_CGS2_MINPER")) : 0;
  const int min_per = min_per_env > 0 ? min_per_env : (live_contexts(ctx->device) > 1 ? 2048 : 512);
  int G = (int)((len + min_per - 1) / min_per);
  if (G > ctx->sms) G = ctx->sms;
  if (G < 1) G = 1;
  size_t per = (len + G - 1) / G;
  per = (per + 1) & ~size_t(1);
  static const int fthreads = getenv("QTT_CGS2_THREADS") ? atoi(getenv("QTT_CGS2_THREADS")) : 512;
  const size_t base = per + (size_t)(fthreads / 32) * 64 + 128;  // doubles: w slice, reduction scratch, h1, h2
  // stash budget: the whole 227 KB of the SM minus the kernel's static shared memory (one CTA per SM anyway); at
  // chi = 256 this holds the CTA's slices of 14 basis vectors (a 200 KB budget and whole groups of 8 held only 8)
  constexpr size_t SMEM_BUDGET = 227 * 1024 - 2048;
  static const int stash_cap = getenv("QTT_CGS2_STASH") ? atoi(getenv("QTT_CGS2_STASH")) : 64;
  int kstash = 0;
  if (stash_cap > 0 && SMEM_BUDGET / sizeof(double) > base) {
    const size_t room = SMEM_BUDGET / sizeof(double) - base;
    kstash = (int)(room / per);
    if (kstash > k) kstash = k;
    if (kstash > stash_cap) kstash = stash_cap;
  }
  const size_t smem = (base + (size_t)kstash * per) * sizeof(double);
  static DevOnce once;
  once.run(ctx->device, [&] { QTT_CUDA(cudaFuncSetAttribute(cgs2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET)); });
  Cgs2Args a;
  a.V = V; a.k = k; a.len = len; a.w = w; a.vnext = vnext; a.S = S;
  a.part = ctx->partial((size_t)3 * G * 64);
  a.bar = ctx->d_sync + 4;
  a.epoch = ctx->bar_epoch;
  a.per = (int)per;
  a.kstash = kstash;
  static const double eta = getenv("QTT_DGKS_ETA") ? atof(getenv("QTT_DGKS_ETA")) : 0.0;
  a.eta2 = eta * eta;
  a.counters = nullptr;
  a.dbg = nullptr;
  static const int dbgmode = getenv("QTT_CGS2_DBGMODE") ? atoi(getenv("QTT_CGS2_DBGMODE")) : 0;
  a.dbgmode = dbgmode;
  static const bool dbg_on = getenv("QTT_CGS2_DBG") != nullptr;
  static long long* dbg_dev = nullptr;
  static int dbg_calls = 0;
  if (dbg_on) {
    if (!dbg_dev) QTT_CUDA(cudaMalloc(&dbg_dev, 16 * sizeof(long long)));
    a.dbg = dbg_dev;
  }
  a.post = post;
  a.a0 = a0; a.a1 = a1; a.tol = tol; a.hstat = hstat;
  a.off_h1 = off_h1; a.off_h2 = off_h2; a.off_nres2 = off_nres2; a.off_done = off_done;
  // cooperative launch: the grid barriers need every CTA resident (measured: no extra launch cost over <<< >>>)
  void* args[] = {&a};
  // The grid barriers need every CTA resident.  A cooperative launch guarantees it but cannot overlap the tail of the previous
  // kernel (measured: ~6.5 us between two cooperative launches).  When this context is the only one alive on the device nothing
  // else competes for the SMs -- the <= 148 CTAs (one per SM) become resident as matvec_b drains -- so the kernel is launched
  // with programmatic stream serialisation instead (it starts with griddepcontrol.wait): Krylov phase 73.0 -> 70.4 ms per sweep,
  // 20.6 -> 18.5 us per step at k = 1.  With several live contexts (concurrent streams of independent instances) two partially
  // resident spinning grids could block each other: those keep the cooperative launch.  QTT_COOP_ALWAYS=1 forces it.
  if (sole_context(ctx)) {
    launch_pdl(ctx, "cgs2_fused", cgs2_fused_kernel, dim3(G), dim3(fthreads), smem, a);
    ctx->launches--;   // (launch_pdl counts; the common tail below counts again)
  } else {
    QTT_CUDA(cudaLaunchCooperativeKernel((void*)cgs2_fused_kernel, dim3(G), dim3(fthreads), args, smem, ctx->stream));
  }
  ctx->launches++;
  if (!dbgmode) ctx->bar_epoch += 3u * (unsigned)G;
  if (dbg_on && (++dbg_calls % 97) == 0) {  // development aid: stage times of CTA 0 (ns since the kernel's first stamp)
    long long h[16];
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    QTT_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[cgs2 dbg] k=%d G=%d kstash=%d: load_w %lld dots1 %lld bar1 %lld red1 %lld axpy1 %lld dots2 %lld bar2 %lld red2+axpy2 %lld norm %lld bar3 %lld store %lld | gap since previous launch's end stamp %lld, prologue %lld ns\n",
            k, G, kstash, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8],
            h[10] - h[9], h[11] - h[10], h[12] - h[13], h[0] - h[12]);
  }
}

}  // namespace qtt

// Device-side helpers shared by the persistent kernels: warp reductions, the grid barrier, and the on-device
// truncation rule (NDTensors.truncate!).
#pragma once
#include <cuda_runtime.h>

namespace qtt {
namespace dev {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// First statement of every kernel that launch_pdl() (common.cuh) may launch programmatically behind its producer: blocks until
// the preceding grid has completed and its memory is visible; a no-op under a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p));
  return v;
}

// Grid barrier on a monotonic counter (cooperative launch guarantees co-residency).  Arrival is ONE release-RED: the
// release is cumulative over the CTA's stores ordered before it by bar.sync, so no separate membar round trip is paid
// (measured: -10 % on the Jacobi kernels); the waiters poll with acquire loads.
__device__ __forceinline__ void grid_arrive_wait(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    while ((int)(ld_acquire(counter) - target) < 0) {
    }
  }
  __syncthreads();
}

// counter starts at 0 for this launch; `target` is the caller's running expectation
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target) {
  target += gridDim.x;
  grid_arrive_wait(counter, target);
}

// NDTensors.truncate!(P; cutoff, maxdim, mindim) on P = sigma^2 sorted descending (relative cutoff on the discarded weight).
// Returns the kept length; *truncerr receives the discarded weight fraction.  Single thread.
__device__ inline int truncate_spectrum(const double* sig2, int len, double cutoff, int maxdim, int mindim, double* truncerr) {
  int n = len;
  double err = 0.0;
  if (maxdim > len) maxdim = len;
  if (mindim < 1) mindim = 1;
  if (len > 1) {
    while (n > maxdim) {
      err += __ldcg(sig2 + n - 1);
      --n;
    }
    double scale = 0.0;
    for (int i = 0; i < len; ++i) scale += __ldcg(sig2 + i);
    if (scale == 0.0) scale = 1.0;
    while (n > mindim && err + __ldcg(sig2 + n - 1) <= cutoff * scale) {
      err += __ldcg(sig2 + n - 1);
      --n;
    }
    err /= scale;
  }
  if (n < 1) n = 1;
  *truncerr = err;
  return n;
}

// Descending rank sort of norms[0..q) into (sig2, perm), ties by index; entries [q, len) are exact zeros with perm = -1.
// Grid-strided: every thread of the grid participates.
__device__ inline void rank_sort_desc(const double* norms, int q, int len, double* sig2, int* perm) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    if (i < q) {
      const double ni = __ldcg(norms + i);
      int rank = 0;
      for (int j = 0; j < q; ++j) {
        const double nj = __ldcg(norms + j);
        rank += (nj > ni) || (nj == ni && j < i);
      }
      perm[rank] = i;
      sig2[rank] = ni;
    } else {
      perm[i] = -1;
      sig2[i] = 0.0;
    }
  }
}

}  // namespace dev
}  // namespace qtt

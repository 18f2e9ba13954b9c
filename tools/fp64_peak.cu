// FP64 issue-rate microbenchmark for B200 (sm_100a): DMMA.8x8x4 vs DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
// Prints TFLOP/s for register-resident loops (no memory traffic) so the number is the pipe's
// issue-rate ceiling; this is the denominator DESIGN.md uses for the "FP64 tensor peak".
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_loop(double* out, int iters, double a0, double b0) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters, double a0, double b0) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s sms %d clock_khz %d\n", prop.name, sms, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int bps = 1; bps <= 2; ++bps) {
      int threads = warps * 32, blocks = sms * bps;
      {
        float ms = time_ms([&] { dmma_loop<8><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 5);
        double flops = 2.0 * 256.0 * 8 * (double)iters * warps * blocks;
        printf("DMMA884 nacc=8 warps/blk=%2d blk/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, flops / ms * 1e-9);
      }
      {
        float ms = time_ms([&] { dmma_loop<16><<<blocks, threads>>>(out, iters, 1.0, 1.0); }, 5);
        double flops = 2.0 * 256.0 * 16 * (double)iters * warps * blocks;
        printf("DMMA884 nacc=16 warps/blk=%2d blk/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, flops / ms * 1e-9);
      }
      {
        float ms = time_ms([&] { dfma_loop<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 32.0 * 16 * (double)iters * warps * blocks;
        printf("DFMA    nacc=16 warps/blk=%2d blk/SM=%d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, flops / ms * 1e-9);
      }
    }
  }
  CK(cudaFree(out));
  return 0;
}

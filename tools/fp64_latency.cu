// Dependent-issue latencies of FP64 operations on sm_100a (single warp, clock64 around an unrolled dependent chain).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void lat(double* out, long long* cyc, double x0, double y0, int iters) {
  double x = x0, y = y0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (OP == 0) x = fma(x, y, y);
      if (OP == 1) x = rsqrt(x + 1.5);
      if (OP == 2) x = sqrt(x + 1.5);
      if (OP == 3) x = y / (x + 1.5);
      if (OP == 4) x = (double)rsqrtf((float)(x + 1.5));
      if (OP == 5) { float f = (float)x; f = fmaf(f, 0.5f, 0.25f); x = (double)f; }
      if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
      if (OP == 7) { __syncthreads(); x = x * y; }
      if (OP == 8) x = x * y;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "rsqrt(double)+dadd", "sqrt(double)+dadd", "div(double)+dadd", "cvt+rsqrtf+cvt+dadd", "cvt f64->f32 fma cvt back",
                         "shfl64+dadd", "bar.sync(threads)+dmul", "dmul"};
  for (int threads : {32, 256, 512}) {
    for (int op = 0; op < 9; ++op) {
      const int iters = 64;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: lat<0><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 1: lat<1><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 2: lat<2><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 3: lat<3><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 4: lat<4><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 5: lat<5><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 6: lat<6><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 7: lat<7><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
          case 8: lat<8><<<1, threads>>>(out, cyc, 0.5, 0.999, iters); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("threads %4d  %-28s %8.1f cycles per op\n", threads, names[op], (double)h / (iters * 16));
    }
  }
  return 0;
}

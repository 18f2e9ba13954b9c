// Floor of a cooperative persistent step on B200: back-to-back cooperative launches (148 CTAs x 512 threads, large dynamic
// shared memory) with 0..3 grid barriers of the kind the fused Krylov kernel uses (one release-RED + acquire polling).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_barrier_probe coop_barrier_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void grid_arrive_wait(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    while ((int)(ld_acquire(counter) - target) < 0) {}
  }
  __syncthreads();
}
__global__ void __launch_bounds__(512) probe(unsigned* bar, unsigned epoch, int nbar, double* out) {
  extern __shared__ double sm[];
  sm[threadIdx.x] = threadIdx.x;
  for (int i = 1; i <= nbar; ++i) grid_arrive_wait(bar, epoch + i * gridDim.x);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = sm[5];
}
int main(int argc, char** argv) {
  int G = argc > 1 ? atoi(argv[1]) : 148;
  size_t smem = argc > 2 ? (size_t)atoi(argv[2]) * 1024 : 220 * 1024;
  unsigned* bar; double* out;
  cudaMalloc(&bar, 4); cudaMemset(bar, 0, 4); cudaMalloc(&out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned epoch = 0;
  for (int nbar = 0; nbar <= 3; ++nbar) {
    for (int coop = 1; coop >= 0; --coop) {
      const int reps = 200;
      for (int pass = 0; pass < 2; ++pass) {
        if (pass) cudaEventRecord(e0, st);
        for (int r = 0; r < reps; ++r) {
          void* args[] = {&bar, &epoch, &nbar, &out};
          if (coop) cudaLaunchCooperativeKernel((void*)probe, dim3(G), dim3(512), args, smem, st);
          else probe<<<G, 512, smem, st>>>(bar, epoch, nbar, out);
          epoch += nbar * G;
        }
        if (pass) cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("G=%d smem=%zuKB barriers=%d %s: %.2f us per launch (%s)\n", G, smem / 1024, nbar, coop ? "cooperative" : "plain <<<>>>", ms * 1e3 / reps,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

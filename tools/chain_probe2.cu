// The look-ahead loop of jacobi_gram_kernel (svd_qr.cu) in isolation: one warp computes, the others sit at the step barrier.
// Variants isolate what costs the cycles.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_probe2 chain_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsqrt_approx(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ void grot(double alpha, double beta, double gamma, double& c, double& s) {
  const double d = beta - alpha, g2 = 2.0 * gamma;
  const double r = rsqrt_approx(d * d + g2 * g2);
  const double c2 = 0.5 + 0.5 * (fabs(d) * r);
  const double rc = rsqrt_approx(c2);
  const double c0 = c2 * rc;
  const double s0 = copysign(0.5 * r * rc, d) * g2;
  const double e = fma(c0, c0, fma(s0, s0, -1.0));
  const double k = fma(e, fma(e, 0.375, -0.5), 1.0);
  c = c0 * k; s = s0 * k;
}
template <int B> __device__ __forceinline__ int rpos(int a) { return ((a & (B / 2 - 1)) << 1) | (a / (B / 2)); }
// V: 0 = as in the kernel; 1 = addresses for the next step computed BEFORE the barrier; 2 = no rotation (c, s from the sums)
template <int B, int V>
__global__ void probe(double* out, long long* cyc, int rounds, double thr, double tol2) {
  constexpr int GR = 2 * B, GLD = GR + 1;
  extern __shared__ double sm[];
  double* G0 = sm; double* G1 = sm + GR * GLD;
  double2* cs2 = reinterpret_cast<double2*>(G1 + GR * GLD);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2 * GR * GLD; i += blockDim.x) {
    const int r = (i % (GR * GLD)) / GLD, c = (i % (GR * GLD)) % GLD;
    sm[i] = (r == c) ? 1.0 + 0.01 * r : 0.001 * ((r * 7 + c * 13) % 17 - 8);
  }
  for (int i = tid; i < (B + 1) * B; i += blockDim.x) cs2[i] = make_double2(0.8, 0.6);
  __syncthreads();
  long long t0 = clock64();
  int anyrot = 0;
  for (int rd = 0; rd < rounds; ++rd) {
    if (warp == 0) {
      const int x_ = lane & (B - 1);
      int st = 0;
      int y_ = B + ((x_ + st + 1) & (B - 1)), px = B + ((x_ + st) & (B - 1)), ay = (y_ - B - st) & (B - 1);
#pragma unroll 1
      for (st = 0; st < B; ++st) {
        const double* __restrict__ Gc = (st & 1) ? G1 : G0;
        const double2* rt = cs2 + st * B;
        if (V != 1) { y_ = B + ((x_ + st + 1) & (B - 1)); px = B + ((x_ + st) & (B - 1)); ay = (y_ - B - st) & (B - 1); }
        const double2 rx = rt[rpos<B>(x_)], ry = rt[rpos<B>(ay)];
        const double kx0 = rx.x, kx1 = -rx.y, ky0 = ry.x, ky1 = ry.y;
        const int xa0 = min(x_, ay), xa1 = max(x_, ay), py0 = min(px, y_), py1 = max(px, y_);
        const double gxx = Gc[x_ * GLD + x_], gxp = Gc[x_ * GLD + px], gpp = Gc[px * GLD + px];
        const double gyy = Gc[y_ * GLD + y_], gya = Gc[ay * GLD + y_], gaa = Gc[ay * GLD + ay];
        const double gxy = Gc[x_ * GLD + y_], gxa = Gc[xa0 * GLD + xa1], gpy = Gc[py0 * GLD + py1], gpa = Gc[ay * GLD + px];
        const double al = kx0 * (kx0 * gxx + kx1 * gxp) + kx1 * (kx0 * gxp + kx1 * gpp);
        const double be = ky0 * (ky0 * gyy + ky1 * gya) + ky1 * (ky0 * gya + ky1 * gaa);
        const double gm = kx0 * (ky0 * gxy + ky1 * gxa) + kx1 * (ky0 * gpy + ky1 * gpa);
        double c, s;
        const bool rot = lane < B && st + 1 < B && (al > thr) && (be > thr) && (gm * gm > tol2 * (al * be));
        if (V == 2) { c = al + be; s = gm; } else grot(al, be, gm, c, s);
        if (lane < B) cs2[(st + 1) * B + rpos<B>(lane)] = rot ? make_double2(0.8 + 1e-12 * c, 0.6 + 1e-12 * s) : make_double2(0.8, 0.6);
        anyrot |= rot;
        if (V == 1) { const int s1 = st + 1; y_ = B + ((x_ + s1 + 1) & (B - 1)); px = B + ((x_ + s1) & (B - 1)); ay = (y_ - B - s1) & (B - 1); }
        asm volatile("bar.sync 0;\n" ::: "memory");
      }
    } else {
#pragma unroll 1
      for (int st = 0; st < B; ++st) asm volatile("bar.sync 0;\n" ::: "memory");
    }
  }
  long long t1 = clock64();
  out[tid] = anyrot;
  if (tid == 0) cyc[0] = t1 - t0;
}
template <int B, int V> void run(const char* name, double* out, long long* cyc) {
  const int rounds = 256;
  const size_t smem = (2 * (2 * B) * (2 * B + 1) + 2 * (B + 1) * B) * sizeof(double);
  cudaFuncSetAttribute(probe<B, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    probe<B, V><<<1, 512, smem>>>(out, cyc, rounds, 1e-30, 1e-29);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  }
  printf("B=%d %-50s %8.1f cycles per step   (%s)\n", B, name, (double)h / (rounds * B), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  run<32, 0>("kernel loop", out, cyc);
  run<32, 1>("addresses before the barrier", out, cyc);
  run<32, 2>("no rotation", out, cyc);
  run<16, 0>("kernel loop", out, cyc);
  run<16, 1>("addresses before the barrier", out, cyc);
  run<16, 2>("no rotation", out, cyc);
  return 0;
}

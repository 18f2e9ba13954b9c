// Latency of the look-ahead chain of the Gram Jacobi kernel in isolation (one warp working, N-1 warps waiting at the barrier).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_probe chain_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int B = 32, GLD = 65;
__device__ __forceinline__ void grot(double alpha, double beta, double gamma, double& c, double& s) {
  const double d = beta - alpha, g2 = 2.0 * gamma;
  const double r = rsqrt(d * d + g2 * g2);
  const double c2 = 0.5 + 0.5 * (fabs(d) * r);
  const double rc = rsqrt(c2);
  c = c2 * rc;
  s = copysign(0.5 * r * rc, d) * g2;
  const double e = fma(c, c, fma(s, s, -1.0));
  c = fma(-0.5 * e, c, c);
  s = fma(-0.5 * e, s, s);
}
// MODE 0: full chain; 1: no grot (c, s from the bilinear values directly); 2: loads only + store; 3: barrier only
template <int MODE>
__global__ void probe(double* out, long long* cyc, int iters, double thr, double tol2) {
  __shared__ double G[2 * B * GLD];
  __shared__ double2 cs2[2][B];
  __shared__ int s_any;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2 * B * GLD; i += blockDim.x) {
    const int r = i / GLD, c = i % GLD;
    G[i] = (r == c) ? 1.0 + 0.01 * r : 0.001 * ((r * 7 + c * 13) % 17 - 8);
  }
  if (tid < B) cs2[0][tid] = make_double2(0.8, 0.6);
  __syncthreads();
  long long t0 = clock64();
  double accum = 0.0;
  for (int it = 0; it < iters; ++it) {
    const int st = it & 31;
    const double2* rt = cs2[it & 1];
    if (warp == 0 && MODE != 3) {
      const int x_ = lane, y_ = B + ((x_ + st + 1) & (B - 1));
      const int px = B + ((x_ + st) & (B - 1));
      const int ay = (y_ - B - st) & (B - 1);
      const double2 rx = rt[x_], ry = rt[ay];
      const double kx0 = rx.x, kx1 = -rx.y, ky0 = ry.x, ky1 = ry.y;
      const int xa0 = x_ < ay ? x_ : ay, xa1 = x_ < ay ? ay : x_;
      const int py0 = px < y_ ? px : y_, py1 = px < y_ ? y_ : px;
      const double gxx = G[x_ * GLD + x_], gxp = G[x_ * GLD + px], gpp = G[px * GLD + px];
      const double gyy = G[y_ * GLD + y_], gya = G[ay * GLD + y_], gaa = G[ay * GLD + ay];
      const double gxy = G[x_ * GLD + y_], gxa = G[xa0 * GLD + xa1], gpy = G[py0 * GLD + py1], gpa = G[ay * GLD + px];
      double c = 1.0, s = 0.0;
      if (MODE == 2) {
        c = gxx + gxp + gpp + gyy + gya + gaa + gxy + gxa + gpy + gpa + kx0 + kx1 + ky0 + ky1;
      } else {
        const double al = kx0 * (kx0 * gxx + kx1 * gxp) + kx1 * (kx0 * gxp + kx1 * gpp);
        const double be = ky0 * (ky0 * gyy + ky1 * gya) + ky1 * (ky0 * gya + ky1 * gaa);
        const double gm = kx0 * (ky0 * gxy + ky1 * gxa) + kx1 * (ky0 * gpy + ky1 * gpa);
        if (MODE == 1) { c = al + be; s = gm; }
        else {
          const bool rot = (al > thr) && (be > thr) && (gm * gm > tol2 * (al * be));
          if (rot) grot(al, be, gm, c, s);
          if (__any_sync(0xffffffffu, rot) && lane == 0) s_any = 1;
        }
      }
      cs2[(it + 1) & 1][lane] = make_double2(0.8 + 1e-9 * c, 0.6 + 1e-9 * s);
      accum += c;
    }
    asm volatile("bar.sync 0;\n" ::: "memory");
  }
  long long t1 = clock64();
  out[tid] = accum;
  if (tid == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char* names[] = {"full chain (loads, bilinear, test, grot, store, barrier)", "no grot", "loads + store + barrier", "barrier only"};
  for (int threads : {32, 128, 512}) {
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 4096;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, threads>>>(out, cyc, iters, 1e-30, 1e-29);
        if (mode == 1) probe<1><<<1, threads>>>(out, cyc, iters, 1e-30, 1e-29);
        if (mode == 2) probe<2><<<1, threads>>>(out, cyc, iters, 1e-30, 1e-29);
        if (mode == 3) probe<3><<<1, threads>>>(out, cyc, iters, 1e-30, 1e-29);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("threads %4d  %-60s %8.1f cycles per step\n", threads, names[mode], (double)h / iters);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

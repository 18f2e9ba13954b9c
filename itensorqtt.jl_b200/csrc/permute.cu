// Index permutes as coalesced shared-memory transposes (replaces NDTensors `permutedims`).
// A general N-d permute is executed as a batch of 32x32 tile transposes over the pair
// (input-fastest dim, output-fastest dim): reads are coalesced along the input's fastest index,
// writes along the output's fastest index, the remaining dims are enumerated by blockIdx.z.
#include "common.cuh"
#include "device_utils.cuh"

namespace qtt {
namespace {

constexpr int MAXD = 6;

struct PermArgs {
  int nd;
  long dim[MAXD];   // input extents (row-major)
  long sin[MAXD];   // input strides
  long sout[MAXD];  // output stride of each INPUT dim
  int dx, dy;       // dx: input-fastest dim, dy: output-fastest dim (as input dim index)
  long rest_count;
  long total;
};

__global__ void permute_tiled_kernel(const double* __restrict__ in, double* __restrict__ out, PermArgs a) {
  dev::pdl_wait();
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long x0 = (long)blockIdx.x * 32, y0 = (long)blockIdx.y * 32;
  const long nx = a.dim[a.dx], ny = a.dim[a.dy];
  for (long rest = blockIdx.z; rest < a.rest_count; rest += gridDim.z) {
    long bi = 0, bo = 0, r = rest;
    for (int d = a.nd - 1; d >= 0; --d) {
      if (d == a.dx || d == a.dy) continue;
      long i = r % a.dim[d];
      r /= a.dim[d];
      bi += i * a.sin[d];
      bo += i * a.sout[d];
    }
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      long y = y0 + ty + j, x = x0 + tx;
      if (y < ny && x < nx) tile[ty + j][tx] = in[bi + y * a.sin[a.dy] + x * a.sin[a.dx]];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      long x = x0 + ty + j, y = y0 + tx;
      if (y < ny && x < nx) out[bo + x * a.sout[a.dx] + y * a.sout[a.dy]] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

// fastest dims coincide: plain gather with coalesced reads and writes
__global__ void permute_rows_kernel(const double* __restrict__ in, double* __restrict__ out, PermArgs a, int perm0, int perm1,
                                    int perm2, int perm3, int perm4, int perm5) {
  dev::pdl_wait();
  const int perm[MAXD] = {perm0, perm1, perm2, perm3, perm4, perm5};
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < a.total; o += (long)gridDim.x * blockDim.x) {
    long r = o, off = 0;
    for (int k = a.nd - 1; k >= 0; --k) {
      int d = perm[k];
      long i = r % a.dim[d];
      r /= a.dim[d];
      off += i * a.sin[d];
    }
    out[o] = in[off];
  }
}

__global__ void deinterleave_kernel(const double2* __restrict__ in, double* __restrict__ re, double* __restrict__ im, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = in[i];
    re[i] = v.x;
    im[i] = v.y;
  }
}
__global__ void interleave_kernel(const double* __restrict__ re, const double* __restrict__ im, double2* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_double2(re[i], im[i]);
}

}  // namespace

void permute(Ctx* ctx, const double* in, double* out, int nd, const int64_t* dims, const int* perm) {
  QTT_REQUIRE(nd >= 1 && nd <= MAXD, "permute: rank %d unsupported", nd);
  PermArgs a{};
  a.nd = nd;
  long total = 1;
  for (int d = 0; d < nd; ++d) {
    a.dim[d] = dims[d];
    total *= dims[d];
  }
  if (total == 0) return;
  a.total = total;
  long s = 1;
  for (int d = nd - 1; d >= 0; --d) {
    a.sin[d] = s;
    s *= dims[d];
  }
  s = 1;
  bool identity = true;
  for (int k = nd - 1; k >= 0; --k) {
    a.sout[perm[k]] = s;
    s *= dims[perm[k]];
    if (perm[k] != k) identity = false;
  }
  if (identity) {
    QTT_CUDA(cudaMemcpyAsync(out, in, sizeof(double) * total, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  if (perm[nd - 1] == nd - 1) {
    int pp[MAXD] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nd; ++k) pp[k] = perm[k];
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(ctx, "permute_rows", permute_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, in, out, a, pp[0], pp[1], pp[2], pp[3], pp[4], pp[5]);
    return;
  }
  a.dx = nd - 1;
  a.dy = perm[nd - 1];
  a.rest_count = total / (a.dim[a.dx] * a.dim[a.dy]);
  dim3 grid((unsigned)((a.dim[a.dx] + 31) / 32), (unsigned)((a.dim[a.dy] + 31) / 32),
            (unsigned)(a.rest_count > 32768 ? 32768 : a.rest_count));
  launch_pdl(ctx, "permute_tiled", permute_tiled_kernel, dim3(grid), dim3(dim3(32, 8)), 0, in, out, a);
}

void deinterleave(Ctx* ctx, const double* in, double* re, double* im, size_t n) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  deinterleave_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const double2*>(in), re, im, n);
  check_launch(ctx, "deinterleave");
}

void interleave(Ctx* ctx, const double* re, const double* im, double* out, size_t n) {
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  interleave_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(re, im, reinterpret_cast<double2*>(out), n);
  check_launch(ctx, "interleave");
}

}  // namespace qtt

// FP64 tensor-core GEMM core for sm_100a: C = alpha * op(A) * op(B) + beta * C, row-major, strided-batched.
//
// Every dense contraction of the sweep engine (the L.W1.W2.R.psi matvec, environment updates, two-site
// merge, U^T.phi after the split, MPO.MPS apply) is lowered onto this kernel (replaces NDTensors'
// permutedims + BLAS dgemm behind every ITensor `*`).
//
// Hardware mapping (measured on B200, profiles/r01_fp64_peak_dmma_dfma.txt): FP64 tensor work on sm_100a is
// DMMA.8x8x4 (mma.sync.m8n8k4.f64), 37.0 TFLOP/s chip-wide, reached with >= 2 warps per SM sub-partition and
// >= 8 independent accumulators per warp.  Each warp therefore owns a register tile of (WTM/8) x (WTN/8)
// m8n8 accumulators; A/B fragments are one double per lane, fetched from shared memory with LDS.64.
// Shared-memory tiles are padded by 4 doubles so that the 16 lanes of each 64-bit wavefront hit 16
// distinct bank pairs in all four operand layouts (row stride = 4 mod 16 doubles).
// Global -> shared staging is a STAGES-deep cp.async (LDGSTS) pipeline with zero-fill predication, so
// ragged edge bonds (chi = 1, 2, 3, ... and odd MPO bond dimensions) take the same code path.
#include "common.cuh"
#include "device_utils.cuh"

namespace qtt {

namespace {

constexpr int BK = 16;

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  long lda, ldb, ldc;
  int batch;
  long sA, sB, sC, sA2, sB2, sC2;
  double alpha, beta;
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Copies a (ROWS x COLS) tile whose rows are contiguous in global memory into a padded smem tile.
template <int ROWS, int COLS, int LD, int VEC, int NT>
__device__ __forceinline__ void load_tile(double* smem, const double* g, long ldg, int row0, int col0, int row_lim, int col_lim,
                                          int tid) {
  constexpr int CPR = COLS / VEC;  // chunks per row
  constexpr int TOTAL = ROWS * CPR;
#pragma unroll
  for (int id0 = 0; id0 < TOTAL; id0 += NT) {
    int id = id0 + tid;
    if (TOTAL % NT != 0 && id >= TOTAL) break;
    int r = id / CPR;
    int c = (id % CPR) * VEC;
    int gr = row0 + r, gc = col0 + c;
    int valid = 0;
    if (gr < row_lim) {
      valid = col_lim - gc;
      valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
    }
    const double* src = valid > 0 ? (g + (long)gr * ldg + gc) : g;
    cp_async<VEC * 8>(smem + r * LD + c, src, valid * 8);
  }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool TA, bool TB, int VEC>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32) gemm_dmma_kernel(GemmArgs p) {
  constexpr int NT = WARPS_M * WARPS_N * 32;
  constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
  constexpr int MT = WTM / 8, NTL = WTN / 8;
  constexpr int A_ROWS = TA ? BK : BM, A_COLS = TA ? BM : BK, A_LD = A_COLS + 4;
  constexpr int B_ROWS = TB ? BN : BK, B_COLS = TB ? BK : BN, B_LD = B_COLS + 4;
  constexpr int A_STAGE = A_ROWS * A_LD, B_STAGE = B_ROWS * B_LD;

  dev::pdl_wait();
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / WARPS_N) * WTM, wn0 = (warp % WARPS_N) * WTN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int b1 = blockIdx.z % p.batch, b2 = blockIdx.z / p.batch;
  const double* A = p.A + b1 * p.sA + b2 * p.sA2;
  const double* B = p.B + b1 * p.sB + b2 * p.sB2;
  double* C = p.C + b1 * p.sC + b2 * p.sC2;

  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (p.K + BK - 1) / BK;

  auto issue = [&](int kt) {
    if (kt < KT) {
      const int s = kt % STAGES;
      const int k0 = kt * BK;
      if constexpr (TA)
        load_tile<A_ROWS, A_COLS, A_LD, VEC, NT>(As + s * A_STAGE, A, p.lda, k0, m0, p.K, p.M, tid);
      else
        load_tile<A_ROWS, A_COLS, A_LD, VEC, NT>(As + s * A_STAGE, A, p.lda, m0, k0, p.M, p.K, tid);
      if constexpr (TB)
        load_tile<B_ROWS, B_COLS, B_LD, VEC, NT>(Bs + s * B_STAGE, B, p.ldb, n0, k0, p.N, p.K, tid);
      else
        load_tile<B_ROWS, B_COLS, B_LD, VEC, NT>(Bs + s * B_STAGE, B, p.ldb, k0, n0, p.K, p.N, tid);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const double* as = As + (kt % STAGES) * A_STAGE;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[MT], b[NTL];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        if constexpr (TA)
          a[i] = as[(kk * 4 + t) * A_LD + wm0 + i * 8 + g];
        else
          a[i] = as[(wm0 + i * 8 + g) * A_LD + kk * 4 + t];
      }
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        if constexpr (TB)
          b[j] = bs[(wn0 + j * 8 + g) * B_LD + kk * 4 + t];
        else
          b[j] = bs[(kk * 4 + t) * B_LD + wn0 + j * 8 + g];
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each lane owns C[row g][cols 2t, 2t+1] of every m8n8 tile
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int row = m0 + wm0 + i * 8 + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NTL; ++j) {
      const int col = n0 + wn0 + j * 8 + 2 * t;
      if (col >= p.N) continue;
      double* cp = C + (long)row * p.ldc + col;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (VEC == 2 && col + 1 < p.N) {
        if (beta != 0.0) {
          double2 old = *reinterpret_cast<const double2*>(cp);
          v0 += beta * old.x;
          v1 += beta * old.y;
        }
        *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
      } else {
        if (beta != 0.0) v0 += beta * cp[0];
        cp[0] = v0;
        if (col + 1 < p.N) {
          if (beta != 0.0) v1 += beta * cp[1];
          cp[1] = v1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- tile-config table
struct TileCfg {
  int bm, bn, threads, stages;
  int occ;       // resident CTAs per SM (shared-memory bound)
  double util;   // fraction of the SM's DMMA rate one resident CTA sustains (calibrated on B200)
  const char* name;
};

constexpr int NUM_TILES = 4;
const TileCfg kTiles[NUM_TILES] = {
    {128, 128, 256, 3, 1, 0.85, "128x128x16_w2x4_s3"},
    {64, 64, 128, 4, 3, 0.45, "64x64x16_w2x2_s4"},
    {32, 32, 128, 4, 5, 0.25, "32x32x16_w2x2_s4"},
    {128, 64, 256, 3, 2, 0.70, "128x64x16_w4x2_s3"},
};

template <int BM, int BN, int STAGES, bool TA, bool TB>
constexpr size_t smem_bytes() {
  constexpr int A_STAGE = TA ? BK * (BM + 4) : BM * (BK + 4);
  constexpr int B_STAGE = TB ? BN * (BK + 4) : BK * (BN + 4);
  return size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
}

template <int BM, int BN, int WM, int WN, int STAGES, bool TA, bool TB, int VEC>
void launch_cfg(Ctx* ctx, const GemmArgs& a, int batch_total) {
  auto kern = gemm_dmma_kernel<BM, BN, WM, WN, STAGES, TA, TB, VEC>;
  constexpr size_t smem = smem_bytes<BM, BN, STAGES, TA, TB>();
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, batch_total);
  launch_pdl(ctx, "gemm_dmma", kern, grid, dim3(WM * WN * 32), smem, a);
}

template <bool TA, bool TB, int VEC>
void launch_layout(Ctx* ctx, int tile, const GemmArgs& a, int batch_total) {
  switch (tile) {
    case 0: launch_cfg<128, 128, 2, 4, 3, TA, TB, VEC>(ctx, a, batch_total); break;
    case 1: launch_cfg<64, 64, 2, 2, 4, TA, TB, VEC>(ctx, a, batch_total); break;
    case 2: launch_cfg<32, 32, 2, 2, 4, TA, TB, VEC>(ctx, a, batch_total); break;
    case 3: launch_cfg<128, 64, 4, 2, 3, TA, TB, VEC>(ctx, a, batch_total); break;
    default: fail(QTT_ERR_ARG, "bad gemm tile index %d", tile);
  }
}

int pick_tile(const Ctx* ctx, int M, int N, int K, int batch_total) {
  // Busiest-SM time model: q tiles on the busiest SM, run min(q, occ) at a time; a lone CTA only
  // sustains `util` of the SM's DMMA rate, so throughput = min(1, concurrency * util).
  (void)K;
  const int sms = ctx->sms > 0 ? ctx->sms : 148;
  double best = 1e300;
  int best_i = 1;
  for (int i = 0; i < NUM_TILES; ++i) {
    const TileCfg& c = kTiles[i];
    long tiles = (long)((M + c.bm - 1) / c.bm) * ((N + c.bn - 1) / c.bn) * batch_total;
    long q = (tiles + sms - 1) / sms;
    double conc = (double)(q < c.occ ? q : c.occ);
    double thr = conc * c.util;
    if (thr > 1.0) thr = 1.0;
    double t = (double)q * c.bm * c.bn / thr;
    if (t < best) {
      best = t;
      best_i = i;
    }
  }
  return best_i;
}

}  // namespace

const char* gemm_tile_name(int tile) { return (tile >= 0 && tile < NUM_TILES) ? kTiles[tile].name : "?"; }
int gemm_num_tiles() { return NUM_TILES; }

void gemm(Ctx* ctx, const GemmDesc& d) {
  if (d.M <= 0 || d.N <= 0 || d.batch <= 0 || d.batch2 <= 0) return;
  QTT_REQUIRE(d.K >= 0 && d.A && d.B && d.C, "gemm: bad arguments");
  GemmArgs a;
  a.A = d.A; a.B = d.B; a.C = d.C;
  a.M = d.M; a.N = d.N; a.K = d.K;
  a.lda = d.lda; a.ldb = d.ldb; a.ldc = d.ldc;
  a.batch = d.batch;
  a.sA = d.sA; a.sB = d.sB; a.sC = d.sC;
  a.sA2 = d.sA2; a.sB2 = d.sB2; a.sC2 = d.sC2;
  a.alpha = d.alpha; a.beta = d.beta;
  const int batch_total = d.batch * d.batch2;
  // 16-byte vector path only when every row start and every batch slice is 16-byte aligned
  auto even = [](long v) { return (v & 1L) == 0; };
  bool vec2 = even(d.lda) && even(d.ldb) && even(d.ldc) && even(d.sA) && even(d.sB) && even(d.sC) && even(d.sA2) &&
              even(d.sB2) && even(d.sC2) && ((reinterpret_cast<uintptr_t>(d.A) & 15) == 0) &&
              ((reinterpret_cast<uintptr_t>(d.B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d.C) & 15) == 0);
  int tile = d.tile >= 0 ? d.tile : pick_tile(ctx, d.M, d.N, d.K, batch_total);
  if (vec2) {
    if (d.transA && d.transB) fail(QTT_ERR_UNSUPPORTED, "gemm: TT layout not instantiated");
    if (d.transA) launch_layout<true, false, 2>(ctx, tile, a, batch_total);
    else if (d.transB) launch_layout<false, true, 2>(ctx, tile, a, batch_total);
    else launch_layout<false, false, 2>(ctx, tile, a, batch_total);
  } else {
    if (d.transA && d.transB) fail(QTT_ERR_UNSUPPORTED, "gemm: TT layout not instantiated");
    if (d.transA) launch_layout<true, false, 1>(ctx, tile, a, batch_total);
    else if (d.transB) launch_layout<false, true, 1>(ctx, tile, a, batch_total);
    else launch_layout<false, false, 1>(ctx, tile, a, batch_total);
  }
}

}  // namespace qtt

// Two-site effective-operator matvec out = L.W1.W2.R.v as TWO fused FP64 tensor-core kernels
// (replaces upstream `product(P::ProjMPO, v)` = noprime(v*L*W1*W2*R), pattern at reference src/linsolve.jl:71-75,283-288).
//
//   K_A  T2[a'][t1 t2][wr][c] = sum_{wl,s1,s2} W12[(wl s1 s2)][(t1 t2 wr)] * ( sum_a L[wl][a'][a] v[a][s1 s2][c] )
//        The inner sum is a DMMA GEMM over a (K = chi_l).  A CTA owns a block of (a', c) and computes it for ALL wl and
//        all four (s1 s2): its tile is (wl*BA) x (4*BC) with gathered rows/columns, and the warp tiling is arranged so
//        that every lane holds all 4*wl partial sums of "its" (a', c) entries.  The skinny contraction with W12 = W1.W2
//        (K = 4*wl: not tensor-core shaped) is therefore a per-lane register epilogue, and the intermediate
//        T1[wl][a'][s12][c] of the unfused path (6 MiB written + read at chi = 256) never exists.
//   K_B  out[(a' t12)][c'] = sum_{(wr c)} T2[(a' t12)][(wr c)] R[(wr c)][c']: a plain GEMM with a small output
//        (4 chi_l x chi_r) and a long K (wr*chi_r).  To fill 148 SMs with 2 warps per scheduler it uses 32 x 64 CTA tiles
//        and splits K two ways INSIDE the CTA (two groups of 4 warps), combining through shared memory in a fixed order:
//        no partial-sum buffer, no extra launch, bit-reproducible.
// Both kernels stage operands with a 4-deep cp.async pipeline into padded shared-memory tiles (row stride = 4 mod 16
// doubles: conflict-free LDS.64 fragment loads) and zero-fill ragged edges, so edge bonds take the same path.
#include "common.cuh"
#include <cstdlib>
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

namespace qtt {
namespace {

constexpr int BK = 16;
constexpr int STAGES = 4;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct MvAArgs {
  const double* L;    // [wl][chil][chil]  (w, a', a)
  const double* v;    // [chil][4][chir]   (a, s12, c)
  const double* W12;  // [4 wl][4 wr]
  double* T2;         // [chil][4][wr][chir]
  int chil, chir, wl, wr;
};

// WL: the left MPO bond dimension (exact); CTA footprint (8*MI*WM) x (8*WN) in (a', c), a warp owns MI a'-subtiles of
// 8 x 8; VEC = 2 when 16-byte copies are legal.  The per-thread copy descriptors (source pointer, smem offset, edge
// predicate) are computed once; per k-tile a copy costs one LDGSTS and two integer instructions.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp16(unsigned dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc));
}
__device__ __forceinline__ void cp16z(unsigned dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp8z(unsigned dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(src_bytes));
}
struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

template <int WL, int MI, int WM, int WN, int VEC>
__global__ void __launch_bounds__(WM* WN * 32) matvec_a_kernel(MvAArgs p) {
  constexpr int NT = WM * WN * 32;
  constexpr int BA = 8 * MI * WM, BC = 8 * WN;
  constexpr int A_LD = BK + 4;            // As[(w, a')][k]
  constexpr int B_LD = 4 * BC + 4;        // Bs[k][(s, c)]
  constexpr int A_STAGE = WL * BA * A_LD;
  constexpr int B_STAGE = BK * B_LD;
  constexpr int A_CPR = BK / VEC, A_TOTAL = WL * BA * A_CPR, NA = (A_TOTAL + NT - 1) / NT;
  constexpr int B_CPS = BC / VEC, B_TOTAL = BK * 4 * B_CPS, NB_ = (B_TOTAL + NT - 1) / NT;
  constexpr bool A_EXACT = (A_TOTAL % NT == 0), B_EXACT = (B_TOTAL % NT == 0);
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  double* Ws = Bs + STAGES * B_STAGE;     // W12 [4 wl][4 wr]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wa0 = (warp / WN) * 8 * MI, wc0 = (warp % WN) * 8;
  const int a0 = blockIdx.x * BA, c0 = blockIdx.y * BC;
  const int wr = p.wr, chil = p.chil, chir = p.chir;

  // programmatic dependent launch: K_B's CTAs may become resident (and run their prologue) while this grid drains
  asm volatile("griddepcontrol.launch_dependents;\n" ::);
  // ... and this grid is itself launched programmatically behind the producer of v
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  for (int i = tid; i < 16 * WL * wr; i += NT) Ws[i] = p.W12[i];

  // ---- copy descriptors, computed once: global source (advanced per k-tile), shared destination of stage 0
  const double* a_g[NA];
  unsigned a_s[NA];
  int a_kleft[NA];   // valid doubles from this chunk's k to the end of the row (0: row out of range / unused slot)
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int id = tid + i * NT;
    const int r = id / A_CPR, ck = (id % A_CPR) * VEC;
    const int w = r / BA, ar = r % BA;
    const int ga = a0 + ar;
    const bool used = A_EXACT || (id < A_TOTAL);
    const bool ok = used && (ga < chil);
    a_g[i] = ok ? (p.L + ((size_t)w * chil + ga) * chil + ck) : p.L;
    a_s[i] = used ? smem_u32(As + r * A_LD + ck) : 0xffffffffu;
    a_kleft[i] = ok ? (chil - ck) : 0;
  }
  const double* b_g[NB_];
  unsigned b_s[NB_];
  int b_kr[NB_], b_cv[NB_];
#pragma unroll
  for (int i = 0; i < NB_; ++i) {
    const int id = tid + i * NT;
    const int kr = id / (4 * B_CPS), rem = id % (4 * B_CPS);
    const int sgm = rem / B_CPS, cc = (rem % B_CPS) * VEC;
    const int gc = c0 + cc;
    const bool used = B_EXACT || (id < B_TOTAL);
    int cv = chir - gc;
    cv = cv < 0 ? 0 : (cv > VEC ? VEC : cv);
    if (!used) cv = 0;
    b_g[i] = cv > 0 ? (p.v + ((size_t)kr * 4 + sgm) * chir + gc) : p.v;
    b_s[i] = used ? smem_u32(Bs + kr * B_LD + sgm * BC + cc) : 0xffffffffu;
    b_kr[i] = kr;
    b_cv[i] = cv;
  }
  const size_t b_step = (size_t)BK * 4 * chir;

  double acc[WL][MI][4][2];
#pragma unroll
  for (int w = 0; w < WL; ++w)
#pragma unroll
    for (int m = 0; m < MI; ++m)
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) acc[w][m][sg][0] = acc[w][m][sg][1] = 0.0;

  const int KT = (chil + BK - 1) / BK;
  const int a_frag = (wa0 + g) * A_LD + t;
  const int b_frag = t * B_LD + wc0 + g;
  // interior CTA with whole k-tiles: unpredicated 16-byte copies
  const bool full = (VEC == 2) && (a0 + BA <= chil) && (c0 + BC <= chir) && (chil % BK == 0);

  auto run = [&](auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    int kt_load = 0, st_load = 0;
    auto issue = [&]() {
      if (kt_load < KT) {
        const unsigned aoff = st_load * (A_STAGE * 8), boff = st_load * (B_STAGE * 8);
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          if (A_EXACT || a_s[i] != 0xffffffffu) {
            if constexpr (FULL) {
              cp16(a_s[i] + aoff, a_g[i]);
            } else {
              int valid = a_kleft[i];
              valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
              const double* src = valid > 0 ? a_g[i] : p.L;
              if constexpr (VEC == 2) cp16z(a_s[i] + aoff, src, valid * 8);
              else cp8z(a_s[i] + aoff, src, valid * 8);
              a_kleft[i] -= BK;
            }
            a_g[i] += BK;
          }
        }
#pragma unroll
        for (int i = 0; i < NB_; ++i) {
          if (B_EXACT || b_s[i] != 0xffffffffu) {
            if constexpr (FULL) {
              cp16(b_s[i] + boff, b_g[i]);
            } else {
              const int valid = (kt_load * BK + b_kr[i] < chil) ? b_cv[i] : 0;
              const double* src = valid > 0 ? b_g[i] : p.v;
              if constexpr (VEC == 2) cp16z(b_s[i] + boff, src, valid * 8);
              else cp8z(b_s[i] + boff, src, valid * 8);
            }
            b_g[i] += b_step;
          }
        }
      }
      cp_async_commit();
      ++kt_load;
      st_load = (st_load + 1 == STAGES) ? 0 : st_load + 1;
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue();
    int st_use = 0;
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue();
      const double* as = As + st_use * A_STAGE + a_frag;
      const double* bs = Bs + st_use * B_STAGE + b_frag;
      st_use = (st_use + 1 == STAGES) ? 0 : st_use + 1;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[WL][MI], b[4];
#pragma unroll
        for (int w = 0; w < WL; ++w)
#pragma unroll
          for (int m = 0; m < MI; ++m) a[w][m] = as[(w * BA + m * 8) * A_LD + kk * 4];
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) b[sg] = bs[(kk * 4) * B_LD + sg * BC];
#pragma unroll
        for (int w = 0; w < WL; ++w)
#pragma unroll
          for (int m = 0; m < MI; ++m)
#pragma unroll
            for (int sg = 0; sg < 4; ++sg) dmma884(acc[w][m][sg][0], acc[w][m][sg][1], a[w][m], b[sg]);
      }
    }
    cp_async_wait<0>();
  };
  if (full) run(TrueT{});
  else run(FalseT{});

  // epilogue: this lane owns (a' = a0 + wa0 + 8m + g, c = c0 + wc0 + 2t, 2t+1) for every (w, s12)
  const int gc = c0 + wc0 + 2 * t;
  if (gc >= chir) return;
  const int nout = 4 * wr;
  const bool pair = (gc + 1 < chir);
#pragma unroll
  for (int m = 0; m < MI; ++m) {
    const int ga = a0 + wa0 + m * 8 + g;
    if (ga >= chil) continue;
    for (int o = 0; o < nout; ++o) {  // o = (t12, wr)
      double o0 = 0.0, o1 = 0.0;
#pragma unroll
      for (int w = 0; w < WL; ++w)
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) {
          const double c = Ws[(w * 4 + sg) * nout + o];
          o0 += c * acc[w][m][sg][0];
          o1 += c * acc[w][m][sg][1];
        }
      double* dst = p.T2 + ((size_t)ga * nout + o) * chir + gc;
      if (VEC == 2 && pair) *reinterpret_cast<double2*>(dst) = make_double2(o0, o1);
      else {
        dst[0] = o0;
        if (pair) dst[1] = o1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K_B: C (M x N) = A (M x K, lda) . B (K x N, ldb), row-major, CTA tile 32 x 64, 8 warps = 2 K-groups x (2 x 2) warps of
// 16 x 32; K-group kg handles the k-tiles kt with kt % 2 == kg; the two partial tiles are combined through shared memory.
struct MvBArgs {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  long lda, ldb, ldc;
};

template <int VEC, int BN>
__global__ void __launch_bounds__(256) matvec_b_kernel(MvBArgs p) {
  constexpr int BM = 32, NT = 256, NJ = BN / 16;  // each warp of a K-group owns 16 rows x BN/2 columns
  constexpr int A_LD = BK + 4, B_LD = BN + 4;
  constexpr int A_STAGE = BM * A_LD, B_STAGE = BK * B_LD;
  constexpr int ST = 6;  // stages; two k-tiles (one per K-group) are consumed per barrier
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + ST * A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kg = warp >> 2;           // K-group
  const int w4 = warp & 3;
  const int wm0 = (w4 >> 1) * 16, wn0 = (w4 & 1) * (BN / 2);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  double acc[2][NJ][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (p.K + BK - 1) / BK;
  constexpr int A_CPR = BK / VEC, A_TOTAL = BM * A_CPR, NA = (A_TOTAL + NT - 1) / NT;
  constexpr int B_CPR = BN / VEC, B_TOTAL = BK * B_CPR, NB_ = (B_TOTAL + NT - 1) / NT;
  constexpr bool A_EXACT = (A_TOTAL % NT == 0), B_EXACT = (B_TOTAL % NT == 0);
  const double* a_g[NA];
  unsigned a_s[NA];
  int a_kleft[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int id = tid + i * NT;
    const int r = id / A_CPR, ck = (id % A_CPR) * VEC;
    const int gr = m0 + r;
    const bool used = A_EXACT || (id < A_TOTAL);
    const bool ok = used && (gr < p.M);
    a_g[i] = ok ? (p.A + (size_t)gr * p.lda + ck) : p.A;
    a_s[i] = used ? smem_u32(As + r * A_LD + ck) : 0xffffffffu;
    a_kleft[i] = ok ? (p.K - ck) : 0;
  }
  const double* b_g[NB_];
  unsigned b_s[NB_];
  int b_kr[NB_], b_cv[NB_];
#pragma unroll
  for (int i = 0; i < NB_; ++i) {
    const int id = tid + i * NT;
    const int r = id / B_CPR, cc = (id % B_CPR) * VEC;
    const int gc = n0 + cc;
    const bool used = B_EXACT || (id < B_TOTAL);
    int cv = p.N - gc;
    cv = cv < 0 ? 0 : (cv > VEC ? VEC : cv);
    if (!used) cv = 0;
    b_g[i] = cv > 0 ? (p.B + (size_t)r * p.ldb + gc) : p.B;
    b_s[i] = used ? smem_u32(Bs + r * B_LD + cc) : 0xffffffffu;
    b_kr[i] = r;
    b_cv[i] = cv;
  }
  const size_t b_step = (size_t)BK * p.ldb;
  const int a_frag = (wm0 + g) * A_LD + t;
  const int b_frag = t * B_LD + wn0 + g;
  const bool full = (VEC == 2) && (m0 + BM <= p.M) && (n0 + BN <= p.N) && (p.K % BK == 0);
  asm volatile("griddepcontrol.wait;\n" ::: "memory");  // T2 is produced by the preceding matvec_a grid

  auto run = [&](auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    int kt_load = 0, st_load = 0;
    auto issue = [&]() {
      if (kt_load < KT) {
        const unsigned aoff = st_load * (A_STAGE * 8), boff = st_load * (B_STAGE * 8);
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          if (A_EXACT || a_s[i] != 0xffffffffu) {
            if constexpr (FULL) {
              cp16(a_s[i] + aoff, a_g[i]);
            } else {
              int valid = a_kleft[i];
              valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
              const double* src = valid > 0 ? a_g[i] : p.A;
              if constexpr (VEC == 2) cp16z(a_s[i] + aoff, src, valid * 8);
              else cp8z(a_s[i] + aoff, src, valid * 8);
              a_kleft[i] -= BK;
            }
            a_g[i] += BK;
          }
        }
#pragma unroll
        for (int i = 0; i < NB_; ++i) {
          if (B_EXACT || b_s[i] != 0xffffffffu) {
            if constexpr (FULL) {
              cp16(b_s[i] + boff, b_g[i]);
            } else {
              const int valid = (kt_load * BK + b_kr[i] < p.K) ? b_cv[i] : 0;
              const double* src = valid > 0 ? b_g[i] : p.B;
              if constexpr (VEC == 2) cp16z(b_s[i] + boff, src, valid * 8);
              else cp8z(b_s[i] + boff, src, valid * 8);
            }
            b_g[i] += b_step;
          }
        }
      }
      cp_async_commit();
      ++kt_load;
      st_load = (st_load + 1 == ST) ? 0 : st_load + 1;
    };
#pragma unroll
    for (int s = 0; s < ST - 2; ++s) issue();
    int st_use = kg;  // stage of this K-group's tile in the current pair
    for (int kt = 0; kt < KT; kt += 2) {
      cp_async_wait<ST - 4>();
      __syncthreads();
      issue();
      issue();
      if (kt + kg < KT) {
        const double* as = As + st_use * A_STAGE + a_frag;
        const double* bs = Bs + st_use * B_STAGE + b_frag;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
          double a[2], b[NJ];
#pragma unroll
          for (int i = 0; i < 2; ++i) a[i] = as[(i * 8) * A_LD + kk * 4];
#pragma unroll
          for (int j = 0; j < NJ; ++j) b[j] = bs[(kk * 4) * B_LD + j * 8];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
      st_use = (st_use + 2 >= ST) ? st_use + 2 - ST : st_use + 2;
    }
    cp_async_wait<0>();
  };
  if (full) run(TrueT{});
  else run(FalseT{});
  __syncthreads();
  // combine the two K-groups: group 1 parks its tile in shared memory, group 0 adds (fixed order) and stores
  double* red = smem;  // [4 warps][2][NJ][32 lanes] double2
  if (kg == 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        *reinterpret_cast<double2*>(red + (((w4 * 2 + i) * NJ + j) * 32 + lane) * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
  }
  __syncthreads();
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = m0 + wm0 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const double2 o = *reinterpret_cast<const double2*>(red + (((w4 * 2 + i) * NJ + j) * 32 + lane) * 2);
        const double v0 = acc[i][j][0] + o.x, v1 = acc[i][j][1] + o.y;
        const int col = n0 + wn0 + j * 8 + 2 * t;
        if (row < p.M && col < p.N) {
          double* cp = p.C + (size_t)row * p.ldc + col;
          if (VEC == 2 && col + 1 < p.N) *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
          else {
            cp[0] = v0;
            if (col + 1 < p.N) cp[1] = v1;
          }
        }
      }
    }
  }
}

// ---- K_B with TMA-staged tiles -------------------------------------------------------------------------------------------
// Same tile shape, warp layout, in-CTA split-K and epilogue as matvec_b_kernel<2, 64>; the operand tiles are fetched by
// `cp.async.bulk.tensor` (SASS UTMALDG) issued by ONE thread and tracked by mbarrier transaction counts instead of 256 threads
// issuing LDGSTS.  Layouts that keep the m8n8k4 fragment loads at the two-wavefront minimum WITHOUT padding (a tensor-map box
// cannot be padded):
//   A tile (32 rows x 16 k, 128-byte rows): one box with the 128-byte swizzle: element (r, c) sits at chunk (c >> 1) ^ (r & 7);
//   B tile (16 k x 64 columns): eight boxes of 16 rows x 8 columns (64-byte rows, no swizzle): the four k rows of a fragment
//   start 64 bytes apart, so the 32 lanes of a warp cover every bank exactly twice.
// Out-of-range rows / columns / k are zero-filled by the TMA unit (edge bonds need no predicates).
constexpr int TMB_ST = 6;
constexpr int TMB_A_BYTES = 32 * BK * 8;        // 4096
constexpr int TMB_B_BYTES = BK * 64 * 8;        // 8192
constexpr int TMB_STAGE = TMB_A_BYTES + TMB_B_BYTES;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* tm, unsigned bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y) : "memory");
}

struct MvBTmaArgs {
  double* C;
  int M, N, K;
  long ldc;
};

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

// 9 warps: warps 0-7 compute (two K-groups of 2 x 2 warps, as in matvec_b_kernel), warp 8 is the TMA producer.  Stage s holds
// k-tile kt with kt % 6 == s, so even stages always belong to K-group 0 and odd ones to K-group 1: the groups run decoupled, each
// stage has a `full` barrier (transaction count) and an `empty` barrier (one arrival per consumer warp of its group); there is no
// CTA-wide barrier in the main loop.
__global__ void __launch_bounds__(288) matvec_b_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                           MvBTmaArgs p) {
  constexpr int BM = 32, BN = 64, NJ = BN / 16;
  extern __shared__ __align__(16) unsigned char raw[];
  // the 128-byte swizzle atom is 1024 bytes: align the stage buffers by hand (the dynamic segment is only 16-byte aligned by contract)
  const unsigned base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char* sbuf = raw + (base - smem_u32(raw));
  __shared__ __align__(8) unsigned long long full[TMB_ST], empty[TMB_ST];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kg = (warp >> 2) & 1;     // K-group of a compute warp
  const int w4 = warp & 3;
  const int wm0 = (w4 >> 1) * 16, wn0 = (w4 & 1) * (BN / 2);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (tid == 0) {
    for (int s = 0; s < TMB_ST; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  const int KT = (p.K + BK - 1) / BK;
  asm volatile("griddepcontrol.wait;\n" ::: "memory");  // T2 is produced by the preceding matvec_a grid

  double acc[2][NJ][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (warp == 8) {
    if (lane == 0) {
      for (int kt = 0; kt < KT; ++kt) {
        const int st = kt % TMB_ST, use = kt / TMB_ST;
        if (use > 0) mbar_wait(smem_u32(&empty[st]), (unsigned)(use - 1) & 1u);
        const unsigned bar = smem_u32(&full[st]);
        const unsigned dstA = base + st * TMB_STAGE, dstB = dstA + TMB_A_BYTES;
        mbar_expect_tx(bar, TMB_STAGE);
        tma_load_2d(dstA, &tmA, bar, kt * BK, m0);
#pragma unroll
        for (int b = 0; b < 8; ++b) tma_load_2d(dstB + b * 1024, &tmB, bar, n0 + 8 * b, kt * BK);
      }
    }
  } else {
    // fragment offsets (doubles) inside a stage
    int a_off[2][BK / 4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) a_off[i][kk] = (wm0 + i * 8 + g) * BK + ((((kk * 2 + (t >> 1)) ^ g) << 1) | (t & 1));
    const int b_off = (wn0 >> 3) * 128 + t * 8 + g;
    for (int kt = kg; kt < KT; kt += 2) {
      const int st = kt % TMB_ST;
      mbar_wait(smem_u32(&full[st]), (unsigned)(kt / TMB_ST) & 1u);
      const double* as = reinterpret_cast<const double*>(sbuf + st * TMB_STAGE);
      const double* bs = reinterpret_cast<const double*>(sbuf + st * TMB_STAGE + TMB_A_BYTES) + b_off;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[2], b[NJ];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = as[a_off[i][kk]];
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = bs[j * 128 + kk * 32];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[st]));
    }
  }
  __syncthreads();
  // combine the two K-groups: group 1 parks its tile in shared memory, group 0 adds (fixed order) and stores
  double* red = reinterpret_cast<double*>(sbuf);  // [4 warps][2][NJ][32 lanes] double2
  if (warp < 8 && kg == 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        *reinterpret_cast<double2*>(red + (((w4 * 2 + i) * NJ + j) * 32 + lane) * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
  }
  __syncthreads();
  if (warp < 8 && kg == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = m0 + wm0 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const double2 o = *reinterpret_cast<const double2*>(red + (((w4 * 2 + i) * NJ + j) * 32 + lane) * 2);
        const double v0 = acc[i][j][0] + o.x, v1 = acc[i][j][1] + o.y;
        const int col = n0 + wn0 + j * 8 + 2 * t;
        if (row < p.M && col < p.N) {
          double* cp = p.C + (size_t)row * p.ldc + col;
          if (col + 1 < p.N) *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
          else cp[0] = v0;
        }
      }
    }
  }
}

// ---- K_A with TMA-staged tiles: the 4-warp 16 x 16 shape of matvec_a_kernel<WL, 1, 2, 2, 2> plus a producer warp.
// Two 3-D boxes per k-tile: L as {k, a', w} with box {16, 16, WL} (all w slices of the (a' block, k-tile) in one copy) and v as
// {c, s, k} with box {16, 4, 16}; both have 128-byte rows and the 128-byte swizzle (fragment loads at the two-wavefront minimum).
constexpr int TMA_ST = 4;
template <int WL>
__global__ void __launch_bounds__(160) matvec_a_tma_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmV,
                                                           MvAArgs p) {
  constexpr int BA = 16, BC = 16;
  constexpr int A_BYTES = WL * BA * BK * 8, B_BYTES = BK * 4 * BC * 8, STAGE = A_BYTES + B_BYTES;
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle atoms");
  extern __shared__ __align__(16) unsigned char raw[];
  const unsigned base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char* sbuf = raw + (base - smem_u32(raw));
  double* Ws = reinterpret_cast<double*>(sbuf + TMA_ST * STAGE);   // W12 [4 wl][4 wr]
  __shared__ __align__(8) unsigned long long full[TMA_ST], empty[TMA_ST];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wa0 = (warp >> 1) * 8, wc0 = (warp & 1) * 8;
  const int a0 = blockIdx.x * BA, c0 = blockIdx.y * BC;
  const int wr = p.wr, chil = p.chil, chir = p.chir;
  if (tid == 0) {
    for (int s = 0; s < TMA_ST; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;\n" ::);
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  for (int i = tid; i < 16 * WL * wr; i += 160) Ws[i] = p.W12[i];
  __syncthreads();
  const int KT = (chil + BK - 1) / BK;
  double acc[WL][4][2];
#pragma unroll
  for (int w = 0; w < WL; ++w)
#pragma unroll
    for (int sg = 0; sg < 4; ++sg) acc[w][sg][0] = acc[w][sg][1] = 0.0;
  if (warp == 4) {
    if (lane == 0) {
      for (int kt = 0; kt < KT; ++kt) {
        const int st = kt % TMA_ST, use = kt / TMA_ST;
        if (use > 0) mbar_wait(smem_u32(&empty[st]), (unsigned)(use - 1) & 1u);
        const unsigned bar = smem_u32(&full[st]);
        const unsigned dstA = base + st * STAGE, dstB = dstA + A_BYTES;
        mbar_expect_tx(bar, STAGE);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                     ::"r"(dstA), "l"(&tmL), "r"(bar), "r"(kt * BK), "r"(a0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                     ::"r"(dstB), "l"(&tmV), "r"(bar), "r"(c0), "r"(0), "r"(kt * BK) : "memory");
      }
    }
  } else {
    int a_off[BK / 4], b_off[BK / 4][4];
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      a_off[kk] = (wa0 + g) * BK + ((((kk * 2 + (t >> 1)) ^ g) << 1) | (t & 1));
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        const int rho = (kk * 4 + t) * 4 + sg;
        b_off[kk][sg] = rho * 16 + (((((wc0 + g) >> 1) ^ (rho & 7)) << 1) | (g & 1));
      }
    }
    for (int kt = 0; kt < KT; ++kt) {
      const int st = kt % TMA_ST;
      mbar_wait(smem_u32(&full[st]), (unsigned)(kt / TMA_ST) & 1u);
      const double* as = reinterpret_cast<const double*>(sbuf + st * STAGE);
      const double* bs = reinterpret_cast<const double*>(sbuf + st * STAGE + A_BYTES);
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[WL], b[4];
#pragma unroll
        for (int w = 0; w < WL; ++w) a[w] = as[w * BA * BK + a_off[kk]];
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) b[sg] = bs[b_off[kk][sg]];
#pragma unroll
        for (int w = 0; w < WL; ++w)
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) dmma884(acc[w][sg][0], acc[w][sg][1], a[w], b[sg]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[st]));
    }
    // epilogue: this lane owns (a' = a0 + wa0 + g, c = c0 + wc0 + 2t, 2t+1) for every (w, s12)
    const int gc = c0 + wc0 + 2 * t;
    const int ga = a0 + wa0 + g;
    if (gc < chir && ga < chil) {
      const int nout = 4 * wr;
      for (int o = 0; o < nout; ++o) {  // o = (t12, wr)
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int w = 0; w < WL; ++w)
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {
            const double c = Ws[(w * 4 + sg) * nout + o];
            o0 += c * acc[w][sg][0];
            o1 += c * acc[w][sg][1];
          }
        *reinterpret_cast<double2*>(p.T2 + ((size_t)ga * nout + o) * chir + gc) = make_double2(o0, o1);   // chir is even on this path
      }
    }
  }
}

typedef CUresult (*TmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmEncodeFn tm_encoder() {
  static TmEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) f = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<TmEncodeFn>(f);
  }();
  return fn;
}
// row-major rows x cols FP64 matrix, box = brows x bcols
bool tm_make(CUtensorMap* tm, const double* ptr, long rows, long cols, long ld, int brows, int bcols, bool swizzle128) {
  TmEncodeFn enc = tm_encoder();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)bcols, (cuuint32_t)brows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// FP64 tensor {d0 (contiguous), d1, d2} with byte strides s1, s2 and box {b0, b1, b2}, 128-byte swizzle (b0 * 8 == 128)
bool tm_make3(CUtensorMap* tm, const double* ptr, long d0, long d1, long d2, long s1, long s2, int b0, int b1, int b2) {
  TmEncodeFn enc = tm_encoder();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)s1, (cuuint64_t)s2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int WL>
bool launch_mva_tma(Ctx* ctx, const MvAArgs& a) {
  struct TmCache { const double* L = nullptr; const double* v = nullptr; int chil = 0, chir = 0; CUtensorMap tl, tv; bool ok = false; };
  static thread_local TmCache tc;
  if (!(tc.ok && tc.L == a.L && tc.v == a.v && tc.chil == a.chil && tc.chir == a.chir)) {
    tc.ok = tm_make3(&tc.tl, a.L, a.chil, a.chil, WL, (long)a.chil * 8, (long)a.chil * a.chil * 8, BK, 16, WL) &&
            tm_make3(&tc.tv, a.v, a.chir, 4, a.chil, (long)a.chir * 8, (long)4 * a.chir * 8, 16, 4, BK);
    tc.L = a.L; tc.v = a.v; tc.chil = a.chil; tc.chir = a.chir;
  }
  if (!tc.ok) return false;
  constexpr size_t smem = (size_t)TMA_ST * (WL * 16 * BK * 8 + BK * 4 * 16 * 8) + 1024 + 16 * WL * 8 * sizeof(double);
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(matvec_a_tma_kernel<WL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.chil + 15) / 16, (a.chir + 15) / 16);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = !(getenv("QTT_PDL") && atoi(getenv("QTT_PDL")) == 0);
  cfg.numAttrs = pdl ? 1 : 0;
  QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_a_tma_kernel<WL>, tc.tl, tc.tv, a));
  check_launch(ctx, "matvec_a_tma");
  return true;
}

template <int WL, int MI, int WM, int WN>
constexpr size_t mva_smem() {
  return (size_t)(STAGES * (WL * 8 * MI * WM * (BK + 4) + BK * (4 * 8 * WN + 4)) + 16 * WL * 8) * sizeof(double);
}

template <int WL, int MI, int WM, int WN, int VEC>
void launch_mva(Ctx* ctx, const MvAArgs& a) {
  auto kern = matvec_a_kernel<WL, MI, WM, WN, VEC>;
  constexpr size_t smem = mva_smem<WL, MI, WM, WN>();
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  dim3 grid((a.chil + 8 * MI * WM - 1) / (8 * MI * WM), (a.chir + 8 * WN - 1) / (8 * WN));
  launch_pdl(ctx, "matvec_a", kern, grid, dim3(WM * WN * 32), smem, a);
}

int g_mva_shape = -1;  // QTT_MVA_SHAPE: tuning override (0: 8 warps 32x16, 1: 4 warps 16x16)

template <int WL, int VEC>
void launch_mva_shape(Ctx* ctx, const MvAArgs& a) {
  int shape = g_mva_shape;
  if (shape < 0) {
    // 4 warps with a 16 x 16 footprint (two to three CTAs resident per SM hide each other's barriers and pipeline fills):
    // measured faster than 8 warps / 32 x 16 at chi = 128 ... 256 and 512 (20.1 vs 20.0, 27.8 vs 26.2 TFLOP/s at w = 3; in the
    // sweep 127.2 vs 128.8 ms) -- except where the 32 x 16 tiles fill exactly one wave of two CTAs per SM (chi = 384:
    // 27.3 vs 24.8 TFLOP/s).  profiles/r01_bench.md
    const long big = (long)((a.chil + 31) / 32) * ((a.chir + 15) / 16);
    const bool one_full_wave = big * 10 > (long)ctx->sms * 2 * 9 && big <= (long)ctx->sms * 2;
    shape = one_full_wave ? 0 : 1;
  }
  // (a third shape, 4 warps with two 16-row fragments each, spilled 10-15 KB of registers at w_l >= 5 and never won: removed)
  static const int tma_env = getenv("QTT_MVA_TMA") ? atoi(getenv("QTT_MVA_TMA")) : 1;   // QTT_MVA_TMA=0: the cp.async kernel (A/B runs)
  if (tma_env && shape == 1 && VEC == 2 && a.wr <= 8 && tm_encoder() && launch_mva_tma<WL>(ctx, a)) return;
  if (shape == 0) launch_mva<WL, 1, 4, 2, VEC>(ctx, a);
  else launch_mva<WL, 1, 2, 2, VEC>(ctx, a);
}

template <int VEC>
void launch_mva_wl(Ctx* ctx, const MvAArgs& a) {
  switch (a.wl) {
    case 1: launch_mva_shape<1, VEC>(ctx, a); break;
    case 2: launch_mva_shape<2, VEC>(ctx, a); break;
    case 3: launch_mva_shape<3, VEC>(ctx, a); break;
    case 4: launch_mva_shape<4, VEC>(ctx, a); break;
    case 5: launch_mva_shape<5, VEC>(ctx, a); break;
    case 6: launch_mva_shape<6, VEC>(ctx, a); break;
    default: fail(QTT_ERR_UNSUPPORTED, "matvec2_fused: wl = %d", a.wl);
  }
}

}  // namespace

bool matvec2_fused_ok(int wl, int wr) { return wl >= 1 && wl <= 6 && wr >= 1 && wr <= 8; }

size_t matvec2_fused_scratch(int chil, int chir, int wr) { return (size_t)4 * chil * chir * wr + 64; }

void matvec2_fused(Ctx* ctx, const double* L, const double* W12, const double* R, const double* v, double* out, int chil, int chir,
                   int wl, int wr, double* scratch) {
  QTT_REQUIRE(matvec2_fused_ok(wl, wr), "matvec2_fused: MPO bond dimension out of range");
  double* T2 = scratch;
  MvAArgs a;
  a.L = L; a.v = v; a.W12 = W12; a.T2 = T2;
  a.chil = chil; a.chir = chir; a.wl = wl; a.wr = wr;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vecA = (chil % 2 == 0) && (chir % 2 == 0) && al16(L) && al16(v) && al16(T2);
  if (g_mva_shape == -1 && getenv("QTT_MVA_SHAPE")) g_mva_shape = atoi(getenv("QTT_MVA_SHAPE"));
  if (vecA) launch_mva_wl<2>(ctx, a);
  else launch_mva_wl<1>(ctx, a);
  MvBArgs b;
  b.A = T2; b.B = R; b.C = out;
  b.M = 4 * chil; b.N = chir; b.K = wr * chir;
  b.lda = (long)wr * chir; b.ldb = chir; b.ldc = chir;
  const bool vecB = (chir % 2 == 0) && al16(T2) && al16(R) && al16(out);
  // 32 x 32 tiles (twice the CTAs, two to three resident per SM) for small outputs, where 32 x 64 tiles leave most SMs
  // idle (chi = 128, w = 4: 18.6 vs 24.7 us), and for multi-wave outputs, where they balance better (chi = 512: 29.6 vs
  // 27.9 TFLOP/s); 32 x 64 tiles for the single-wave sizes in between (chi = 256: 21.1 vs 17.4 TFLOP/s).  For the
  // 32 x 64 single-wave case the request is padded to > half of an SM's shared memory so that at most ONE CTA is
  // resident per SM: with programmatic dependent launch the early CTAs would otherwise pile up two-deep on the SMs that
  // matvec_a leaves idle.
  constexpr size_t smemB64 = (size_t)6 * (32 * (BK + 4) + BK * (64 + 4)) * sizeof(double);
  constexpr size_t smemB32 = (size_t)6 * (32 * (BK + 4) + BK * (32 + 4)) * sizeof(double);
  constexpr size_t smemB_max = smemB64 > 116 * 1024 ? smemB64 : 116 * 1024;
  static DevOnce once;
  once.run(ctx->device, [&] {
    QTT_CUDA(cudaFuncSetAttribute(matvec_b_kernel<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB_max));
    QTT_CUDA(cudaFuncSetAttribute(matvec_b_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB_max));
    QTT_CUDA(cudaFuncSetAttribute(matvec_b_kernel<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB_max));
    QTT_CUDA(cudaFuncSetAttribute(matvec_b_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB_max));
  });
  {
    const long ctas64 = (long)((b.M + 31) / 32) * ((b.N + 63) / 64);
    static const int bn_env = getenv("QTT_MVB_BN") ? atoi(getenv("QTT_MVB_BN")) : 0;
    const int bn = bn_env ? bn_env : ((ctas64 * 4 < (long)ctx->sms * 3 || ctas64 >= 2L * ctx->sms) ? 32 : 64);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((b.M + 31) / 32, (b.N + bn - 1) / bn);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = bn == 32 ? smemB32 : (ctas64 <= ctx->sms ? smemB_max : smemB64);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool pdl = !(getenv("QTT_PDL") && atoi(getenv("QTT_PDL")) == 0);
    cfg.numAttrs = pdl ? 1 : 0;
    // TMA-staged variant of the 32 x 64 kernel (default; A/B in profiles/r02_matvec_tma.md): needs 16-byte global strides
    static const int tma_env = getenv("QTT_MVB_TMA") ? atoi(getenv("QTT_MVB_TMA")) : 1;   // QTT_MVB_TMA=0: the cp.async kernel (A/B runs)
    // (single-wave grids only: with several CTAs per SM the 64-byte B boxes lose to cp.async, chi = 384: 24.4 vs 28.2 TFLOP/s)
    if (tma_env && bn == 64 && ctas64 <= ctx->sms && vecB && (b.lda % 2 == 0) && (b.ldb % 2 == 0) && tm_encoder()) {
      // the two tensor maps depend on the operand addresses and shape only: one-entry cache (31 matvecs per bond share them)
      struct TmCache { const double* A = nullptr; const double* B = nullptr; int M = 0, N = 0, K = 0; long lda = 0, ldb = 0; CUtensorMap ta, tb; bool ok = false; };
      static thread_local TmCache tc;
      if (!(tc.ok && tc.A == b.A && tc.B == b.B && tc.M == b.M && tc.N == b.N && tc.K == b.K && tc.lda == b.lda && tc.ldb == b.ldb)) {
        tc.ok = tm_make(&tc.ta, b.A, b.M, b.K, b.lda, 32, BK, true) && tm_make(&tc.tb, b.B, b.K, b.N, b.ldb, BK, 8, false);
        tc.A = b.A; tc.B = b.B; tc.M = b.M; tc.N = b.N; tc.K = b.K; tc.lda = b.lda; tc.ldb = b.ldb;
      }
      if (tc.ok) {
        static DevOnce once_t;
        once_t.run(ctx->device, [&] {
          QTT_CUDA(cudaFuncSetAttribute(matvec_b_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB_max));
        });
        const size_t need = (size_t)TMB_ST * TMB_STAGE + 1024;
        cfg.dynamicSmemBytes = smemB_max > need ? smemB_max : need;   // one CTA per SM (see the padding note above)
        cfg.blockDim = dim3(288);
        MvBTmaArgs ta;
        ta.C = b.C; ta.M = b.M; ta.N = b.N; ta.K = b.K; ta.ldc = b.ldc;
        QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_b_tma_kernel, tc.ta, tc.tb, ta));
        check_launch(ctx, "matvec_b_tma");
        return;
      }
    }
    if (bn == 32) {
      if (vecB) QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_b_kernel<2, 32>, b));
      else QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_b_kernel<1, 32>, b));
    } else {
      if (vecB) QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_b_kernel<2, 64>, b));
      else QTT_CUDA(cudaLaunchKernelEx(&cfg, matvec_b_kernel<1, 64>, b));
    }
  }
  check_launch(ctx, "matvec_b");
}

}  // namespace qtt

// MPO.MPS `apply` (density-matrix algorithm), `inner`, and the residual metric, for Float64 and ComplexF64 chains.
//
// Replaces ITensorMPS `apply(A::MPO, psi::MPS; cutoff, maxdim, mindim)` = `contract(::Algorithm"densitymatrix", ...)`
// as called by reference src/dft.jl:124,131 and src/prolongation.jl:44,79, and `inner` /
// `sqeuclidean[_normalized]` (reference src/mps_utils.jl:109-132).
//
// Complex tensors are PLANAR on the device (real plane, imaginary plane), so every contraction is 1-4 REAL GEMMs on
// the FP64 DMMA core; a general pairwise contraction is TTGT: at most one coalesced shared-memory permute per operand
// (skipped when the operand already is M x K, K x M, K x N or N x K in memory), then the GEMM(s).
// The Hermitian eigen-decomposition of the density matrix rho is done with the same one-sided Jacobi kernel as the
// site split (rows of rho converge to lambda_i q_i^H), which keeps SVD-grade relative accuracy of small eigenvalues
// (eigen-based truncation at cutoff 1e-15 sits at the noise floor of rho).
#include "common.cuh"
#include <algorithm>
#include <numeric>

namespace qtt {
namespace {

struct DT {  // device tensor view, row-major, planar complex
  double* re = nullptr;
  double* im = nullptr;
  int nd = 0;
  int64_t dim[6] = {1, 1, 1, 1, 1, 1};
  size_t size() const {
    size_t n = 1;
    for (int d = 0; d < nd; ++d) n *= (size_t)dim[d];
    return n;
  }
  bool cplx() const { return im != nullptr; }
};

DT make_dt(double* re, double* im, std::initializer_list<int64_t> dims) {
  DT t;
  t.re = re; t.im = im; t.nd = (int)dims.size();
  int i = 0;
  for (int64_t d : dims) t.dim[i++] = d;
  return t;
}

DT arena_dt(Ctx* ctx, bool cplx, const std::vector<int64_t>& dims) {
  DT t;
  t.nd = (int)dims.size();
  size_t n = 1;
  for (int i = 0; i < t.nd; ++i) { t.dim[i] = dims[i]; n *= (size_t)dims[i]; }
  t.re = ctx->arena.alloc(n);
  t.im = cplx ? ctx->arena.alloc(n) : nullptr;
  return t;
}

DT mps_core(const Mps& x, int j) {
  const int64_t a = x.chi[j], c = x.chi[j + 1];
  double* p = x.core[j].p;
  return make_dt(p, x.dtype == QTT_C128 ? p + (size_t)a * 2 * c : nullptr, {a, 2, c});
}
DT mpo_core(const Mpo& A, int j) {
  const int64_t l = A.w[j], r = A.w[j + 1];
  double* p = A.core[j].p;
  return make_dt(p, A.dtype == QTT_C128 ? p + (size_t)l * r * 4 : nullptr, {l, r, 2, 2});
}

DT permuted(Ctx* ctx, const DT& t, const std::vector<int>& order) {
  std::vector<int64_t> nd(order.size());
  for (size_t k = 0; k < order.size(); ++k) nd[k] = t.dim[order[k]];
  DT o = arena_dt(ctx, t.cplx(), nd);
  permute(ctx, t.re, o.re, t.nd, t.dim, order.data());
  if (t.cplx()) permute(ctx, t.im, o.im, t.nd, t.dim, order.data());
  return o;
}

bool is_order(const std::vector<int>& first, const std::vector<int>& second, int nd) {
  std::vector<int> cat(first);
  cat.insert(cat.end(), second.begin(), second.end());
  if ((int)cat.size() != nd) return false;
  for (int i = 0; i < nd; ++i)
    if (cat[i] != i) return false;
  return true;
}

void gemm_plane(Ctx* ctx, const double* A, bool ta, const double* B, bool tb, double* C, int M, int N, int K, double alpha, double beta) {
  GemmDesc g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.transA = ta; g.transB = tb;
  g.lda = ta ? M : K;
  g.ldb = tb ? K : N;
  g.ldc = N;
  g.alpha = alpha; g.beta = beta;
  gemm(ctx, g);
}

// C[free(A)..., free(B)...] = sum over paired axes of (conj?)A * (conj?)B
DT contract(Ctx* ctx, const DT& A, const std::vector<int>& axa, const DT& B, const std::vector<int>& axb, bool conjA = false,
            bool conjB = false) {
  QTT_REQUIRE(axa.size() == axb.size(), "contract: axis lists differ in length");
  std::vector<int> fa, fb;
  for (int d = 0; d < A.nd; ++d)
    if (std::find(axa.begin(), axa.end(), d) == axa.end()) fa.push_back(d);
  for (int d = 0; d < B.nd; ++d)
    if (std::find(axb.begin(), axb.end(), d) == axb.end()) fb.push_back(d);
  int64_t M = 1, N = 1, K = 1;
  for (int d : fa) M *= A.dim[d];
  for (int d : fb) N *= B.dim[d];
  for (size_t i = 0; i < axa.size(); ++i) {
    QTT_REQUIRE(A.dim[axa[i]] == B.dim[axb[i]], "contract: extent mismatch (%lld vs %lld)", (long long)A.dim[axa[i]], (long long)B.dim[axb[i]]);
    K *= A.dim[axa[i]];
  }
  // operand layouts
  bool ta = false, tb = false;
  DT a = A, b = B;
  if (is_order(fa, axa, A.nd)) ta = false;
  else if (is_order(axa, fa, A.nd)) ta = true;
  else {
    std::vector<int> ord(fa);
    ord.insert(ord.end(), axa.begin(), axa.end());
    a = permuted(ctx, A, ord);
  }
  if (is_order(axb, fb, B.nd)) tb = false;
  else if (is_order(fb, axb, B.nd) && !ta) tb = true;
  else {
    std::vector<int> ord(axb);
    ord.insert(ord.end(), fb.begin(), fb.end());
    b = permuted(ctx, B, ord);
  }
  std::vector<int64_t> cd;
  for (int d : fa) cd.push_back(A.dim[d]);
  for (int d : fb) cd.push_back(B.dim[d]);
  if (cd.empty()) cd.push_back(1);
  const bool cplx = A.cplx() || B.cplx();
  DT C = arena_dt(ctx, cplx, cd);
  const double sa = conjA ? -1.0 : 1.0, sb = conjB ? -1.0 : 1.0;
  const int m = (int)M, n = (int)N, k = (int)K;
  gemm_plane(ctx, a.re, ta, b.re, tb, C.re, m, n, k, 1.0, 0.0);
  if (a.cplx() && b.cplx()) gemm_plane(ctx, a.im, ta, b.im, tb, C.re, m, n, k, -sa * sb, 1.0);
  if (cplx) {
    bool first = true;
    if (b.cplx()) { gemm_plane(ctx, a.re, ta, b.im, tb, C.im, m, n, k, sb, 0.0); first = false; }
    if (a.cplx()) gemm_plane(ctx, a.im, ta, b.re, tb, C.im, m, n, k, sa, first ? 0.0 : 1.0);
  }
  return C;
}

__global__ void fill_kernel(double* p, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void negate_copy_kernel(const double* in, double* out, size_t n, double sign) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = sign * in[i];
}

DT ones_dt(Ctx* ctx, bool cplx, int nd) {
  std::vector<int64_t> d(nd, 1);
  DT t = arena_dt(ctx, cplx, d);
  fill_kernel<<<1, 32, 0, ctx->stream>>>(t.re, 1, 1.0);
  check_launch(ctx, "fill");
  if (cplx) {
    fill_kernel<<<1, 32, 0, ctx->stream>>>(t.im, 1, 0.0);
    check_launch(ctx, "fill");
  }
  return t;
}

void copy_to(Ctx* ctx, const DT& t, double* re, double* im, double im_sign = 1.0) {
  const size_t n = t.size();
  QTT_CUDA(cudaMemcpyAsync(re, t.re, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (im) {
    if (t.cplx()) {
      unsigned g = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sms * 8);
      negate_copy_kernel<<<g ? g : 1, 256, 0, ctx->stream>>>(t.im, im, n, im_sign);
      check_launch(ctx, "negate_copy");
    } else {
      QTT_CUDA(cudaMemsetAsync(im, 0, n * sizeof(double), ctx->stream));
    }
  }
}

struct Persist {  // persistent (non-arena) copy of a tensor
  DBuf buf;
  DT t;
};

void persist(Ctx* ctx, const DT& src, Persist& dst) {
  const size_t n = src.size();
  dbuf_reserve(dst.buf, n * (src.cplx() ? 2 : 1));
  dst.t = src;
  dst.t.re = dst.buf.p;
  dst.t.im = src.cplx() ? dst.buf.p + n : nullptr;
  copy_to(ctx, src, dst.t.re, dst.t.im);
}

void scalar_out(Ctx* ctx, const DT& t, double* out2) {
  out2[0] = out2[1] = 0.0;
  QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 90, t.re, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (t.cplx()) QTT_CUDA(cudaMemcpyAsync(ctx->h_scal + 91, t.im, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QTT_CUDA(cudaStreamSynchronize(ctx->stream));
  out2[0] = ctx->h_scal[90];
  if (t.cplx()) out2[1] = ctx->h_scal[91];
}

size_t max_chi(const Mps& x) { return (size_t)*std::max_element(x.chi.begin(), x.chi.end()); }
size_t max_w(const Mpo& A) { return (size_t)*std::max_element(A.w.begin(), A.w.end()); }

}  // namespace

namespace {

// generic "advance an environment by one site" driver with persistent storage of the running environment
struct EnvRunner {
  Ctx* ctx;
  Persist cur;
  explicit EnvRunner(Ctx* c) : ctx(c) {}
  ~EnvRunner() { dbuf_free(cur.buf); }
  void init(bool cplx, int nd) {
    size_t mark = ctx->arena.mark();
    DT one = ones_dt(ctx, cplx, nd);
    persist(ctx, one, cur);
    ctx->arena.release(mark);
  }
};

// <y|A|x>
void expect_impl(Ctx* ctx, const Mps& y, const Mpo& A, const Mps& x, double* out2) {
  const bool cplx = y.dtype == QTT_C128 || A.dtype == QTT_C128 || x.dtype == QTT_C128;
  EnvRunner er(ctx);
  er.init(cplx, 3);  // (cy, cx, w)
  for (int j = 0; j < x.n; ++j) {
    size_t mark = ctx->arena.mark();
    DT T = contract(ctx, er.cur.t, {1}, mps_core(x, j), {0});                 // (cy, w, s, cx')
    DT T2 = contract(ctx, T, {1, 2}, mpo_core(A, j), {0, 2});                  // (cy, cx', w', t)
    DT En = contract(ctx, mps_core(y, j), {0, 1}, T2, {0, 3}, true);           // (cy', cx', w')
    Persist nxt;
    persist(ctx, En, nxt);
    ctx->arena.release(mark);
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    dbuf_free(er.cur.buf);
    er.cur = nxt;
  }
  scalar_out(ctx, er.cur.t, out2);
}

// <A x | B y>
void inner4_impl(Ctx* ctx, const Mpo& A, const Mps& x, const Mpo& B, const Mps& y, double* out2) {
  const bool cplx = y.dtype == QTT_C128 || A.dtype == QTT_C128 || x.dtype == QTT_C128 || B.dtype == QTT_C128;
  EnvRunner er(ctx);
  er.init(cplx, 4);  // (ax, ay, wb, wa)
  for (int j = 0; j < x.n; ++j) {
    size_t mark = ctx->arena.mark();
    DT T = contract(ctx, er.cur.t, {1}, mps_core(y, j), {0});                  // (ax, wb, wa, s, cy)
    DT T2 = contract(ctx, T, {1, 3}, mpo_core(B, j), {0, 2});                   // (ax, wa, cy, wb', t)
    DT T3 = contract(ctx, T2, {1, 4}, mpo_core(A, j), {0, 3}, false, true);     // (ax, cy, wb', wa', s)
    DT En = contract(ctx, mps_core(x, j), {0, 1}, T3, {0, 4}, true);            // (cx, cy, wb', wa')
    Persist nxt;
    persist(ctx, En, nxt);
    ctx->arena.release(mark);
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    dbuf_free(er.cur.buf);
    er.cur = nxt;
  }
  scalar_out(ctx, er.cur.t, out2);
}

}  // namespace

// <x|y>
void mps_inner(Ctx* ctx, const Mps& x, const Mps& y, double* out2) {
  QTT_REQUIRE(x.n == y.n, "inner: length mismatch");
  size_t cx = max_chi(x), cy = max_chi(y);
  ctx->ensure_arena((size_t)64 * (cx * cy * 4 + 1024) * sizeof(double));
  const bool cplx = x.dtype == QTT_C128 || y.dtype == QTT_C128;
  EnvRunner er(ctx);
  er.init(cplx, 2);  // (ax, by)
  for (int j = 0; j < x.n; ++j) {
    size_t mark = ctx->arena.mark();
    DT T = contract(ctx, er.cur.t, {1}, mps_core(y, j), {0});                // (ax, s, b')
    DT En = contract(ctx, mps_core(x, j), {0, 1}, T, {0, 1}, true);           // (a', b')
    Persist nxt;
    persist(ctx, En, nxt);
    ctx->arena.release(mark);
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    dbuf_free(er.cur.buf);
    er.cur = nxt;
  }
  scalar_out(ctx, er.cur.t, out2);
}

// sqeuclidean((A,x), b) = <Ax,Ax> - 2 Re <b|A|x> + <b,b>;  normalized: 1 + (<Ax,Ax>/bb - 2 Re<b|A|x>/bb)
void mps_residual(Ctx* ctx, const Mpo& A, const Mps& x, const Mps& b, double* sq, double* sqn) {
  QTT_REQUIRE(A.n == x.n && b.n == x.n, "residual: length mismatch");
  size_t cx = max_chi(x), cb = max_chi(b), w = max_w(A);
  size_t big = std::max(cx, cb);
  ctx->ensure_arena((size_t)96 * (big * big * w * w * 4 + 4096) * sizeof(double));
  double axax[2], bax[2], bb[2];
  inner4_impl(ctx, A, x, A, x, axax);
  expect_impl(ctx, b, A, x, bax);
  mps_inner(ctx, b, b, bb);
  if (sq) *sq = axax[0] - 2.0 * bax[0] + bb[0];
  if (sqn) *sqn = 1.0 + (axax[0] / bb[0] - 2.0 * bax[0] / bb[0]);
}

// apply(A, psi; cutoff, maxdim, mindim): density-matrix algorithm (oracle/apply.py apply_mpo_mps)
void apply_mpo_mps(Ctx* ctx, const Mpo& A, const Mps& psi, double cutoff, int maxdim, int mindim, Mps& out) {
  const int n = A.n;
  QTT_REQUIRE(psi.n == n, "apply: MPS length %d != MPO length %d", psi.n, n);
  QTT_REQUIRE(A.w[0] == 1 && A.w[n] == 1 && psi.chi[0] == 1 && psi.chi[n] == 1, "apply: boundary bond dimensions must be 1");
  const bool cplx = A.dtype == QTT_C128 || psi.dtype == QTT_C128;
  if (mindim < 1) mindim = 1;
  const size_t chim = max_chi(psi), wm = max_w(A);
  int requested = maxdim > 0 ? maxdim : (int)(chim * wm);
  out.n = n;
  out.dtype = cplx ? QTT_C128 : QTT_F64;
  out.chi.assign(n + 1, 1);
  out.core.resize(n);
  const size_t kcap = std::min<size_t>((size_t)requested, chim * wm);
  // workspace: largest intermediates are (w w chi 2 chi) for the environments and (2k)^2 for rho
  {
    size_t e1 = wm * wm * chim * chim * 2;
    size_t e2 = (2 * kcap) * (2 * kcap) * 4 + wm * chim * 2 * kcap * 4;
    size_t elems = std::max(e1, e2) + 65536;
    ctx->ensure_arena((size_t)40 * elems * sizeof(double) * (cplx ? 2 : 1));
  }
  auto store_core = [&](int j, const DT& t, int64_t a, int64_t c, double im_sign) {
    const size_t ne = (size_t)a * 2 * c;
    dbuf_reserve(out.core[j], ne * (cplx ? 2 : 1));
    copy_to(ctx, t, out.core[j].p, cplx ? out.core[j].p + ne : nullptr, im_sign);
    out.chi[j] = a;
    out.chi[j + 1] = c;
  };
  if (n == 1) {
    size_t mark = ctx->arena.mark();
    DT T = contract(ctx, mpo_core(A, 0), {2}, mps_core(psi, 0), {1});  // (1,1,t,1,1)
    store_core(0, T, 1, 1, 1.0);
    QTT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->arena.release(mark);
    return;
  }
  // left environments E[j] (p, a, a~, p~) for j = 0..n-2, kept in their own buffers
  std::vector<Persist> E(n - 1);
  auto free_env = [&]() {
    for (auto& e : E) dbuf_free(e.buf);
  };
  Persist carry;  // Rk (a, p, k): the running right factor of the density-matrix pass
  try {
    {
      Persist e0;
      size_t mark = ctx->arena.mark();
      DT one = ones_dt(ctx, cplx, 4);
      persist(ctx, one, e0);
      ctx->arena.release(mark);
      DT cur = e0.t;
      for (int j = 0; j < n - 1; ++j) {
        size_t mk = ctx->arena.mark();
        DT T = contract(ctx, cur, {0}, mps_core(psi, j), {0});                      // (a, a~, p~, s, p')
        DT T2 = contract(ctx, T, {0, 3}, mpo_core(A, j), {0, 2});                    // (a~, p~, p', a', t)
        DT T3 = contract(ctx, T2, {0, 4}, mpo_core(A, j), {0, 3}, false, true);      // (p~, p', a', a~', s~)
        DT T4 = contract(ctx, T3, {0, 4}, mps_core(psi, j), {0, 1}, false, true);    // (p', a', a~', p~')
        persist(ctx, T4, E[j]);
        ctx->arena.release(mk);
        cur = E[j].t;
      }
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      dbuf_free(e0.buf);
    }
    // right-to-left: rho -> eigen -> carry
    // (carry is declared before the try block so that the error path can release it)
    int k = 1;
    {
      size_t mk = ctx->arena.mark();
      DT R0 = contract(ctx, mpo_core(A, n - 1), {2}, mps_core(psi, n - 1), {1});  // (a, 1, t, p, 1)
      DT R = make_dt(R0.re, R0.im, {R0.dim[0], 2, R0.dim[3]});                     // (a, t, p)
      DT T = contract(ctx, E[n - 2].t, {1, 0}, R, {0, 2});                         // (a~, p~, t)
      DT rho = contract(ctx, T, {0, 1}, R, {0, 2}, false, true);                   // (t, t~)
      const int d = 2;
      RowSvd ev = svd_rows_ex(ctx, rho.re, rho.im, d, d, cutoff, std::min(requested, d), mindim, nullptr, 0, true);
      k = ev.k;
      DT Wn = make_dt(ev.Wn, ev.Wni, {k, 2});
      // out core [k][t][1] = U^T = conj(Wn)
      store_core(n - 1, Wn, k, 1, -1.0);
      DT Rk = contract(ctx, R, {1}, Wn, {1});  // (a, p, k)
      persist(ctx, Rk, carry);
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->arena.release(mk);
    }
    for (int j = n - 2; j >= 1; --j) {
      size_t mk = ctx->arena.mark();
      DT M = contract(ctx, mps_core(psi, j), {2}, carry.t, {1});      // (p_l, s, a, k)
      DT M2 = contract(ctx, mpo_core(A, j), {1, 2}, M, {2, 1});        // (a_l, t, p_l, k)
      DT T = contract(ctx, E[j - 1].t, {1, 0}, M2, {0, 2});            // (a~, p~, t, k)
      DT rho = contract(ctx, T, {0, 1}, M2, {0, 2}, false, true);      // (t, k, t~, k~)
      const int d = 2 * k;
      const int64_t prod_dims = M2.dim[0] * M2.dim[2];
      int md = (int)std::min<int64_t>(prod_dims, requested);
      RowSvd ev = svd_rows_ex(ctx, rho.re, rho.im, d, d, cutoff, std::min(md, d), mindim, nullptr, 0, true);
      const int kn = ev.k;
      DT Wn3 = make_dt(ev.Wn, ev.Wni, {kn, 2, k});
      store_core(j, Wn3, kn, k, -1.0);
      DT Rk = contract(ctx, M2, {1, 3}, Wn3, {1, 2});                  // (a_l, p_l, kn)
      Persist nxt;
      persist(ctx, Rk, nxt);
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->arena.release(mk);
      dbuf_free(carry.buf);
      carry = nxt;
      k = kn;
    }
    {
      size_t mk = ctx->arena.mark();
      DT M = contract(ctx, mps_core(psi, 0), {2}, carry.t, {1});      // (1, s, a, k)
      DT M2 = contract(ctx, mpo_core(A, 0), {1, 2}, M, {2, 1});        // (1, t, 1, k)
      store_core(0, M2, 1, k, 1.0);
      QTT_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->arena.release(mk);
    }
    dbuf_free(carry.buf);
  } catch (...) {
    dbuf_free(carry.buf);
    free_env();
    throw;
  }
  free_env();
}

}  // namespace qtt

"""Host-side mirror of the reference's public API for the hot path (same names, argument meaning and
error behaviour as ITensorQTT / ITensorMPS), bound to the C ABI.  A Julia user keeps calling
`linsolve(A, b, x0, a0, a1; nsweeps, maxdim, cutoff, ...)`; INTEGRATION.md shows the `ccall` shim.

Everything here runs on the GPU through libqtt_b200.so.  Nothing in this package imports `oracle/`."""
import ctypes as C

import numpy as np

from ._lib import lib, check, default_context, SweepParams, SweepInfo, as_f, dptr
from .chains import DeviceMPS, DeviceMPO

__all__ = ["linsolve", "linsolve_pseudoinverse", "b_linsolve", "dmrg_target", "vec_to_mps", "function_to_mps", "apply_dft_mpo", "apply_idft_mpo",
           "fourier_interpolation", "repeat_function", "apply", "add", "mps_to_vec", "sample", "prolongate", "retract", "multigrid_linsolve", "inner", "sqeuclidean", "sqeuclidean_normalized", "matvec2",
           "env_left", "env_right", "split2", "gemm", "krylov_op", "matvec2_flops", "env_flops",
           "SOLVER_GMRES", "SOLVER_LANCZOS", "SOLVER_SHIFT_INVERT", "SOLVER_DENSE_QR", "SOLVER_DENSE_EIGEN", "SOLVER_PG_LSQ"]

SOLVER_GMRES, SOLVER_LANCZOS, SOLVER_SHIFT_INVERT, SOLVER_DENSE_QR, SOLVER_DENSE_EIGEN, SOLVER_PG_LSQ = 0, 1, 2, 3, 4, 5


def _per_sweep(v, nsweeps, ctype, default):
    if v is None:
        return None, None
    if np.isscalar(v):
        v = [v]
    v = list(v)
    vals = [v[min(i, len(v) - 1)] for i in range(nsweeps)]  # tdvp `process_sweeps`: last value repeated
    arr = (ctype * nsweeps)(*vals)
    return arr, vals


def _as_dev_mps(x, ctx):
    return (x, False) if isinstance(x, DeviceMPS) else (DeviceMPS(x, ctx), True)


def _as_dev_mpo(A, ctx):
    return (A, False) if isinstance(A, DeviceMPO) else (DeviceMPO(A, ctx), True)


def _params(nsweeps, maxdim, mindim, cutoff, solver, a0, a1, tol, krylovdim, maxiter, fixed_matvecs, normalize, outputlevel,
            target, x0_canonical, svd_max_sweeps=0):
    if nsweeps is None:
        raise TypeError("nsweeps is required (as in ITensorMPS `tdvp`)")
    p = SweepParams()
    keep = []
    p.nsweeps = int(nsweeps)
    for name, val, ct in (("maxdim", maxdim, C.c_int32), ("mindim", mindim, C.c_int32), ("cutoff", cutoff, C.c_double)):
        arr, _ = _per_sweep(val, p.nsweeps, ct, None)
        keep.append(arr)
        setattr(p, name, arr if arr is not None else C.cast(None, C.POINTER(ct)))
    p.solver = solver
    if np.iscomplexobj(a0) or np.iscomplexobj(a1):
        raise NotImplementedError("complex a0/a1 are not supported by this build")
    p.a0, p.a1, p.tol = float(a0), float(a1), float(tol)
    p.krylovdim, p.maxiter = int(krylovdim), int(maxiter)
    p.fixed_matvecs = int(fixed_matvecs or 0)
    p.normalize = int(bool(normalize))
    p.outputlevel = int(outputlevel)
    p.target = float(target)
    p.svd_max_sweeps = int(svd_max_sweeps)
    p.x0_canonical = int(bool(x0_canonical))
    return p, keep


def linsolve(A, b, x0, a0=0, a1=1, *, nsweeps=None, maxdim=None, mindim=None, cutoff=None, tol=1e-5, krylovdim=20,
             maxiter=10, ishermitian=False, normalize=False, outputlevel=0, fixed_matvecs=None, x0_canonical=False,
             ctx=None, return_info=False, on_device=False, host_order="C", out=None, _solver=SOLVER_GMRES):
    """`linsolve(A::MPO, b::MPS, x0::MPS, a0=0, a1=1; nsweeps, maxdim, cutoff, tol, krylovdim, maxiter, ishermitian)`
    -- solves (a0 + a1 A) x = b by two-site sweeps with a local GMRES
    (reference examples/airy_equation/src/linsolve.jl:24-34; call sites airy_equation.jl:429, src/eigsolve_target.jl:4).
    `ishermitian` is accepted for signature compatibility (KrylovKit still picks GMRES unless isposdef).
    A, b, x0: lists of host cores or Device* handles.  Returns host cores (or a DeviceMPS if `on_device`).  `out`: caller-owned
    column-major host buffers (e.g. pinned) that receive the result, see `DeviceMPS.download`."""
    ctx = ctx or default_context()
    dA, fa = _as_dev_mpo(A, ctx)
    db, fb = _as_dev_mps(b, ctx)
    dx = x0.clone() if isinstance(x0, DeviceMPS) else DeviceMPS(x0, ctx)
    p, keep = _params(nsweeps, maxdim, mindim, cutoff, _solver, a0, a1, tol, krylovdim, maxiter, fixed_matvecs,
                      normalize, outputlevel, 0.0, x0_canonical)
    infos = (SweepInfo * max(p.nsweeps, 1))()
    try:
        check(ctx.handle, lib.qtt_linsolve(ctx.handle, dA.handle, db.handle, dx.handle, C.byref(p), infos))
    except Exception:
        dx.free()      # the working copy of x0 must not outlive a failed solve
        raise
    finally:
        if fa:
            dA.free()
        if fb:
            db.free()
    try:
        res = dx if on_device else dx.download(order=host_order, out=out)
    finally:
        if not on_device:
            dx.free()
    if return_info:
        return res, [infos[i].as_dict() for i in range(p.nsweeps)]
    return res


def b_linsolve(A, b, x0, a0=0, a1=1, *, nsweeps=None, maxdim=None, mindim=None, cutoff=None, **kwargs):
    """`b_linsolve(A, b, x0, a0=0, a1=1; kwargs...)` (reference src/linsolve.jl:263-307, live solver `b_linsolve_exact_solver`):
    the two-site sweep with Petrov-Galerkin environments -- the right-hand side b is the BRA of the operator environments
    (`ProjMPOLinsolve`, :175-253, b re-orthogonalised to every bond) -- and a dense SVD least-squares local solve on the
    (4 chi_b^2) x (4 chi_x^2) projected operator with `bvec` (:255-261) as the right-hand side.  a0 / a1 are accepted and
    ignored, as in the reference body."""
    return linsolve(A, b, x0, 0.0, 1.0, nsweeps=nsweeps, maxdim=maxdim, mindim=mindim, cutoff=cutoff, _solver=SOLVER_PG_LSQ, **kwargs)


def linsolve_pseudoinverse(A, b, x0, a0=0, a1=1, *, nsweeps=None, maxdim=None, mindim=None, cutoff=None, tol=1e-5, krylovdim=20,
                           maxiter=10, ishermitian=False, **kwargs):
    """`linsolve_pseudoinverse(A, b, x0, a0=0, a1=1; kwargs...)` (reference src/linsolve.jl:54-97): the same two-site sweep
    with a DENSE local solve -- the projected operator as a (4 chi_l chi_r)^2 matrix, `qr!(T_mat) \\ b_vec` (:88-89).
    As in the reference, `a0`, `a1` and the Krylov keywords are accepted and IGNORED (the body never uses them, :67-92):
    the local problem is T x = b.  Memory grows as chi^4 (reference's own warning, :50-52); the device path accepts
    4 chi_l chi_r <= 8192 and raises otherwise."""
    return linsolve(A, b, x0, 0.0, 1.0, nsweeps=nsweeps, maxdim=maxdim, mindim=mindim, cutoff=cutoff, _solver=SOLVER_DENSE_QR,
                    **kwargs)


def dmrg_target(A, psi0, *, target_eigenvalue, nsweeps=None, maxdim=None, mindim=None, cutoff=None, tol=1e-12, krylovdim=30,
                maxiter=1, normalize=True, outputlevel=0, fixed_matvecs=None, x0_canonical=False, ctx=None,
                return_info=False, on_device=False, local_solver="lanczos"):
    """`dmrg_target(A, psi0; target_eigenvalue, nsweeps, maxdim, cutoff, ...)` (reference src/dmrg_target.jl:1-12) with
    the Krylov local solver the reference sketches at :14-21 (dense `eigen` of the (4 chi^2)^2 effective operator is
    infeasible beyond chi ~ 64).  local_solver="lanczos": restarted Lanczos, Ritz value nearest `target_eigenvalue` (extremal
    targets); local_solver="shift_invert": `eigsolve_target` (reference src/eigsolve_target.jl:2-9) =
    eigsolve(x -> linsolve(P, x, x, -target)), which also resolves interior eigenvalues; local_solver="dense": the reference
    body itself (:3-7) -- dense effective operator, eigenvector of the eigenvalue nearest the target (4 chi_l chi_r <= 8192)."""
    solver = {"lanczos": SOLVER_LANCZOS, "shift_invert": SOLVER_SHIFT_INVERT, "dense": SOLVER_DENSE_EIGEN}[local_solver]
    ctx = ctx or default_context()
    dA, fa = _as_dev_mpo(A, ctx)
    dx = psi0.clone() if isinstance(psi0, DeviceMPS) else DeviceMPS(psi0, ctx)
    p, keep = _params(nsweeps, maxdim, mindim, cutoff, solver, 0.0, 1.0, tol, krylovdim, maxiter, fixed_matvecs,
                      normalize, outputlevel, target_eigenvalue, x0_canonical)
    infos = (SweepInfo * max(p.nsweeps, 1))()
    try:
        check(ctx.handle, lib.qtt_dmrg_target(ctx.handle, dA.handle, dx.handle, C.byref(p), infos))
    except Exception:
        dx.free()
        raise
    finally:
        if fa:
            dA.free()
    out = dx if on_device else dx.download()
    if not on_device:
        dx.free()
    if return_info:
        return out, [infos[i].as_dict() for i in range(p.nsweeps)]
    return out


def apply(A, psi, *, cutoff=1e-13, maxdim=None, mindim=1, ctx=None, on_device=False):
    """`apply(A::MPO, psi::MPS; cutoff, maxdim, mindim)` (upstream density-matrix algorithm; reference call sites
    src/dft.jl:124,131, src/prolongation.jl:44,79).  Output cores carry the MPO's output site index."""
    ctx = ctx or default_context()
    dA, fa = _as_dev_mpo(A, ctx)
    dp, fp = _as_dev_mps(psi, ctx)
    out = C.c_void_p()
    try:
        check(ctx.handle, lib.qtt_apply(ctx.handle, dA.handle, dp.handle, float(cutoff), int(maxdim or 0), int(mindim), C.byref(out)))
    finally:
        if fa:
            dA.free()
        if fp:
            dp.free()
    res = DeviceMPS(ctx=ctx, _handle=out)
    if on_device:
        return res
    cores = res.download()
    res.free()
    return cores


def vec_to_mps(v, *, cutoff=1e-15, maxdim=None, mindim=1, ctx=None, on_device=False, return_info=False):
    """`vec_to_mps(v, s; cutoff, maxdim, mindim)` = `MPS(vec_to_tensor(v), s; ...)` (reference src/function_to_mps.jl:170-177):
    TT-SVD of 2^n samples on the GPU, site 1 = most significant bit, orthogonality centre at site n.
    v: host array of length 2^n, or a CUDA tensor / object with `data_ptr()` and `numel()` already on the context's device
    (read in place)."""
    ctx = ctx or default_context()
    if hasattr(v, "data_ptr"):
        total = int(v.numel())
        ptr, dev = C.c_void_p(int(v.data_ptr())), 1
        hold = v
    else:
        hold = np.ascontiguousarray(np.asarray(v), dtype=np.float64).reshape(-1)
        if np.iscomplexobj(v):
            raise NotImplementedError("vec_to_mps: complex samples are not supported by this build")
        total = hold.size
        ptr, dev = C.c_void_p(hold.ctypes.data), 0
    n = int(total).bit_length() - 1
    if total < 2 or (1 << n) != total:
        raise ValueError(f"vec_to_mps: length {total} is not a power of two >= 2")
    out = C.c_void_p()
    err = C.c_double(0.0)
    check(ctx.handle, lib.qtt_vec_to_mps(ctx.handle, ptr, n, dev, float(cutoff), int(maxdim or 0), int(mindim), C.byref(out), C.byref(err)))
    del hold
    res = DeviceMPS(ctx=ctx, _handle=out)
    if not on_device:
        cores = res.download()
        res.free()
        res = cores
    return (res, {"maxtruncerr": err.value}) if return_info else res


def function_to_mps(f, n, xstart, xstop, *, cutoff=1e-15, alg="factorize", **kwargs):
    """`function_to_mps(f, s, xstart, xstop; alg="factorize", cutoff=1e-15)` (reference src/function_to_mps.jl:22-52): sample
    f on `qtt_xrange(n, xstart, xstop)` (vectorised numpy callable) and TT-SVD the samples on the GPU."""
    if alg != "factorize":
        raise NotImplementedError('only alg="factorize" runs on the device; the polynomial construction (:88-133) is host-side')
    from .operators import qtt_xrange
    return vec_to_mps(f(qtt_xrange(n, xstart, xstop)), cutoff=cutoff, **kwargs)


def add(x, y, *, alpha=1.0, beta=1.0, cutoff=1e-15, maxdim=None, mindim=1, ctx=None, on_device=False):
    """`+(x, y; cutoff, maxdim)` of ITensorMPS (reference call sites src/fourier_interpolation.jl:8, src/function_to_mps.jl:123),
    generalised to alpha x + beta y: the direct sum is assembled and truncated on the GPU (`qtt_mps_add`)."""
    ctx = ctx or default_context()
    dx, fx = _as_dev_mps(x, ctx)
    dy, fy = _as_dev_mps(y, ctx)
    out = C.c_void_p()
    try:
        check(ctx.handle, lib.qtt_mps_add(ctx.handle, dx.handle, dy.handle, float(alpha), float(beta), float(cutoff), int(maxdim or 0),
                                          int(mindim), C.byref(out)))
    finally:
        if fx:
            dx.free()
        if fy:
            dy.free()
    res = DeviceMPS(ctx=ctx, _handle=out)
    if on_device:
        return res
    cores = res.download()
    res.free()
    return cores


def mps_to_vec(psi, ctx=None, out=None):
    """`mps_to_discrete_function(psi)` (reference src/function_to_mps.jl:136-145, one dimension) on the GPU -- the inverse of
    `vec_to_mps`: the 2^n grid values, site 1 = most significant bit, expanded by `qtt_mps_to_vec` and returned as a host array.
    (`operators.mps_to_discrete_function` is the host-side restatement used for operand specs.)
    out: optional CUDA tensor / object with `data_ptr()` and `numel()` on the context's device that receives the values in
    place (2^n float64, or 2^n complex128 for a complex chain); it is returned instead of a host array."""
    ctx = ctx or default_context()
    dp, fp = _as_dev_mps(psi, ctx)
    try:
        n = len(dp)
        if out is not None and hasattr(out, "data_ptr"):
            need = (1 << n)
            if int(out.numel()) != need:
                raise ValueError(f"mps_to_vec: out has {int(out.numel())} elements, the chain has 2^{n} = {need} values")
            check(ctx.handle, lib.qtt_mps_to_vec(ctx.handle, dp.handle, C.c_void_p(int(out.data_ptr())), 1))
            return out
        res = np.empty(1 << n, dtype=np.complex128 if dp.is_complex else np.float64)
        check(ctx.handle, lib.qtt_mps_to_vec(ctx.handle, dp.handle, C.c_void_p(res.ctypes.data), 0))
        return res
    finally:
        if fp:
            dp.free()


def sample(psi, idx, ctx=None):
    """Values of the QTT at the grid indices `idx` (0-based, site 1 = most significant bit): `project_bits(psi, bits)` with every
    site projected (reference src/project_bits.jl:1-15, scalar return :11-13), batched on the GPU (`qtt_mps_sample`)."""
    ctx = ctx or default_context()
    ii = np.ascontiguousarray(np.asarray(idx, dtype=np.int64).reshape(-1))
    dp, fp = _as_dev_mps(psi, ctx)
    try:
        out = np.empty(ii.size, dtype=np.complex128 if dp.is_complex else np.float64)
        check(ctx.handle, lib.qtt_mps_sample(ctx.handle, dp.handle, ii.ctypes.data_as(C.POINTER(C.c_int64)), int(ii.size),
                                             C.c_void_p(out.ctypes.data)))
    finally:
        if fp:
            dp.free()
    return out.reshape(np.shape(idx))


def inner(x, y, ctx=None):
    """`inner(x, y)` = <x|y>."""
    ctx = ctx or default_context()
    dx, fx = _as_dev_mps(x, ctx)
    dy, fy = _as_dev_mps(y, ctx)
    out = (C.c_double * 2)()
    cplx = dx.is_complex or dy.is_complex
    try:
        check(ctx.handle, lib.qtt_inner(ctx.handle, dx.handle, dy.handle, out))
    finally:
        if fx:
            dx.free()
        if fy:
            dy.free()
    return complex(out[0], out[1]) if cplx else out[0]


def _residual(A, x, b, ctx):
    ctx = ctx or default_context()
    dA, fa = _as_dev_mpo(A, ctx)
    dx, fx = _as_dev_mps(x, ctx)
    db, fb = _as_dev_mps(b, ctx)
    sq, sqn = C.c_double(), C.c_double()
    try:
        check(ctx.handle, lib.qtt_residual(ctx.handle, dA.handle, dx.handle, db.handle, C.byref(sq), C.byref(sqn)))
    finally:
        for d, f in ((dA, fa), (dx, fx), (db, fb)):
            if f:
                d.free()
    return sq.value, sqn.value


def sqeuclidean(Ax, b, ctx=None):
    """`sqeuclidean((A, x), b)` (reference src/mps_utils.jl:109-111)."""
    A, x = Ax
    return _residual(A, x, b, ctx)[0]


def sqeuclidean_normalized(Ax, b, ctx=None):
    """`sqeuclidean_normalized((A, x), b)` (reference src/mps_utils.jl:113-116)."""
    A, x = Ax
    return _residual(A, x, b, ctx)[1]


# ------------------------------------------------------------------ fine-grained kernels (tests, ncu, roofline)
def matvec2_flops(chil, chir, wl, wm, wr, d=2):
    """F_mv."""
    return 2 * (chil * (wl * chil) * (d * d * chir) + (wl * d) * (chil * d * chir) * (wm * d)
                + (wm * d) * (chil * d * chir) * (wr * d) + (wr * chir) * (chil * d * d) * chir)


def env_flops(chi, chi2, w, w2, d=2):
    return 2 * (chi * (w * chi) * (d * chi2) + (w * d) * (chi * chi2) * (w2 * d) + (chi * d) * (w2 * chi2) * chi2)


def matvec2(L, W1, W2, R, v, reps=1, ctx=None):
    """`product(P, v)` = noprime(v*L*W1*W2*R).  L (a, w, a'), R (c, w, c'), W (wl, wr, s, s'), v (a, s1, s2, c).
    Returns (out, ms_per_rep)."""
    ctx = ctx or default_context()
    L, W1, W2, R, v = (as_f(t) for t in (L, W1, W2, R, v))
    chil, wl = L.shape[0], L.shape[1]
    chir, wr = R.shape[0], R.shape[1]
    wm = W1.shape[1]
    assert L.shape == (chil, wl, chil) and R.shape == (chir, wr, chir)
    assert W1.shape == (wl, wm, 2, 2) and W2.shape == (wm, wr, 2, 2) and v.shape == (chil, 2, 2, chir)
    out = np.empty_like(v, order="F")
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_matvec2(ctx.handle, dptr(L), dptr(W1), dptr(W2), dptr(R), dptr(v), dptr(out), chil, chir, wl, wm, wr,
                                      int(reps), C.byref(ms)))
    return np.ascontiguousarray(out), ms.value


def env_left(L, x, W, reps=1, ctx=None):
    ctx = ctx or default_context()
    L, x, W = (as_f(t) for t in (L, x, W))
    chia, w = L.shape[0], L.shape[1]
    chib, w2 = x.shape[2], W.shape[1]
    out = np.empty((chib, w2, chib), order="F")
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_env_left(ctx.handle, dptr(L), dptr(x), dptr(W), dptr(out), chia, chib, w, w2, int(reps), C.byref(ms)))
    return np.ascontiguousarray(out), ms.value


def env_right(R, x, W, reps=1, ctx=None):
    ctx = ctx or default_context()
    R, x, W = (as_f(t) for t in (R, x, W))
    chic, w2 = R.shape[0], R.shape[1]
    chia, w = x.shape[0], W.shape[0]
    out = np.empty((chia, w, chia), order="F")
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_env_right(ctx.handle, dptr(R), dptr(x), dptr(W), dptr(out), chia, chic, w, w2, int(reps), C.byref(ms)))
    return np.ascontiguousarray(out), ms.value


def split2(phi, ortho="left", cutoff=0.0, maxdim=None, mindim=1, reps=1, ctx=None):
    """`replacebond!`: truncated SVD split of phi (a, 2, 2, c).  Returns dict(A, B, sigma, rank, sweeps, truncerr, ms)."""
    ctx = ctx or default_context()
    phi = as_f(phi)
    a, c = phi.shape[0], phi.shape[3]
    kmax = min(2 * a, 2 * c)
    A = np.zeros((a, 2, kmax), order="F")
    B = np.zeros((kmax, 2, c), order="F")
    sig = np.zeros(kmax)
    info = (C.c_int32 * 2)()
    err = C.c_double(0)
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_split2(ctx.handle, dptr(phi), a, c, 0 if ortho == "left" else 1, float(cutoff), int(maxdim or 0),
                                     int(mindim), dptr(A), dptr(B), dptr(sig), info, C.byref(err), int(reps), C.byref(ms)))
    k = info[0]
    # the library wrote packed (a, 2, k) / (k, 2, c) column-major arrays into the front of the buffers
    Ak = np.ascontiguousarray(A.reshape(-1, order="F")[: a * 2 * k].reshape((a, 2, k), order="F"))
    Bk = np.ascontiguousarray(B.reshape(-1, order="F")[: k * 2 * c].reshape((k, 2, c), order="F"))
    return dict(A=Ak, B=Bk, sigma=sig[:k].copy(), rank=k, sweeps=info[1], truncerr=err.value, ms=ms.value)


def gemm(A, B, Cin=None, alpha=1.0, beta=0.0, transA=False, transB=False, tile=-1, reps=1, ctx=None):
    """Row-major C = alpha op(A) op(B) + beta C on the DMMA core (test hook)."""
    ctx = ctx or default_context()
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    Cm = np.zeros((M, N)) if Cin is None else np.ascontiguousarray(Cin, dtype=np.float64).copy()
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_gemm(ctx.handle, int(transA), int(transB), M, N, K, float(alpha), dptr(A), A.shape[1], dptr(B), B.shape[1],
                                   float(beta), dptr(Cm), N, int(tile), int(reps), C.byref(ms)))
    return Cm, ms.value


def krylov_op(op, V, w, h=None, reps=1, ctx=None):
    """op 0: h = V^T w; 1: w -= V h; 2: ||w||; 3: one CGS2 step (returns w_orth, h, norm)."""
    ctx = ctx or default_context()
    V = np.ascontiguousarray(V, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64).copy()
    nvec, ln = V.shape
    hh = np.zeros(nvec) if h is None else np.ascontiguousarray(h, dtype=np.float64).copy()
    out = np.zeros(2)
    ms = C.c_float(0)
    check(ctx.handle, lib.qtt_krylov_op(ctx.handle, int(op), dptr(V), nvec, ln, dptr(w), dptr(hh), dptr(out), int(reps), C.byref(ms)))
    return w, hh, out[0], ms.value


# ---------------------------------------------------------------- prolongation / retraction / multigrid (host glue around qtt_apply)
def _host_cores(psi):
    return psi.download() if isinstance(psi, DeviceMPS) else [np.asarray(c) for c in psi]


def prolongate(psi, nnew=1, *, cutoff=1e-8, maxdim=None, ctx=None):
    """`prolongate(psi, s_new; cutoff=1e-8)` (reference src/prolongation.jl:29-60, 1-D): append a dummy site, apply the
    prolongation MPO (linear interpolation onto the grid with one more bit) on the GPU; repeated `nnew` times."""
    from .operators import prolongation_mpo
    cores = _host_cores(psi)
    for _ in range(nnew):
        dummy = np.zeros((1, 2, 1), dtype=cores[0].dtype)
        dummy[0, 0, 0] = 1.0
        cores = apply(prolongation_mpo(len(cores) + 1), list(cores) + [dummy], cutoff=cutoff, maxdim=maxdim, ctx=ctx)
    return cores


def retract(psi, nretract=1, *, cutoff=1e-8, maxdim=None, ctx=None):
    """`retract(psi, nretract=1; cutoff=1e-8)` (reference src/prolongation.jl:66-102, 1-D): apply the transposed prolongation
    MPO on the GPU, absorb the dummy output site into its neighbour, divide by 2."""
    from .operators import retraction_mpo
    cores = _host_cores(psi)
    for _ in range(nretract):
        out = apply(retraction_mpo(len(cores)), cores, cutoff=cutoff, maxdim=maxdim, ctx=ctx)
        last = out[-1][:, 0, :]
        out[-2] = np.tensordot(out[-2], last, axes=([2], [0]))
        cores = [c.copy() for c in out[:-1]]
        cores[0] = 0.5 * cores[0]
    return cores


def apply_dft_mpo(psi, *, cutoff=1e-13, maxdim=None, ctx=None):
    """`apply_dft_mpo(psi; kwargs...)` = reverse(apply(dft_mpo(s), psi; kwargs...)) (reference src/dft.jl:121-126): site 1 of the
    result is the most significant bit of the frequency index.  The MPO.MPS apply runs on the GPU (ComplexF64, planar);
    the DFT MPO comes from the host cache (`operators.dft_mpo`)."""
    from .operators import dft_mpo
    cores = [np.asarray(c, dtype=np.complex128) for c in _host_cores(psi)]
    out = apply(dft_mpo(len(cores)), cores, cutoff=cutoff, maxdim=maxdim, ctx=ctx)
    return [np.ascontiguousarray(np.transpose(c, (2, 1, 0))) for c in reversed(out)]


def apply_idft_mpo(psi, *, cutoff=1e-13, maxdim=None, ctx=None):
    """`apply_idft_mpo(psi; kwargs...)` = apply(idft_mpo(s), reverse(psi); kwargs...) (reference src/dft.jl:128-133)."""
    from .operators import idft_mpo
    cores = [np.asarray(c, dtype=np.complex128) for c in _host_cores(psi)]
    rpsi = [np.ascontiguousarray(np.transpose(c, (2, 1, 0))) for c in reversed(cores)]
    return apply(idft_mpo(len(cores)), rpsi, cutoff=cutoff, maxdim=maxdim, ctx=ctx)


def _basis_site(bit):
    c = np.zeros((1, 2, 1), dtype=np.complex128)
    c[0, bit, 0] = 1.0
    return c


def fourier_interpolation(psi, k=1, *, cutoff=1e-15, ctx=None):
    """`fourier_interpolation(psi, k; cutoff=1e-15)` (reference src/fourier_interpolation.jl:1-19, arXiv:1909.06619): QTT of length n
    -> length n + k.  Both transforms (the O(n w^2 chi^3) work) run on the GPU.  The step between them --
    `project_bits(F, [0])`, `project_bits(F, [1])`, k + 1 basis sites in front of each, `+(...; cutoff=1e-15)` (:4-8) -- is
    written out exactly: the two summands share every core after the frequency MSB, so their sum is the chain whose first
    k + 1 sites are the copy tensor delta(l, s, r) on a bond of 2 and whose next core stacks the two projected cores; the
    inverse-DFT `apply` truncates with `cutoff` exactly as it does after the reference's compression of that sum."""
    F = apply_dft_mpo(psi, cutoff=cutoff, ctx=ctx)
    n = len(F)
    if n < 2:
        raise ValueError("fourier_interpolation needs at least 2 sites")
    head = []
    for i in range(k + 1):
        l, r = (1 if i == 0 else 2), 2
        c = np.zeros((l, 2, r), dtype=np.complex128)
        for b in range(2):
            c[b if l == 2 else 0, b, b] = 1.0
        head.append(c)
    # bond value b selects the projection of the frequency MSB: row b of the stacked core = F[0][0, b, :] . F[1]
    nxt = np.stack([np.tensordot(F[0][0, b, :], F[1], axes=([0], [0])) for b in range(2)], axis=0)
    Fi = head + [nxt] + [c for c in F[2:]]
    out = apply_idft_mpo(Fi, cutoff=cutoff, ctx=ctx)
    out[0] = out[0] * 2.0 ** (k / 2)
    return out


def repeat_function(psi, k=1, *, cutoff=1e-15, ctx=None):
    """`repeat_function(psi, k; cutoff=1e-15)` (reference src/fourier_interpolation.jl:21-38): repeat the function 2^k times =
    DFT, k |0> sites appended as the low frequency bits, inverse DFT, times 2^(k/2).  Both transforms on the GPU."""
    F = apply_dft_mpo(psi, cutoff=cutoff, ctx=ctx)
    Fk = F + [_basis_site(0) for _ in range(k)]
    out = apply_idft_mpo(Fk, cutoff=cutoff, ctx=ctx)
    out[0] = out[0] * 2.0 ** (k / 2)
    return out


def multigrid_linsolve(problem, levels, x0, a0=0, a1=1, *, prolong_cutoff=1e-15, ctx=None, return_history=False, **kw):
    """The multigrid driver of the reference examples (examples/airy_equation/airy_equation.jl:453-516,
    examples/helmholtz_equation/helmholtz_equation.jl:153-224): solve on n bits, prolongate the solution to n + 1 bits
    (`prolongate(u, s_new; cutoff=1e-15)`, :497) and use it as the starting state of the next level.
    problem(n) -> (A, b); levels: increasing bit counts; x0: starting MPS on levels[0] bits; kw: `linsolve` keywords."""
    x = x0
    hist = []
    for i, n in enumerate(levels):
        A, b = problem(n)
        if i > 0:
            x = prolongate(x, n - levels[i - 1], cutoff=prolong_cutoff, ctx=ctx)
        x, info = linsolve(A, b, x, a0, a1, ctx=ctx, return_info=True, **kw)
        hist.append(dict(n=n, maxlinkdim=info[-1]["maxlinkdim"], matvecs=sum(s["matvecs"] for s in info),
                         residual=sqeuclidean_normalized((A, x), b, ctx=ctx) if (a0 == 0 and a1 == 1) else None))
    return (x, hist) if return_history else x

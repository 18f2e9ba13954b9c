"""GPU parity of the sweep engine through the C ABI vs the oracle on identical explicit inputs.
Gates (north-star): fidelity 1 - |<x_ref|x>| <= 1e-10; relative residual within 1.05x of the reference's."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _airy(n, xi=1.0, xf=8.0):
    A = O.airy_mpo(n, xi, xf)
    al = be = 1 / np.sqrt(2)
    ui, uf = O.airy_solution(xi, al, be), O.airy_solution(xf, al, be)
    return A, O.boundary_value_mps(n, ui, uf), O.airy_matrix(n, xi, xf), O.boundary_value_vector(n, ui, uf)


def test_upload_download_roundtrip(lib):
    x = O.random_mps(9, 7, 5)
    d = lib.DeviceMPS(x)
    assert d.dims() == [1] + O.linkdims(x) + [1]
    y = d.download()
    for a, b in zip(x, y):
        assert np.array_equal(a, b)       # bit-exact: layout maps only
    xc = [c.astype(complex) * (1 + 0.5j) for c in x]
    yc = lib.DeviceMPS(xc).download()
    for a, b in zip(xc, yc):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [6, 10])
def test_airy_linsolve_parity(lib, n):
    A, b, Am, bv = _airy(n)
    x0 = O.random_mps(n, 10, 1234)
    kw = dict(nsweeps=3, maxdim=32, cutoff=1e-12, tol=1e-10, krylovdim=30, maxiter=10)
    xo, _ = O.linsolve(A, b, x0, **kw)
    xg, info = lib.linsolve(A, b, x0, return_info=True, **kw)
    assert O.fidelity_defect(xo, xg) <= 1e-10
    vo = O.mps_to_discrete_function(xo); vg = O.mps_to_discrete_function(xg)
    ro = np.linalg.norm(Am @ vo - bv) / np.linalg.norm(bv)
    rg = np.linalg.norm(Am @ vg - bv) / np.linalg.norm(bv)
    assert rg <= 1.05 * ro + 1e-14
    assert O.linkdims(xg) == O.linkdims(xo)
    xd = np.linalg.solve(Am, bv)
    assert 1 - abs(vg @ xd) / np.linalg.norm(vg) / np.linalg.norm(xd) < 1e-10
    assert info[-1]["maxlinkdim"] == max(O.linkdims(xg))


def test_helmholtz_shifted_linsolve_parity(lib):
    """(a0 + a1 A) x = b with the shift passed as a0 = -lambda (SURVEY 8d cfg 2 at CPU-checkable size)."""
    n, nk = 10, 3
    A = O.laplacian_mpo(n, 1.0)
    lam = O.helmholtz_lambda(n, nk)
    b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
    x0 = O.random_mps(n, 8, 2)
    kw = dict(nsweeps=3, maxdim=16, cutoff=1e-13, tol=1e-11, krylovdim=30, maxiter=20)
    xo, _ = O.linsolve(A, b, x0, -lam, 1.0, **kw)
    xg = lib.linsolve(A, b, x0, -lam, 1.0, **kw)
    assert O.fidelity_defect(xo, xg) <= 1e-10
    Am = O.laplacian_matrix(2 ** n) - lam * np.eye(2 ** n)
    bv = O.mps_to_discrete_function(b)
    ro = np.linalg.norm(Am @ O.mps_to_discrete_function(xo) - bv)
    rg = np.linalg.norm(Am @ O.mps_to_discrete_function(xg) - bv)
    assert rg <= 1.05 * ro + 1e-13


def test_fixed_work_mode_matches_oracle(lib):
    """Benchmark mode: exactly one GMRES cycle of k Arnoldi steps per bond (no tolerance test)."""
    n = 8
    A, b, Am, bv = _airy(n, 1.0, 2.0)
    x0 = O.random_mps(n, 6, 77)
    kw = dict(nsweeps=2, maxdim=12, cutoff=1e-14, fixed_matvecs=6)
    st = {}
    xo, _ = O.linsolve(A, b, x0, stats=st, **kw)
    xg, info = lib.linsolve(A, b, x0, return_info=True, **kw)
    assert sum(i["matvecs"] for i in info) == st["matvecs"] == 2 * 2 * (n - 1) * 7
    assert O.fidelity_defect(xo, xg) <= 1e-10


def test_dmrg_target_lanczos(lib):
    n = 8
    Al = O.laplacian_mpo(n, 1.0)
    Lm = O.laplacian_matrix(2 ** n)
    D = np.linalg.eigvalsh(Lm)
    target = D[0]           # most negative eigenvalue: extremal, well separated relative to the local problem
    psi0 = O.random_mps(n, 6, 4)
    psi, info = lib.dmrg_target(Al, psi0, target_eigenvalue=target, nsweeps=4, maxdim=8, cutoff=1e-14, krylovdim=20, maxiter=4,
                                tol=1e-12, return_info=True)
    v = O.mps_to_discrete_function(psi)
    ev = (v @ Lm @ v) / (v @ v)
    assert abs(ev - target) < 1e-8
    assert abs(info[-1]["energy"] - target) < 1e-8


def test_multi_kernel_path_large_bonds(lib):
    """Bonds with 4 chi_l chi_r > 1024 take the multi-launch solver (fused DMMA matvec + cooperative CGS2 kernel) instead of
    the single-CTA small-bond kernel: same answer as the oracle in fixed-work mode and in converged mode."""
    n = 10
    A, b, Am, bv = _airy(n, 1.0, 3.0)
    x0 = O.random_mps(n, 24, 5)                       # interior bonds 24: len = 2304
    kw = dict(nsweeps=1, maxdim=24, cutoff=1e-13, fixed_matvecs=10)
    xo, _ = O.linsolve(A, b, x0, **kw)
    xg = lib.linsolve(A, b, x0, **kw)
    assert O.fidelity_defect(xo, xg) <= 1e-10
    kw = dict(nsweeps=2, maxdim=24, cutoff=1e-13, tol=1e-11, krylovdim=30, maxiter=6)
    xo, _ = O.linsolve(A, b, x0, **kw)
    xg = lib.linsolve(A, b, x0, **kw)
    assert O.fidelity_defect(xo, xg) <= 1e-10
    vg = O.mps_to_discrete_function(xg); vo = O.mps_to_discrete_function(xo)
    assert np.linalg.norm(Am @ vg - bv) <= 1.05 * np.linalg.norm(Am @ vo - bv) + 1e-13


def test_dmrg_target_shift_invert_interior_eigenvalue(lib):
    """`eigsolve_target` as the local solver (reference src/eigsolve_target.jl:2-9, src/dmrg_target.jl:14-21): an INTERIOR
    eigenvalue of the Laplacian (the Helmholtz mode nk = 2, helmholtz_equation.jl:48), which a Ritz-value pick cannot isolate."""
    n = 7
    Al = O.laplacian_mpo(n, 1.0)
    Lm = O.laplacian_matrix(2 ** n)
    D, U = np.linalg.eigh(Lm)
    target = O.helmholtz_lambda(n, 2) * 1.001          # near, not at, an eigenvalue
    near = int(np.argmin(np.abs(D - target)))
    psi0 = O.vec_to_mps(U[:, near] + 0.3 * np.random.default_rng(0).standard_normal(2 ** n), cutoff=1e-12, maxdim=6)
    psi, info = lib.dmrg_target(Al, psi0, target_eigenvalue=target, nsweeps=3, maxdim=8, cutoff=1e-13, krylovdim=20, maxiter=3,
                                tol=1e-10, return_info=True, local_solver="shift_invert")
    v = O.mps_to_discrete_function(psi)
    v /= np.linalg.norm(v)
    assert 1 - abs(v @ U[:, near]) < 1e-8
    assert abs(info[-1]["energy"] - D[near]) < 1e-8 * abs(D[near])


@pytest.mark.parametrize("n,chi", [(6, 4), (9, 10), (10, 16)])
def test_linsolve_pseudoinverse_dense_qr_parity(lib, n, chi):
    """`linsolve_pseudoinverse` (reference src/linsolve.jl:54-97): dense projected operator + Householder QR on the device
    vs the oracle's restatement (numpy QR + triangular solve) on the same inputs; bonds up to len = 4 * 10 * 10 = 400
    exercise several 32-reflector panels."""
    A, b, Am, bv = _airy(n)
    x0 = O.random_mps(n, chi, 1234)
    kw = dict(nsweeps=2, maxdim=chi, cutoff=1e-13)
    xo, _ = O.linsolve_pseudoinverse(A, b, x0, **kw)
    xg, info = lib.linsolve_pseudoinverse(A, b, x0, return_info=True, **kw)
    assert O.fidelity_defect(xo, xg) <= 1e-10
    vo = O.mps_to_discrete_function(xo); vg = O.mps_to_discrete_function(xg)
    ro = np.linalg.norm(Am @ vo - bv) / np.linalg.norm(bv)
    rg = np.linalg.norm(Am @ vg - bv) / np.linalg.norm(bv)
    assert rg <= 1.05 * ro + 1e-12
    assert info[-1]["matvecs"] == 2 * (n - 1)          # one residual-check matvec per bond, none for the solve
    # a0 / a1 are accepted and ignored, as in the reference body (:67-92)
    xs = lib.linsolve_pseudoinverse(A, b, x0, 0.5, 2.0, **kw)
    assert O.fidelity_defect(xg, xs) <= 1e-13


def test_dense_qr_local_solver_honours_shift_at_the_c_abi(lib):
    """At the C ABI QTT_SOLVER_DENSE_QR solves (a0 + a1 T) x = b (the Python/Julia mirrors pin a0 = 0, a1 = 1): same answer as
    converged GMRES on the shifted Helmholtz problem."""
    n = 8
    Al = O.laplacian_mpo(n, 1.0)
    lam = O.helmholtz_lambda(n, 2)
    b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
    x0 = O.random_mps(n, 6, 2)
    kw = dict(nsweeps=2, maxdim=8, cutoff=1e-13)
    xk = lib.linsolve(Al, b, x0, -lam, 1.0, tol=1e-13, krylovdim=30, maxiter=20, **kw)
    xd = lib.linsolve(Al, b, x0, -lam, 1.0, _solver=lib.SOLVER_DENSE_QR, **kw)
    assert O.fidelity_defect(xk, xd) <= 1e-10


def test_dmrg_target_dense_eigen_parity(lib):
    """`dmrg_target` as written (reference src/dmrg_target.jl:1-12): dense effective operator, eigenvector nearest the target.
    Interior target (Helmholtz mode nk = 2) -> same state as the oracle's `eigh` + argmin, and the known eigenpair."""
    n = 7
    Al = O.laplacian_mpo(n, 1.0)
    Lm = O.laplacian_matrix(2 ** n)
    D, U = np.linalg.eigh(Lm)
    target = O.helmholtz_lambda(n, 2) * 1.001
    near = int(np.argmin(np.abs(D - target)))
    psi0 = O.vec_to_mps(U[:, near] + 0.3 * np.random.default_rng(0).standard_normal(2 ** n), cutoff=1e-12, maxdim=6)
    kw = dict(nsweeps=2, maxdim=8, cutoff=1e-13)
    po, _ = O.dmrg_target(Al, psi0, target, normalize=True, **kw)
    pg, info = lib.dmrg_target(Al, psi0, target_eigenvalue=target, return_info=True, local_solver="dense", **kw)
    assert O.fidelity_defect(po, pg) <= 1e-10
    v = O.mps_to_discrete_function(pg)
    v /= np.linalg.norm(v)
    assert 1 - abs(v @ U[:, near]) < 1e-9
    assert abs(info[-1]["energy"] - D[near]) < 1e-10 * abs(D[near])


def test_dense_solver_refuses_large_bonds(lib, monkeypatch):
    """The dense modes are chi^4 in memory (reference's own warning, src/linsolve.jl:50-52): beyond the cap the call fails
    loudly instead of falling back to anything."""
    n = 9
    A, b, _, _ = _airy(n)
    x0 = O.random_mps(n, 10, 3)                          # interior bonds: 4 * 10 * 10 = 400
    monkeypatch.setenv("QTT_DENSE_LEN_MAX", "256")
    with pytest.raises(Exception, match="dense local solver"):
        lib.linsolve_pseudoinverse(A, b, x0, nsweeps=1, maxdim=10, mindim=10, cutoff=0.0)


@pytest.mark.parametrize("n,chi,nsw", [(3, 2, 1), (4, 3, 1), (8, 6, 2), (10, 8, 2)])
def test_b_linsolve_parity(lib, n, chi, nsw):
    """`b_linsolve` (reference src/linsolve.jl:263-307): Petrov-Galerkin environments with b as the bra, dense SVD least squares.
    Same state as the oracle's restatement (numpy lstsq); the gauge of the re-orthogonalised b differs (Jacobi vs QR), which
    the minimum-norm least-squares solution does not see."""
    A, b, Am, bv = _airy(n)
    x0 = O.random_mps(n, chi, 11)
    kw = dict(nsweeps=nsw, maxdim=12, cutoff=1e-12)
    xo, _ = O.b_linsolve(A, b, x0, **kw)
    xg, info = lib.b_linsolve(A, b, x0, return_info=True, **kw)
    assert O.linkdims(xg) == O.linkdims(xo)
    assert O.fidelity_defect(xo, xg) <= 1e-9
    assert info[-1]["maxlinkdim"] == max(O.linkdims(xg))
    if n >= 4:
        # the Petrov-Galerkin conditions are NOT the Galerkin ones: the two entry points must disagree on this system,
        # i.e. the C-ABI really took the b_linsolve branch
        xl = lib.linsolve(A, b, x0, **kw)
        assert O.fidelity_defect(xl, xg) > 1e-3


def test_linsolve_into_caller_buffers(lib):
    """`out=`: the result is written into caller-owned column-major buffers (pinned memory in bench.py's end-to-end leg);
    same bits as the plain call, wrong shapes are rejected."""
    n = 8
    A, b, Am, bv = _airy(n)
    x0 = O.random_mps(n, 4, 5)
    kw = dict(nsweeps=2, maxdim=8, mindim=8, cutoff=0.0, tol=1e-12, krylovdim=20, maxiter=4)
    ref = lib.linsolve(A, b, x0, host_order="F", **kw)
    out = [np.empty(c.shape, dtype=np.float64, order="F") for c in ref]
    got = lib.linsolve(A, b, x0, host_order="F", out=out, **kw)
    assert all(g is o for g, o in zip(got, out))
    assert all(np.array_equal(g, r) for g, r in zip(got, ref))
    bad = [np.empty((1, 2, 1), order="F") for _ in ref]
    with pytest.raises(ValueError):
        lib.linsolve(A, b, x0, host_order="F", out=bad, **kw)

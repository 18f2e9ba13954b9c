"""Sweep-engine oracle anchored on the substitutes the reference's examples use:
dense solves of the same system, analytic solutions, and the residual metric."""
import numpy as np

import oracle as O


def _airy(n, xi=1.0, xf=8.0):
    A = O.airy_mpo(n, xi, xf)
    al = be = 1 / np.sqrt(2)
    ui, uf = O.airy_solution(xi, al, be), O.airy_solution(xf, al, be)
    return A, O.boundary_value_mps(n, ui, uf), O.airy_matrix(n, xi, xf), O.boundary_value_vector(n, ui, uf)


def test_gmres_krylovkit_dense():
    rng = np.random.default_rng(5)
    M = rng.standard_normal((40, 40)); M = (M + M.T) / 4 + 12 * np.eye(40)
    b = rng.standard_normal(40); x0 = rng.standard_normal(40)
    x, info = O.gmres_krylovkit(lambda v: M @ v, b, x0, 0.3, 1.7, tol=1e-12, krylovdim=10, maxiter=50)
    assert info["converged"] == 1
    assert np.linalg.norm((0.3 * np.eye(40) + 1.7 * M) @ x - b) < 1e-11
    # fixed-work mode applies the operator exactly 1 + k times
    cnt = [0]
    def mv(v):
        cnt[0] += 1
        return M @ v
    O.gmres_krylovkit(mv, b, x0, fixed_matvecs=7)
    assert cnt[0] == 8


def test_gmres_restart_residual_recurrence():
    """The operator-free restart residual must equal the true residual."""
    rng = np.random.default_rng(6)
    M = rng.standard_normal((30, 30)) + 6 * np.eye(30)
    b = rng.standard_normal(30)
    x1, i1 = O.gmres_krylovkit(lambda v: M @ v, b, np.zeros(30), tol=1e-13, krylovdim=5, maxiter=200)
    assert i1["converged"] == 1 and np.linalg.norm(M @ x1 - b) < 1e-12


def test_eigsolve_krylovkit():
    rng = np.random.default_rng(7)
    M = rng.standard_normal((60, 60)); M = M + M.T
    D = np.linalg.eigvalsh(M)
    val, vec, info = O.eigsolve_krylovkit(lambda v: M @ v, rng.standard_normal(60), which="LR", krylovdim=12, maxiter=200, tol=1e-10)
    assert abs(val - D[-1]) < 1e-9 and np.linalg.norm(M @ vec - val * vec) < 1e-8


def test_airy_linsolve_matches_dense_solve():
    n = 10
    A, b, Am, bv = _airy(n)
    xd = np.linalg.solve(Am, bv)
    x0 = O.random_mps(n, 10, 1234)
    for fn, kw in ((O.linsolve_pseudoinverse, {}), (O.linsolve, dict(tol=1e-10, krylovdim=30, maxiter=10))):
        x, rec = fn(A, b, x0, nsweeps=3, maxdim=32, cutoff=1e-12, **kw)
        xv = O.mps_to_discrete_function(x)
        assert 1 - abs(xv @ xd) / np.linalg.norm(xv) / np.linalg.norm(xd) < 1e-10
        assert np.linalg.norm(Am @ xv - bv) / np.linalg.norm(bv) < 5e-5
        assert rec[-1]["maxlinkdim"] <= 32
        # exact Airy function up to the discretisation error of the reference scheme (b[N] = u(xf) on the [xi, xf) grid: O(h))
        xs = O.qtt_xrange(n, 1.0, 8.0)
        ex = O.airy_solution(xs, 1 / np.sqrt(2), 1 / np.sqrt(2))
        assert np.abs(xv - ex).max() < 1e-2


def test_b_linsolve_petrov_galerkin():
    n = 8
    A, b, Am, bv = _airy(n, 1.0, 2.0)
    x0 = O.random_mps(n, 6, 11)
    x, _ = O.b_linsolve(A, b, x0, nsweeps=2, maxdim=16, cutoff=1e-14)
    xv = O.mps_to_discrete_function(x)
    # Petrov-Galerkin with test space span(b): <b|A x> = <b|b>
    assert abs(bv @ (Am @ xv) - bv @ bv) < 1e-8 * abs(bv @ bv)


def test_dmrg_target_helmholtz():
    """n=10, nk=2 converges to the nearest exact Dirichlet eigenvalue -6.011878981140e-04, bonds 2."""
    n, nk = 10, 2
    Al = O.laplacian_mpo(n, 1.0)
    lam = O.helmholtz_lambda(n, nk)
    psi, rec = O.dmrg_target(Al, O.random_mps(n, 10, 7), lam, nsweeps=3, cutoff=1e-15, maxdim=16)
    v = O.mps_to_discrete_function(psi)
    Lm = O.laplacian_matrix(2 ** n)
    ev = (v @ Lm @ v) / (v @ v)
    assert abs(ev - (-6.011878981140e-04)) < 1e-12
    assert np.linalg.norm(Lm @ v - ev * v) < 1e-12
    assert O.linkdims(psi) == [2] * (n - 1)


def test_matvec_matches_dense_operator():
    rng = np.random.default_rng(8)
    L = rng.standard_normal((5, 3, 5)); R = rng.standard_normal((4, 3, 4))
    W1 = rng.standard_normal((3, 4, 2, 2)); W2 = rng.standard_normal((4, 3, 2, 2))
    v = rng.standard_normal((5, 2, 2, 4))
    T = O.effective_operator_dense(L, W1, W2, R)
    assert np.allclose(T @ v.reshape(-1), O.matvec2(L, W1, W2, R, v).reshape(-1))
    assert O.matvec2_flops(256, 256, 3, 3, 3) == 4 * 3 * 4 * 256 ** 3 + 4 * 9 * 8 * 256 ** 2

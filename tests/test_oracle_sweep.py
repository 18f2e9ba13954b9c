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


def _captured_local_problem(n=8, chi=6, bond=3, seed=4):
    """The projected two-site problem of an Airy chain at one bond, as the sweep sees it: dense effective operator, projected
    right-hand side, current two-site tensor (reference src/linsolve.jl:13-22, 71-75)."""
    from oracle.core import orthogonalize
    from oracle.sweep import env_left_update, env_right_update, benv_left_update, benv_right_update, effective_operator_dense, project_rhs
    A, b, _, _ = _airy(n)
    x = orthogonalize(O.random_mps(n, chi, seed), bond)
    one3, one2 = np.ones((1, 1, 1)), np.ones((1, 1))
    L, Lb = one3, one2
    for i in range(bond):
        L, Lb = env_left_update(L, x[i], A[i]), benv_left_update(Lb, x[i], b[i])
    R, Rb = one3, one2
    for i in range(n - 1, bond + 1, -1):
        R, Rb = env_right_update(R, x[i], A[i]), benv_right_update(Rb, x[i], b[i])
    T = effective_operator_dense(L, A[bond], A[bond + 1], R)
    rhs = project_rhs(Lb, b[bond], b[bond + 1], Rb).reshape(-1)
    phi = np.tensordot(x[bond], x[bond + 1], axes=([2], [0])).reshape(-1)
    return T, rhs, phi


def test_gmres_cycle_is_the_krylov_minimiser_and_matches_scipy():
    """Hardening of the (Julia-unpinned) KrylovKit restatement on captured local problems of a sweep: after ONE cycle of k
    Arnoldi steps from x0, GMRES returns THE minimiser of |b - (a0 + a1 T) x| over x0 + K_k(r0) -- unique, independent of the
    orthogonalisation scheme and of how the shift is folded in.  Checked against (1) an explicit least-squares solve on the
    power basis of the Krylov space and (2) scipy.sparse.linalg.gmres with restart = k, one outer iteration."""
    import scipy.sparse.linalg as sla
    for bond, a0, a1, k in ((3, 0.0, 1.0, 6), (2, -0.37, 1.0, 9), (4, 0.8, -1.3, 5)):
        T, rhs, phi = _captured_local_problem(bond=bond)
        N = T.shape[0]
        Aeff = a0 * np.eye(N) + a1 * T
        x, info = O.gmres_krylovkit(lambda v: T @ v, rhs, phi, a0, a1, fixed_matvecs=k)
        assert info["numops"] == k + 1
        r0 = rhs - Aeff @ phi
        K = np.empty((N, k))
        v = r0 / np.linalg.norm(r0)
        for j in range(k):       # orthonormalised power basis (QR keeps it well conditioned)
            K[:, j] = v
            v = Aeff @ v
            v -= K[:, : j + 1] @ (K[:, : j + 1].T @ v)
            v /= np.linalg.norm(v)
        y = np.linalg.lstsq(Aeff @ K, r0, rcond=None)[0]
        x_min = phi + K @ y
        scale = np.linalg.norm(x_min)
        assert np.linalg.norm(x - x_min) < 1e-10 * scale
        xs, _ = sla.gmres(Aeff, rhs, x0=phi, rtol=1e-300, atol=0.0, restart=k, maxiter=1)
        assert np.linalg.norm(x - xs) < 1e-9 * scale
        # the residual estimate the solver reports is the true residual of its iterate
        assert abs(info["normres"] - np.linalg.norm(rhs - Aeff @ x)) < 1e-9 * np.linalg.norm(rhs)


def test_gmres_converged_mode_matches_dense_solve_on_local_problem():
    """tol mode (the reference's linsolve_kwargs: tol, krylovdim, maxiter with restarts): converges to the dense solution of the
    captured local problem, restart residual recomputed from the operator."""
    T, rhs, phi = _captured_local_problem(bond=3)
    a0, a1 = 0.25, 1.0
    Aeff = a0 * np.eye(T.shape[0]) + a1 * T
    x, info = O.gmres_krylovkit(lambda v: T @ v, rhs, phi, a0, a1, tol=1e-11, krylovdim=12, maxiter=200)
    assert info["converged"] == 1
    assert np.linalg.norm(x - np.linalg.solve(Aeff, rhs)) < 1e-8 * np.linalg.norm(x)

"""Oracle builders: operators (Laplacian, prolongation, DFT layers, Airy) and analytic QTTs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to /root/reference.
"""
import numpy as np
from math import comb

from .core import (mpo_add_directsum, mpo_scale, mps_add_directsum, mps_scale, mps_truncate,
                   interleave_mpo, mpo_identity, qtt_xrange)

__all__ = [
    "laplacian_tensor", "laplacian_matrix", "laplacian_mpo", "laplacian_mpo_2d", "prolongation_tensor",
    "prolongation_mpo", "retraction_mpo", "boundary_value_mps", "boundary_value_vector", "qtt_exp",
    "qtt_sin", "qtt_cos", "polynomial_tensor", "function_to_mps_polynomial", "s_plus_tensor", "dft_matrix",
    "phase_mpo", "hadamard_phase_mpo", "dft_layer", "airy_mpo", "airy_matrix", "airy_solution",
    "helmholtz_lambda",
]


# ------------------------------------------------------------------ Laplacian (src/laplacian.jl)
def laplacian_matrix(xlength, xstep=1.0):
    """src/laplacian.jl:9-14: SymTridiagonal(-2, 1) / xstep^2."""
    A = np.diag(np.full(xlength, -2.0)) + np.diag(np.ones(xlength - 1), 1) + np.diag(np.ones(xlength - 1), -1)
    return A / (xstep ** 2)


def laplacian_tensor():
    """src/laplacian.jl:17-26 (1-based (l, r, s, s') -> 0-based)."""
    T = np.zeros((3, 3, 2, 2))
    T[0, 0, 0, 0] = 1.0
    T[0, 0, 1, 1] = 1.0
    T[1, 1, 1, 0] = 1.0
    T[0, 1, 0, 1] = 1.0
    T[0, 2, 1, 0] = 1.0
    T[2, 2, 0, 1] = 1.0
    return T


def laplacian_mpo(n, xstep=1.0):
    """src/laplacian.jl:29-43: left boundary onehot(l=1) (:34-35), right boundary (-2, 1, 1)
    (:37-41), every core divided by xstep^(2/L) (:42)."""
    T = laplacian_tensor()
    cores = [T.copy() for _ in range(n)]
    cores[0] = cores[0][0:1, :, :, :]
    right = np.array([-2.0, 1.0, 1.0])
    cores[n - 1] = np.tensordot(cores[n - 1], right, axes=([1], [0]))[:, None, :, :]
    scale = xstep ** (2.0 / n)
    return [c / scale for c in cores]


def laplacian_mpo_2d(n1, n2, xstep=(1.0, 1.0)):
    """src/laplacian.jl:45-53: interleave(Delta_1, I) (+) interleave(I, Delta_2) by direct sum."""
    assert n1 == n2
    D1 = interleave_mpo(laplacian_mpo(n1, xstep[0]), mpo_identity(n2))
    D2 = interleave_mpo(mpo_identity(n1), laplacian_mpo(n2, xstep[1]))
    return mpo_add_directsum(D1, D2)


def s_plus_tensor():
    """src/laplacian.jl:68-78: T[l,r,s,s'] = 1 if s' = s xor r and l = s and r."""
    T = np.zeros((2, 2, 2, 2))
    for l in range(2):
        for r in range(2):
            for s in range(2):
                for sp in range(2):
                    if sp == (s ^ r) and l == (s & r):
                        T[l, r, s, sp] = 1.0
    return T


# ------------------------------------------------------------------ Prolongation (src/prolongation.jl)
def prolongation_tensor():
    """src/prolongation.jl:2-9."""
    P = np.zeros((2, 2, 2, 2))
    P[0, 0, 0, 0] = 1.0
    P[0, 0, 1, 1] = 1.0
    P[0, 1, 1, 0] = 1.0
    P[1, 1, 0, 1] = 1.0
    return P


def prolongation_mpo(n_fine):
    """src/prolongation.jl:12-23 for a fine grid of `n_fine` sites (coarse = n_fine-1).  The last core
    has only an output leg (:18-21); following `prolongate` (:34-42) it gets a dummy input leg of
    dimension 1, stored here as s_in = 0 with s_in = 1 zero-padded so every core is (w, w, 2, 2)
    and the coarse MPS is padded with a trailing (chi, [1, 0]) site."""
    L = n_fine - 1
    P = prolongation_tensor()
    cores = [P.copy() for _ in range(L)]
    cores[0] = cores[0][0:1]
    last = np.zeros((2, 1, 2, 2))
    last[0, 0, 0, 0] = 1.0
    last[0, 0, 0, 1] = 0.5
    last[1, 0, 0, 1] = 0.5
    cores.append(last)
    return cores


def retraction_mpo(n_fine):
    """src/prolongation.jl:66-68: swapprime(prolongation_mpo, 0 => 1) = transpose of each core."""
    return [np.transpose(c, (0, 1, 3, 2)).copy() for c in prolongation_mpo(n_fine)]


# ------------------------------------------------------------------ boundary RHS / analytic QTTs (src/qtt_utils.jl)
def boundary_value_mps(n, ui, uf):
    """src/qtt_utils.jl:1-14: b[1]=ui, b[N]=uf, zero elsewhere, as a bond-2 MPS."""
    dt = np.result_type(type(ui), type(uf), np.float64)
    cores = []
    first = np.zeros((1, 2, 2), dtype=dt)
    first[0] = np.eye(2)
    cores.append(first)
    a = np.zeros((2, 2, 2), dtype=dt)
    a[0, 0, 0] = 1.0
    a[1, 1, 1] = 1.0
    for _ in range(1, n - 1):
        cores.append(a.copy())
    last = np.zeros((2, 2, 1), dtype=dt)
    last[0, 0, 0] = ui
    last[1, 1, 0] = uf
    cores.append(last)
    return cores


def boundary_value_vector(n, ui, uf):
    """src/qtt_utils.jl:16-23."""
    b = np.zeros(2 ** n, dtype=np.result_type(type(ui), type(uf), np.float64))
    b[0] = ui
    b[-1] = uf
    return b


def qtt_exp(alpha, beta, n):
    """src/qtt_utils.jl:34-37: exp(beta) * prod_j [1, exp(alpha 2^-j)]."""
    dt = np.result_type(type(alpha), type(beta), np.float64)
    cores = [np.array([1.0, np.exp(alpha * 2.0 ** (-(j + 1)))], dtype=dt).reshape(1, 2, 1) for j in range(n)]
    cores[0] = cores[0] * np.exp(beta)
    return cores


def qtt_sin(alpha, n, beta=0.0):
    """src/qtt_utils.jl:49-58: (qtt(exp, i alpha, beta) - qtt(exp, -i alpha, beta)) / 2i  [direct sum]."""
    p1 = qtt_exp(1j * alpha, beta, n)
    p2 = mps_scale(qtt_exp(-1j * alpha, beta, n), -1.0)
    return mps_scale(mps_add_directsum(p1, p2), 1.0 / 2j)


def qtt_cos(alpha, n, beta=0.0):
    """src/qtt_utils.jl:67-76."""
    p1 = qtt_exp(1j * alpha, beta, n)
    p2 = qtt_exp(-1j * alpha, beta, n)
    return mps_scale(mps_add_directsum(p1, p2), 0.5)


# ------------------------------------------------------------------ polynomial QTT (src/function_to_mps.jl:55-96)
def polynomial_tensor(j, kappa):
    """src/function_to_mps.jl:55-67 (j is the 1-based site)."""
    Q = np.zeros((kappa + 1, kappa + 1, 2))
    for a in range(kappa + 1):
        Q[a, a, 0] = 1.0
        Q[a, a, 1] = 1.0
    for b in range(1, kappa + 1):
        for a in range(b):
            Q[a, b, 1] = comb(b, a) * 2.0 ** (-(b - a) * j)
    return Q


def function_to_mps_polynomial(f, n, xstart, xstop, degree, length=None, cutoff=1e-15):
    """src/function_to_mps.jl:71-96: least-squares polynomial fit on `length` grid points of
    [xstart, xstop) (Polynomials.fit), evaluated at t = sum_j b_j 2^-j, then truncate!."""
    length = 2 ** n if length is None else length
    xr = (xstart + (xstop - xstart) * np.arange(length) / length)
    V = np.vander(xr, degree + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V, np.array([f(x) for x in xr]), rcond=None)
    cores = []
    for j in range(1, n + 1):
        Q = polynomial_tensor(j, degree)          # (alpha, beta, s)
        cores.append(np.transpose(Q, (0, 2, 1)))  # (l, s, r)
    cores[0] = cores[0][0:1]
    cores[-1] = np.tensordot(cores[-1], c, axes=([2], [0]))[:, :, None]
    return mps_truncate(cores, cutoff=cutoff)


# ------------------------------------------------------------------ DFT layers (src/dft.jl)
def dft_matrix(n):
    """src/dft.jl:35-39: omega^(jk)/sqrt(N), omega = exp(-2 pi i / N)."""
    N = 2 ** n
    j = np.arange(N)
    return np.exp(-2j * np.pi * np.outer(j, j) / N) / np.sqrt(N)


def _op_phase(phi):
    return np.array([[1.0, 0.0], [0.0, np.exp(1j * phi)]], dtype=np.complex128)


def phase_mpo(n):
    """src/dft.jl:51-66.  Operator matrices are stored [s_in, s_out]; 'I' and 'Phase' are
    diagonal so orientation does not matter here."""
    if n == 1:
        return mpo_identity(1, np.complex128)
    cores = []
    first = np.zeros((1, 2, 2, 2), dtype=np.complex128)
    for s in range(2):
        first[0, s, s, s] = 1.0  # delta(s, s', l)
    cores.append(first)
    for j in range(2, n):
        c = np.zeros((2, 2, 2, 2), dtype=np.complex128)
        c[0, 0] = np.eye(2)
        c[1, 1] = _op_phase(-np.pi / 2 ** (j - 1))
        cores.append(c)
    c = np.zeros((2, 1, 2, 2), dtype=np.complex128)
    c[0, 0] = np.eye(2)
    c[1, 0] = _op_phase(-np.pi / 2 ** (n - 1))
    cores.append(c)
    return cores


def hadamard_phase_mpo(n):
    """src/dft.jl:68-74: apply H on the output of site 1 after the controlled phases."""
    U = phase_mpo(n)
    H = np.array([[1.0, 1.0], [1.0, -1.0]], dtype=np.complex128) / np.sqrt(2.0)
    # U[1][l, r, s, s'] -> sum_{s'} U[..., s, s'] H[s'', s']   (H symmetric)
    U[0] = np.einsum("lrst,ut->lrsu", U[0], H)
    return U


def dft_layer(n, j):
    """src/dft.jl:84-93: Us[j] = [MPO(s[1:j-1], "I"); hadamard_phase_mpo(s[j:n])]  (j 1-based)."""
    return mpo_identity(j - 1, np.complex128) + hadamard_phase_mpo(n - j + 1)


# ------------------------------------------------------------------ Airy / Helmholtz systems (examples/)
def airy_solution(x, alpha=1.0, beta=0.0):
    """examples/airy_equation/src/airy_utils.jl:8: alpha Ai(-x) + beta Bi(-x)."""
    from scipy.special import airy
    ai, _, bi, _ = airy(-np.asarray(x, dtype=np.float64))
    return alpha * ai + beta * bi


def airy_mpo(n, xi, xf):
    """examples/airy_equation/src/airy_utils.jl:10-27:  A = (-Laplacian(s, 1)) (+) h^2 diag(q),
    q(t) = -((xf - xi) t + xi) as a degree-1 polynomial QTT (bond 2) -> MPO bond 5."""
    A1 = mpo_scale(laplacian_mpo(n, 1.0), -1.0)
    h = (xf - xi) / 2 ** n
    f = lambda x: -((xf - xi) * x + xi)
    # NOTE the reference passes (xi, xf) as the fit interval but the polynomial is evaluated at
    # t in [0, 1); for a degree-1 fit the fitted line is exact, so the fit
    # interval is irrelevant.
    q = function_to_mps_polynomial(f, n, xi, xf, degree=1, length=1000, cutoff=1e-8)
    A2 = []
    for c in q:
        a, _, b = c.shape
        W = np.zeros((a, b, 2, 2))
        for s in range(2):
            W[:, :, s, s] = c[:, s, :]
        A2.append(W)
    A2 = mpo_scale(A2, h ** 2)
    return mpo_add_directsum(A1, A2)


def airy_matrix(n, xi, xf):
    """examples/airy_equation/src/airy_utils.jl:41-48."""
    N = 2 ** n
    h = (xf - xi) / N
    A = np.diag(np.full(N, 2.0)) - np.diag(np.ones(N - 1), 1) - np.diag(np.ones(N - 1), -1)
    return A + h ** 2 * np.diag(-(xi + np.arange(N) * h))


def helmholtz_lambda(n, nk, xi=0.0, xf=1.0):
    """examples/helmholtz_equation/helmholtz_equation.jl:48: -(2 pi (xf - xi) 2^(nk - n))^2."""
    return -((2 * np.pi * (xf - xi) * 2.0 ** (nk - n)) ** 2)

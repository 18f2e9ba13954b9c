"""Oracle MPO.MPS `apply` (density-matrix algorithm), MPO.MPO zip-up, and the front-ends built on them
(DFT, prolongation, retraction, Fourier interpolation).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to /root/reference;
"App. B.x" refers to SURVEY.md (upstream ITensorMPS semantics, restated).
"""
import numpy as np

from .core import (eigh_truncated, svd_truncated, orthogonalize, mps_reverse, mps_scale, mps_add,
                   project_bits, mpo_linkdims)
from .builders import dft_layer, prolongation_mpo, retraction_mpo

__all__ = [
    "mpo_orthogonalize", "apply_mpo_mpo", "apply_mpo_mps", "apply_mpo_mps_svd", "dft_mpo", "idft_mpo",
    "apply_dft_mpo", "apply_idft_mpo", "prolongate", "retract", "fourier_interpolation", "repeat_function",
    "mpo_truncate",
]


# ------------------------------------------------------------------ MPO . MPO  (zip-up, App. B.4)
def mpo_orthogonalize(M, center=0):
    M = [A.copy() for A in M]
    n = len(M)
    for j in range(n - 1, center, -1):
        a, c, d1, d2 = M[j].shape
        X = np.transpose(M[j], (0, 2, 3, 1)).reshape(a, d1 * d2 * c)
        Q, R = np.linalg.qr(X.T)
        k = Q.shape[1]
        M[j] = np.transpose(Q.T.reshape(k, d1, d2, c), (0, 3, 1, 2))
        M[j - 1] = np.transpose(np.tensordot(M[j - 1], R.T, axes=([1], [0])), (0, 3, 1, 2))
    for j in range(center):
        a, c, d1, d2 = M[j].shape
        X = np.transpose(M[j], (0, 2, 3, 1)).reshape(a * d1 * d2, c)
        Q, R = np.linalg.qr(X)
        k = Q.shape[1]
        M[j] = np.transpose(Q.reshape(a, d1, d2, k), (0, 3, 1, 2))
        M[j + 1] = np.tensordot(R, M[j + 1], axes=([1], [0]))
    return M


def mpo_truncate(M, cutoff=1e-15, maxdim=None):
    """ITensorMPS `truncate!(::MPO; cutoff)`: orthogonalize to the last site, SVD sweep leftwards."""
    n = len(M)
    M = mpo_orthogonalize(M, n - 1)
    for j in range(n - 1, 0, -1):
        a, c, d1, d2 = M[j].shape
        X = np.transpose(M[j], (0, 2, 3, 1)).reshape(a, d1 * d2 * c)
        U, S, Vh, _ = svd_truncated(X, cutoff, maxdim)
        k = S.size
        M[j] = np.transpose(Vh.reshape(k, d1, d2, c), (0, 3, 1, 2))
        M[j - 1] = np.transpose(np.tensordot(M[j - 1], U * S[None, :], axes=([1], [0])), (0, 3, 1, 2))
    return M


def apply_mpo_mpo(A, B, cutoff=1e-14, maxdim=None):
    """`apply(A::MPO, B::MPO; cutoff)` = operator product A o B (input -> B -> A), zip-up algorithm
    (App. B.4): orthogonalize both to site 1, sweep left to right factorising
    (R, A[i], B[i]) with `factorize(...; ortho="left")`, last two sites together, then truncate!."""
    n = len(A)
    assert len(B) == n
    if n == 1:
        return [np.einsum("abtu,cdst->acbdsu", A[0], B[0]).reshape(
            A[0].shape[0] * B[0].shape[0], A[0].shape[1] * B[0].shape[1], 2, 2)]
    A = mpo_orthogonalize(A, 0)
    B = mpo_orthogonalize(B, 0)
    if maxdim is None:
        maxdim = max(mpo_linkdims(A) + [1]) * max(mpo_linkdims(B) + [1])
    C = []
    R = np.ones((1, 1, 1), dtype=np.result_type(A[0], B[0]))  # (lC, lA, lB)
    for i in range(n - 2):
        # T[lC, s, s'', rA, rB] = R[lC, lA, lB] B[i][lB, rB, s, s'] A[i][lA, rA, s', s'']
        T = np.einsum("cab,bfst,aetu->csuef", R, B[i], A[i])
        lc, _, _, ra, rb = T.shape
        U, S, Vh, _ = svd_truncated(T.reshape(lc * 4, ra * rb), cutoff, maxdim)
        k = S.size
        C.append(np.transpose(U.reshape(lc, 2, 2, k), (0, 3, 1, 2)))
        R = (S[:, None] * Vh).reshape(k, ra, rb)
    i = n - 2
    T = np.einsum("cab,bfst,aetu,fhvw,egwx->csuvxgh", R, B[i], A[i], B[i + 1], A[i + 1])
    lc = T.shape[0]
    ra, rb = T.shape[5], T.shape[6]
    U, S, Vh, _ = svd_truncated(T.reshape(lc * 4, 4 * ra * rb), cutoff, maxdim)
    k = S.size
    C.append(np.transpose((U * S[None, :]).reshape(lc, 2, 2, k), (0, 3, 1, 2)))
    C.append(np.transpose(Vh.reshape(k, 2, 2, ra * rb), (0, 3, 1, 2)))
    return mpo_truncate(C, cutoff=cutoff, maxdim=maxdim)


# ------------------------------------------------------------------ MPO . MPS  (density matrix, App. B.4)
def apply_mpo_mps(A, psi, cutoff=1e-13, maxdim=None, mindim=1, normalize=False):
    """`apply(A::MPO, psi::MPS; alg="densitymatrix", cutoff, maxdim)`:
    E[1] = psi1 A1 conj(A1) conj(psi1), E[j] = E[j-1] psi_j A_j conj(A_j) conj(psi_j) (site outputs traced);
    from the right rho = E[n-1] (psi_n A_n)(conj) -> Hermitian eigen with truncation on the
    eigenvalues -> out[n] = U^dagger; carry R <- R U psi_{j} A_{j}; out[1] = R (not normalised).
    Output cores use the MPO's *output* site index (replaceprime 1 => 0)."""
    n = len(A)
    assert len(psi) == n
    if n == 1:
        T = np.einsum("asc,wvst->awtcv", psi[0], A[0])
        return [T.reshape(1, 2, 1)]
    mindim = max(mindim, 1)
    if maxdim is None:
        maxdim = max(mpo_linkdims(A) + [1]) * max([p.shape[2] for p in psi[:-1]] + [1])
    requested_maxdim = maxdim
    dt = np.result_type(A[0], psi[0])
    # left environments E[j] with legs (p, a, a~, p~): psi link, A link, conj A link, conj psi link
    E = []
    Ej = np.ones((1, 1, 1, 1), dtype=dt)
    for j in range(n - 1):
        T = np.tensordot(Ej, psi[j], axes=([0], [0]))                 # (a, a~, p~, s, p')
        T = np.tensordot(T, A[j], axes=([0, 3], [0, 2]))              # (a~, p~, p', a', t)
        T = np.tensordot(T, A[j].conj(), axes=([0, 4], [0, 3]))       # (p~, p', a', a~', s~)
        T = np.tensordot(T, psi[j].conj(), axes=([0, 4], [0, 1]))     # (p', a', a~', p~')
        Ej = T
        E.append(Ej)
    out = [None] * n
    # R carries (p, a, t_n..., r) -> we keep R as (p_left, a_left, k) where k is the kept right link
    R = np.einsum("psc,avst->patcv", psi[n - 1], A[n - 1]).reshape(psi[n - 1].shape[0], A[n - 1].shape[0], 2)
    # rho[(t, .), (t~, .)] for the last site: no right link
    Eprev = E[n - 2]
    T = np.tensordot(Eprev, R, axes=([0, 1], [0, 1]))                 # (a~, p~, t)
    rho = np.tensordot(T, R.conj(), axes=([0, 1], [1, 0]))            # (t, t~)
    D, U, _ = eigh_truncated(rho, cutoff, min(maxdim, rho.shape[0]), mindim)
    k = D.size
    out[n - 1] = U.T.reshape(k, 2, 1)                                 # F.Vt = V relabelled (not dagged)
    # R <- R * dag(Ut) * psi[n-1] * A[n-1]  : (p,a,t) U[t,k] -> (p,a,k)
    Rk = np.tensordot(R, U.conj(), axes=([2], [0]))                   # (p, a, k)
    for j in range(n - 2, 0, -1):
        # M[p_l, a_l, t, k] = psi_j[p_l, s, p] A_j[a_l, a, s, t] Rk[p, a, k]
        M = np.tensordot(psi[j], Rk, axes=([2], [0]))                 # (p_l, s, a, k)
        M = np.tensordot(A[j], M, axes=([1, 2], [2, 1]))              # (a_l, t, p_l, k)
        M = np.transpose(M, (2, 0, 1, 3))                             # (p_l, a_l, t, k)
        prod_dims = M.shape[0] * M.shape[1]
        md = min(prod_dims, requested_maxdim)
        Eprev = E[j - 1]
        T = np.tensordot(Eprev, M, axes=([0, 1], [0, 1]))             # (a~, p~, t, k)
        rho = np.tensordot(T, M.conj(), axes=([0, 1], [1, 0]))        # (t, k, t~, k~)
        d = rho.shape[0] * rho.shape[1]
        D, U, _ = eigh_truncated(rho.reshape(d, d), cutoff, md, mindim)
        knew = D.size
        out[j] = U.T.reshape(knew, 2, k)
        Rk = np.tensordot(M, U.conj().reshape(2, k, knew), axes=([2, 3], [0, 1]))  # (p_l, a_l, knew)
        k = knew
    M = np.tensordot(psi[0], Rk, axes=([2], [0]))                     # (1, s, a, k)
    M = np.tensordot(A[0], M, axes=([1, 2], [2, 1]))                  # (1, t, 1, k)
    first = M.reshape(2, k)[None, :, :]
    if normalize:
        first = first / np.linalg.norm(first)
    out[0] = first
    return out


def apply_mpo_mps_svd(A, psi, cutoff=1e-13, maxdim=None):
    """Exact product (bond w*chi) followed by SVD truncation: a numerically more robust route to
    the same state; used by tests to bound the density-matrix algorithm's own reproducibility."""
    from .core import mps_truncate
    out = []
    for W, X in zip(A, psi):
        T = np.einsum("asc,wvst->awtcv", X, W)
        a, w, _, c, v = T.shape
        out.append(T.reshape(a * w, 2, c * v))
    return mps_truncate(out, cutoff=cutoff, maxdim=maxdim)


# ------------------------------------------------------------------ DFT front-ends (src/dft.jl)
def dft_mpo(n, cutoff=1e-15):
    """src/dft.jl:81-99 (alg="mpo"): U = Us[1]; U = apply(U, Us[j]; cutoff) for j = 2..n;
    then swapprime(0 <-> 1) and reverse_output_bits.  The output-bit reversal is an Index
    relabelling in the reference (:76-79): primed site j is *named* site n+1-j.  In the dense
    convention the relabelling is carried separately: this function returns the MPO *before*
    relabelling (F_raw) whose output site q carries frequency bit n+1-q, exactly the object
    `apply` sees, and `mpo_to_mat(F_raw, reverse_output_sites=True)` is the DFT matrix."""
    U = dft_layer(n, 1)
    for j in range(2, n + 1):
        U = apply_mpo_mpo(U, dft_layer(n, j), cutoff=cutoff)
    return [np.transpose(c, (0, 1, 3, 2)).copy() for c in U]  # swapprime(U, 0 => 1)


def idft_mpo(n, cutoff=1e-15):
    """src/dft.jl:45-47: swapprime(dag(dft_mpo), 0 => 1) (raw, i.e. before bit relabelling)."""
    return [np.transpose(c.conj(), (0, 1, 3, 2)).copy() for c in dft_mpo(n, cutoff)]


def apply_dft_mpo(psi, cutoff=1e-13, maxdim=None, F=None):
    """src/dft.jl:121-126: reverse(apply(F, psi; kwargs...)).  Because of reverse_output_bits the
    raw apply result has output site q = frequency bit n+1-q; `reverse` flips the chain so site 1
    is the MSB of the frequency index."""
    n = len(psi)
    F = dft_mpo(n) if F is None else F
    out = apply_mpo_mps(F, [p.astype(np.complex128) for p in psi], cutoff=cutoff, maxdim=maxdim)
    return mps_reverse(out)


def apply_idft_mpo(psi, cutoff=1e-13, maxdim=None, Finv=None):
    """src/dft.jl:128-133: apply(idft_mpo, reverse(psi)).  With the relabelling made explicit:
    idft (relabelled) input site j is raw *output* leg n+1-j of dag(F_raw); feeding reverse(psi)
    through the transposed-conjugated raw cores in reversed chain order realises that."""
    n = len(psi)
    if Finv is None:
        Finv = idft_mpo(n)
    # Finv raw: core q maps input leg (old output, frequency bit n+1-q) -> output leg (time bit q).
    # psi site j = frequency bit j (MSB first) must meet core q = n+1-j: reverse psi.
    rpsi = mps_reverse([p.astype(np.complex128) for p in psi])
    return apply_mpo_mps(Finv, rpsi, cutoff=cutoff, maxdim=maxdim)


# ------------------------------------------------------------------ prolongation / retraction (src/prolongation.jl)
def prolongate(psi, nnew=1, cutoff=1e-8, maxdim=None):
    """src/prolongation.jl:29-60 (1-D): append a dummy site, apply the prolongation MPO, repeat."""
    for _ in range(nnew):
        n = len(psi) + 1
        P = prolongation_mpo(n)
        dummy = np.zeros((1, 2, 1), dtype=psi[0].dtype)
        dummy[0, 0, 0] = 1.0
        psi = apply_mpo_mps(P, list(psi) + [dummy], cutoff=cutoff, maxdim=maxdim)
    return psi


def retract(psi, nretract=1, cutoff=1e-8, maxdim=None):
    """src/prolongation.jl:70-102 (1-D): apply the transposed prolongation MPO, absorb the last
    (dummy-output) site into its neighbour, divide by 2."""
    for _ in range(nretract):
        n = len(psi)
        R = retraction_mpo(n)
        out = apply_mpo_mps(R, psi, cutoff=cutoff, maxdim=maxdim)
        last = out[-1][:, 0, :]                       # dummy output leg = index 0
        out[-2] = np.tensordot(out[-2], last, axes=([2], [0]))
        out = out[:-1]
        psi = mps_scale(out, 0.5)
    return psi


# ------------------------------------------------------------------ Fourier interpolation (src/fourier_interpolation.jl)
def _basis_site(bit, dtype=np.complex128):
    c = np.zeros((1, 2, 1), dtype=dtype)
    c[0, bit, 0] = 1.0
    return c


def fourier_interpolation(psi, k=1, cutoff=1e-15):
    """src/fourier_interpolation.jl:1-11: DFT, split on the frequency MSB, pad k+1 sites of 0s / 1s
    in front of each half, sum, inverse DFT, scale by 2^(k/2)."""
    F = apply_dft_mpo(psi, cutoff=cutoff)
    F0 = project_bits(F, [0])
    F1 = project_bits(F, [1])
    F0s = [_basis_site(0) for _ in range(k + 1)] + F0
    F1s = [_basis_site(1) for _ in range(k + 1)] + F1
    Fi = mps_add(F0s, F1s, cutoff=1e-15)
    return mps_scale(apply_idft_mpo(Fi, cutoff=cutoff), 2.0 ** (k / 2))


def repeat_function(psi, k=1, cutoff=1e-15):
    """src/fourier_interpolation.jl:21-31: DFT, append k |0> sites (low frequency bits), inverse DFT."""
    F = apply_dft_mpo(psi, cutoff=cutoff)
    Fk = F + [_basis_site(0) for _ in range(k)]
    return mps_scale(apply_idft_mpo(Fk, cutoff=cutoff), 2.0 ** (k / 2))

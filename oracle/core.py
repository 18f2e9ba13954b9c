"""Oracle core: truncation rule, factorisations, MPS/MPO algebra, grid <-> QTT conversions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference file:line
(relative to /root/reference) or the upstream routine it restates.
"""
import numpy as np
import scipy.linalg as sla

__all__ = [
    "truncate_spectrum", "svd_truncated", "eigh_truncated", "qtt_xrange", "vec_to_tensor", "vec_to_mps",
    "mps_to_tensor", "mps_to_discrete_function", "mpo_to_mat", "mat_to_mpo", "project_bits",
    "interleave_bits", "float_to_bits", "bits_to_index", "index_to_bits", "mps_inner", "mps_norm",
    "mpo_expect", "mpo_mps_inner4", "sqeuclidean", "sqeuclidean_normalized", "sqeuclidean_mps",
    "sqeuclidean_normalized_mps", "fidelity_defect", "orthogonalize", "mps_truncate", "mps_add_directsum",
    "mps_add", "mps_scale", "maxlinkdim", "linkdims", "random_mps", "insert_identity_links", "interleave_mps",
    "mps_reverse", "apply_mpo_dense", "mps_contract_dense", "integrate_mps", "mpo_add_directsum", "mpo_scale",
    "interleave_mpo", "mpo_identity", "mpo_linkdims",
]


# --------------------------------------------------------------------------------------------
# Truncation (NDTensors.truncate!)
# --------------------------------------------------------------------------------------------
def truncate_spectrum(P, cutoff=0.0, maxdim=None, mindim=1, use_absolute_cutoff=False, use_relative_cutoff=True):
    """Restates `NDTensors.truncate!(P; cutoff, maxdim, mindim)` on P = sigma^2 sorted descending.

    Returns (n_keep, truncerr).  Negative tail -> 0; discard while n>maxdim; then (relative mode)
    discard while n>mindim and err+P[n] <= cutoff*sum(P).  Keeps at least one value.
    """
    P = np.array(P, dtype=np.float64, copy=True)
    origm = P.size
    if origm == 0:
        return 0, 0.0
    mindim = max(int(mindim), 1)
    maxdim = origm if maxdim is None else min(int(maxdim), origm)
    if origm == 1:
        return 1, 0.0
    # zero out negative weight (can appear from eigen of a density matrix)
    P[P < 0.0] = 0.0
    truncerr = 0.0
    n = origm
    while n > maxdim:
        truncerr += P[n - 1]
        n -= 1
    if use_absolute_cutoff:
        while n > mindim and truncerr + P[n - 1] <= cutoff:
            truncerr += P[n - 1]
            n -= 1
    else:
        scale = 1.0
        if use_relative_cutoff:
            scale = P.sum()
            if scale == 0.0:
                scale = 1.0
        while n > mindim and truncerr + P[n - 1] <= cutoff * scale:
            truncerr += P[n - 1]
            n -= 1
        truncerr /= scale
    n = max(n, 1)
    return n, truncerr


def _svd(M):
    """LAPACK gesdd with gesvd fallback (NDTensors dense `svd`)."""
    try:
        return sla.svd(M, full_matrices=False, lapack_driver="gesdd")
    except Exception:  # pragma: no cover
        return sla.svd(M, full_matrices=False, lapack_driver="gesvd")


def svd_truncated(M, cutoff=0.0, maxdim=None, mindim=1):
    """ITensors `svd(T, linds; cutoff, maxdim, mindim)`: returns U, S, Vh, truncerr (kept part)."""
    U, S, Vh = _svd(M)
    k, err = truncate_spectrum(S * S, cutoff, maxdim, mindim)
    return U[:, :k], S[:k], Vh[:k, :], err


def eigh_truncated(rho, cutoff=0.0, maxdim=None, mindim=1):
    """ITensors `eigen(rho; ishermitian=true, cutoff, maxdim, mindim)`: descending eigenvalues,
    truncation applied to the eigenvalues (= sigma^2).  Returns (D_kept, U_kept, truncerr)."""
    rho = 0.5 * (rho + rho.conj().T)
    D, U = sla.eigh(rho)
    D = D[::-1]
    U = U[:, ::-1]
    k, err = truncate_spectrum(D, cutoff, maxdim, mindim)
    return D[:k], U[:, :k], err


# --------------------------------------------------------------------------------------------
# Grid and bit maps (bit-exact contract)
# --------------------------------------------------------------------------------------------
def qtt_xrange(n, xstart, xstop):
    """src/function_to_mps.jl:10-12: range(start, stop, length=2^n+1)[1:2^n].

    Julia's `range(start, stop; length)` evaluates element j as a two-sided linear interpolation
    (`lerpi`): x_j = (1 - j/L)*start + (j/L)*stop with L = 2^n.  For power-of-two L both
    fractions are exact, so this equals start + j*(stop-start)/L up to the final rounding."""
    L = 2 ** n
    j = np.arange(L, dtype=np.float64)
    t = j / L
    return (1.0 - t) * xstart + t * xstop


def vec_to_tensor(v):
    """src/function_to_mps.jl:170-173: permutedims(reshape(v, 2,...,2), n:-1:1) == C-order reshape:
    tensor index k (site k) is bit k of the grid index counted from the most significant bit."""
    v = np.asarray(v)
    n = int(round(np.log2(v.size)))
    assert 2 ** n == v.size
    return v.reshape((2,) * n)


def bits_to_index(bits):
    """Grid index j = sum_k b_k 2^(n-k) (site 1 = MSB), src/function_to_mps.jl:170-173."""
    j = 0
    for b in bits:
        j = (j << 1) | int(b)
    return j


def index_to_bits(j, n):
    return [(j >> (n - 1 - k)) & 1 for k in range(n)]


def float_to_bits(x, n):
    """examples/helmholtz_equation/helmholtz_equation.jl:630-641: reverse(digits(round(Int, x*2^n), base=2, pad=n))."""
    return index_to_bits(int(np.round(x * 2 ** n)), n)


def interleave_bits(x, y):
    """src/mps_utils.jl:72-81: z[2j-1]=x[j], z[2j]=y[j]."""
    assert len(x) == len(y)
    z = []
    for a, b in zip(x, y):
        z += [a, b]
    return z


def vec_to_mps(v, cutoff=1e-15, maxdim=None, mindim=1):
    """src/function_to_mps.jl:175-177 -> ITensors `MPS(array, sites; cutoff)`: left-to-right
    sequence of `factorize(...; ortho="left")` (SVD for cutoff <= 1e-12),
    orthogonality centre on the last site."""
    T = vec_to_tensor(np.asarray(v))
    n = T.ndim
    cores = []
    rest = T.reshape(1, -1)
    chi = 1
    for j in range(n - 1):
        M = rest.reshape(chi * 2, -1)
        U, S, Vh, _ = svd_truncated(M, cutoff, maxdim, mindim)
        k = S.size
        cores.append(U.reshape(chi, 2, k))
        rest = S[:, None] * Vh
        chi = k
    cores.append(rest.reshape(chi, 2, 1))
    return cores


def mps_to_tensor(psi):
    """contract(psi) as an ndarray with one axis per site (site order)."""
    acc = psi[0]
    for A in psi[1:]:
        acc = np.tensordot(acc, A, axes=([acc.ndim - 1], [0]))
    assert acc.shape[0] == 1 and acc.shape[-1] == 1
    return acc.reshape(acc.shape[1:-1])


def mps_contract_dense(psi):
    return mps_to_tensor(psi)


def mps_to_discrete_function(psi, ndims=1):
    """src/function_to_mps.jl:136-145.  1-D: C-order flatten of the site tensor (site 1 = MSB).
    N-D: sites interleaved s[j:ndims:n] per dimension; output array dims = (grid dim 1, grid dim 2, ...)."""
    T = mps_to_tensor(psi)
    n = T.ndim
    if ndims == 1:
        return T.reshape(-1)
    assert n % ndims == 0
    perm = []
    for d in range(ndims):
        perm += list(range(d, n, ndims))
    nb = n // ndims
    return np.transpose(T, perm).reshape((2 ** nb,) * ndims)


def mpo_to_mat(M, reverse_input_sites=False, reverse_output_sites=False):
    """src/function_to_mps.jl:183-192: primed (output) indices on rows, site 1 = MSB."""
    n = len(M)
    acc = M[0]  # (wl, wr, s, s')
    acc = acc.reshape(acc.shape[0], acc.shape[1], 2, 2)
    T = np.transpose(acc, (0, 2, 3, 1))  # (wl, s1, s1', wr)
    for A in M[1:]:
        T = np.tensordot(T, np.transpose(A, (0, 2, 3, 1)), axes=([T.ndim - 1], [0]))
    assert T.shape[0] == 1 and T.shape[-1] == 1
    T = T.reshape(T.shape[1:-1])  # (s1, s1', s2, s2', ...)
    ins = list(range(0, 2 * n, 2))
    outs = list(range(1, 2 * n, 2))
    if reverse_input_sites:
        ins = ins[::-1]
    if reverse_output_sites:
        outs = outs[::-1]
    return np.transpose(T, outs + ins).reshape(2 ** n, 2 ** n)


def mat_to_mpo(m, cutoff=1e-15, maxdim=None):
    """src/function_to_mps.jl:194-201: reshape rows/cols into bits (MSB first), pair (s_j, s_j'),
    then ITensors `MPO(array, sites; cutoff)` = sequential left-to-right SVD."""
    m = np.asarray(m)
    n = int(round(np.log2(m.shape[0])))
    t = m.reshape((2,) * (2 * n))  # (out bits..., in bits...)
    perm = []
    for j in range(n):
        perm += [n + j, j]  # (s_j (in), s_j' (out))
    t = np.transpose(t, perm)
    cores = []
    rest = t.reshape(1, -1)
    w = 1
    for j in range(n - 1):
        Mj = rest.reshape(w * 4, -1)
        U, S, Vh, _ = svd_truncated(Mj, cutoff, maxdim)
        k = S.size
        cores.append(np.transpose(U.reshape(w, 2, 2, k), (0, 3, 1, 2)))
        rest = S[:, None] * Vh
        w = k
    cores.append(np.transpose(rest.reshape(w, 2, 2, 1), (0, 3, 1, 2)))
    return cores


def project_bits(psi, bits, right_bits=None):
    """src/project_bits.jl:1-15 (leading bits) and :25-44 (trailing bits via reverse; an integer
    `right_bits` means that many zero bits, :32-36).  `right_bits` are listed in site order.
    Returns a shorter MPS, or a scalar if every site is projected (:11-13)."""
    left = [int(b) for b in bits]
    if right_bits is None:
        right = []
    elif np.isscalar(right_bits):
        right = [0] * int(right_bits)
    else:
        right = [int(b) for b in right_bits]
    n = len(psi)
    nl, nr = len(left), len(right)
    assert nl + nr <= n
    dt = psi[0].dtype
    pl = np.ones((1, 1), dtype=dt)
    for j, b in enumerate(left):
        pl = pl @ psi[j][:, b, :]
    pr = np.ones((1, 1), dtype=dt)
    for k in range(nr):
        pr = psi[n - 1 - k][:, right[nr - 1 - k], :] @ pr
    rest = [A.copy() for A in psi[nl:n - nr]]
    if len(rest) == 0:
        return (pl @ pr)[0, 0]
    rest[0] = np.tensordot(pl, rest[0], axes=([1], [0]))
    rest[-1] = np.tensordot(rest[-1], pr, axes=([2], [0]))
    return rest


# --------------------------------------------------------------------------------------------
# MPS / MPO algebra (ITensorMPS inner, norm, orthogonalize!, truncate!, +)
# --------------------------------------------------------------------------------------------
def mps_inner(x, y):
    """inner(x, y) = <x|y> (conjugates x)."""
    E = np.ones((1, 1), dtype=np.result_type(x[0], y[0]))
    for A, B in zip(x, y):
        T = np.tensordot(E, B, axes=([1], [0]))            # (ax, s, by)
        E = np.tensordot(A.conj(), T, axes=([0, 1], [0, 1]))  # (ax', by')
    return E.reshape(-1)[0]


def mps_norm(x):
    return float(np.sqrt(abs(mps_inner(x, x).real)))


def mpo_expect(y, A, x):
    """inner(y', A, x) = <y|A|x>."""
    E = np.ones((1, 1, 1), dtype=np.result_type(y[0], A[0], x[0]))
    for Y, W, X in zip(y, A, x):
        T = np.tensordot(E, X, axes=([2], [0]))                  # (ay, w, s, cx)
        T = np.tensordot(T, W, axes=([1, 2], [0, 2]))            # (ay, cx, w', s')
        E = np.tensordot(Y.conj(), T, axes=([0, 1], [0, 3]))     # (cy, cx, w')
        E = np.transpose(E, (0, 2, 1))
    return E.reshape(-1)[0]


def mpo_mps_inner4(A, x, B, y):
    """inner(A, x, B, y) = <A x | B y>."""
    E = np.ones((1, 1, 1, 1), dtype=np.result_type(A[0], x[0], B[0], y[0]))  # (ax, wa, wb, ay)
    for WA, X, WB, Y in zip(A, x, B, y):
        T = np.tensordot(E, Y, axes=([3], [0]))                   # (ax, wa, wb, s, cy)
        T = np.tensordot(T, WB, axes=([2, 3], [0, 2]))            # (ax, wa, cy, wb', s')
        T = np.tensordot(T, WA.conj(), axes=([1, 4], [0, 3]))     # (ax, cy, wb', wa', s)
        E = np.tensordot(X.conj(), T, axes=([0, 1], [0, 4]))      # (cx, cy, wb', wa')
        E = np.transpose(E, (0, 3, 2, 1))
    return E.reshape(-1)[0]


def sqeuclidean_mps(A, x, b):
    """src/mps_utils.jl:109-111: inner(A,x,A,x) - 2 real(inner(b',A,x)) + inner(b,b)."""
    return (mpo_mps_inner4(A, x, A, x) - 2 * np.real(mpo_expect(b, A, x)) + mps_inner(b, b)).real


def sqeuclidean_normalized_mps(A, x, b):
    """src/mps_utils.jl:113-116: 1 + (inner(A,x,A,x)/bb - 2 real(inner(b',A,x))/bb)."""
    bb = mps_inner(b, b)
    return (1.0 + (mpo_mps_inner4(A, x, A, x) / bb - 2 * np.real(mpo_expect(b, A, x)) / bb)).real


def sqeuclidean(x, y):
    """src/mps_utils.jl:134-136 (vector form)."""
    x = np.asarray(x); y = np.asarray(y)
    return (np.vdot(x, x) - 2 * np.real(np.vdot(y, x)) + np.vdot(y, y)).real


def sqeuclidean_normalized(x, y):
    """src/mps_utils.jl:138-141 (vector form)."""
    x = np.asarray(x); y = np.asarray(y)
    yy = np.vdot(y, y)
    return (1.0 + (np.vdot(x, x) / yy - 2 * np.real(np.vdot(y, x)) / yy)).real


def fidelity_defect(x, y):
    """1 - |<x|y>| / (|x||y|)  (north-star fidelity gate; `normalized_inner`, src/mps_utils.jl:1)."""
    return 1.0 - abs(mps_inner(x, y)) / (mps_norm(x) * mps_norm(y))


def maxlinkdim(psi):
    return max([A.shape[-1] if A.ndim == 3 else A.shape[1] for A in psi[:-1]] + [1])


def linkdims(psi):
    return [A.shape[2] for A in psi[:-1]]


def mpo_linkdims(M):
    return [A.shape[1] for A in M[:-1]]


def orthogonalize(psi, center):
    """ITensorMPS `orthogonalize!(psi, center)` (0-based centre): QR sweeps from both ends."""
    psi = [A.copy() for A in psi]
    n = len(psi)
    for j in range(center):
        a, d, c = psi[j].shape
        Q, R = np.linalg.qr(psi[j].reshape(a * d, c))
        psi[j] = Q.reshape(a, d, Q.shape[1])
        psi[j + 1] = np.tensordot(R, psi[j + 1], axes=([1], [0]))
    for j in range(n - 1, center, -1):
        a, d, c = psi[j].shape
        Q, R = np.linalg.qr(psi[j].reshape(a, d * c).T)
        psi[j] = Q.T.reshape(Q.shape[1], d, c)
        psi[j - 1] = np.tensordot(psi[j - 1], R.T, axes=([2], [0]))
    return psi


def mps_truncate(psi, cutoff=1e-15, maxdim=None, mindim=1):
    """ITensorMPS `truncate!(psi; cutoff, maxdim)`: orthogonalize to the last site, then sweep
    right-to-left with truncated SVDs."""
    n = len(psi)
    psi = orthogonalize(psi, n - 1)
    for j in range(n - 1, 0, -1):
        a, d, c = psi[j].shape
        U, S, Vh, _ = svd_truncated(psi[j].reshape(a, d * c), cutoff, maxdim, mindim)
        psi[j] = Vh.reshape(S.size, d, c)
        psi[j - 1] = np.tensordot(psi[j - 1], U * S[None, :], axes=([2], [0]))
    return psi


def mps_scale(psi, alpha):
    out = [A.copy() for A in psi]
    out[0] = out[0] * alpha
    return out


def mps_add_directsum(*psis):
    """`+(psi...; alg="directsum")`: block-diagonal cores, first core
    row-concatenated, last core column-concatenated."""
    n = len(psis[0])
    dt = np.result_type(*[p[0] for p in psis])
    out = []
    for j in range(n):
        if n == 1:
            out.append(sum(p[0].astype(dt) for p in psis))
            continue
        if j == 0:
            out.append(np.concatenate([p[j].astype(dt) for p in psis], axis=2))
        elif j == n - 1:
            out.append(np.concatenate([p[j].astype(dt) for p in psis], axis=0))
        else:
            L = sum(p[j].shape[0] for p in psis)
            R = sum(p[j].shape[2] for p in psis)
            C = np.zeros((L, 2, R), dtype=dt)
            l0 = r0 = 0
            for p in psis:
                a, _, c = p[j].shape
                C[l0:l0 + a, :, r0:r0 + c] = p[j]
                l0 += a; r0 += c
            out.append(C)
    return out


def mps_add(*psis, cutoff=1e-15, maxdim=None):
    """`+(psi...; cutoff)`.  ITensorMPS' default is the density-matrix sum; the oracle uses
    direct sum followed by SVD truncation, which yields the same state to O(sqrt(cutoff))."""
    return mps_truncate(mps_add_directsum(*psis), cutoff=cutoff, maxdim=maxdim)


def mpo_add_directsum(*Ms):
    n = len(Ms[0])
    dt = np.result_type(*[m[0] for m in Ms])
    out = []
    for j in range(n):
        if n == 1:
            out.append(sum(m[0].astype(dt) for m in Ms)); continue
        if j == 0:
            out.append(np.concatenate([m[j].astype(dt) for m in Ms], axis=1))
        elif j == n - 1:
            out.append(np.concatenate([m[j].astype(dt) for m in Ms], axis=0))
        else:
            L = sum(m[j].shape[0] for m in Ms); R = sum(m[j].shape[1] for m in Ms)
            C = np.zeros((L, R, 2, 2), dtype=dt)
            l0 = r0 = 0
            for m in Ms:
                a, c = m[j].shape[:2]
                C[l0:l0 + a, r0:r0 + c] = m[j]
                l0 += a; r0 += c
            out.append(C)
    return out


def mpo_scale(M, alpha):
    out = [A.copy() for A in M]
    out[0] = out[0] * alpha
    return out


def mpo_identity(n, dtype=np.float64):
    """MPO(s, "I")."""
    return [np.eye(2, dtype=dtype).reshape(1, 1, 2, 2).copy() for _ in range(n)]


def random_mps(n, chi, seed, dtype=np.float64, normalize=True):
    """Synthetic stand-in for `random_mps(s; linkdims)`: standard-normal cores from numpy
    default_rng(seed), right-orthonormalised (Julia's RNG stream is not reproducible here;
    SURVEY.md section 8c: identical inputs are exchanged as explicit arrays)."""
    rng = np.random.default_rng(seed)
    dims = [1]
    for j in range(1, n):
        dims.append(int(min(chi, 2 ** j, 2 ** (n - j))))
    dims.append(1)
    psi = []
    for j in range(n):
        shp = (dims[j], 2, dims[j + 1])
        A = rng.standard_normal(shp)
        if np.issubdtype(dtype, np.complexfloating):
            A = A + 1j * rng.standard_normal(shp)
        psi.append(A.astype(dtype))
    psi = orthogonalize(psi, 0)
    if normalize:
        psi[0] = psi[0] / np.linalg.norm(psi[0])
    return psi


def mps_reverse(psi):
    """Base.reverse(psi) (src/mps_utils.jl:12): reverse the chain; each core's link legs swap."""
    return [np.transpose(A, (2, 1, 0)).copy() for A in psi[::-1]]


def insert_identity_links(psi):
    """src/mps_utils.jl:26-41 specialised to the dense-core convention: length 2n-1 chain whose
    even sites are bare link identities (no site leg), represented as (chi, 1, chi) cores."""
    raise NotImplementedError("only used through interleave_mps")


def interleave_mps(psi1, psi2):
    """src/mps_utils.jl:55-66: sites alternate psi1[1], psi2[1], psi1[2], ...; the link carried
    across the other chain's site is the Kronecker product with an identity."""
    n = len(psi1)
    assert len(psi2) == n
    out = []
    for j in range(n):
        A = psi1[j]; B = psi2[j]
        a1, _, c1 = A.shape
        b1, _, c2 = B.shape
        # site 2j (from psi1): links (a1 x b1) -> (c1 x b1), identity on psi2's link
        T = np.einsum("asc,bd->abscd", A, np.eye(b1, dtype=B.dtype))
        out.append(T.reshape(a1 * b1, 2, c1 * b1))
        # site 2j+1 (from psi2): links (c1 x b1) -> (c1 x c2), identity on psi1's link
        T = np.einsum("ce,bsd->cbsed", np.eye(c1, dtype=A.dtype), B)
        out.append(T.reshape(c1 * b1, 2, c1 * c2))
    return out


def interleave_mpo(M1, M2):
    """interleave for MPOs (src/mps_utils.jl:55-66), used by the 2-D Laplacian (src/laplacian.jl:48-49)."""
    n = len(M1)
    out = []
    for j in range(n):
        A = M1[j]; B = M2[j]
        a1, c1 = A.shape[:2]
        b1, c2 = B.shape[:2]
        T = np.einsum("acst,bd->abcdst", A, np.eye(b1, dtype=B.dtype))
        out.append(T.reshape(a1 * b1, c1 * b1, 2, 2))
        T = np.einsum("ce,bdst->cbedst", np.eye(c1, dtype=A.dtype), B)
        out.append(T.reshape(c1 * b1, c1 * c2, 2, 2))
    return out


def apply_mpo_dense(M, v):
    """Dense check helper: mpo_to_mat(M) @ v."""
    return mpo_to_mat(M) @ np.asarray(v)


def integrate_mps(psi):
    """src/integration.jl:1-5: inner product with the all-(1/2,1/2) product state."""
    E = np.ones((1,), dtype=psi[0].dtype)
    for A in psi:
        E = E @ (0.5 * (A[:, 0, :] + A[:, 1, :]))
    return E.reshape(-1)[0]

"""Host-side operand builders of the drop-in: the tensors the reference's L4 builders produce
(O(n w^2) set-up work that stays on the host), as dense cores in the C-ABI's logical
index order.  In the Julia integration these come from stock ITensorQTT (`laplacian_mpo`,
`boundary_value_mps`, `prolongation_mpo`, `dft_mpo`); this module is their Python mirror so that the
parity tests read like the reference's own.  Bit-exact operand spec: SURVEY.md Appendix A."""
import numpy as np

__all__ = ["qtt_xrange", "laplacian_mpo", "boundary_value_mps", "prolongation_mpo", "retraction_mpo", "shift_mpo",
           "helmholtz_lambda", "index_to_bits", "bits_to_index", "project_bits", "mps_to_discrete_function",
           "random_mps", "qtt_exp", "qtt_sin", "qtt_cos", "mps_add_directsum", "dft_mpo", "idft_mpo"]


def qtt_xrange(n, xstart, xstop):
    """reference src/function_to_mps.jl:10-12."""
    L = 2 ** n
    t = np.arange(L, dtype=np.float64) / L
    return (1.0 - t) * xstart + t * xstop


def laplacian_mpo(n, xstep=1.0):
    """reference src/laplacian.jl:17-43 (six unit entries, left onehot(1), right (-2, 1, 1), / xstep^(2/L))."""
    T = np.zeros((3, 3, 2, 2))
    for idx in ((0, 0, 0, 0), (0, 0, 1, 1), (1, 1, 1, 0), (0, 1, 0, 1), (0, 2, 1, 0), (2, 2, 0, 1)):
        T[idx] = 1.0
    cores = [T.copy() for _ in range(n)]
    cores[0] = cores[0][0:1]
    cores[-1] = np.einsum("lrst,r->lst", cores[-1], np.array([-2.0, 1.0, 1.0]))[:, None]
    scale = xstep ** (2.0 / n)
    return [c / scale for c in cores]


def boundary_value_mps(n, ui, uf):
    """reference src/qtt_utils.jl:1-14."""
    cores = [np.eye(2).reshape(1, 2, 2)]
    a = np.zeros((2, 2, 2))
    a[0, 0, 0] = a[1, 1, 1] = 1.0
    cores += [a.copy() for _ in range(n - 2)]
    last = np.zeros((2, 2, 1))
    last[0, 0, 0] = ui
    last[1, 1, 0] = uf
    cores.append(last)
    return cores


def prolongation_mpo(n_fine):
    """reference src/prolongation.jl:2-23 (+ the dummy input leg of `prolongate`, :34-42, zero-padded to d = 2)."""
    P = np.zeros((2, 2, 2, 2))
    for idx in ((0, 0, 0, 0), (0, 0, 1, 1), (0, 1, 1, 0), (1, 1, 0, 1)):
        P[idx] = 1.0
    cores = [P.copy() for _ in range(n_fine - 1)]
    cores[0] = cores[0][0:1]
    last = np.zeros((2, 1, 2, 2))
    last[0, 0, 0, 0] = 1.0
    last[0, 0, 0, 1] = 0.5
    last[1, 0, 0, 1] = 0.5
    cores.append(last)
    return cores


def retraction_mpo(n_fine):
    """reference src/prolongation.jl:66-68."""
    return [np.transpose(c, (0, 1, 3, 2)).copy() for c in prolongation_mpo(n_fine)]


# ---- N-D grids: dimensions interleaved site by site (reference src/mps_utils.jl:55-66 `interleave`) -----------------------
def _identity_mpo(n):
    return [np.eye(2).reshape(1, 1, 2, 2) for _ in range(n)]


def interleave_mpo(*chains):
    """`interleave(M_1, ..., M_D)`: site j of chain d lands on site j*D + d; while the other chains' sites pass, a chain's
    open link rides along as an identity factor of the combined link (Kronecker structure)."""
    D, n = len(chains), len(chains[0])
    assert all(len(c) == n for c in chains)
    links = [1] * D                     # open link of every chain to the left of the current site
    out = []
    for j in range(n):
        for d in range(D):
            core = chains[d][j]
            wl, wr = core.shape[:2]
            assert wl == links[d]
            left = int(np.prod(links))
            before = int(np.prod(links[:d])); after = int(np.prod(links[d + 1:]))
            T = np.einsum("ab,lrst,cd->alcbrdst", np.eye(before), core, np.eye(after))
            links[d] = wr
            out.append(T.reshape(left, int(np.prod(links)), 2, 2))
    return out


def interleave_mps(*chains):
    """`interleave(psi_1, ..., psi_D)` for states (same rule as interleave_mpo)."""
    D, n = len(chains), len(chains[0])
    links = [1] * D
    out = []
    for j in range(n):
        for d in range(D):
            core = chains[d][j]
            al, _, ar = core.shape
            assert al == links[d]
            left = int(np.prod(links))
            before = int(np.prod(links[:d])); after = int(np.prod(links[d + 1:]))
            T = np.einsum("ab,lsr,cd->alcsbrd", np.eye(before), core, np.eye(after))
            links[d] = ar
            out.append(T.reshape(left, 2, int(np.prod(links))))
    return out


def mpo_directsum(A, B):
    """`+(A, B; alg="directsum")` for two MPOs of equal length: block-diagonal links, shared boundary links of dimension 1."""
    n = len(A)
    assert len(B) == n
    out = []
    for j, (a, b) in enumerate(zip(A, B)):
        al, ar = a.shape[:2]; bl, br = b.shape[:2]
        L = 1 if j == 0 else al + bl
        R = 1 if j == n - 1 else ar + br
        T = np.zeros((L, R, 2, 2), dtype=np.result_type(a, b))
        if j == 0:
            T[0:1, 0:ar] = a; T[0:1, ar:ar + br] = b
        elif j == n - 1:
            T[0:al, 0:1] = a; T[al:al + bl, 0:1] = b
        else:
            T[0:al, 0:ar] = a; T[al:, ar:] = b
        out.append(T)
    return out


def laplacian_mpo_nd(nbits, xstep=None):
    """`laplacian_mpo(s::Tuple{...}, xstep::Tuple)` (reference src/laplacian.jl:45-53), any number of equal-length
    dimensions: sum over d of interleave(I, ..., Delta_d, ..., I), added by direct sum."""
    D = len(nbits)
    assert len(set(nbits)) == 1, "interleaved dimensions need equal bit counts"
    xstep = xstep or (1.0,) * D
    total = None
    for d in range(D):
        parts = [laplacian_mpo(nbits[k], xstep[k]) if k == d else _identity_mpo(nbits[k]) for k in range(D)]
        term = interleave_mpo(*parts)
        total = term if total is None else mpo_directsum(total, term)
    return total


def prolongation_mpo_nd(nbits_fine):
    """`prolongation_mpo(s::Tuple)` (reference src/prolongation.jl:25-27): interleave of the 1-D prolongation MPOs."""
    return interleave_mpo(*[prolongation_mpo(n) for n in nbits_fine])


def mps_to_discrete_function_nd(cores, ndims):
    """`mps_to_discrete_function(Val(ndims), psi)` (reference src/function_to_mps.jl:140-145): sites j, j + ndims, ... belong to
    dimension j; the result has one axis of 2^(n / ndims) points per dimension (first site = most significant bit)."""
    n = len(cores)
    assert n % ndims == 0
    T = cores[0]
    for c in cores[1:]:
        T = np.tensordot(T, c, axes=([-1], [0]))
    T = T.reshape((2,) * n)
    perm = [j for d in range(ndims) for j in range(d, n, ndims)]
    return np.transpose(T, perm).reshape((2 ** (n // ndims),) * ndims)


def shift_mpo(A, a0, a1=1.0):
    """(a0 + a1 A) as an MPO by direct sum with the identity (reference examples/helmholtz_equation/helmholtz_equation.jl:122)."""
    n = len(A)
    out = []
    for j, c in enumerate(A):
        wl, wr = c.shape[:2]
        I = np.eye(2).reshape(1, 1, 2, 2)
        if n == 1:
            out.append(a1 * c + a0 * I)
            continue
        if j == 0:
            out.append(np.concatenate([a1 * c, a0 * I], axis=1))
        elif j == n - 1:
            out.append(np.concatenate([c, I], axis=0))
        else:
            W = np.zeros((wl + 1, wr + 1, 2, 2), dtype=c.dtype)
            W[:wl, :wr] = c
            W[wl, wr] = np.eye(2)
            out.append(W)
    return out


def helmholtz_lambda(n, nk, xi=0.0, xf=1.0):
    """reference examples/helmholtz_equation/helmholtz_equation.jl:48."""
    return -((2 * np.pi * (xf - xi) * 2.0 ** (nk - n)) ** 2)


def index_to_bits(j, n):
    return [(j >> (n - 1 - k)) & 1 for k in range(n)]


def bits_to_index(bits):
    j = 0
    for b in bits:
        j = (j << 1) | int(b)
    return j


def project_bits(cores, bits, right_bits=None):
    """reference src/project_bits.jl:1-15,25-44 on host cores (O(n chi^2), stays on the host)."""
    left = [int(b) for b in bits]
    right = [] if right_bits is None else ([0] * int(right_bits) if np.isscalar(right_bits) else [int(b) for b in right_bits])
    n = len(cores)
    pl = np.ones((1, 1), dtype=cores[0].dtype)
    for j, b in enumerate(left):
        pl = pl @ cores[j][:, b, :]
    pr = np.ones((1, 1), dtype=cores[0].dtype)
    for k in range(len(right)):
        pr = cores[n - 1 - k][:, right[len(right) - 1 - k], :] @ pr
    rest = [c.copy() for c in cores[len(left):n - len(right)]]
    if not rest:
        return (pl @ pr)[0, 0]
    rest[0] = np.tensordot(pl, rest[0], axes=([1], [0]))
    rest[-1] = np.tensordot(rest[-1], pr, axes=([2], [0]))
    return rest


def mps_to_discrete_function(cores):
    """reference src/function_to_mps.jl:136-145 (1-D): site 1 = most significant bit."""
    acc = cores[0]
    for c in cores[1:]:
        acc = np.tensordot(acc, c, axes=([acc.ndim - 1], [0]))
    return acc.reshape(-1)


def random_mps(n, chi, seed):
    """Synthetic start vector: standard-normal cores from numpy default_rng(seed), right-orthonormalised
    (QR), unit norm.  Same recipe as the oracle's, so both arms can be fed identical arrays."""
    rng = np.random.default_rng(seed)
    dims = [1] + [int(min(chi, 2 ** j, 2 ** (n - j))) for j in range(1, n)] + [1]
    cores = [rng.standard_normal((dims[j], 2, dims[j + 1])) for j in range(n)]
    for j in range(n - 1, 0, -1):
        a, d, c = cores[j].shape
        Q, R = np.linalg.qr(cores[j].reshape(a, d * c).T)
        cores[j] = Q.T.reshape(Q.shape[1], d, c)
        cores[j - 1] = np.tensordot(cores[j - 1], R.T, axes=([2], [0]))
    cores[0] = cores[0] / np.linalg.norm(cores[0])
    return cores


def qtt_exp(alpha, beta, n):
    """reference src/qtt_utils.jl:34-37."""
    dt = np.result_type(type(alpha), type(beta), np.float64)
    cores = [np.array([1.0, np.exp(alpha * 2.0 ** (-(j + 1)))], dtype=dt).reshape(1, 2, 1) for j in range(n)]
    cores[0] = cores[0] * np.exp(beta)
    return cores


def mps_add_directsum(*psis):
    """`+(psi...; alg="directsum")` (upstream): block-diagonal cores."""
    n = len(psis[0])
    dt = np.result_type(*[p[0] for p in psis])
    out = []
    for j in range(n):
        if j == 0:
            out.append(np.concatenate([p[j].astype(dt) for p in psis], axis=2))
        elif j == n - 1:
            out.append(np.concatenate([p[j].astype(dt) for p in psis], axis=0))
        else:
            Cc = np.zeros((sum(p[j].shape[0] for p in psis), 2, sum(p[j].shape[2] for p in psis)), dtype=dt)
            l0 = r0 = 0
            for p in psis:
                a, _, c = p[j].shape
                Cc[l0:l0 + a, :, r0:r0 + c] = p[j]
                l0 += a
                r0 += c
            out.append(Cc)
    return out


def _scale(cores, alpha):
    out = [c.copy() for c in cores]
    out[0] = out[0] * alpha
    return out


def qtt_sin(alpha, n, beta=0.0):
    """reference src/qtt_utils.jl:49-58."""
    return _scale(mps_add_directsum(qtt_exp(1j * alpha, beta, n), _scale(qtt_exp(-1j * alpha, beta, n), -1.0)), 1.0 / 2j)


def qtt_cos(alpha, n, beta=0.0):
    """reference src/qtt_utils.jl:67-76."""
    return _scale(mps_add_directsum(qtt_exp(1j * alpha, beta, n), qtt_exp(-1j * alpha, beta, n)), 0.5)


# ------------------------------------------------------------------ DFT operator (reference src/dft.jl)
def _truncate_rule(p, cutoff, maxdim=None):
    """NDTensors.truncate! on descending weights p: drop from the tail while the dropped weight <= cutoff * sum(p)."""
    n = len(p)
    if maxdim is not None:
        n = min(n, int(maxdim))
    err = float(np.sum(p[n:]))
    scale = float(np.sum(p)) or 1.0
    while n > 1 and err + p[n - 1] <= cutoff * scale:
        err += p[n - 1]
        n -= 1
    return n


def _mpo_compress(M, cutoff):
    """Canonical compression of a small MPO [l, r, s, t]: QR sweep to the right end, truncated-SVD sweep back."""
    M = [c.copy() for c in M]
    n = len(M)
    for j in range(n - 1):
        a, c, d1, d2 = M[j].shape
        Q, R = np.linalg.qr(np.transpose(M[j], (0, 2, 3, 1)).reshape(a * d1 * d2, c))
        k = Q.shape[1]
        M[j] = np.transpose(Q.reshape(a, d1, d2, k), (0, 3, 1, 2))
        M[j + 1] = np.tensordot(R, M[j + 1], axes=([1], [0]))
    for j in range(n - 1, 0, -1):
        a, c, d1, d2 = M[j].shape
        U, S, Vh = np.linalg.svd(np.transpose(M[j], (0, 2, 3, 1)).reshape(a, d1 * d2 * c), full_matrices=False)
        k = _truncate_rule(S ** 2, cutoff)
        M[j] = np.transpose(Vh[:k].reshape(k, d1, d2, c), (0, 3, 1, 2))
        M[j - 1] = np.transpose(np.tensordot(M[j - 1], U[:, :k] * S[None, :k], axes=([1], [0])), (0, 3, 1, 2))
    return M


def _hadamard_phase_layer(n, j):
    """reference src/dft.jl:51-74,84-93: identity on sites 1..j-1, then delta(s, s', l) with a Hadamard on the output of site
    j and the controlled phases exp(-i pi / 2^(m-1)) on the m-th site of the block (cores [l, r, s_in, s_out])."""
    cores = [np.eye(2, dtype=np.complex128).reshape(1, 1, 2, 2) for _ in range(j - 1)]
    m = n - j + 1
    H = np.array([[1.0, 1.0], [1.0, -1.0]], dtype=np.complex128) / np.sqrt(2.0)
    if m == 1:
        cores.append(H.reshape(1, 1, 2, 2).copy())          # apply(H, I)
        return cores
    first = np.zeros((1, 2, 2, 2), dtype=np.complex128)
    for b in range(2):
        first[0, b, b, :] = H[:, b]                           # delta(s, s', l) followed by H on the output leg
    cores.append(first)
    for i in range(2, m + 1):
        c = np.zeros((2, 2 if i < m else 1, 2, 2), dtype=np.complex128)
        c[0, 0] = np.eye(2)
        c[1, 1 if i < m else 0] = np.diag([1.0, np.exp(-1j * np.pi / 2 ** (i - 1))])
        cores.append(c)
    return cores


_DFT_CACHE = {}


def dft_mpo(n, cutoff=1e-15):
    """`dft_mpo(s; cutoff=1e-15)` (reference src/dft.jl:81-99, alg="mpo"): the product of the n Hadamard+phase layers,
    compressed with `cutoff` after every product, then swapprime(0 <-> 1).  Returned BEFORE `reverse_output_bits` (an
    Index relabelling in the reference, :76-79): output site q of these cores carries frequency bit n+1-q, which is what
    `apply` sees; `apply_dft_mpo` reverses the chain afterwards (:124).  Tiny (bond <= 11): built on the host once per
    (n, cutoff) and cached -- the reference rebuilds it on every `apply_dft_mpo` call (:123)."""
    key = (int(n), float(cutoff))
    if key not in _DFT_CACHE:
        U = _hadamard_phase_layer(n, 1)
        for j in range(2, n + 1):
            V = _hadamard_phase_layer(n, j)
            # operator product U o V (input -> V -> U): bonds multiply, then compress
            prod = [np.einsum("aetu,bfst->abefsu", U[i], V[i]).reshape(U[i].shape[0] * V[i].shape[0], U[i].shape[1] * V[i].shape[1], 2, 2)
                    for i in range(n)]
            U = _mpo_compress(prod, cutoff)
        _DFT_CACHE[key] = [np.ascontiguousarray(np.transpose(c, (0, 1, 3, 2))) for c in U]
    return [c.copy() for c in _DFT_CACHE[key]]


def idft_mpo(n, cutoff=1e-15):
    """`idft_mpo(s)` = swapprime(dag(dft_mpo(s)), 0 => 1) (reference src/dft.jl:44-46), raw (before bit relabelling)."""
    return [np.ascontiguousarray(np.transpose(c.conj(), (0, 1, 3, 2))) for c in dft_mpo(n, cutoff)]

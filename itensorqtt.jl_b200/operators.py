"""Host-side operand builders of the drop-in: the tensors the reference's L4 builders produce
(O(n w^2) set-up work that stays on the host), as dense cores in the C-ABI's logical
index order.  In the Julia integration these come from stock ITensorQTT (`laplacian_mpo`,
`boundary_value_mps`, `prolongation_mpo`, `dft_mpo`); this module is their Python mirror so that the
parity tests read like the reference's own.  Bit-exact operand spec: SURVEY.md Appendix A."""
import numpy as np

__all__ = ["qtt_xrange", "laplacian_mpo", "boundary_value_mps", "prolongation_mpo", "retraction_mpo", "shift_mpo",
           "helmholtz_lambda", "index_to_bits", "bits_to_index", "project_bits", "mps_to_discrete_function",
           "random_mps", "qtt_exp", "qtt_sin", "qtt_cos", "mps_add_directsum"]


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

"""Known-answer tests pinning the oracle: every assertion of the reference's own test-suite
(/root/reference/test/runtests.jl) that touches the hot path, restated on dense-core MPS/MPO.
`isapprox` (Julia `≈`) = relative tolerance sqrt(eps) in the 2-norm unless the reference states rtol."""
import numpy as np
import pytest

import oracle as O

RTOL = np.sqrt(np.finfo(np.float64).eps)


def approx(a, b, rtol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def mps_approx(x, y, rtol=RTOL):
    """ITensorMPS isapprox(x::MPS, y::MPS): norm(x - y) <= rtol * max(norm(x), norm(y)) via inner products."""
    d2 = (O.mps_inner(x, x) + O.mps_inner(y, y) - 2 * np.real(O.mps_inner(x, y))).real
    return np.sqrt(abs(d2)) <= rtol * max(O.mps_norm(x), O.mps_norm(y))


def test_interleave_runtests_25_33():
    n = 3
    psi = O.random_mps(n, 2, 1)
    phi = O.random_mps(n, 2, 2)
    rho = O.interleave_mps(psi, phi)
    assert len(rho) == 2 * n
    T = O.mps_to_tensor(rho)               # axes (p1, f1, p2, f2, p3, f3)
    want = np.einsum("abc,def->adbecf", O.mps_to_tensor(psi), O.mps_to_tensor(phi))
    assert approx(T, want)


@pytest.mark.parametrize("cutoff", [1e-15, 1e-8])
def test_function_to_mps_1d_runtests_35_80(cutoff):
    nbits = 8
    x = O.qtt_xrange(nbits, 0.0, 1.0)
    f = np.sin(np.pi * x)
    psi = O.vec_to_mps(f, cutoff=cutoff)
    assert len(psi) == nbits and O.maxlinkdim(psi) == 2
    assert approx(O.mps_to_discrete_function(psi), f)


@pytest.mark.parametrize("length", [None, 50])
def test_function_to_mps_polynomial_runtests_59_71(length):
    nbits = 8
    x = O.qtt_xrange(nbits, 0.0, 1.0)
    psi = O.function_to_mps_polynomial(lambda t: np.sin(np.pi * t), nbits, 0.0, 1.0, degree=8, length=length, cutoff=1e-8)
    assert len(psi) == nbits and O.maxlinkdim(psi) == 2
    assert approx(O.mps_to_discrete_function(psi), np.sin(np.pi * x), rtol=1e-6)


def test_function_to_mps_2d_runtests_82_142():
    nb = 4
    x = O.qtt_xrange(nb, 0.0, 1.0)
    f1 = np.sin(2 * np.pi * x); f2 = np.sin(np.pi * x)
    psi = O.interleave_mps(O.vec_to_mps(f1), O.vec_to_mps(f2))
    assert len(psi) == 2 * nb and O.maxlinkdim(psi) == 4
    F = O.mps_to_discrete_function(psi, ndims=2)
    assert approx(F, np.outer(f1, f2))


def test_laplacian_mpo_runtests_144_158():
    nbits = 3
    assert approx(O.mpo_to_mat(O.laplacian_mpo(nbits)), O.laplacian_matrix(2 ** nbits))
    h = 1 / 2 ** nbits
    assert approx(O.mpo_to_mat(O.laplacian_mpo(nbits, h)), O.laplacian_matrix(2 ** nbits, h))


def test_laplacian_exact_entries():
    """Bit-exact operand spec: six ones, boundary vectors, xstep scaling."""
    T = O.laplacian_tensor()
    nz = sorted(map(tuple, np.argwhere(T != 0)))
    assert nz == sorted([(0, 0, 0, 0), (0, 0, 1, 1), (1, 1, 1, 0), (0, 1, 0, 1), (0, 2, 1, 0), (2, 2, 0, 1)])
    M = O.laplacian_mpo(6, 1.0)
    assert M[0].shape == (1, 3, 2, 2) and M[-1].shape == (3, 1, 2, 2)
    assert np.array_equal(M[0][0], T[0]) and np.array_equal(M[2], T)
    assert np.array_equal(M[-1][:, 0], np.tensordot(T, np.array([-2.0, 1.0, 1.0]), axes=([1], [0])))
    M2 = O.laplacian_mpo(4, 0.5)
    assert np.array_equal(M2[1], T / 0.5 ** (2.0 / 4))


def test_laplacian_2d_matches_kron():
    n = 3
    M = O.laplacian_mpo_2d(n, n, (1.0, 0.5))
    D = O.mpo_to_mat(M)
    # interleaved bits -> (x1, x2) grid
    N = 2 ** n
    perm = np.array([O.bits_to_index(O.interleave_bits(O.index_to_bits(i, n), O.index_to_bits(j, n)))
                     for i in range(N) for j in range(N)])
    Dk = np.kron(O.laplacian_matrix(N, 1.0), np.eye(N)) + np.kron(np.eye(N), O.laplacian_matrix(N, 0.5))
    assert approx(D[np.ix_(perm, perm)], Dk)


def test_integration_runtests_160_171():
    from scipy.integrate import quad
    f = lambda x: x ** 2 * np.sin(np.pi * x)
    exact = quad(f, 0.0, 1.0, epsabs=1e-14)[0]
    psi = O.vec_to_mps(f(O.qtt_xrange(12, 0.0, 1.0)))
    assert abs(O.integrate_mps(psi) - exact) <= 1e-7 * abs(exact) * 3e3  # rectangle rule on 2^12 points: O(h)
    assert abs(O.integrate_mps(psi) - exact) <= 1e-3


def test_prolongation_retraction_runtests_184_215():
    f = lambda x: np.sin(np.pi * x)
    nbits = 12
    p0, p1, p2 = (O.vec_to_mps(f(O.qtt_xrange(nb, 0.0, 1.0))) for nb in (nbits - 2, nbits - 1, nbits))
    assert mps_approx(O.prolongate(p1, 1), p2, rtol=1e-6)
    assert mps_approx(O.prolongate(p0, 2), p2, rtol=1e-6)
    assert mps_approx(O.retract(p2, 1), p1, rtol=1e-5)
    assert mps_approx(O.retract(p2, 2), p0, rtol=1e-5)


def test_prolongation_is_linear_interpolation():
    """dense 4 -> 8 prolongation: fine[2i] = c[i], fine[2i+1] = (c[i]+c[i+1])/2, zero extension."""
    nc = 2
    P = O.prolongation_mpo(nc + 1)
    D = O.mpo_to_mat(P)                      # rows: fine (3 bits); cols: coarse bits + dummy
    cols = [2 * i for i in range(2 ** nc)]   # dummy input bit = 0
    Pm = D[:, cols]
    want = np.zeros((8, 4))
    for i in range(4):
        want[2 * i, i] = 1.0
        want[2 * i + 1, i] = 0.5
        if i + 1 < 4:
            want[2 * i + 1, i + 1] = 0.5
    assert np.array_equal(Pm, want)


def test_dft_runtests_249_274():
    n = 5
    F = O.dft_mpo(n)
    assert max(O.mpo_linkdims(F)) < 10
    assert approx(O.mpo_to_mat(F, reverse_output_sites=True), O.dft_matrix(n))
    psi = O.mps_add_directsum(*[O.qtt_sin(a * np.pi, n) for a in (2, 4, 8, 16)])
    f = O.mps_to_discrete_function(psi)
    Fpsi = O.mps_reverse(O.apply_mpo_mps(F, psi, cutoff=1e-15))
    assert approx(np.fft.fft(f) / 2 ** (n / 2), O.mps_to_discrete_function(Fpsi))
    assert mps_approx(O.apply_idft_mpo(Fpsi), [c.astype(complex) for c in psi])
    Fpsi = O.apply_dft_mpo(psi, cutoff=1e-15)
    assert approx(np.fft.fft(f) / 2 ** (n / 2), O.mps_to_discrete_function(Fpsi))
    assert mps_approx(O.apply_idft_mpo(Fpsi), [c.astype(complex) for c in psi])


def test_qtt_analytic():
    n = 7
    x = O.qtt_xrange(n, 0.0, 1.0)
    assert approx(O.mps_to_discrete_function(O.qtt_exp(-1.3, 0.2, n)), np.exp(-1.3 * x + 0.2))
    assert approx(O.mps_to_discrete_function(O.qtt_sin(5.0, n)), np.sin(5.0 * x))
    assert approx(O.mps_to_discrete_function(O.qtt_cos(5.0, n)), np.cos(5.0 * x))


def test_fourier_interpolation_runtests_276_296():
    n = 5
    g = lambda x: np.exp(-(x - 0.5) ** 2 / 0.1 ** 2)
    psi = O.vec_to_mps(g(O.qtt_xrange(n, 0.0, 1.0)), cutoff=1e-15)
    Fpsi = O.apply_dft_mpo(psi)
    assert approx(np.fft.fft(O.mps_to_discrete_function(psi)) / 2 ** (n / 2), O.mps_to_discrete_function(Fpsi))
    assert mps_approx(O.apply_idft_mpo(Fpsi), [c.astype(complex) for c in psi])
    k = 2
    q = O.fourier_interpolation(psi, k)
    assert len(q) == n + k
    want = O.vec_to_mps(g(O.qtt_xrange(n + k, 0.0, 1.0)), cutoff=1e-15)
    assert mps_approx(q, [c.astype(complex) for c in want])


def test_repeat_function_runtests_298_311():
    n = 3
    g = lambda x: np.exp(-(x - 0.5) ** 2 / 0.1 ** 2)
    psi = O.vec_to_mps(g(O.qtt_xrange(n, 0.0, 1.0)), cutoff=1e-15)
    v = O.mps_to_discrete_function(psi)
    for k in (1, 2, 3):
        assert approx(O.mps_to_discrete_function(O.repeat_function(psi, k)), np.tile(v, 2 ** k))


def test_bit_maps_exact():
    """Bit-exact contract: site 1 = MSB; vec_to_tensor == C-order reshape;
    project_bits picks exactly the grid sample."""
    n = 6
    rng = np.random.default_rng(0)
    v = rng.standard_normal(2 ** n)
    T = O.vec_to_tensor(v)
    for j in (0, 1, 13, 42, 63):
        bits = O.index_to_bits(j, n)
        assert O.bits_to_index(bits) == j
        assert T[tuple(bits)] == v[j]
    psi = O.vec_to_mps(v, cutoff=0.0)
    for j in (0, 5, 33, 63):
        assert abs(O.project_bits(psi, O.index_to_bits(j, n)) - v[j]) < 1e-13
    # leading + trailing projection
    sub = O.project_bits(psi, [1, 0], [1])
    want = T[1, 0, :, :, :, 1].reshape(-1)
    assert approx(O.mps_to_discrete_function(sub), want)
    assert O.float_to_bits(0.625, 3) == [1, 0, 1]
    assert O.interleave_bits([1, 0], [0, 1]) == [1, 0, 0, 1]
    x = O.qtt_xrange(4, 1.0, 3.0)
    assert x[0] == 1.0 and x[8] == 2.0 and np.array_equal(x, 1.0 + np.arange(16) * 0.125)


def test_truncation_rule_app_b3():
    P = np.array([1.0, 1e-2, 1e-14, 1e-17, -1e-18])
    assert O.truncate_spectrum(P, cutoff=1e-15)[0] == 3
    assert O.truncate_spectrum(P, cutoff=0.0)[0] == 4          # negative tail -> 0 is discarded at cutoff 0
    assert O.truncate_spectrum(P, cutoff=1e-15, maxdim=2)[0] == 2
    k, err = O.truncate_spectrum(P, cutoff=1e-15, maxdim=2)
    assert abs(err - (1e-14 + 1e-17) / P[:4].sum()) < 1e-20
    assert O.truncate_spectrum(P, cutoff=1.0, mindim=2)[0] == 2
    assert O.truncate_spectrum(np.array([0.0, 0.0]), cutoff=0.0)[0] == 1


def test_airy_system_matches_dense():
    """examples/airy_equation/src/airy_utils.jl:10-27 vs :41-48 (equal to 1e-16, bond 5)."""
    for n, xf in ((8, 2.0), (10, 8.0)):
        A = O.airy_mpo(n, 1.0, xf)
        assert O.mpo_linkdims(A) == [5] * (n - 1)
        assert np.abs(O.mpo_to_mat(A) - O.airy_matrix(n, 1.0, xf)).max() < 1e-14
    b = O.boundary_value_mps(6, 0.3, -0.7)
    assert np.array_equal(O.mps_to_discrete_function(b), O.boundary_value_vector(6, 0.3, -0.7))


def test_residual_metric():
    n = 6
    A = O.airy_mpo(n, 1.0, 2.0); Am = O.mpo_to_mat(A)
    x = O.random_mps(n, 4, 3); b = O.random_mps(n, 3, 4)
    xv = O.mps_to_discrete_function(x); bv = O.mps_to_discrete_function(b)
    assert abs(O.sqeuclidean_mps(A, x, b) - np.sum((Am @ xv - bv) ** 2)) < 1e-12
    assert abs(O.sqeuclidean_normalized_mps(A, x, b) - np.sum((Am @ xv - bv) ** 2) / (bv @ bv)) < 1e-12


def test_mirror_dft_mpo_builder_is_the_dft_matrix():
    """Host operand builder of the drop-in (`itensorqtt.jl_b200/operators.py:dft_mpo`, reference src/dft.jl:81-99): the MPO is
    the DFT matrix after the output-bit relabelling, to the accuracy its cutoff allows, with the reference's bond profile."""
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "itensorqtt.jl_b200", "operators.py")
    spec = importlib.util.spec_from_file_location("qtt_operators_standalone", path)
    ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ops)
    for n in (1, 2, 4, 9):
        F = ops.dft_mpo(n)
        m = O.mpo_to_mat(F, reverse_output_sites=True)
        assert np.abs(m - O.dft_matrix(n)).max() * 2 ** (n / 2) < 1e-6
        Fi = ops.idft_mpo(n)
        mi = O.mpo_to_mat(Fi, reverse_input_sites=True)
        assert np.abs(mi @ m - np.eye(2 ** n)).max() < 1e-6
    assert max(O.mpo_linkdims(ops.dft_mpo(20))) <= 11          # SURVEY: DFT MPO bond ~8-11

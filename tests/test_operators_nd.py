"""N-D (interleaved) operand specs of the product's host mirror (itensorqtt.jl_b200/operators.py), CPU only: the builders are
plain numpy and load without the GPU library.  Checked against the oracle's restatement of the same reference lines
(src/laplacian.jl:45-53, src/prolongation.jl:25-27, src/mps_utils.jl:55-66, src/function_to_mps.jl:140-145) and against
Kronecker-product ground truth; the integer-valued cores must agree bit for bit."""
import importlib.util
import os

import numpy as np

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("qtt_operators", os.path.join(ROOT, "itensorqtt.jl_b200", "operators.py"))
P = importlib.util.module_from_spec(spec)
spec.loader.exec_module(P)


def test_laplacian_2d_cores_match_the_oracle_bit_for_bit():
    for n, xs in ((3, (1.0, 1.0)), (4, (1.0, 0.5))):
        a = P.laplacian_mpo_nd((n, n), xs)
        b = O.laplacian_mpo_2d(n, n, xs)
        assert len(a) == len(b) == 2 * n
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)


def test_laplacian_2d_is_the_kronecker_sum():
    n = 3
    M = O.mpo_to_mat(P.laplacian_mpo_nd((n, n), (1.0, 0.5)))
    N = 2 ** n
    L1, L2 = O.laplacian_matrix(N, 1.0), O.laplacian_matrix(N, 0.5)
    perm = np.array([O.bits_to_index(O.interleave_bits(O.index_to_bits(i, n), O.index_to_bits(j, n))) for i in range(N) for j in range(N)])
    K = np.kron(L1, np.eye(N)) + np.kron(np.eye(N), L2)
    assert np.allclose(M[np.ix_(perm, perm)], K, atol=1e-12)


def test_laplacian_3d_is_the_kronecker_sum():
    n = 2
    M = O.mpo_to_mat(P.laplacian_mpo_nd((n, n, n)))
    N = 2 ** n
    L = O.laplacian_matrix(N, 1.0)
    I = np.eye(N)
    K = np.kron(np.kron(L, I), I) + np.kron(np.kron(I, L), I) + np.kron(np.kron(I, I), L)

    def inter(i, j, k):
        bi, bj, bk = O.index_to_bits(i, n), O.index_to_bits(j, n), O.index_to_bits(k, n)
        bits = [b for t in zip(bi, bj, bk) for b in t]
        return O.bits_to_index(bits)
    perm = np.array([inter(i, j, k) for i in range(N) for j in range(N) for k in range(N)])
    assert np.allclose(M[np.ix_(perm, perm)], K, atol=1e-12)


def test_interleave_mps_and_nd_expansion():
    rng = np.random.default_rng(0)
    n = 4
    f1, f2 = rng.standard_normal(2 ** n), rng.standard_normal(2 ** n)
    p1, p2 = O.vec_to_mps(f1), O.vec_to_mps(f2)
    a = P.interleave_mps(p1, p2)
    b = O.interleave_mps(p1, p2)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)
    F = P.mps_to_discrete_function_nd(a, 2)
    assert F.shape == (2 ** n, 2 ** n) and np.allclose(F, np.outer(f1, f2), atol=1e-12)
    assert np.allclose(F, O.mps_to_discrete_function(b, ndims=2))


def test_prolongation_2d_is_the_interleave_of_the_1d_operators():
    n = 3
    a = P.prolongation_mpo_nd((n, n))
    b = O.interleave_mpo(O.prolongation_mpo(n), O.prolongation_mpo(n))
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)

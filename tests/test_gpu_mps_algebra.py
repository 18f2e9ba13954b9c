"""GPU parity of the device-resident MPS algebra around the apply path, through the C ABI vs the oracle:
`+(x, y; cutoff)` (reference src/fourier_interpolation.jl:8), `mps_to_discrete_function` (src/function_to_mps.jl:140-145)
and batched `project_bits` point values (src/project_bits.jl:1-15).

Tolerances: dense expansion and point values are plain chains of products -> 1e-12 relative to max|f|; the truncated
sum is compared as a function on the grid, 1e-9 relative to max|f| (north-star tolerance for sampled values)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def dense(psi):
    return O.mps_to_discrete_function(psi)


def cplx_mps(n, chi, seed):
    rng = np.random.default_rng(seed)
    return [c + 1j * rng.standard_normal(c.shape) for c in O.random_mps(n, chi, seed)]


@pytest.mark.parametrize("n,chi", [(1, 1), (2, 2), (5, 3), (10, 7), (14, 24)])
def test_mps_to_vec_real(lib, n, chi):
    psi = O.random_mps(n, chi, 10 + n)
    want = dense(psi)
    got = lib.mps_to_vec(psi)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


@pytest.mark.parametrize("n,chi", [(1, 1), (3, 2), (9, 6), (13, 17)])
def test_mps_to_vec_complex(lib, n, chi):
    psi = cplx_mps(n, chi, 20 + n)
    want = dense(psi)
    got = lib.mps_to_vec(psi)
    assert got.dtype == np.complex128
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


def test_mps_to_vec_inverts_vec_to_mps(lib):
    """vec_to_mps -> mps_to_vec round trip on the device (reference runtests.jl:43-79 round trips), n = 18."""
    n = 18
    x = O.qtt_xrange(n, 0.0, 1.0)
    f = np.sin(7.0 * x) * np.exp(-x) + 0.25 * x ** 2
    psi = lib.vec_to_mps(f, cutoff=1e-15, on_device=True)
    got = lib.mps_to_vec(psi)
    cores = psi.download()
    psi.free()
    # exact expansion of the chain that is on the device ...
    assert np.max(np.abs(got - dense(cores))) <= 1e-12 * np.max(np.abs(f))
    # ... which reproduces the samples to the TT-SVD truncation floor: cutoff 1e-15 on sigma^2 -> sqrt(1e-15) ||f||_2 in norm
    assert np.linalg.norm(got - f) <= 1e-7 * np.linalg.norm(f)


@pytest.mark.parametrize("n,chi", [(1, 1), (4, 2), (12, 9), (20, 33)])
def test_sample_matches_project_bits(lib, n, chi):
    psi = O.random_mps(n, chi, 30 + n)
    rng = np.random.default_rng(n)
    idx = np.unique(np.concatenate([[0, (1 << n) - 1], rng.integers(0, 1 << n, size=200)]))
    got = lib.sample(psi, idx)
    want = np.array([O.project_bits(psi, O.index_to_bits(int(j), n)) for j in idx])
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) <= 1e-12 * scale


def test_sample_complex_and_bit_order(lib):
    """Bit-exact index map: a product state whose value at grid index j is exactly j (site 1 = most significant bit)."""
    n = 16
    # sum_j 2^(n-1-j) b_j as a bond-2 chain: rows (1, acc)
    cores = []
    for j in range(n):
        w = float(1 << (n - 1 - j))
        c = np.zeros((2, 2, 2))
        for b in range(2):
            c[:, b, :] = np.array([[1.0, b * w], [0.0, 1.0]])
        cores.append(c)
    cores[0] = cores[0][:1]
    cores[-1] = cores[-1][:, :, 1:]
    idx = np.array([0, 1, 2, 3, 255, 256, 32768, 65535, 43690, 21845], dtype=np.int64)
    got = lib.sample(cores, idx)
    assert np.array_equal(got, idx.astype(np.float64))
    psi = cplx_mps(9, 5, 77)
    want = dense(psi)
    idc = np.arange(1 << 9)[::7]
    assert np.max(np.abs(lib.sample(psi, idc) - want[idc])) <= 1e-12 * np.max(np.abs(want))
    assert lib.sample(psi, np.zeros((0,), dtype=np.int64)).size == 0


def test_sample_rejects_out_of_range(lib):
    psi = O.random_mps(6, 3, 1)
    with pytest.raises(lib.QttError):
        lib.sample(psi, [64])
    with pytest.raises(lib.QttError):
        lib.sample(psi, [-1])


@pytest.mark.parametrize("n,cx,cy", [(1, 1, 1), (2, 2, 2), (8, 4, 6), (12, 10, 3)])
def test_add_matches_oracle(lib, n, cx, cy):
    x = O.random_mps(n, cx, 40 + n)
    y = O.random_mps(n, cy, 50 + n)
    want = dense(x) - 0.5 * dense(y)
    got = lib.add(x, y, alpha=1.0, beta=-0.5, cutoff=1e-15)
    assert np.max(np.abs(dense(got) - want)) <= 1e-9 * np.max(np.abs(want))
    ref = O.mps_add(x, O.mps_scale(y, -0.5), cutoff=1e-15)
    assert np.max(np.abs(dense(got) - dense(ref))) <= 1e-9 * np.max(np.abs(want))
    if n > 1:
        # exact ranks are recovered: the bond dimension never exceeds the direct sum, nor the oracle's by more than noise
        assert all(g.shape[2] <= a.shape[2] + b.shape[2] for g, a, b in zip(got[:-1], x[:-1], y[:-1]))


def test_add_truncates_to_the_rank_of_the_sum(lib):
    """sin + sin of the same frequency has rank 2, not 4, and `maxdim` caps the bond (reference runtests.jl:255-257 builds
    its inputs with the same `+`)."""
    n = 14
    a = O.qtt_sin(2 * np.pi * 3, n)
    b = O.qtt_sin(2 * np.pi * 3, n, 0.4)
    got = lib.add(a, b, cutoff=1e-14)
    assert max(c.shape[2] for c in got[:-1]) == 2
    want = dense(a) + dense(b)
    assert np.max(np.abs(dense(got) - want)) <= 1e-9 * np.max(np.abs(want))
    c = O.qtt_cos(2 * np.pi * 5, n)
    got = lib.add(a, c, cutoff=1e-14, maxdim=3)
    assert max(k.shape[2] for k in got[:-1]) <= 3


def test_add_complex_and_mixed(lib):
    n = 9
    x = cplx_mps(n, 4, 5)
    y = O.random_mps(n, 3, 6)
    got = lib.add(x, y, alpha=2.0, beta=1.0, cutoff=1e-15)
    want = 2.0 * dense(x) + dense(y)
    assert np.iscomplexobj(got[0])
    assert np.max(np.abs(dense(got) - want)) <= 1e-9 * np.max(np.abs(want))


def test_add_is_the_sum_step_of_fourier_interpolation(lib):
    """The `+(F0s, F1s; cutoff=1e-15)` of reference src/fourier_interpolation.jl:4-8 on the device equals the host-written sum."""
    n, k = 8, 2
    psi = O.vec_to_mps(np.cos(2 * np.pi * 3 * O.qtt_xrange(n, 0.0, 1.0)), cutoff=1e-15)
    F = lib.apply_dft_mpo(psi, cutoff=1e-15)

    def basis(bit):
        c = np.zeros((1, 2, 1), dtype=np.complex128)
        c[0, bit, 0] = 1.0
        return c

    parts = []
    for b in range(2):
        rest = O.project_bits(F, [b])
        parts.append([basis(b) for _ in range(k + 1)] + list(rest))
    got = lib.add(parts[0], parts[1], cutoff=1e-15)
    want = dense(parts[0]) + dense(parts[1])
    assert np.max(np.abs(dense(got) - want)) <= 1e-9 * np.max(np.abs(want))

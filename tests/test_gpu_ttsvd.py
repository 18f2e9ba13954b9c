"""GPU TT-SVD (`vec_to_mps` / `function_to_mps(...; alg="factorize")`, reference src/function_to_mps.jl:41-52,170-177) through
the C ABI vs the oracle's restatement (numpy SVD + the ITensors truncation rule) on the same samples."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _dense(cores):
    return O.mps_to_discrete_function(cores)


def _left_orthogonal(cores, tol=1e-12):
    for c in cores[:-1]:
        m = c.reshape(-1, c.shape[2])
        assert np.abs(m.T @ m - np.eye(m.shape[1])).max() < tol


def test_known_ranks_sin_exp(lib):
    """runtests.jl-style known answers: sin has QTT rank 2, exp rank 1 (src/qtt_utils.jl:37-58)."""
    n = 14
    x = O.qtt_xrange(n, 0.0, 1.0)
    for f, rank in ((lambda t: np.sin(2 * np.pi * 5 * t + 0.3), 2), (lambda t: np.exp(-3.0 * t), 1)):
        v = f(x)
        g = lib.vec_to_mps(v)
        o = O.vec_to_mps(v)
        assert O.linkdims(g) == O.linkdims(o)
        assert max(O.linkdims(g)) == rank
        assert np.abs(_dense(g) - v).max() <= 1e-12 * np.abs(v).max()
        _left_orthogonal(g)


def test_random_vector_is_reproduced_exactly(lib):
    """Full-rank input, cutoff below the noise: bonds 2,4,...,2^(n/2),...,4,2 and an exact reconstruction; bonds up to 64
    exercise the DMMA GEMM branch (q = 2 r >= 32) as well as the streaming small-q kernels."""
    n = 12
    v = np.random.default_rng(0).standard_normal(2 ** n)
    g = lib.vec_to_mps(v, cutoff=0.0)
    assert O.linkdims(g) == [min(2 ** j, 2 ** (n - j)) for j in range(1, n)]
    assert np.abs(_dense(g) - v).max() <= 1e-12
    _left_orthogonal(g)


@pytest.mark.parametrize("cutoff", [1e-15, 1e-10, 1e-6])
def test_smooth_function_truncation_matches_oracle(lib, cutoff):
    n = 16
    f = lambda t: np.exp(-((t - 0.4) ** 2) / 0.02) * np.cos(37.0 * t) + 0.2 * t ** 3
    v = f(O.qtt_xrange(n, 0.0, 1.0))
    g, info = lib.function_to_mps(f, n, 0.0, 1.0, cutoff=cutoff, return_info=True)
    o = O.vec_to_mps(v, cutoff=cutoff)
    lg, lo = O.linkdims(g), O.linkdims(o)
    # the rule is applied to the true discarded weight of an orthogonal projection: same ranks as the SVD up to one
    # marginal singular value per bond
    assert all(abs(a - b) <= 1 for a, b in zip(lg, lo)), (lg, lo)
    eg = np.linalg.norm(_dense(g) - v) / np.linalg.norm(v)
    eo = np.linalg.norm(_dense(o) - v) / np.linalg.norm(v)
    assert eg <= max(2.0 * eo, np.sqrt(n * cutoff)) + 1e-14
    assert info["maxtruncerr"] <= cutoff * 1.0000001 or cutoff == 0
    assert O.fidelity_defect(g, o) <= 1e-10 + 4 * n * cutoff


def test_maxdim_and_mindim(lib):
    n = 13
    rng = np.random.default_rng(3)
    x = O.qtt_xrange(n, 0.0, 1.0)
    v = sum(rng.standard_normal() * np.sin(2 * np.pi * rng.integers(1, 2 ** (n - 2)) * x + rng.uniform(0, 6)) for _ in range(12))
    g = lib.vec_to_mps(v, cutoff=1e-14, maxdim=8)
    o = O.vec_to_mps(v, cutoff=1e-14, maxdim=8)
    assert max(O.linkdims(g)) <= 8 and O.linkdims(g) == O.linkdims(o)
    eg = np.linalg.norm(_dense(g) - v); eo = np.linalg.norm(_dense(o) - v)
    assert eg <= 1.05 * eo + 1e-12
    full = lib.vec_to_mps(v, cutoff=1e-14)
    ofull = O.vec_to_mps(v, cutoff=1e-14)
    assert max(O.linkdims(full)) == 24 and O.linkdims(full) == O.linkdims(ofull)   # 12 sines -> rank 24
    assert np.linalg.norm(_dense(full) - v) <= 1.05 * np.linalg.norm(_dense(ofull) - v) + 1e-12
    # far below the resolution of a Gram matrix (sigma / sigma_1 ~ 1e-10): the block refinement still finds the rank to +-1
    tiny = lib.vec_to_mps(v, cutoff=1e-20)
    assert all(abs(a - b) <= 1 for a, b in zip(O.linkdims(tiny), O.linkdims(O.vec_to_mps(v, cutoff=1e-20))))
    assert np.abs(_dense(tiny) - v).max() <= 1e-9 * np.abs(v).max()
    m = lib.vec_to_mps(np.ones(2 ** 6), cutoff=1e-15, mindim=2)
    assert O.linkdims(m) == [2] * 5                          # rank-1 input padded by mindim, still exact
    assert np.abs(_dense(m) - 1.0).max() <= 1e-13


def test_device_resident_samples_and_edges(lib):
    import torch
    n = 18
    t = torch.linspace(0, 1, 2 ** n + 1, dtype=torch.float64, device="cuda")[:-1]
    v = torch.sin(40.0 * t) * torch.exp(-t)
    keep = v.clone()
    g = lib.vec_to_mps(v, cutoff=1e-15)
    assert torch.equal(v, keep)                              # input read in place, not modified
    assert max(O.linkdims(g)) == 2
    assert np.abs(_dense(g) - v.cpu().numpy()).max() <= 1e-12
    two = lib.vec_to_mps(np.array([3.0, -4.0]))
    assert len(two) == 1 and two[0].shape == (1, 2, 1) and np.array_equal(two[0].ravel(), [3.0, -4.0])
    z = lib.vec_to_mps(np.zeros(2 ** 5))
    assert np.abs(_dense(z)).max() == 0.0
    with pytest.raises(ValueError):
        lib.vec_to_mps(np.ones(12))
    # the residency flag is checked against the pointer instead of faulting inside a kernel
    import ctypes as C
    out, err = C.c_void_p(), C.c_double()
    host = np.ones(8)
    ctx = lib.default_context()
    rc = lib.lib.qtt_vec_to_mps(ctx.handle, C.c_void_p(host.ctypes.data), 3, 1, 1e-15, 0, 1, C.byref(out), C.byref(err))
    assert rc != 0 and b"not a device pointer" in lib.lib.qtt_last_error(ctx.handle)

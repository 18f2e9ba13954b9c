"""GPU parity of MPO.MPS `apply` (density-matrix algorithm), `inner` and the residual metric through the C ABI vs the
oracle, on the reference's own known-answer cases (test/runtests.jl:184-215 prolongation, :249-274 QFT/DFT/FFT).

Tolerances: sampled function values within 1e-9 of the oracle's relative to max|f| (north-star); DFT vs numpy FFT at
the input-truncation floor (~3.5e-9 norm-relative at cutoff 1e-15)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def dense(psi):
    return O.mps_to_discrete_function(psi)


def sines(n, freqs=(1, 2, 4, 8)):
    """psi = sum_f sin(2 pi f x) built like runtests.jl:255-257: qtt(sin, ...) + directsum."""
    return O.mps_add_directsum(*[O.qtt_sin(2 * np.pi * f, n) for f in freqs])


@pytest.mark.parametrize("n", [4, 7])
def test_inner_real_and_complex(lib, n):
    x = O.random_mps(n, 5, 1)
    y = O.random_mps(n, 3, 2)
    assert abs(lib.inner(x, y) - O.mps_inner(x, y)) < 1e-13
    rng = np.random.default_rng(0)
    xc = [c + 1j * rng.standard_normal(c.shape) for c in x]
    yc = [c * (0.3 - 0.8j) for c in y]
    got, want = lib.inner(xc, yc), O.mps_inner(xc, yc)
    assert abs(got - want) < 1e-13 * max(1.0, abs(want))
    got, want = lib.inner(xc, y), O.mps_inner(xc, y)
    assert abs(got - want) < 1e-13 * max(1.0, abs(want))


def test_residual_metric(lib):
    n = 8
    A = O.airy_mpo(n, 1.0, 4.0)
    b = O.boundary_value_mps(n, 0.3, -0.2)
    x = O.random_mps(n, 6, 3)
    sq = lib.sqeuclidean((A, x), b)
    sqn = lib.sqeuclidean_normalized((A, x), b)
    assert abs(sq - O.sqeuclidean_mps(A, x, b)) < 1e-11 * abs(sq)
    assert abs(sqn - O.sqeuclidean_normalized_mps(A, x, b)) < 1e-11 * abs(sqn)
    v = dense(x)
    r = O.mpo_to_mat(A) @ v - dense(b)
    assert abs(sq - r @ r) < 1e-10 * (r @ r)


@pytest.mark.parametrize("n", [3, 6, 10, 11])
def test_prolongation_apply_matches_oracle_and_interpolation(lib, n):
    """apply(P, [psi; dummy]; cutoff=1e-8) (reference src/prolongation.jl:29-45); known answer of runtests.jl:184-203:
    prolongate(function_to_mps(sin(pi x), n bits)) ~ function_to_mps(sin(pi x), n+1 bits), MPS `isapprox` (norm) rtol 1e-6."""
    for shift in (0.0, 0.25):
        psi = O.vec_to_mps(np.sin(np.pi * O.qtt_xrange(n, 0.0, 1.0)) + shift, cutoff=1e-15)
        P = O.prolongation_mpo(n + 1)
        dummy = np.zeros((1, 2, 1)); dummy[0, 0, 0] = 1.0
        ext = list(psi) + [dummy]
        got = lib.apply(P, ext, cutoff=1e-8)
        want = O.apply_mpo_mps(P, ext, cutoff=1e-8)
        assert O.linkdims(got) == O.linkdims(want)
        fg, fw = dense(got), dense(want)
        assert np.abs(fg - fw).max() <= 1e-9 * np.abs(fw).max()
        c = dense(psi)
        fine = np.empty(2 * c.size)
        fine[0::2] = c
        fine[1::2] = 0.5 * (c + np.append(c[1:], 0.0))
        # the cutoff acts on sigma^2, so the truncated interpolant is sqrt(1e-8)-accurate in norm
        assert np.linalg.norm(fg - fine) < 2e-4 * np.linalg.norm(fine)
    if n >= 10:  # the reference's own assertion: rank-2 sin(pi x), nothing near the cutoff, interpolation error O(h^2)
        exact = np.sin(np.pi * O.qtt_xrange(n + 1, 0.0, 1.0))
        psi = O.vec_to_mps(np.sin(np.pi * O.qtt_xrange(n, 0.0, 1.0)), cutoff=1e-15)
        got = lib.apply(O.prolongation_mpo(n + 1), list(psi) + [dummy], cutoff=1e-8)
        assert np.linalg.norm(dense(got) - exact) < 1e-6 * np.linalg.norm(exact) * (4.0 ** (12 - n))


def test_apply_untruncated_is_exact(lib):
    """cutoff = 0: the result equals the dense operator applied to the dense vector to rounding."""
    n = 6
    A = O.airy_mpo(n, 1.0, 3.0)
    psi = O.random_mps(n, 4, 5)
    got = lib.apply(A, psi, cutoff=0.0)
    want = O.mpo_to_mat(A) @ dense(psi)
    assert np.abs(dense(got) - want).max() < 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("n", [5, 8])
def test_dft_apply_matches_fft_and_oracle(lib, n):
    """runtests.jl:249-274: reverse(apply(F, psi; cutoff=1e-15)) ~ fft(f)/2^(n/2), and apply_idft_mpo inverts it."""
    F = O.dft_mpo(n)
    psi = sines(n)
    raw = lib.apply(F, psi, cutoff=1e-15)
    got = O.mps_reverse(raw)
    f = dense(psi)
    want_fft = np.fft.fft(f) / 2 ** (n / 2)
    fg = dense(got)
    assert np.linalg.norm(fg - want_fft) < 5e-8 * np.linalg.norm(want_fft)
    fo = dense(O.apply_dft_mpo(psi, cutoff=1e-15, F=F))
    assert np.abs(fg - fo).max() <= 1e-9 * np.abs(fo).max()
    # inverse: apply(idft_mpo, reverse(F psi)) ~ psi   (reference src/dft.jl:128-133)
    Finv = O.idft_mpo(n)
    back = lib.apply(Finv, O.mps_reverse(got), cutoff=1e-15)
    assert np.linalg.norm(dense(back) - f) < 5e-8 * np.linalg.norm(f)


def test_dft_apply_real_input_complex_operator_bond_dims(lib):
    """Mixed dtypes (Float64 psi, ComplexF64 MPO) and the kept ranks of the density-matrix truncation."""
    n = 8
    F = O.dft_mpo(n)
    xg = O.qtt_xrange(n, 0.0, 1.0)
    psi = O.vec_to_mps(np.exp(-40.0 * (xg - 0.4) ** 2) + 0.1 * np.cos(6 * np.pi * xg), cutoff=1e-14)
    assert all(c.dtype == np.float64 for c in psi)
    got = lib.apply(F, psi, cutoff=1e-13)
    want = O.apply_mpo_mps(F, [c.astype(complex) for c in psi], cutoff=1e-13)
    assert max(O.linkdims(got)) <= max(O.linkdims(want)) + 1
    fw = dense(want)
    assert np.abs(dense(got) - fw).max() <= 1e-9 * np.abs(fw).max()


def test_prolongate_and_retract_known_answers(lib):
    """test/runtests.jl:184-215 ("Prolongation and retraction - 1-dimensional"): f(x) = sin(pi x) on 10, 11, 12 bits."""
    f = lambda n: O.vec_to_mps(np.sin(np.pi * O.qtt_xrange(n, 0.0, 1.0)), cutoff=1e-15)
    psi0, psi1, psi2 = f(10), f(11), f(12)
    nrm = lambda a: np.linalg.norm(dense(a))
    close = lambda a, b, rtol: np.linalg.norm(dense(a) - dense(b)) <= rtol * max(nrm(a), nrm(b))
    assert close(lib.prolongate(psi1), psi2, 1e-6)              # prolongate by one
    assert close(lib.prolongate(psi0, 2), psi2, 1e-6)           # prolongate by two
    assert close(lib.retract(psi2), psi1, 1e-5)                 # retract by one
    assert close(lib.retract(psi2, 2), psi0, 1e-5)              # retract by two
    # the GPU path agrees with the oracle's apply-based restatement
    assert close(lib.prolongate(psi1), O.prolongate(psi1), 1e-9) and close(lib.retract(psi2), O.retract(psi2), 1e-9)


def test_multigrid_linsolve_airy(lib):
    """Multigrid driver (examples/airy_equation/airy_equation.jl:453-516): the prolongated coarse solution is the starting
    state of the next level; the final level reaches the dense solution."""
    xi, xf = 1.0, 4.0
    alpha = beta = 1 / np.sqrt(2)

    def problem(n):
        h = (xf - xi) / 2 ** n
        return O.airy_mpo(n, xi, xf), O.boundary_value_mps(n, O.airy_solution(xi - h, alpha, beta), O.airy_solution(xf, alpha, beta))

    x, hist = lib.multigrid_linsolve(problem, [4, 5, 6, 7, 8], O.random_mps(4, 4, 1), nsweeps=2, maxdim=16, cutoff=1e-13, tol=1e-12,
                                     krylovdim=30, maxiter=10, return_history=True)
    A, b = problem(8)
    xd = np.linalg.solve(O.mpo_to_mat(A), dense(b))
    v = dense(x)
    assert np.linalg.norm(v - xd) <= 1e-5 * np.linalg.norm(xd)     # cutoff 1e-13 on sigma^2: ~3e-7 in norm
    assert [h["n"] for h in hist] == [4, 5, 6, 7, 8] and abs(hist[-1]["residual"]) < 1e-9
    # the prolongated coarse solution is already close to the fine one (what makes the warm start worthwhile)
    A7, b7 = problem(7)
    x7 = lib.linsolve(A7, b7, O.random_mps(7, 4, 1), nsweeps=3, maxdim=16, cutoff=1e-13, tol=1e-12, krylovdim=30, maxiter=10)
    guess = dense(lib.prolongate(x7, cutoff=1e-15))
    assert 1 - abs(guess @ xd) / np.linalg.norm(guess) / np.linalg.norm(xd) < 1e-3   # last point interpolates towards 0


def test_mirror_dft_front_ends_match_fft(lib):
    """`apply_dft_mpo` / `apply_idft_mpo` of the host mirror (reference src/dft.jl:121-133) with the cached host-built DFT MPO."""
    n = 10
    psi = sines(n)
    f = dense(psi)
    Fpsi = lib.apply_dft_mpo(psi, cutoff=1e-15)
    want = np.fft.fft(f) / 2 ** (n / 2)
    assert np.linalg.norm(dense(Fpsi) - want) < 5e-8 * np.linalg.norm(want)
    back = lib.apply_idft_mpo(Fpsi, cutoff=1e-15)
    assert np.linalg.norm(dense(back) - f) < 5e-8 * np.linalg.norm(f)


@pytest.mark.parametrize("k", [1, 3])
def test_fourier_interpolation_and_repeat_function(lib, k):
    """reference src/fourier_interpolation.jl:1-38 (runtests.jl fourier-interpolation block): a band-limited function on 2^n
    points is interpolated EXACTLY onto 2^(n+k) points; `repeat_function` tiles it 2^k times.  Compared with the oracle's
    restatement and with the analytic values."""
    n = 8
    x = O.qtt_xrange(n, 0.0, 1.0)
    f = lambda t: 0.7 * np.cos(2 * np.pi * 3 * t + 0.4) + 0.2 * np.sin(2 * np.pi * 11 * t) + 0.5
    psi = O.vec_to_mps(f(x), cutoff=1e-15)
    got = lib.fourier_interpolation(psi, k, cutoff=1e-15)
    assert len(got) == n + k
    xf = O.qtt_xrange(n + k, 0.0, 1.0)
    g = dense(got)
    assert np.abs(g.imag).max() < 1e-7
    assert np.abs(g.real - f(xf)).max() < 5e-7
    want = dense(O.fourier_interpolation(psi, k, cutoff=1e-15))
    assert np.abs(g - want).max() < 5e-7
    rep = dense(lib.repeat_function(psi, k, cutoff=1e-15))
    assert np.abs(rep - np.tile(f(x), 2 ** k)).max() < 5e-7
    assert np.abs(rep - dense(O.repeat_function(psi, k, cutoff=1e-15))).max() < 5e-7


def _multi_sine(lib, n, nf, seed=3):
    rng = np.random.default_rng(seed)
    freqs = rng.choice(2 ** (n - 2), size=nf, replace=False).astype(np.int64) + 1
    amps = rng.standard_normal(nf)
    psi = O.mps_add_directsum(*[[c * (amps[i] if j == 0 else 1.0) for j, c in enumerate(O.qtt_sin(2 * np.pi * f, n))]
                                for i, f in enumerate(freqs)])
    return psi, freqs, amps


@pytest.mark.parametrize("n,nf", [(16, 16), (20, 32)])
def test_dft_apply_multi_frequency_vs_fft(lib, n, nf):
    """BASELINE config 3's recipe below its full size (reference src/dft.jl:121-126, test/runtests.jl:249-274): the whole
    spectrum of a sum of `nf` sines against numpy's FFT.  Tolerance 1e-6 norm-relative: the floor is the double-precision
    phase 2 pi f x of the input construction (measured 4e-8)."""
    psi, freqs, amps = _multi_sine(lib, n, nf)
    got = lib.apply_dft_mpo(psi, cutoff=1e-15, maxdim=2 * nf)
    f = dense(psi)
    want = np.fft.fft(f) / 2 ** (n / 2)
    g = dense(got)
    assert np.linalg.norm(g - want) <= 1e-6 * np.linalg.norm(want)
    assert max(O.linkdims(got)) <= 2 * nf


def test_dft_apply_config3_full_size_analytic_spikes(lib):
    """BASELINE config 3 at its full size (n = 30 sites, chi <= 128): the DFT of sum_i a_i sin(2 pi f_i x) is -i a_i sqrt(N)/2 at
    k = f_i, +i a_i sqrt(N)/2 at k = N - f_i and zero elsewhere; the operator is unitary.  Sampled through `sample`."""
    n, nf = 30, 64
    N = 2 ** n
    psi, freqs, amps = _multi_sine(lib, n, nf)
    got = lib.apply_dft_mpo(psi, cutoff=1e-15, maxdim=2 * nf)
    assert max(O.linkdims(got)) <= 2 * nf
    spikes = np.concatenate([freqs, N - freqs])
    want = np.concatenate([-0.5j * amps * np.sqrt(N), 0.5j * amps * np.sqrt(N)])
    scale = np.abs(want).max()
    assert np.abs(lib.sample(got, spikes) - want).max() <= 1e-6 * scale
    rng = np.random.default_rng(5)
    off = rng.integers(0, N, size=2048)
    off = off[~np.isin(off, spikes)]
    assert np.abs(lib.sample(got, off)).max() <= 1e-6 * scale
    nin, nout = lib.inner(psi, psi), lib.inner(got, got)
    assert abs(abs(nout) / abs(nin) - 1.0) <= 1e-9

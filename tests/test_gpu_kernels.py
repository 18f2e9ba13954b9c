"""GPU parity tests of the individual kernels, through the C ABI, against the numpy oracle on the
same seeded inputs.  Tolerances are stated per test: FP64 contractions agree to a few ulps times the
contraction length (no reduced precision anywhere)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("tile", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 256, 96), (37, 53, 29), (1, 7, 3), (200, 130, 257), (768, 1024, 256)])
def test_gemm_nn(lib, tile, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)); C0 = rng.standard_normal((M, N))
    C, _ = lib.gemm(A, B, C0, alpha=0.7, beta=-0.3, tile=tile)
    assert rel(C, 0.7 * A @ B - 0.3 * C0) < 1e-14


@pytest.mark.parametrize("shape", [(64, 64, 64), (33, 129, 70), (5, 3, 2), (256, 256, 512)])
def test_gemm_transposed_layouts(lib, shape):
    M, N, K = shape
    rng = np.random.default_rng(N)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N))
    C, _ = lib.gemm(np.ascontiguousarray(A.T), B, transA=True)
    assert rel(C, A @ B) < 1e-14
    C, _ = lib.gemm(A, np.ascontiguousarray(B.T), transB=True)
    assert rel(C, A @ B) < 1e-14


CASES = [  # (chil, chir, wl, wm, wr)
    (4, 4, 3, 3, 3), (32, 32, 3, 3, 3), (16, 8, 5, 5, 5), (1, 2, 1, 3, 3), (2, 1, 3, 3, 1), (7, 5, 4, 3, 2),
    (128, 128, 4, 4, 4), (64, 96, 3, 5, 4), (3, 3, 9, 2, 6),
    # TMA-staged kernels (even bond dimensions): ragged edges are zero-filled by the TMA unit; K_B takes its TMA kernel for
    # single-wave grids only (4 chil / 32 x chir / 64 tiles between 112 and 148)
    (250, 254, 3, 3, 3), (256, 230, 3, 3, 3), (254, 256, 6, 2, 8), (130, 258, 2, 3, 5), (192, 320, 1, 2, 3), (18, 6, 5, 4, 3),
    (255, 257, 3, 3, 3),   # odd dimensions: tensor maps need 16-byte strides, the cp.async kernels take over
]


@pytest.mark.parametrize("case", CASES)
def test_matvec2_matches_oracle(lib, case):
    chil, chir, wl, wm, wr = case
    rng = np.random.default_rng(sum(case))
    L = rng.standard_normal((chil, wl, chil)); R = rng.standard_normal((chir, wr, chir))
    W1 = rng.standard_normal((wl, wm, 2, 2)); W2 = rng.standard_normal((wm, wr, 2, 2))
    v = rng.standard_normal((chil, 2, 2, chir))
    out, _ = lib.matvec2(L, W1, W2, R, v)
    assert rel(out, O.matvec2(L, W1, W2, R, v)) < 1e-13


def test_matvec2_headline_shape(lib):
    """chi = 256, w = 3 (BASELINE config 2 interior bond): parity + linearity at full size."""
    chi, w = 256, 3
    rng = np.random.default_rng(1)
    L = rng.standard_normal((chi, w, chi)); R = rng.standard_normal((chi, w, chi))
    W = O.laplacian_tensor()
    v1 = rng.standard_normal((chi, 2, 2, chi)); v2 = rng.standard_normal((chi, 2, 2, chi))
    o1, _ = lib.matvec2(L, W, W, R, v1)
    assert rel(o1, O.matvec2(L, W, W, R, v1)) < 1e-13
    o2, _ = lib.matvec2(L, W, W, R, v2)
    o12, _ = lib.matvec2(L, W, W, R, 2.0 * v1 - 0.5 * v2)
    assert rel(o12, 2.0 * o1 - 0.5 * o2) < 1e-13


@pytest.mark.parametrize("dims", [(4, 6, 3, 3), (32, 32, 3, 3), (1, 2, 1, 3), (2, 1, 3, 1), (64, 48, 5, 4), (7, 9, 2, 5)])
def test_env_updates_match_oracle(lib, dims):
    a, b, w, w2 = dims
    rng = np.random.default_rng(sum(dims))
    L = rng.standard_normal((a, w, a)); x = rng.standard_normal((a, 2, b)); W = rng.standard_normal((w, w2, 2, 2))
    Ln, _ = lib.env_left(L, x, W)
    assert rel(Ln, O.env_left_update(L, x, W)) < 1e-13
    R = rng.standard_normal((b, w2, b))
    Rn, _ = lib.env_right(R, x, W)
    assert rel(Rn, O.env_right_update(R, x, W)) < 1e-13


@pytest.mark.parametrize("nvec,ln", [(1, 64), (5, 1000), (30, 4096), (31, 65536), (7, 333), (30, 262144), (3, 8)])
def test_krylov_ops(lib, nvec, ln):
    rng = np.random.default_rng(nvec + ln)
    V = rng.standard_normal((nvec, ln)); w = rng.standard_normal(ln)
    _, h, _, _ = lib.krylov_op(0, V, w)
    assert rel(h, V @ w) < 1e-13
    w2, _, _, _ = lib.krylov_op(1, V, w, h=V @ w / ln)
    assert rel(w2, w - V.T @ (V @ w / ln)) < 1e-13
    _, _, nrm, _ = lib.krylov_op(2, V, w)
    assert abs(nrm - np.linalg.norm(w)) < 1e-12 * np.linalg.norm(w)
    # CGS2 against an orthonormal basis leaves an orthogonal remainder
    Q = np.linalg.qr(V.T)[0].T.copy()
    wo, hh, nrm, _ = lib.krylov_op(3, Q, w)
    assert np.abs(Q @ wo).max() < 1e-13 * np.linalg.norm(w)
    assert rel(hh, Q @ w) < 1e-12 and abs(nrm - np.linalg.norm(wo)) < 1e-12 * nrm
    if ln % 2 == 0:
        # the fused single-launch CGS2 Arnoldi step (op 4) returns the same coefficients, remainder and norm
        wf, hf, nf, _ = lib.krylov_op(4, Q, w)
        assert np.abs(Q @ wf).max() < 1e-13 * np.linalg.norm(w)
        assert rel(hf, Q @ w) < 1e-12 and abs(nf - np.linalg.norm(wf)) < 1e-12 * nf
        assert rel(wf, wo) < 1e-12
        wf2, hf2, nf2, _ = lib.krylov_op(4, Q, w)   # bit-reproducible: fixed-order reductions
        assert np.array_equal(wf, wf2) and np.array_equal(hf, hf2) and nf == nf2


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("dims", [(4, 4), (16, 16), (1, 2), (2, 1), (8, 3), (3, 8), (64, 64), (32, 128)])
def test_split2_full_rank(lib, ortho, dims):
    """Un-truncated split reproduces phi, the orthonormal factor is an isometry, sigma matches LAPACK."""
    a, c = dims
    rng = np.random.default_rng(a * 31 + c)
    phi = rng.standard_normal((a, 2, 2, c))
    r = lib.split2(phi, ortho=ortho, cutoff=0.0)
    k = r["rank"]
    assert k == min(2 * a, 2 * c)
    A = r["A"].reshape(2 * a, k); B = r["B"].reshape(k, 2 * c)
    assert rel(A @ B, phi.reshape(2 * a, 2 * c)) < 1e-13
    S = np.linalg.svd(phi.reshape(2 * a, 2 * c), compute_uv=False)
    assert np.abs(r["sigma"] - S[:k]).max() < 1e-13 * S[0]
    if ortho == "left":
        assert np.abs(A.T @ A - np.eye(k)).max() < 1e-13
    else:
        assert np.abs(B @ B.T - np.eye(k)).max() < 1e-13


@pytest.mark.parametrize("ortho", ["left", "right"])
def test_split2_truncation_rule(lib, ortho):
    """Same kept rank and truncerr as the oracle's restatement of NDTensors.truncate! on a graded spectrum,
    including sigma^2-relative cutoffs down to 1e-15 (needs high relative accuracy of small sigma)."""
    a = c = 24
    rng = np.random.default_rng(3)
    U = np.linalg.qr(rng.standard_normal((2 * a, 2 * a)))[0]
    V = np.linalg.qr(rng.standard_normal((2 * c, 2 * c)))[0]
    s = 10.0 ** (-np.arange(2 * a) * 0.45)
    phi = ((U * s) @ V.T).reshape(a, 2, 2, c)
    for cutoff, maxdim, mindim in ((1e-15, None, 1), (1e-12, None, 1), (1e-8, None, 1), (0.0, 10, 1), (1e-3, None, 7), (1e-12, 12, 1)):
        r = lib.split2(phi, ortho=ortho, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        Ao, Bo, err, So = O.split_two_site(phi, ortho, cutoff, maxdim, mindim)
        assert r["rank"] == So.size, (cutoff, maxdim, mindim)
        assert abs(r["truncerr"] - err) <= 1e-9 * max(err, 1e-300) + 1e-30
        assert np.abs(r["sigma"] / So - 1).max() < 1e-9
        k = r["rank"]
        got = r["A"].reshape(2 * a, k) @ r["B"].reshape(k, 2 * c)
        want = Ao.reshape(2 * a, k) @ Bo.reshape(k, 2 * c)
        assert rel(got, want) < 1e-11


def test_split2_headline_size(lib):
    """512 x 512 split (chi = 256): round trip + isometry at BASELINE size; prints the time for the log."""
    a = c = 256
    rng = np.random.default_rng(9)
    phi = rng.standard_normal((a, 2, 2, c))
    r = lib.split2(phi, ortho="left", cutoff=0.0, maxdim=256, reps=2)
    k = r["rank"]
    assert k == 256
    A = r["A"].reshape(2 * a, k)
    assert np.abs(A.T @ A - np.eye(k)).max() < 1e-12
    S = np.linalg.svd(phi.reshape(2 * a, 2 * c), compute_uv=False)
    assert np.abs(r["sigma"] - S[:k]).max() < 1e-12 * S[0]
    want_err = (S[k:] ** 2).sum() / (S ** 2).sum()
    assert abs(r["truncerr"] - want_err) < 1e-12
    print(f"split2 512x512: {r['ms']:.2f} ms, {r['sweeps']} Jacobi sweeps")

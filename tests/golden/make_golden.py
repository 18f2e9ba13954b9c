"""Generates the committed golden vectors tests/golden/*.npz from the CPU oracle.

Run from the repository root:  python tests/golden/make_golden.py
The oracle is first pinned against every assertion of the reference's own test-suite (tests/test_oracle_kat.py,
restating /root/reference/test/runtests.jl) and against dense / analytic solves (tests/test_oracle_sweep.py); only
then are its outputs frozen here.  The Julia reference itself cannot run in the build container (no julia, deps not
vendored), so these vectors are oracle outputs on explicit seeded inputs, not Julia outputs.
The GPU tests compare the CUDA path with these files; the CPU tests check that the oracle still reproduces them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack(prefix, cores):
    return {f"{prefix}_{j}": np.asarray(c) for j, c in enumerate(cores)}


def operators():
    d = {}
    n = 6
    d.update(pack("laplacian_n6_h0p5", O.laplacian_mpo(n, 0.5)))
    d.update(pack("prolongation_n5", O.prolongation_mpo(5)))
    d.update(pack("boundary_n6", O.boundary_value_mps(n, 0.3, -0.2)))
    d.update(pack("airy_n6", O.airy_mpo(n, 1.0, 4.0)))
    d.update(pack("dft_n5", O.dft_mpo(5)))
    d["xrange_n6_1_8"] = O.qtt_xrange(6, 1.0, 8.0)
    d["laplacian_dense_n4"] = O.mpo_to_mat(O.laplacian_mpo(4, 1.0))
    # bit maps (bit-exact contract): grid index <-> site bits, projection of leading / trailing bits
    f = np.arange(64, dtype=np.float64) ** 2
    psi = O.vec_to_mps(f, cutoff=0.0)
    d["project_bits_lead_101"] = O.mps_to_discrete_function(O.project_bits(psi, [1, 0, 1]))
    d["project_bits_trail_01"] = O.mps_to_discrete_function(O.project_bits(psi, [], right_bits=[0, 1]))
    d["index_bits_37_n6"] = np.array(O.index_to_bits(37, 6))
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **d)


def linsolve_airy():
    n = 8
    A = O.airy_mpo(n, 1.0, 4.0)
    b = O.boundary_value_mps(n, 0.3, -0.2)
    x0 = O.random_mps(n, 6, 11)
    kw = dict(nsweeps=3, maxdim=16, cutoff=1e-12, tol=1e-12, krylovdim=30, maxiter=10)
    x, rec = O.linsolve(A, b, x0, **kw)
    d = {}
    d.update(pack("x0", x0))
    d.update(pack("x", x))
    d["dense_x"] = O.mps_to_discrete_function(x)
    d["residual_normalized"] = np.array(O.sqeuclidean_normalized_mps(A, x, b))
    d["linkdims"] = np.array(O.linkdims(x))
    d["params"] = np.array([n, 1.0, 4.0, 0.3, -0.2, kw["nsweeps"], kw["maxdim"], kw["cutoff"], kw["tol"], kw["krylovdim"], kw["maxiter"]])
    np.savez_compressed(os.path.join(OUT, "linsolve_airy_n8.npz"), **d)


def apply_dft():
    n = 6
    psi = O.mps_add_directsum(*[O.qtt_sin(2 * np.pi * f, n) for f in (1, 2, 4, 8)])
    F = O.dft_mpo(n)
    out = O.apply_mpo_mps(F, psi, cutoff=1e-15)
    d = {}
    d.update(pack("psi", psi))
    d.update(pack("F", F))
    d["dense_out_reversed"] = O.mps_to_discrete_function(O.mps_reverse(out))
    d["fft_over_sqrtN"] = np.fft.fft(O.mps_to_discrete_function(psi)) / 2 ** (n / 2)
    d["linkdims"] = np.array(O.linkdims(out))
    np.savez_compressed(os.path.join(OUT, "apply_dft_n6.npz"), **d)


def split_rule():
    rng = np.random.default_rng(5)
    a = c = 8
    U = np.linalg.qr(rng.standard_normal((2 * a, 2 * a)))[0]
    V = np.linalg.qr(rng.standard_normal((2 * c, 2 * c)))[0]
    s = 10.0 ** (-np.arange(2 * a) * 0.9)
    phi = ((U * s) @ V.T).reshape(a, 2, 2, c)
    d = {"phi": phi, "sigma": s}
    for k, (cutoff, maxdim, mindim) in enumerate([(1e-12, None, 1), (1e-6, 5, 1), (0.0, 9, 3), (1e-30, None, 12)]):
        _, S, _, err = O.svd_truncated(phi.reshape(2 * a, 2 * c), cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        d[f"case{k}"] = np.array([cutoff, -1 if maxdim is None else maxdim, mindim, len(S), err])
    np.savez_compressed(os.path.join(OUT, "split_rule.npz"), **d)


if __name__ == "__main__":
    operators()
    linsolve_airy()
    apply_dft()
    split_rule()
    print("golden vectors written to", OUT)

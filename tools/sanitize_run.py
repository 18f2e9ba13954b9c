"""Smallest shapes that reach every persistent / cooperative kernel of the sweep path, for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py
Kernels reached: matvec_a/b (fused pair), cgs2_fused_kernel, gmres_small_kernel, qr_panel/qr_apply (Householder), gs_select /
gs_panel / gs_update (Gram-Schmidt QR), jacobi_gram_kernel<16|32>, jacobi_qr_block_kernel, jacobi_qr_kernel, gather_rows."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import itensorqtt_jl_b200 as q

rng = np.random.default_rng(0)
# splits: chi = 40 (80 x 80: Gram-Schmidt QR + blocked Gram Jacobi), rank-deficient with mindim (completion / Householder path),
# chi = 8 (row-cyclic kernels)
for chi, rank, mind in ((40, 80, 1), (40, 50, 80), (40, 3, 80), (8, 16, 1)):
    m = 2 * chi
    U = np.linalg.qr(rng.standard_normal((m, m)))[0]; V = np.linalg.qr(rng.standard_normal((m, m)))[0]
    s = np.r_[10.0 ** (-np.arange(rank) * 0.1), np.zeros(m - rank)]
    phi = ((U * s) @ V.T).reshape(chi, 2, 2, chi)
    for ortho in ("left", "right"):
        r = q.split2(phi, ortho=ortho, cutoff=0.0, maxdim=m, mindim=mind)
        k = r["rank"]
        A = r["A"].reshape(m, k); B = r["B"].reshape(k, m)
        assert np.linalg.norm(A @ B - phi.reshape(m, m)) <= 1e-12 * np.linalg.norm(phi), (chi, rank, ortho)
print("splits ok")
# sweeps: multi-kernel path (fused matvec pair + cooperative CGS2) and the single-CTA small-bond solver
n = 10
A = O.laplacian_mpo(n, 1.0); lam = O.helmholtz_lambda(n, 3)
b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
x0 = O.random_mps(n, 24, 5)
xg = q.linsolve(A, b, x0, -lam, 1.0, nsweeps=1, maxdim=24, cutoff=1e-13, fixed_matvecs=10)
xo, _ = O.linsolve(A, b, x0, -lam, 1.0, nsweeps=1, maxdim=24, cutoff=1e-13, fixed_matvecs=10)
assert O.fidelity_defect(xo, xg) <= 1e-10
xg = q.linsolve(A, b, O.random_mps(n, 4, 1), -lam, 1.0, nsweeps=1, maxdim=6, cutoff=1e-13, tol=1e-10, krylovdim=20, maxiter=4)
print("sweeps ok")

import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import itensorqtt_jl_b200 as lib
what = sys.argv[1]
if what == "iso":
    n, chi, nk = 30, 256, 18
    A = O.laplacian_mpo(n, 1.0); lam = -((2 * np.pi * (2.0 ** nk) / 2.0 ** n) ** 2)
    b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2)); x0 = O.random_mps(n, chi, 2)
    xg = lib.linsolve(A, b, x0, -lam, 1.0, nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30)
    d = []
    for c in xg[1:]:
        m = c.reshape(c.shape[0], -1)
        d.append(float(np.max(np.abs(m @ m.T - np.eye(m.shape[0])))))
    print("isometry defect per core:", ["%.1e" % v for v in d])
else:
    n, chi = 24, 128
    A = O.laplacian_mpo(n, 1.0); x0 = O.random_mps(n, chi, 5)
    target = -4.0 * 4.0 ** n
    for cutoff in (1e-8, 1e-12):
        kw = dict(nsweeps=1, maxdim=chi, mindim=1, cutoff=cutoff, tol=1e-300, krylovdim=30, maxiter=1)
        xo, _ = O.dmrg_target_krylov(A, x0, target, mode="lanczos", **kw)
        xg, info = lib.dmrg_target(A, x0, target_eigenvalue=target, return_info=True, **kw)
        eo = (O.mpo_expect(xo, A, xo) / O.mps_inner(xo, xo)).real
        eg = (O.mpo_expect(xg, A, xg) / O.mps_inner(xg, xg)).real
        print(cutoff, "energies", eo, eg, info[-1]["energy"], "fid", O.fidelity_defect(xo, xg), O.linkdims(xo)[:12], O.linkdims(xg)[:12])

"""Numbers behind tests/test_gpu_baseline_sizes.py: GPU vs oracle after one fixed-work sweep at several sizes."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import itensorqtt_jl_b200 as lib
cases = [(14, 256, 6, 1e-12), (18, 256, 8, 1e-12), (30, 256, 18, 1e-12), (30, 256, 18, 1e-8)]
for n, chi, nk, cutoff in cases:
    A = O.laplacian_mpo(n, 1.0)
    lam = -((2 * np.pi * (2.0 ** nk) / 2.0 ** n) ** 2)
    b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
    x0 = O.random_mps(n, chi, 2)
    kw = dict(nsweeps=1, maxdim=chi, mindim=1, cutoff=cutoff, fixed_matvecs=30)
    t0 = time.time(); xo, _ = O.linsolve(A, b, x0, -lam, 1.0, **kw); to = time.time() - t0
    xg = lib.linsolve(A, b, x0, -lam, 1.0, **kw)
    ro = O.sqeuclidean_normalized_shift(A, xo, b, -lam, 1.0) if hasattr(O, "sqeuclidean_normalized_shift") else None
    print(n, chi, cutoff, "maxlink", max(O.linkdims(xo)), max(O.linkdims(xg)), "oracle %.1f s" % to, "fidelity defect %.3e" % O.fidelity_defect(xo, xg), "linkdims equal", O.linkdims(xo) == O.linkdims(xg), flush=True)

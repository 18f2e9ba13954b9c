import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
import oracle as O
import itensorqtt_jl_b200 as lib
n, chi, nsw, nk = 30, 256, 16, 18
ctx = lib.default_context(0)
ctx.set_profiling(True)      # phase timers (one stream drain per phase), no per-kernel trace
A = O.laplacian_mpo(n, 1.0)
lam = -((2 * np.pi * (2.0 ** nk) / 2.0 ** n) ** 2)
b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
x0 = O.random_mps(n, chi, 2)
xg, info = lib.linsolve(A, b, x0, -lam, 1.0, return_info=True, nsweeps=nsw, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30)
for i, s in enumerate(info):
    print("sweep %d device_ms %.1f krylov %.1f svd %.1f env %.1f" % (i, s["device_ms"], s["t_krylov"], s["t_svd"], s["t_env"]))

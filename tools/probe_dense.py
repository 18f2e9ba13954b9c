"""Dense local solvers at the largest size configs[0] allows (4 * 32 * 32 = 4096): one sweep of a short chain, timed."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import oracle as O   # operand builders only

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
A = O.airy_mpo(n, 1.0, 8.0)
b = O.boundary_value_mps(n, 0.3, 0.7)
x0 = O.random_mps(n, chi, 7)
ctx = q.default_context(0)
kw = dict(nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, ctx=ctx, return_info=True)
q.linsolve_pseudoinverse(A, b, x0, **kw)
t0 = time.perf_counter()
x, info = q.linsolve_pseudoinverse(A, b, x0, **kw)
print(json.dumps({"mode": "dense_qr", "n": n, "chi": chi, "seconds": time.perf_counter() - t0, "device_ms": info[0]["device_ms"],
                  "solver_resid": info[0]["solver_resid"]}))
Al = O.laplacian_mpo(n, 1.0)
t0 = time.perf_counter()
psi, info = q.dmrg_target(Al, x0, target_eigenvalue=-1e-3, local_solver="dense", **kw)
print(json.dumps({"mode": "dense_eigen", "n": n, "chi": chi, "seconds": time.perf_counter() - t0, "device_ms": info[0]["device_ms"],
                  "energy": info[0]["energy"], "resid": info[0]["solver_resid"]}))

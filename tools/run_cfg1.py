"""BASELINE configs[0]: Airy ODE QTT linsolve, n = 20 bits, maxdim = 32, cutoff = 1e-12, Float64 (examples/airy_equation),
on the GPU through the public API, checked against the analytic solution u(x) = (Ai(-x) + Bi(-x)) / sqrt(2) sampled through
project_bits-style dense evaluation (2^20 points fit in memory)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import oracle as O   # operand builders / analytic solution only (tools, not the product path)

n, xi, xf = 20, 1.0, 8.0
alpha = beta = 1 / np.sqrt(2)
A = O.airy_mpo(n, xi, xf)
h = (xf - xi) / 2 ** n
ui = O.airy_solution(xi - h, alpha, beta); uf = O.airy_solution(xf, alpha, beta)
b = O.boundary_value_mps(n, ui, uf)
x0 = O.random_mps(n, 10, 1234)
ctx = q.default_context(0)
for mode, kw in (("gmres tol=1e-15 kd=30 maxiter=30 (example settings)", dict(tol=1e-15, krylovdim=30, maxiter=30)),
                 ("gmres tol=1e-10 kd=30 maxiter=4", dict(tol=1e-10, krylovdim=30, maxiter=4))):
    t0 = time.perf_counter()
    x, info = q.linsolve(A, b, x0, nsweeps=10, maxdim=32, cutoff=1e-12, ctx=ctx, return_info=True, **kw)
    dt = time.perf_counter() - t0
    u = q.mps_to_discrete_function(x)
    xs = O.qtt_xrange(n, xi, xf)
    ua = O.airy_solution(xs, alpha, beta)
    res = q.sqeuclidean_normalized((A, x), b, ctx=ctx)
    print(json.dumps({"cfg": 1, "mode": mode, "seconds": dt, "sweeps": 10, "matvecs": int(sum(i["matvecs"] for i in info)),
                      "maxlinkdim": int(info[-1]["maxlinkdim"]), "rel_err_vs_analytic_max": float(np.abs(u - ua).max() / np.abs(ua).max()),
                      "sqeuclidean_normalized": float(res)}), flush=True)

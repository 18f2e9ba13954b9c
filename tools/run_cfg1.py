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
                 ("gmres tol=1e-10 kd=30 maxiter=4", dict(tol=1e-10, krylovdim=30, maxiter=4)),
                 ("dense QR local solve (linsolve_pseudoinverse, the example's default solver)", None)):
    t0 = time.perf_counter()
    if kw is None:
        x, info = q.linsolve_pseudoinverse(A, b, x0, nsweeps=10, maxdim=32, cutoff=1e-12, ctx=ctx, return_info=True)
    else:
        x, info = q.linsolve(A, b, x0, nsweeps=10, maxdim=32, cutoff=1e-12, ctx=ctx, return_info=True, **kw)
    dt = time.perf_counter() - t0
    u = q.mps_to_discrete_function(x)
    xs = O.qtt_xrange(n, xi, xf)
    ua = O.airy_solution(xs, alpha, beta)
    res = q.sqeuclidean_normalized((A, x), b, ctx=ctx)
    print(json.dumps({"cfg": 1, "mode": mode, "seconds": dt, "sweeps": 10, "matvecs": int(sum(i["matvecs"] for i in info)),
                      "maxlinkdim": int(info[-1]["maxlinkdim"]), "rel_err_vs_analytic_max": float(np.abs(u - ua).max() / np.abs(ua).max()),
                      "sqeuclidean_normalized": float(res)}), flush=True)

# dense solver at the largest local dimension configs[0] allows (maxdim = 32 -> 4 * 32 * 32 = 4096): one sweep from a bond-32 start
x32 = O.random_mps(n, 32, 7)
q.linsolve_pseudoinverse(A, b, x32, nsweeps=1, maxdim=32, mindim=32, cutoff=0.0, ctx=ctx)   # warm (allocator)
t0 = time.perf_counter()
x, info = q.linsolve_pseudoinverse(A, b, x32, nsweeps=1, maxdim=32, mindim=32, cutoff=0.0, ctx=ctx, return_info=True)
print(json.dumps({"cfg": 1, "mode": "dense QR, one sweep from a bond-32 start, mindim=32 (local matrices up to 4096^2)", "seconds": time.perf_counter() - t0,
                  "device_ms": info[0]["device_ms"], "maxlinkdim": int(info[0]["maxlinkdim"]), "solver_resid": info[0]["solver_resid"]}), flush=True)

"""Per-bond phase timings (QTT_TRACE) of the late sweeps of the bench workload: python tools/trace_late_sweep.py [n chi nsweeps [notrace]]
(notrace: same sweeps without the trace counters, for ncu captures)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TRACE = not (len(sys.argv) > 4 and sys.argv[4] == "notrace")
if TRACE:
    os.environ["QTT_TRACE"] = "1"
import oracle as O
import itensorqtt_jl_b200 as lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 12
nk = 18
ctx = lib.default_context(0)
ctx.set_profiling(TRACE)
A = O.laplacian_mpo(n, 1.0)
lam = -((2 * np.pi * (2.0 ** nk) / 2.0 ** n) ** 2)
b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
x0 = O.random_mps(n, chi, 2)
kw = dict(nsweeps=nsw, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30)
xg, info = lib.linsolve(A, b, x0, -lam, 1.0, return_info=True, **kw)
for i, s in enumerate(info):
    print("sweep %d device_ms %.1f krylov %.1f svd %.1f env %.1f" % (i, s["device_ms"], s["t_krylov"], s["t_svd"], s["t_env"]))

import json, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import itensorqtt_jl_b200 as q
from concurrent.futures import ThreadPoolExecutor
n, chi, ninst = 28, int(sys.argv[2]) if len(sys.argv) > 2 else 128, 16
T = int(sys.argv[1])
ctxs = [q.Context(0) for _ in range(T)]
A = q.laplacian_mpo(n, 1.0)
b = q.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
dev = [(q.DeviceMPO(A, c), q.DeviceMPS(b, c)) for c in ctxs]
x0s = [q.random_mps(n, chi, 1000 + i) for i in range(ninst)]
def solve(i, t):
    c = ctxs[t]; dA, db = dev[t]
    k = 2 * np.pi * (2 ** 16 + i); lam = -((k / 2 ** n) ** 2)
    x, info = q.linsolve(dA, db, x0s[i], -lam, 1.0, nsweeps=2, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30, x0_canonical=True,
                         ctx=c, return_info=True, on_device=True)
    return info[-1]["solver_resid"]
for t in range(T): solve(0, t)
t0 = time.perf_counter()
def worker(t):
    return [solve(i, t) for i in range(t, ninst, T)]
with ThreadPoolExecutor(T) as ex:
    res = list(ex.map(worker, range(T)))
dt = time.perf_counter() - t0
print(json.dumps({"threads": T, "chi": chi, "instances": ninst, "wall_s": dt, "instances_per_s": ninst / dt}))

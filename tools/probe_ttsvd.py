import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import oracle as O

def report(name, v, **kw):
    g = q.vec_to_mps(v, **kw); o = O.vec_to_mps(v, **kw)
    dg = O.mps_to_discrete_function(g); do = O.mps_to_discrete_function(o)
    orth = max(np.abs(c.reshape(-1, c.shape[2]).T @ c.reshape(-1, c.shape[2]) - np.eye(c.shape[2])).max() for c in g[:-1])
    print(name, kw, "\n  gpu", O.linkdims(g), "\n  ora", O.linkdims(o), "\n  err gpu %.3e ora %.3e orth %.2e" % (
        np.linalg.norm(dg - v) / np.linalg.norm(v), np.linalg.norm(do - v) / np.linalg.norm(v), orth), flush=True)

n = 13
rng = np.random.default_rng(3)
x = O.qtt_xrange(n, 0.0, 1.0)
v = sum(rng.standard_normal() * np.sin(2 * np.pi * rng.integers(1, 2 ** (n - 2)) * x + rng.uniform(0, 6)) for _ in range(12))
report("12 sines", v, cutoff=1e-14)
report("12 sines", v, cutoff=1e-20)
n = 16
f = lambda t: np.exp(-((t - 0.4) ** 2) / 0.02) * np.cos(37.0 * t) + 0.2 * t ** 3
v = f(O.qtt_xrange(n, 0.0, 1.0))
for c in (1e-15, 1e-12, 1e-10):
    report("smooth", v, cutoff=c)

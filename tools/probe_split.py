import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
rng = np.random.default_rng(0)
chi = 256
kind = sys.argv[1] if len(sys.argv) > 1 else "graded"
U = np.linalg.qr(rng.standard_normal((2 * chi, 2 * chi)))[0]; V = np.linalg.qr(rng.standard_normal((2 * chi, 2 * chi)))[0]
s = 10.0 ** (-np.arange(2 * chi) * 0.07) if kind == "graded" else np.abs(rng.standard_normal(2 * chi)) + 0.1
phi = ((U * s) @ V.T).reshape(chi, 2, 2, chi)
r = q.split2(phi, ortho="right", cutoff=0.0, maxdim=chi, reps=2)
print(json.dumps({"split2": kind, "ms": r["ms"], "jacobi_sweeps": r["sweeps"], "rank": r["rank"]}))

import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
rng = np.random.default_rng(0)
if os.environ.get("QTT_TRACE"):
    q.default_context(0).set_profiling(True)
chi = 256; m = 2*chi
U = np.linalg.qr(rng.standard_normal((m, m)))[0]; V = np.linalg.qr(rng.standard_normal((m, m)))[0]
s = np.abs(rng.standard_normal(m)) + 0.1
phi = ((U * s) @ V.T).reshape(chi, 2, 2, chi)
try:
    r = q.split2(phi, ortho="right", cutoff=0.0, maxdim=chi, mindim=chi, reps=2)
    print(r["ms"], r["sweeps"])
except Exception as e:   # measurement modes that break convergence (QTT_GRAM_SKIP)
    print("split failed:", str(e)[:100])

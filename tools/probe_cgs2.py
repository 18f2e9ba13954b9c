"""Fused CGS2 step (op 4) time vs number of basis vectors k at len = 4 chi^2, back-to-back launches (CUDA events)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ln = 4 * chi * chi
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.standard_normal((ln, 31)))
V = np.ascontiguousarray(Q.T)
w = rng.standard_normal(ln)
res = {}
for k in (1, 2, 4, 8, 12, 14, 16, 20, 24, 30):
    _, _, _, ms = q.krylov_op(4, V[:k], w, reps=50)
    res[k] = round(ms * 1e3, 2)
print(json.dumps({"chi": chi, "len": ln, "env": {e: os.environ.get(e) for e in ("QTT_CGS2_STASH", "QTT_CGS2_THREADS", "QTT_CGS2_MINPER")}, "us_per_step_vs_k": res}))

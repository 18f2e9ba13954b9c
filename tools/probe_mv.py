"""Minimal matvec probe for ncu: one (chi, w) case, few reps."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
c = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ww = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.default_rng(0)
L = rng.standard_normal((c, ww, c)); R = rng.standard_normal((c, ww, c)); W = rng.standard_normal((ww, ww, 2, 2))
v = rng.standard_normal((c, 2, 2, c))
_, ms = q.matvec2(L, W, W, R, v, reps=reps)
print(json.dumps({"matvec2": [c, ww], "ms": ms, "tflops": q.matvec2_flops(c, c, ww, ww, ww) / ms * 1e-9}))

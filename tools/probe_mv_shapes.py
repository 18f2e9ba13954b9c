"""Two-site matvec time for the bond shapes of one cfg2 sweep (chi_l, chi_r), w = 3, and a few K_A tile shapes."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
rng = np.random.default_rng(0)
ww = 3
W = rng.standard_normal((ww, ww, 2, 2))
out = {}
for cl, cr in ((256, 256), (128, 256), (64, 256), (32, 128), (16, 64), (256, 128), (256, 64), (128, 32), (64, 16)):
    L = rng.standard_normal((cl, ww, cl)); R = rng.standard_normal((cr, ww, cr)); v = rng.standard_normal((cl, 2, 2, cr))
    _, ms = q.matvec2(L, W, W, R, v, reps=50)
    out["%dx%d" % (cl, cr)] = {"us": round(ms * 1e3, 2), "tflops": round(q.matvec2_flops(cl, cr, ww, ww, ww) / ms * 1e-9, 2)}
print(json.dumps({"shape_env": os.environ.get("QTT_MVA_SHAPE"), "matvec": out}))

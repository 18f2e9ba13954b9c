"""Real `apply` probe (prolongation MPO on a random real QTT): one warm-up + one timed apply; dumps sampled values so that
two builds / modes (QTT_EIGEN_QR=0/1) can be compared."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tag = sys.argv[3] if len(sys.argv) > 3 else "x"
ctx = q.default_context(0)
psi = q.random_mps(n, chi, 7)
dummy = np.zeros((1, 2, 1)); dummy[0, 0, 0] = 1.0
ext = q.DeviceMPS(list(psi) + [dummy], ctx)
P = q.DeviceMPO(q.prolongation_mpo(n + 1), ctx)
q.apply(P, ext, cutoff=1e-8, on_device=True).free()
ctx.sync()
l0 = ctx.launches
t0 = time.perf_counter()
out = q.apply(P, ext, cutoff=1e-8, on_device=True)
ctx.sync()
dt = time.perf_counter() - t0
idx = np.random.default_rng(0).integers(0, 1 << (n + 1), size=2048)
vals = q.sample(out, idx, ctx=ctx)
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"prolong_{tag}.npy"), vals)
print(json.dumps({"n": n, "chi": chi, "tag": tag, "apply_ms": dt * 1e3, "launches": ctx.launches - l0, "chi_out": max(out.dims()),
                  "norm2": q.inner(out, out, ctx=ctx)}))

"""Diagnostic (GPU): per-phase breakdown of consecutive bench-config sweeps and the spectrum of a mid-chain two-site tensor."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import bench

nsw = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n, chi = 30, 256
ctx = q.default_context(0)
A, b, x0, lam = bench.workload(0, n, chi, 18)
ctx.set_profiling(True)
dA, db, dx = q.DeviceMPO(A, ctx), q.DeviceMPS(b, ctx), q.DeviceMPS(x0, ctx)
x = dx
for i in range(nsw):
    x, info = q.linsolve(dA, db, x, -lam, 1.0, nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30, x0_canonical=True,
                         ctx=ctx, return_info=True, on_device=True)
    print(json.dumps({"sweep": i, **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in info[0].items()}}))
ctx.set_profiling(False)
cores = x.download()
j = n // 2 - 1
phi = np.einsum("asm,mtc->astc", cores[j], cores[j + 1]).reshape(2 * chi, 2 * chi)
s = np.linalg.svd(phi, compute_uv=False)
print(json.dumps({"bond": j, "sigma_rel": {str(k): float(s[k] / s[0]) for k in (0, 1, 2, 3, 5, 10, 20, 50, 100, 200, 255, 256, 300, 400, 511)}}))
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/phi_mid.npy", phi)

"""Timing probe of the device-resident MPS algebra (qtt_mps_to_vec / qtt_mps_sample / qtt_mps_add); wall clock around the
synchronous C calls, best of `reps`."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itensorqtt_jl_b200 as q

reps = 5
ctx = q.default_context(0)


def best(f):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts), r


for n, chi in ((26, 32), (28, 64), (28, 128)):
    psi = q.DeviceMPS(q.random_mps(n, chi, 1), ctx)
    out = torch.empty(1 << n, dtype=torch.float64, device="cuda")
    t, _ = best(lambda: q.mps_to_vec(psi, ctx=ctx, out=out))
    dims = psi.dims()
    print(json.dumps({"op": "mps_to_vec", "n": n, "chi": chi, "ms": t * 1e3, "out_GB_per_s": 8.0 * (1 << n) / t / 1e9,
                      "final_gemm_tflops_lower_bound": 2.0 * (1 << n) * min(dims[n // 2 - 3:n // 2 + 4]) / t / 1e12}))
    # spot check against point sampling
    idx = np.random.default_rng(0).integers(0, 1 << n, size=4096)
    ts, vals = best(lambda: q.sample(psi, idx, ctx=ctx))
    ref = out[torch.from_numpy(idx).cuda()].cpu().numpy()
    print(json.dumps({"op": "mps_sample", "n": n, "chi": chi, "count": int(idx.size), "ms": ts * 1e3,
                      "us_per_point": ts * 1e6 / idx.size, "max_abs_diff_vs_to_vec": float(np.max(np.abs(vals - ref)))}))
    del out
    psi.free()

n = 30
for cx, cy in ((32, 32), (64, 64)):
    x = q.DeviceMPS(q.random_mps(n, cx, 2), ctx); y = q.DeviceMPS(q.random_mps(n, cy, 3), ctx)
    t, r = best(lambda: q.add(x, y, cutoff=1e-15, maxdim=cx + cy, ctx=ctx, on_device=True))
    print(json.dumps({"op": "mps_add", "n": n, "chi": [cx, cy], "ms": t * 1e3, "maxlinkdim": max(r.dims())}))
    x.free(); y.free()

"""One LATE sweep of the bench workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (the earlier sweeps run
un-instrumented): python tools/late_launch_list.py [nwarm]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itensorqtt_jl_b200 as q
import bench
nwarm = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ctx = q.default_context(0)
A, b, x0, lam = bench.workload(0, 30, 256, 18)
dA, db, x = q.DeviceMPO(A, ctx), q.DeviceMPS(b, ctx), q.DeviceMPS(x0, ctx)
kw = dict(nsweeps=1, maxdim=256, mindim=256, cutoff=0.0, fixed_matvecs=30, x0_canonical=True, ctx=ctx, return_info=True, on_device=True)
for i in range(nwarm):
    xn, info = q.linsolve(dA, db, x, -lam, 1.0, **kw)
    x.free(); x = xn
ctx.sync()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
xn, info = q.linsolve(dA, db, x, -lam, 1.0, **kw)
ctx.sync()
rt.cudaProfilerStop()
print("sweep", nwarm + 1, "device_ms %.1f" % info[0]["device_ms"], "launches", ctx.launches)

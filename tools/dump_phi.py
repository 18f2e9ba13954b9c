"""Dump the solved two-site tensor of one bond of the benchmark workload at a given sweep (QTT_DUMP facility of sweep.cu).
usage: QTT_TRACE=1 QTT_DUMP=<sweep>,<dir>,<bond>,<path> python tools/dump_phi.py <nsweeps>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import bench
nsw = int(sys.argv[1])
ctx = q.default_context(0)
A, b, x0, lam = bench.workload(0, 30, 256, 18)
ctx.set_profiling(True)
dA, db, x = q.DeviceMPO(A, ctx), q.DeviceMPS(b, ctx), q.DeviceMPS(x0, ctx)
for i in range(nsw):
    x, info = q.linsolve(dA, db, x, -lam, 1.0, nsweeps=1, maxdim=256, mindim=256, cutoff=0.0, fixed_matvecs=30, x0_canonical=True,
                         ctx=ctx, return_info=True, on_device=True)

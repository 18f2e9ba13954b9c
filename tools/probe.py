"""Kernel-level timing probe (GPU): GEMM tile sweep at the matvec stage shapes, Krylov ops, split2 on a realistic
two-site tensor, and the per-phase breakdown of one sweep.  Output: JSON lines."""
import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

what = set(sys.argv[1:]) or {"gemm", "krylov", "split", "sweep", "matvec"}
ctx = q.default_context(0)
rng = np.random.default_rng(0)
chi, w = 256, 3

if "gemm" in what:
    for name, (M, N, K, batch_note) in {"stage1": (w * chi, 4 * chi, chi, ""), "stage4_full": (4 * chi, chi, w * chi, ""),
                                         "stage4_perw": (4 * chi, chi, chi, "x3 batch"), "big": (4096, 4096, 4096, "")}.items():
        A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N))
        for tile in range(4):
            _, ms = q.gemm(A, B, tile=tile, reps=50)
            print(json.dumps({"gemm": name, "mnk": [M, N, K], "tile": tile, "ms": ms, "tflops": 2 * M * N * K / ms * 1e-9}))
if "matvec" in what:
    for c, ww in ((256, 3), (256, 4), (128, 4), (512, 3)):
        L = rng.standard_normal((c, ww, c)); R = rng.standard_normal((c, ww, c)); W = rng.standard_normal((ww, ww, 2, 2))
        v = rng.standard_normal((c, 2, 2, c))
        _, ms = q.matvec2(L, W, W, R, v, reps=100)
        f = q.matvec2_flops(c, c, ww, ww, ww)
        print(json.dumps({"matvec2": [c, ww], "ms": ms, "tflops": f / ms * 1e-9}))
    for c, ww in ((256, 3), (512, 3)):
        L = rng.standard_normal((c, ww, c)); x = rng.standard_normal((c, 2, c)); W = rng.standard_normal((ww, ww, 2, 2))
        _, ms = q.env_left(L, x, W, reps=50)
        print(json.dumps({"env_left": [c, ww], "ms": ms, "tflops": q.env_flops(c, c, ww, ww) / ms * 1e-9}))
        _, ms = q.env_right(L, x, W, reps=50)
        print(json.dumps({"env_right": [c, ww], "ms": ms, "tflops": q.env_flops(c, c, ww, ww) / ms * 1e-9}))
if "krylov" in what:
    ln = 4 * chi * chi
    for nvec in (1, 8, 16, 30):
        V = rng.standard_normal((nvec, ln)); wv = rng.standard_normal(ln)
        for op, nm in ((0, "dots"), (1, "axpys"), (3, "cgs2"), (4, "cgs2_fused")):
            _, _, _, ms = q.krylov_op(op, V, wv, h=np.ones(nvec) * 1e-3, reps=100)
            passes = {0: 1, 1: 1, 3: 4, 4: 4}[op]
            print(json.dumps({"krylov": nm, "nvec": nvec, "len": ln, "ms": ms, "GBps_alg": passes * (nvec + 1) * ln * 8 / ms * 1e-6}))
if "split" in what:
    # graded spectrum like a converged two-site tensor, and a flat random one
    for name, decay in (("random", 0.0), ("graded", 0.07), ("steep", 0.3)):
        a = c = chi
        U = np.linalg.qr(rng.standard_normal((2 * a, 2 * a)))[0]; V = np.linalg.qr(rng.standard_normal((2 * c, 2 * c)))[0]
        s = 10.0 ** (-np.arange(2 * a) * decay) if decay else np.abs(rng.standard_normal(2 * a)) + 0.1
        phi = ((U * s) @ V.T).reshape(a, 2, 2, c)
        for ortho in ("left", "right"):
            r = q.split2(phi, ortho=ortho, cutoff=0.0, maxdim=chi, reps=3)
            print(json.dumps({"split2": name, "ortho": ortho, "ms": r["ms"], "jacobi_sweeps": r["sweeps"], "rank": r["rank"]}))
if "sweep" in what:
    import bench
    n = 30
    A, b, x0, lam = bench.workload(0, n, chi, 18)
    ctx.set_profiling(True)
    dA, db, dx = q.DeviceMPO(A, ctx), q.DeviceMPS(b, ctx), q.DeviceMPS(x0, ctx)
    x = dx
    for i in range(3):
        x, info = q.linsolve(dA, db, x, -lam, 1.0, nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30, x0_canonical=True,
                             ctx=ctx, return_info=True, on_device=True)
        print(json.dumps({"sweep": i, **info[0]}))
    ctx.set_profiling(False)

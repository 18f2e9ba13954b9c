"""Runs the BASELINE.json parity / stress configurations 3-5 once on the GPU and prints JSON lines (profiles/).
cfg3: DFT apply (complex128) on a multi-frequency signal; cfg4: Laplacian lowest eigenpair, chi = 512 Lanczos;
cfg5: batched Helmholtz wavenumber sweep (instances sharded over ranks when launched under torchrun)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

what = set(sys.argv[1:]) or {"cfg3", "cfg4", "cfg5"}
ctx = q.default_context(int(os.environ.get("LOCAL_RANK", "0")))


def sync():
    ctx.sync()


if "cfg3" in what:
    for n, nf in ((16, 16), (20, 32), (24, 64)):
        rng = np.random.default_rng(3)
        freqs = rng.choice(2 ** (n - 2), size=nf, replace=False) + 1
        amps = rng.standard_normal(nf)
        psi = q.mps_add_directsum(*[[c * (amps[i] if j == 0 else 1.0) for j, c in enumerate(q.qtt_sin(2 * np.pi * f, n))] for i, f in enumerate(freqs)])
        t0 = time.perf_counter(); F = q.dft_mpo(n); t_build = time.perf_counter() - t0
        dpsi, dF = q.DeviceMPS(psi, ctx), q.DeviceMPO(F, ctx)
        out = q.apply(dF, dpsi, cutoff=1e-15, maxdim=2 * nf, on_device=True); sync()
        t0 = time.perf_counter()
        out = q.apply(dF, dpsi, cutoff=1e-15, maxdim=2 * nf, on_device=True); sync()
        dt = time.perf_counter() - t0
        dims = out.dims()
        rec = {"cfg": 3, "n": n, "n_freq": nf, "chi_in": max(dpsi.dims()), "w_dft": max(dF.dims()), "chi_out": max(dims), "apply_s": dt,
               "dft_mpo_build_host_s": t_build}
        if n <= 20:
            f = q.mps_to_discrete_function(psi)
            got = q.mps_to_discrete_function([np.ascontiguousarray(np.transpose(c, (2, 1, 0))) for c in reversed(out.download())])
            want = np.fft.fft(f) / 2 ** (n / 2)
            rec["rel_err_vs_fft"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        print(json.dumps(rec), flush=True)

if "cfg4" in what:
    n = 24
    for chi in (128, 512):
        A = q.laplacian_mpo(n, 1.0)
        N = 2 ** n
        target = -4 * np.sin(np.pi / (2 * (N + 1))) ** 2
        x0 = q.random_mps(n, chi, 4)
        dA, dx = q.DeviceMPO(A, ctx), q.DeviceMPS(x0, ctx)
        ctx.set_profiling(True)
        x, info = q.dmrg_target(dA, dx, target_eigenvalue=target, nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30,
                                x0_canonical=True, ctx=ctx, return_info=True, on_device=True)
        ctx.set_profiling(False)
        i0 = info[0]
        print(json.dumps({"cfg": 4, "n": n, "chi": chi, "solver": "lanczos fixed 30", "sweep_s": i0["seconds"], "t_krylov_ms": i0["t_krylov"],
                          "t_svd_ms": i0["t_svd"], "t_env_ms": i0["t_env"], "matvecs": i0["matvecs"],
                          "matvec_tflops_incl_krylov": i0["flops_matvec"] / (i0["t_krylov"] * 1e-3) / 1e12, "ritz": i0["energy"], "target": target}), flush=True)

if "cfg5" in what:
    import torch, torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n, chi, ninst = 28, 128, int(os.environ.get("QTT_NINST", "16"))
    A = q.laplacian_mpo(n, 1.0)
    b = q.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
    dA, db = q.DeviceMPO(A, ctx), q.DeviceMPS(b, ctx)
    inst = []
    for i in range(ninst):
        k = 2 * np.pi * (2 ** 16 + i)
        lam = -((k / 2 ** n) ** 2)
        inst.append((dA, db, None, -lam, 1.0))
    def solve_one(i):
        x0 = q.random_mps(n, chi, 1000 + i)
        t0 = time.perf_counter()
        x, info = q.linsolve(dA, db, x0, inst[i][3], 1.0, nsweeps=2, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30, x0_canonical=True,
                             ctx=ctx, return_info=True, on_device=True)
        return None, (info[-1]["solver_resid"], sum(j["matvecs"] for j in info), time.perf_counter() - t0)
    solve_one(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table, _ = q.batch_solve(ninst, solve_one, device="cuda" if world > 1 else None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"cfg": 5, "n": n, "chi": chi, "instances": ninst, "world": world, "wall_s": dt, "instances_per_s": ninst / dt,
                          "resid_max": float(np.nanmax(table[:, 0])), "s_per_instance_mean": float(np.nanmean(table[:, 2]))}), flush=True)
    if world > 1:
        dist.destroy_process_group()

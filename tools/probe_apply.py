"""DFT apply probe for the ncu launch list: one warm-up apply + one timed apply of cfg3's recipe at a reduced size."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = q.default_context(0)
rng = np.random.default_rng(3)
freqs = rng.choice(2 ** (n - 2), size=nf, replace=False) + 1
amps = rng.standard_normal(nf)
psi = q.mps_add_directsum(*[[c * (amps[i] if j == 0 else 1.0) for j, c in enumerate(q.qtt_sin(2 * np.pi * f, n))] for i, f in enumerate(freqs)])
F = q.dft_mpo(n)
dpsi, dF = q.DeviceMPS(psi, ctx), q.DeviceMPO(F, ctx)
for _ in range(warm):
    q.apply(dF, dpsi, cutoff=1e-15, maxdim=2 * nf, on_device=True).free()
ctx.sync()
l0 = ctx.launches
t0 = time.perf_counter()
out = q.apply(dF, dpsi, cutoff=1e-15, maxdim=2 * nf, on_device=True)
ctx.sync()
dt = time.perf_counter() - t0
print(json.dumps({"n": n, "n_freq": nf, "chi_in": max(dpsi.dims()), "w_dft": max(dF.dims()), "chi_out": max(out.dims()), "apply_ms": dt * 1e3,
                  "launches": ctx.launches - l0}))

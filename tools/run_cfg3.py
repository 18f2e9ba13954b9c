"""BASELINE config 3: `apply_dft_mpo` on a multi-frequency signal (reference src/dft.jl:121-133), ComplexF64 on the GPU.
n = 16, 20: the whole spectrum against numpy FFT.  n = 24, 30: the analytic spikes of sum_i a_i sin(2 pi f_i x)
(-i a_i sqrt(N)/2 at k = f_i, +i a_i sqrt(N)/2 at k = N - f_i, zero elsewhere) sampled through `sample`, plus unitarity.
Prints one JSON line per size (profiles/).  Usage: python tools/run_cfg3.py [n:nfreq ...]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

ctx = q.default_context(0)
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(16, 16), (20, 32), (24, 64), (30, 64)]
for n, nf in cases:
    N = 2 ** n
    rng = np.random.default_rng(3)
    freqs = rng.choice(2 ** (n - 2), size=nf, replace=False).astype(np.int64) + 1
    amps = rng.standard_normal(nf)
    psi = q.mps_add_directsum(*[[c * (amps[i] if j == 0 else 1.0) for j, c in enumerate(q.qtt_sin(2 * np.pi * f, n))] for i, f in enumerate(freqs)])
    t0 = time.perf_counter(); F = q.dft_mpo(n); t_build = time.perf_counter() - t0
    dpsi, dF = q.DeviceMPS(psi, ctx), q.DeviceMPO(F, ctx)
    kw = dict(cutoff=1e-15, maxdim=2 * nf, on_device=True)
    out = q.apply(dF, dpsi, **kw); ctx.sync(); out.free()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = q.apply(dF, dpsi, **kw); ctx.sync()
        ts.append(time.perf_counter() - t0)
        if _ < 2: out.free()
    res = [np.ascontiguousarray(np.transpose(c, (2, 1, 0))) for c in reversed(out.download())]   # reverse(...) of src/dft.jl:124
    rec = {"cfg": 3, "n": n, "n_freq": nf, "chi_in": max(dpsi.dims()), "w_dft": max(dF.dims()), "chi_out": max(out.dims()),
           "apply_ms_best_of_3": 1e3 * min(ts), "apply_ms_all": [1e3 * t for t in ts], "dft_mpo_build_host_s": t_build}
    nin, nout = q.inner(psi, psi), q.inner(res, res)
    rec["norm_ratio_minus_1"] = float(abs(np.sqrt(abs(nout) / abs(nin)) - 1.0))
    if n <= 20:
        f = q.mps_to_discrete_function(psi)
        got = q.mps_to_discrete_function(res)
        want = np.fft.fft(f) / 2 ** (n / 2)
        rec["rel_err_vs_fft"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    # analytic spikes, any n
    spikes = np.concatenate([freqs, N - freqs])
    want_s = np.concatenate([-0.5j * amps * np.sqrt(N), 0.5j * amps * np.sqrt(N)])
    got_s = q.sample(res, spikes)
    off = rng.integers(0, N, size=4096)
    off = off[~np.isin(off, spikes)]
    got_o = q.sample(res, off)
    scale = np.abs(want_s).max()
    rec["spike_max_err_rel"] = float(np.abs(got_s - want_s).max() / scale)
    rec["offspike_max_abs_rel"] = float(np.abs(got_o).max() / scale)
    print(json.dumps(rec), flush=True)
    out.free(); dpsi.free(); dF.free()

"""SURVEY 8(f) row 1 / cfg3 input (ii): TT-SVD of 2^n samples of psi = sum_{m<64} a_m sin(2 pi f_m x + phi_m) (QTT rank 128)
on the GPU (`vec_to_mps`, cutoff 1e-15, maxdim 128), samples generated on the device.  Reports time, the HBM traffic the
algorithm needs (per step: 2 reads of the q x N remainder + 1 write, q = 2 r) against the measured HBM peak, and the error
against the analytic samples at random grid points."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
nterms = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(3)
f = rng.choice(2 ** min(n - 1, 29), size=nterms, replace=False).astype(np.int64)
a = rng.standard_normal(nterms); ph = rng.uniform(0, 2 * np.pi, nterms)
N = 2 ** n
dev = torch.device("cuda:0")
v = torch.zeros(N, dtype=torch.float64, device=dev)
j = torch.arange(N, dtype=torch.int64, device=dev)
for m in range(nterms):
    v += a[m] * torch.sin((2 * np.pi / N) * ((j * int(f[m])) & (N - 1)).to(torch.float64) + ph[m])
del j
torch.cuda.synchronize()
ctx = q.default_context(0)
res = {}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, info = q.vec_to_mps(v, cutoff=1e-15, maxdim=2 * nterms, ctx=ctx, on_device=True, return_info=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    dims = x.dims()
    if rep == 0:
        x.free()
chi = dims[1:-1]
# algorithmic HBM bytes: per step j read C twice (Gram, projection) and write B once
byt = 0
rprev = 1
for jj in range(1, n):
    qq = 2 * rprev; Nn = 2 ** (n - jj)
    byt += 3 * 8 * qq * Nn
    rprev = chi[jj - 1]
cores = x.download(); x.free()
# check random grid points against the analytic samples
idx = rng.integers(0, N, size=200)
vals = []
for i in idx:
    bits = [(int(i) >> (n - 1 - k)) & 1 for k in range(n)]
    m = np.ones((1, 1))
    for k, c in enumerate(cores):
        m = m @ c[:, bits[k], :]
    vals.append(m[0, 0])
ref = v[torch.as_tensor(idx, device=dev)].cpu().numpy()
peak = 6532.0
print(json.dumps({"n": n, "terms": nterms, "seconds": dt, "maxlinkdim": int(max(chi)), "maxtruncerr": info["maxtruncerr"],
                  "algorithmic_GB": byt / 1e9, "GBps": byt / 1e9 / dt, "frac_of_hbm_peak_6532": byt / 1e9 / dt / peak,
                  "max_abs_err_at_200_points": float(np.abs(np.array(vals) - ref).max()), "max_abs_value": float(np.abs(ref).max())}))

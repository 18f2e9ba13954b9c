"""Split (truncated SVD) A/B probe: timing + accuracy on flat / graded / steep / rank-deficient two-site tensors.
Run once per variant (the environment switches are read at the first call):
  python tools/probe_split_ab.py [chi ...]                 # default path
  QTT_JACOBI=cyclic python tools/probe_split_ab.py [chi ...]   # round-1 block-cyclic row kernel
Output: JSON lines."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q

chis = [int(v) for v in sys.argv[1:]] or [256]
rng = np.random.default_rng(0)
ctx = q.default_context(0)
if os.environ.get("QTT_TRACE"):
    ctx.set_profiling(True)
for chi in chis:
    m = 2 * chi
    U = np.linalg.qr(rng.standard_normal((m, m)))[0]; V = np.linalg.qr(rng.standard_normal((m, m)))[0]
    cases = {"random": np.abs(rng.standard_normal(m)) + 0.1, "graded": 10.0 ** (-np.arange(m) * 0.07), "steep": 10.0 ** (-np.arange(m) * 0.3),
             "rank2": np.r_[1.0, 0.5, np.zeros(m - 2)], "noise_floor": np.r_[10.0 ** (-np.arange(40) * 0.2), 1e-9 * (np.abs(rng.standard_normal(m - 40)) + 0.1)]}
    for name, s in cases.items():
        phi = ((U * s) @ V.T).reshape(chi, 2, 2, chi)
        Z = phi.reshape(m, m)
        sref = np.linalg.svd(Z, compute_uv=False)
        for ortho in ("left", "right"):
            r = q.split2(phi, ortho=ortho, cutoff=0.0, maxdim=chi, mindim=chi, reps=3)
            k = r["rank"]
            A = r["A"].reshape(m, k); B = r["B"].reshape(k, m)
            rec = np.linalg.norm(A @ B - Z) / np.linalg.norm(Z)
            best = np.sqrt(np.sum(sref[k:] ** 2)) / np.linalg.norm(Z)
            iso = (A.T @ A if ortho == "left" else B @ B.T) - np.eye(k)
            sg = r["sigma"]
            print(json.dumps({"chi": chi, "case": name, "ortho": ortho, "ms": round(r["ms"], 4), "sweeps": r["sweeps"], "rank": k,
                              "recon_err": rec, "best_rank_k_err": best, "isometry_err": float(np.max(np.abs(iso))),
                              "sigma_abs_err_rel_s1": float(np.max(np.abs(sg - sref[:k])) / sref[0])}), flush=True)

"""Random-shape stress of the site split (QR + Jacobi paths) against LAPACK: isometry, reconstruction, singular values, kept rank.
usage: python tools/fuzz_split.py [ncases seed]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
import oracle as O
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {"iso": 0.0, "rec": 0.0, "sig": 0.0}
bad = 0
for case in range(ncases):
    a = int(rng.choice([17, 32, 33, 48, 64, 65, 96, 100, 127, 128, 160, 200, 255, 256, 300]))
    c = int(rng.choice([17, 32, 33, 48, 64, 65, 96, 100, 127, 128, 160, 200, 255, 256, 300]))
    m, n = 2 * a, 2 * c
    k = min(m, n)
    kind = rng.integers(0, 4)
    U = np.linalg.qr(rng.standard_normal((m, k)))[0]; V = np.linalg.qr(rng.standard_normal((n, k)))[0]
    if kind == 0: s = np.abs(rng.standard_normal(k)) + 0.1                       # flat
    elif kind == 1: s = 10.0 ** (-np.arange(k) * rng.uniform(0.02, 0.3))         # graded
    elif kind == 2:                                                              # rank deficient
        r = int(rng.integers(1, k)); s = np.r_[10.0 ** (-np.arange(r) * 0.05), np.zeros(k - r)]
    else:                                                                        # clustered / repeated values
        s = np.repeat(10.0 ** (-np.arange((k + 3) // 4) * 0.1), 4)[:k]
    phi = ((U * s) @ V.T).reshape(a, 2, 2, c)
    maxdim = int(rng.choice([k, max(1, k // 2), max(1, k // 3)]))
    mindim = int(rng.choice([1, maxdim]))
    cutoff = float(rng.choice([0.0, 1e-12, 1e-8]))
    ortho = "left" if rng.integers(0, 2) else "right"
    try:
        r = q.split2(phi, ortho=ortho, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
    except Exception as e:
        print("CASE", case, (a, c, kind, maxdim, mindim, cutoff, ortho), "FAILED:", str(e)[:120]); bad += 1; continue
    kk = r["rank"]
    _, S, _, err = O.svd_truncated(phi.reshape(m, n), cutoff=cutoff, maxdim=maxdim, mindim=mindim)
    A = r["A"].reshape(m, kk); B = r["B"].reshape(kk, n)
    iso = np.max(np.abs((A.T @ A if ortho == "left" else B @ B.T) - np.eye(kk)))
    rec2 = np.linalg.norm(A @ B - phi.reshape(m, n)) ** 2 / np.linalg.norm(phi) ** 2
    sig = np.max(np.abs(np.asarray(r["sigma"])[:min(kk, len(S))] - S[:min(kk, len(S))])) / S[0]
    ok = (kk == len(S)) and iso < 1e-12 and abs(rec2 - err) < 1e-6 * err + 1e-24 and sig < 1e-12
    worst["iso"] = max(worst["iso"], iso); worst["sig"] = max(worst["sig"], sig); worst["rec"] = max(worst["rec"], abs(rec2 - err))
    if not ok:
        bad += 1
        print("CASE", case, (a, c, kind, maxdim, mindim, cutoff, ortho), "k", kk, "want", len(S), "iso %.1e" % iso, "rec2 %.3e err %.3e" % (rec2, err), "sig %.1e" % sig)
print("cases", ncases, "bad", bad, "worst", {k: "%.2e" % v for k, v in worst.items()})

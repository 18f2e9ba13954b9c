import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
rng = np.random.default_rng(1)
for (a, c, rank) in [(128, 32, 40), (128, 32, 10), (128, 32, 64), (64, 64, 50), (256, 64, 100), (64, 32, 30)]:
    m, n = 2 * a, 2 * c
    k = min(m, n)
    U = np.linalg.qr(rng.standard_normal((m, k)))[0]; V = np.linalg.qr(rng.standard_normal((n, k)))[0]
    s = np.r_[10.0 ** (-np.arange(rank) * 0.2), np.zeros(k - rank)]
    phi = ((U * s) @ V.T).reshape(a, 2, 2, c)
    for ortho in ("left", "right"):
        r = q.split2(phi, ortho=ortho, cutoff=0.0, maxdim=256, mindim=256)
        kk = r["rank"]
        A = r["A"].reshape(m, kk); B = r["B"].reshape(kk, n)
        iso = (A.T @ A if ortho == "left" else B @ B.T) - np.eye(kk)
        print(a, c, rank, ortho, "k", kk, "iso %.2e" % np.max(np.abs(iso)), "rec %.2e" % (np.linalg.norm(A @ B - phi.reshape(m, n)) / np.linalg.norm(phi)))

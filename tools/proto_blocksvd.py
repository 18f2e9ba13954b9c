"""numpy prototype of the blocked one-sided Jacobi (Gram inner solve) used to fix the design of csrc/svd_block.cu.
Counts outer sweeps for flat / graded / steep / rank-deficient 512 x 512 inputs."""
import sys
import numpy as np

def rot(alpha, beta, gamma):
    d = beta - alpha; g2 = 2 * gamma
    h = np.hypot(d, g2)
    t = g2 / (d + np.copysign(h, d))
    c = 1 / np.sqrt(1 + t * t)
    return c, c * t

def inner(G, pairs_steps, tol, thr):
    """two-sided Jacobi on G over the given steps (list of (I, Jx) index arrays of disjoint pairs); returns J, nrot."""
    n = G.shape[0]
    J = np.eye(n)
    nrot = 0
    for I, K in pairs_steps:
        a = G[I, I]; b = G[K, K]; g = G[I, K]
        do = (a > thr) & (b > thr) & (g * g > tol * tol * a * b)
        if not do.any():
            continue
        nrot += int(do.sum())
        c = np.ones(len(I)); s = np.zeros(len(I))
        cc, ss = rot(a[do], b[do], g[do])
        c[do] = cc; s[do] = ss
        # rows: G[I,:], G[K,:]
        gi = G[I, :].copy(); gk = G[K, :].copy()
        G[I, :] = c[:, None] * gi - s[:, None] * gk
        G[K, :] = s[:, None] * gi + c[:, None] * gk
        gi = G[:, I].copy(); gk = G[:, K].copy()
        G[:, I] = c[None, :] * gi - s[None, :] * gk
        G[:, K] = s[None, :] * gi + c[None, :] * gk
        ji = J[:, I].copy(); jk = J[:, K].copy()
        J[:, I] = c[None, :] * ji - s[None, :] * jk
        J[:, K] = s[None, :] * ji + c[None, :] * jk
    return J, nrot

def rr_steps(n):
    """round-robin: n-1 steps of n/2 disjoint pairs"""
    steps = []
    for r in range(n - 1):
        I = []; K = []
        for i in range(n // 2):
            if i == 0: a, b = r, n - 1
            else: a, b = (r + i) % (n - 1), (r - i) % (n - 1)
            I.append(min(a, b)); K.append(max(a, b))
        steps.append((np.array(I), np.array(K)))
    return steps

def cross_steps(b):
    return [(np.arange(b), b + (np.arange(b) + t) % b) for t in range(b)]

def block_jacobi(Y, ndot, b=32, tol=None, thr=0.0, max_sweeps=40, full_inner=False, verbose=False):
    r = Y.shape[0]
    nblk = (r + b - 1) // b
    if nblk % 2: nblk += 1
    rp = nblk * b
    Yp = np.zeros((rp, Y.shape[1])); Yp[:r] = Y
    intra = rr_steps(b); cross = cross_steps(b)
    full = rr_steps(2 * b)
    sweeps = 0
    for sw in range(max_sweeps):
        nrot = 0
        if not full_inner:
            for B in range(nblk):
                sl = slice(B * b, (B + 1) * b)
                G = Yp[sl, :ndot] @ Yp[sl, :ndot].T
                J, k = inner(G, intra, tol, thr); nrot += k
                if k: Yp[sl] = J.T @ Yp[sl]
        for R in range(nblk - 1):
            for i in range(nblk // 2):
                if i == 0: bp, bm = R, nblk - 1
                else: bp, bm = (R + i) % (nblk - 1), (R - i) % (nblk - 1)
                idx = np.r_[bp * b:(bp + 1) * b, bm * b:(bm + 1) * b]
                G = Yp[idx, :ndot] @ Yp[idx, :ndot].T
                J, k = inner(G, full if full_inner else cross, tol, thr); nrot += k
                if k: Yp[idx] = J.T @ Yp[idx]
        sweeps = sw + 1
        if verbose: print("  sweep", sweeps, "rotations", nrot)
        if nrot == 0: break
    return Yp[:r], sweeps

def scalar_jacobi_sweeps(Y, ndot, tol, thr, max_sweeps=40):
    r = Y.shape[0]; Y = Y.copy()
    steps = rr_steps(r + (r % 2))
    for sw in range(max_sweeps):
        nrot = 0
        for I, K in steps:
            ok = K < r; I = I[ok]; K = K[ok]
            a = np.einsum('ij,ij->i', Y[I, :ndot], Y[I, :ndot]); bb = np.einsum('ij,ij->i', Y[K, :ndot], Y[K, :ndot])
            g = np.einsum('ij,ij->i', Y[I, :ndot], Y[K, :ndot])
            do = (a > thr) & (bb > thr) & (g * g > tol * tol * a * bb)
            if not do.any(): continue
            nrot += int(do.sum())
            c, s = rot(a[do], bb[do], g[do])
            yi = Y[I[do]].copy(); yk = Y[K[do]].copy()
            Y[I[do]] = c[:, None] * yi - s[:, None] * yk
            Y[K[do]] = s[:, None] * yi + c[:, None] * yk
        if nrot == 0: return sw + 1
    return max_sweeps

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    for name, decay in (("random", 0.0), ("graded", 0.07), ("steep", 0.3)):
        U = np.linalg.qr(rng.standard_normal((m, m)))[0]; V = np.linalg.qr(rng.standard_normal((m, m)))[0]
        s = 10.0 ** (-np.arange(m) * decay) if decay else np.abs(rng.standard_normal(m)) + 0.1
        Z = (U * s) @ V.T
        # QR with column pivoting of Z^T as the preconditioner (stand-in for the rank-revealing panel QR)
        import scipy.linalg as sla
        Q, R, piv = sla.qr(Z.T, pivoting=True)
        tolr = np.sqrt(m) * 1.11e-16
        thr = (8 * tolr) ** 2 * np.sum(Z * Z)
        rank = int(np.sum(np.abs(np.diag(R)) ** 2 > thr))
        Rr = np.zeros((rank, m)); Rr[:, piv] = R[:rank]
        Y = np.hstack([Rr, Q[:, :rank].T])
        for full_inner in (False, True):
            Yc, sw = block_jacobi(Y, m, b=b, tol=tolr, thr=thr, full_inner=full_inner)
            sig = np.sort(np.sqrt(np.sum(Yc[:, :m] ** 2, axis=1)))[::-1]
            sref = np.linalg.svd(Z, compute_uv=False)[:rank]
            W = Yc[:, m:]
            print(name, "rank", rank, "full_inner" if full_inner else "cross+intra", "sweeps", sw, "max rel sigma err (top half)",
                  np.max(np.abs(sig[:rank // 2] - sref[:rank // 2]) / sref[:rank // 2]), "abs err/s1", np.max(np.abs(sig - sref)) / sref[0],
                  "orth", np.max(np.abs(W @ W.T - np.eye(rank))))
        if m <= 256:
            print(name, "scalar sweeps", scalar_jacobi_sweeps(Y, m, tolr, thr))

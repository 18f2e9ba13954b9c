"""Oracle two-site sweep engine: environments, effective-operator matvec, local solvers
(GMRES / dense QR / dense eigen / Krylov eigensolver), truncated-SVD site split, and the front-ends
`linsolve`, `linsolve_pseudoinverse`, `b_linsolve`, `dmrg_target`, `eigsolve_target`.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED against Julia for this file
(no reference test exists and Julia/ITensorMPS/KrylovKit are absent); anchored on dense solves and
analytic solutions.  Citations are relative to /root/reference; "App. B.x" = SURVEY.md.

Tensor conventions: L[a, w, a'] (ket link, MPO link, bra link), R[c, w, c'], two-site tensor
phi[a, s1, s2, c], MPO core W[wl, wr, s_in, s_out].
"""
import time
import numpy as np
import scipy.linalg as sla

from .core import svd_truncated, orthogonalize, mps_norm

__all__ = [
    "env_left_update", "env_right_update", "benv_left_update", "benv_right_update", "matvec2",
    "effective_operator_dense", "project_rhs", "gmres_krylovkit", "eigsolve_krylovkit", "split_two_site",
    "two_site_sweeps", "linsolve", "linsolve_pseudoinverse", "b_linsolve", "dmrg_target", "eigsolve_target",
    "dmrg_target_krylov", "matvec2_flops", "env_flops", "SweepRecord",
]


# ------------------------------------------------------------------ environments (App. B.2; src/linsolve.jl:200-246)
def env_left_update(L, x, W, y=None):
    """makeL!: L_new[b, w', b'] = L[a, w, a'] x[a, s, b] W[w, w', s, s'] conj(y[a', s', b'])
    (y = x for the Galerkin ProjMPO; y = b' for ProjMPOLinsolve, src/linsolve.jl:215)."""
    y = x if y is None else y
    T = np.tensordot(L, x, axes=([0], [0]))                  # (w, a', s, b)
    T = np.tensordot(T, W, axes=([0, 2], [0, 2]))            # (a', b, w', s')
    T = np.tensordot(T, y.conj(), axes=([0, 3], [0, 1]))     # (b, w', b')
    return T


def env_right_update(R, x, W, y=None):
    """makeR!: R_new[a, w, a'] = x[a, s, c] W[w, w', s, s'] conj(y[a', s', c']) R[c, w', c']
    (src/linsolve.jl:240)."""
    y = x if y is None else y
    T = np.tensordot(x, R, axes=([2], [0]))                  # (a, s, w', c')
    T = np.tensordot(T, W, axes=([1, 2], [2, 1]))            # (a, c', w, s')
    T = np.tensordot(T, y.conj(), axes=([1, 3], [2, 1]))     # (a, w, a')
    return T


def benv_left_update(Lb, x, b):
    """ProjMPS makeL!: Lb_new[c_x, c_b] = Lb[a_x, a_b] x[a_x, s, c_x] conj(b[a_b, s, c_b]) (App. B.2)."""
    T = np.tensordot(Lb, x, axes=([0], [0]))                 # (a_b, s, c_x)
    return np.tensordot(T, b.conj(), axes=([0, 1], [0, 1]))  # (c_x, c_b)


def benv_right_update(Rb, x, b):
    T = np.tensordot(x, Rb, axes=([2], [0]))                 # (a_x, s, c_b)
    return np.tensordot(T, b.conj(), axes=([1, 2], [1, 2]))  # (a_x, a_b)


def matvec2(L, W1, W2, R, v):
    """`product(P, v)` = noprime(v*L*W_j*W_{j+1}*R) (App. B.2; pattern at src/linsolve.jl:71-75):
    out[a', t1, t2, c'] = L[a, wl, a'] W1[wl, wm, s1, t1] W2[wm, wr, s2, t2] R[c, wr, c'] v[a, s1, s2, c]."""
    T = np.tensordot(L, v, axes=([0], [0]))                  # (wl, a', s1, s2, c)
    T = np.tensordot(T, W1, axes=([0, 2], [0, 2]))           # (a', s2, c, wm, t1)
    T = np.tensordot(T, W2, axes=([3, 1], [0, 2]))           # (a', c, t1, wr, t2)
    T = np.tensordot(T, R, axes=([1, 3], [0, 1]))            # (a', t1, t2, c')
    return T


def effective_operator_dense(L, W1, W2, R):
    """`contract(PH, ITensor(1.0))` (src/dmrg_target.jl:3; src/linsolve.jl:71-75): dense matrix
    T[(a', t1, t2, c'), (a, s1, s2, c)]."""
    T = np.einsum("awb,wmst,mruv,crd->btvdasuc", L, W1, W2, R, optimize=True)
    d = T.shape[0] * 4 * T.shape[3]
    return T.reshape(d, -1)


def project_rhs(Lb, b1, b2, Rb):
    """`proj(P::ProjMPS)` src/linsolve.jl:13-22: dag(Lb * dag(b_j) * dag(b_{j+1}) * Rb):
    beff[a, s1, s2, c] = conj(Lb)[a, ab] b1[ab, s1, m] b2[m, s2, cb] conj(Rb)[c, cb]."""
    T = np.tensordot(Lb.conj(), b1, axes=([1], [0]))         # (a, s1, m)
    T = np.tensordot(T, b2, axes=([2], [0]))                 # (a, s1, s2, cb)
    return np.tensordot(T, Rb.conj(), axes=([3], [1]))       # (a, s1, s2, c)


def matvec2_flops(chil, chir, wl, wm, wr, d=2):
    """F_mv, real FP64."""
    return 2 * (chil * (wl * chil) * (d * d * chir) + (wl * d) * (chil * d * chir) * (wm * d)
                + (wm * d) * (chil * d * chir) * (wr * d) + (wr * chir) * (chil * d * d) * chir)


def env_flops(chi, chi2, w, w2, d=2):
    """F_env."""
    return 2 * (chi * (w * chi) * (d * chi2) + (w * d) * (chi * chi2) * (w2 * d) + (chi * d) * (w2 * chi2) * chi2)


# ------------------------------------------------------------------ KrylovKit.linsolve -> GMRES (App. B.5)
def _givens(f, g):
    """LinearAlgebra.givensAlgorithm (real or complex): returns (c, s, r) with
    [c s; -conj(s) c] [f; g] = [r; 0]."""
    if g == 0:
        return 1.0, 0.0 * g, f
    if f == 0:
        return 0.0, np.conj(g) / abs(g) if np.iscomplexobj(g) else np.sign(g), abs(g)
    r = np.hypot(abs(f), abs(g))
    c = abs(f) / r
    sgn = f / abs(f)
    s = sgn * np.conj(g) / r
    return c, s, sgn * r


def gmres_krylovkit(matvec, b, x0, a0=0.0, a1=1.0, tol=1e-12, krylovdim=30, maxiter=100, fixed_matvecs=None):
    """Restarted GMRES as in `KrylovKit.linsolve(f, b, x0, GMRES(...), a0, a1)` (KrylovKit 0.5-0.7,
    App. B.5): solves (a0 + a1 f) x = b.  Arnoldi with ModifiedGramSchmidt2 (KrylovDefaults.orth),
    shift folded into the Hessenberg, Givens rotations, absolute tolerance on the residual norm,
    explicit residual recomputation only when the estimate signals convergence.

    `fixed_matvecs` (benchmark mode, SURVEY.md section 7 'fixed-work'): run exactly one restart
    cycle of that many Arnoldi steps, ignoring tol (krylovdim = fixed_matvecs, maxiter = 1); the
    operator is applied 1 + fixed_matvecs times.
    Returns (x, info) with info = dict(converged, normres, numiter, numops)."""
    if fixed_matvecs is not None:
        # a Krylov space cannot exceed the dimension of the local problem
        krylovdim, maxiter, tol = int(min(fixed_matvecs, b.size)), 1, -1.0
    shape = b.shape
    bv = b.reshape(-1)
    x = x0.reshape(-1).copy()
    mv = lambda u: matvec(u.reshape(shape)).reshape(-1)
    y0 = mv(x)
    dt = np.result_type(bv, y0, type(a0), type(a1))
    r = bv.astype(dt) - a0 * x - a1 * y0
    x = x.astype(dt)
    beta = np.linalg.norm(r)
    numops = 1
    if beta < tol:
        return x.reshape(shape), dict(converged=1, normres=beta, numiter=0, numops=numops)
    numiter = 0
    while numiter < maxiter:
        numiter += 1
        # --- Arnoldi initialize (applies the operator once)
        V = [r / beta]
        w = mv(V[0]); numops += 1
        h = [np.vdot(V[0], w)]
        w = w - h[0] * V[0]
        dh = np.vdot(V[0], w); h[0] += dh; w = w - dh * V[0]
        nres = np.linalg.norm(w)
        H = np.zeros((krylovdim + 1, krylovdim), dtype=dt)
        H[0, 0] = h[0]
        Rm = np.zeros((krylovdim, krylovdim), dtype=dt)
        y = np.zeros(krylovdim + 1, dtype=dt)
        gs = []
        y[0] = beta
        k = 1
        Rm[0, 0] = a0 + a1 * H[0, 0]
        c, s, rr = _givens(Rm[0, 0], a1 * nres)
        gs.append((c, s)); Rm[0, 0] = rr
        y[1] = 0.0
        y[0], y[1] = c * y[0] + s * y[1], -np.conj(s) * y[0] + c * y[1]
        beta = abs(y[1])
        while beta > tol and k < krylovdim:
            # --- expand!: normalise residual, apply operator, MGS2
            V.append(w / nres)
            H[k, k - 1] = nres
            w = mv(V[k]); numops += 1
            hk = np.zeros(k + 1, dtype=dt)
            for i in range(k + 1):
                sdot = np.vdot(V[i], w); w = w - sdot * V[i]; hk[i] = sdot
            for i in range(k + 1):
                sdot = np.vdot(V[i], w); w = w - sdot * V[i]; hk[i] += sdot
            nres = np.linalg.norm(w)
            k += 1
            H[:k, k - 1] = hk
            Rm[:k - 1, k - 1] = a1 * H[:k - 1, k - 1]
            Rm[k - 1, k - 1] = a0 + a1 * H[k - 1, k - 1]
            for i in range(k - 1):
                ci, si = gs[i]
                t1, t2 = Rm[i, k - 1], Rm[i + 1, k - 1]
                Rm[i, k - 1], Rm[i + 1, k - 1] = ci * t1 + si * t2, -np.conj(si) * t1 + ci * t2
            c, s, rr = _givens(Rm[k - 1, k - 1], a1 * nres)
            gs.append((c, s)); Rm[k - 1, k - 1] = rr
            y[k] = 0.0
            y[k - 1], y[k] = c * y[k - 1] + s * y[k], -np.conj(s) * y[k - 1] + c * y[k]
            beta = abs(y[k])
        # --- solve the triangular system and update x
        yk = sla.solve_triangular(Rm[:k, :k], y[:k], lower=False)
        for i in range(k):
            x = x + yk[i] * V[i]
        if beta > tol:
            # residual without re-evaluating the operator: r = y[k+1] * (G_1^H ... G_k^H e_{k+1}) in basis [V, w/nres]
            Vext = V + [w / nres]
            coef = np.zeros(k + 1, dtype=dt); coef[k] = 1.0
            for i in range(k - 1, -1, -1):
                ci, si = gs[i]
                t1, t2 = coef[i], coef[i + 1]
                # apply G_i^H = [c -s; conj(s) c]
                coef[i], coef[i + 1] = ci * t1 - si * t2, np.conj(si) * t1 + ci * t2
            r = sum(coef[i] * Vext[i] for i in range(k + 1)) * y[k]
        else:
            r = bv - (a0 * x + a1 * mv(x)); numops += 1
            beta = np.linalg.norm(r)
            if beta < tol:
                return x.reshape(shape), dict(converged=1, normres=beta, numiter=numiter, numops=numops)
    return x.reshape(shape), dict(converged=0, normres=beta, numiter=numiter, numops=numops)


# ------------------------------------------------------------------ KrylovKit.eigsolve (Hermitian), App. B.5
def eigsolve_krylovkit(matvec, x0, which="LM", target=None, tol=1e-12, krylovdim=30, maxiter=100):
    """Krylov-Schur style restarted Lanczos with full orthogonalisation, restating
    `KrylovKit.eigsolve(f, x0, 1, which, Lanczos(...))` for a Hermitian operator: expand to
    krylovdim, Ritz-decompose the projected matrix, converged when |beta_k * u_k[last]| < tol,
    otherwise thick-restart keeping div(3*krylovdim + 2*converged, 5) Ritz vectors.
    `which`: "LM" (largest magnitude), "SR", "LR", or "target" (closest to `target`).
    Returns (value, vector, info)."""
    shape = x0.shape
    mv = lambda u: matvec(u.reshape(shape)).reshape(-1)
    v = x0.reshape(-1)
    v = v / np.linalg.norm(v)
    w_first = mv(v)
    dt = np.result_type(v, w_first)
    numops = 0

    def order(D):
        if which == "LM":
            return np.argsort(-np.abs(D), kind="stable")
        if which == "SR":
            return np.argsort(D, kind="stable")
        if which == "LR":
            return np.argsort(-D, kind="stable")
        return np.argsort(np.abs(D - target), kind="stable")

    V = [v.astype(dt)]
    Hm = np.zeros((krylovdim + 1, krylovdim + 1), dtype=dt)
    k = 0
    numiter = 0
    beta = 0.0
    w = None
    krylovdim = int(min(krylovdim, v.size))
    invariant = False
    while True:
        # expand until krylovdim
        while k < krylovdim:
            if w_first is not None:
                w, w_first = w_first, None
            else:
                w = mv(V[k])
            numops += 1
            for _ in range(2):
                for i in range(k + 1):
                    sdot = np.vdot(V[i], w); w = w - sdot * V[i]; Hm[i, k] += sdot
            beta = np.linalg.norm(w)
            k += 1
            # invariant subspace (KrylovKit stops expanding when normres <= tol); also guards w / beta
            if beta <= max(tol, 1e-13 * np.linalg.norm(Hm[:k, k - 1])):
                invariant = True
                break
            if k < krylovdim:
                V.append(w / beta); Hm[k, k - 1] = beta
        numiter += 1
        Hk = Hm[:k, :k]
        Hk = 0.5 * (Hk + Hk.conj().T)
        D, U = np.linalg.eigh(Hk)
        p = order(D)
        D = D[p]; U = U[:, p]
        f = beta * U[k - 1, :]
        converged = 0
        while converged < k and abs(f[converged]) < tol:
            converged += 1
        if converged >= 1 or numiter >= maxiter or invariant:
            vec = sum(U[i, 0] * V[i] for i in range(k))
            return D[0], vec.reshape(shape), dict(converged=converged, normres=abs(f[0]), numiter=numiter, numops=numops)
        keep = (3 * krylovdim + 2 * converged) // 5
        Vnew = [sum(U[i, j] * V[i] for i in range(k)) for j in range(keep)]
        Vnew.append(w / beta)
        V = Vnew
        Hm[:] = 0
        for j in range(keep):
            Hm[j, j] = D[j]
            Hm[keep, j] = f[j]      # upper part Hm[j, keep] is produced by the next orthogonalisation
        k = keep


# ------------------------------------------------------------------ site split (replacebond!, App. B.2)
def split_two_site(phi, ortho, cutoff=0.0, maxdim=None, mindim=1):
    """`replacebond!(psi, b, phi; ortho, cutoff, maxdim, mindim)`: SVD of phi[(a s1), (s2 c)],
    truncation rule App. B.3; ortho "left": (U, S V); "right": (U S, V).  Returns (A, B, truncerr, S)."""
    a, _, _, c = phi.shape
    U, S, Vh, err = svd_truncated(phi.reshape(a * 2, 2 * c), cutoff, maxdim, mindim)
    k = S.size
    if ortho == "left":
        A = U.reshape(a, 2, k)
        B = (S[:, None] * Vh).reshape(k, 2, c)
    else:
        A = (U * S[None, :]).reshape(a, 2, k)
        B = Vh.reshape(k, 2, c)
    return A, B, err, S


class SweepRecord(dict):
    """Per-sweep log mirroring the `tdvp` output line (App. B.1): maxlinkdim, maxtruncerr, seconds."""


def _per_sweep(v, nsweeps, default):
    if v is None:
        v = default
    if np.isscalar(v):
        return [v] * nsweeps
    v = list(v)
    return [v[min(i, len(v) - 1)] for i in range(nsweeps)]


def two_site_sweeps(A, x0, solver, nsweeps, maxdim=None, mindim=1, cutoff=1e-16, b=None, normalize=False,
                    bra=None, timers=None, bond_window=None):
    """`tdvp(solver, P, t=Inf, x0; reverse_step=false, nsite=2, nsweeps, maxdim, mindim, cutoff)`
    (App. B.1): per sweep a forward half-sweep (bonds 1..n-1, ortho "left") then a reverse
    half-sweep (bonds n-1..1, ortho "right"); per bond position!, phi = x[b]*x[b+1],
    phi = solver(...), replacebond!.

    solver(env, phi) -> phi', where env = dict(L, W1, W2, R, Lb, Rb, b1, b2, bond, n).
    `b`  : optional MPS for the ProjMPS overlap environments (ProjMPO_MPS(A, [b])).
    `bra`: optional MPS used as the bra of the operator environments (ProjMPOLinsolve, src/linsolve.jl:175-253).
    `bond_window=(j0, m)` (bench sampling only): process just the m forward bond updates j0..j0+m-1 of the
    first half-sweep, starting from x0 brought to orthogonality centre j0 with the environments it needs.
    Returns (x, records)."""
    n = len(A)
    assert len(x0) == n and n >= 2
    if bond_window is not None:
        return _window_sweep(A, x0, solver, bond_window, maxdim, mindim, cutoff, b, timers)
    maxdims = _per_sweep(maxdim, nsweeps, None)
    mindims = _per_sweep(mindim, nsweeps, 1)
    cutoffs = _per_sweep(cutoff, nsweeps, 1e-16)
    x = orthogonalize(x0, 0)
    dt = np.result_type(x[0], A[0]) if b is None else np.result_type(x[0], A[0], b[0])
    x = [c.astype(dt) for c in x]
    brav = None if bra is None else [c.copy() for c in bra]
    one3 = np.ones((1, 1, 1), dtype=dt)
    one2 = np.ones((1, 1), dtype=dt)
    LA = [None] * (n + 1); RA = [None] * (n + 1)
    LB = [None] * (n + 1); RB = [None] * (n + 1)
    LA[0] = one3; RA[n] = one3; LB[0] = one2; RB[n] = one2
    tm = timers if timers is not None else {}
    for key in ("env", "solver", "svd"):
        tm.setdefault(key, 0.0)

    def ybra(j):
        return None if brav is None else brav[j]

    t0 = time.perf_counter()
    for j in range(n - 1, 1, -1):
        RA[j] = env_right_update(RA[j + 1], x[j], A[j], ybra(j))
        if b is not None:
            RB[j] = benv_right_update(RB[j + 1], x[j], b[j])
    tm["env"] += time.perf_counter() - t0
    records = []
    for sw in range(nsweeps):
        tsw = time.perf_counter()
        maxerr = 0.0
        for direction in ("forward", "reverse"):
            bonds = range(0, n - 1) if direction == "forward" else range(n - 2, -1, -1)
            ortho = "left" if direction == "forward" else "right"
            for j in bonds:
                if brav is not None:
                    # ProjMPOLinsolve.position! re-orthogonalises b to the bond (src/linsolve.jl:248-253);
                    # changes of gauge in b change L/R built from it, so rebuild what depends on it.
                    brav = orthogonalize(brav, j)
                    t0 = time.perf_counter()
                    for i in range(j):
                        LA[i + 1] = env_left_update(LA[i], x[i], A[i], brav[i])
                    for i in range(n - 1, j + 1, -1):
                        RA[i] = env_right_update(RA[i + 1], x[i], A[i], brav[i])
                    tm["env"] += time.perf_counter() - t0
                phi = np.tensordot(x[j], x[j + 1], axes=([2], [0]))
                env = dict(L=LA[j], W1=A[j], W2=A[j + 1], R=RA[j + 2], bond=j, n=n,
                           Lb=LB[j] if b is not None else None, Rb=RB[j + 2] if b is not None else None,
                           b1=b[j] if b is not None else None, b2=b[j + 1] if b is not None else None,
                           bra1=None if brav is None else brav[j], bra2=None if brav is None else brav[j + 1])
                t0 = time.perf_counter()
                phi = solver(env, phi)
                tm["solver"] += time.perf_counter() - t0
                if normalize:
                    phi = phi / np.linalg.norm(phi)
                t0 = time.perf_counter()
                Aj, Bj, err, _ = split_two_site(phi, ortho, cutoffs[sw], maxdims[sw], mindims[sw])
                tm["svd"] += time.perf_counter() - t0
                maxerr = max(maxerr, err)
                x[j] = Aj; x[j + 1] = Bj
                t0 = time.perf_counter()
                if direction == "forward":
                    LA[j + 1] = env_left_update(LA[j], x[j], A[j], ybra(j))
                    if b is not None:
                        LB[j + 1] = benv_left_update(LB[j], x[j], b[j])
                else:
                    RA[j + 1] = env_right_update(RA[j + 2], x[j + 1], A[j + 1], ybra(j + 1))
                    if b is not None:
                        RB[j + 1] = benv_right_update(RB[j + 2], x[j + 1], b[j + 1])
                tm["env"] += time.perf_counter() - t0
        records.append(SweepRecord(sweep=sw + 1, maxlinkdim=max(c.shape[2] for c in x[:-1]),
                                   maxtruncerr=maxerr, seconds=time.perf_counter() - tsw))
    return x, records


def _window_sweep(A, x0, solver, bond_window, maxdim, mindim, cutoff, b, timers):
    n = len(A)
    j0, m = bond_window
    assert 0 <= j0 and j0 + m <= n - 1
    md = _per_sweep(maxdim, 1, None)[0]; mn = _per_sweep(mindim, 1, 1)[0]; co = _per_sweep(cutoff, 1, 1e-16)[0]
    x = orthogonalize(x0, j0)
    dt = np.result_type(x[0], A[0]) if b is None else np.result_type(x[0], A[0], b[0])
    LA = [None] * (n + 1); RA = [None] * (n + 1); LB = [None] * (n + 1); RB = [None] * (n + 1)
    LA[0] = np.ones((1, 1, 1), dtype=dt); RA[n] = np.ones((1, 1, 1), dtype=dt)
    LB[0] = np.ones((1, 1), dtype=dt); RB[n] = np.ones((1, 1), dtype=dt)
    for j in range(j0):
        LA[j + 1] = env_left_update(LA[j], x[j], A[j])
        if b is not None:
            LB[j + 1] = benv_left_update(LB[j], x[j], b[j])
    for j in range(n - 1, j0 + 1, -1):
        RA[j] = env_right_update(RA[j + 1], x[j], A[j])
        if b is not None:
            RB[j] = benv_right_update(RB[j + 1], x[j], b[j])
    tm = timers if timers is not None else {}
    for key in ("env", "solver", "svd"):
        tm.setdefault(key, 0.0)
    t_start = time.perf_counter()
    maxerr = 0.0
    for j in range(j0, j0 + m):
        phi = np.tensordot(x[j], x[j + 1], axes=([2], [0]))
        env = dict(L=LA[j], W1=A[j], W2=A[j + 1], R=RA[j + 2], bond=j, n=n,
                   Lb=LB[j] if b is not None else None, Rb=RB[j + 2] if b is not None else None,
                   b1=b[j] if b is not None else None, b2=b[j + 1] if b is not None else None, bra1=None, bra2=None)
        t0 = time.perf_counter(); phi = solver(env, phi); tm["solver"] += time.perf_counter() - t0
        t0 = time.perf_counter(); Aj, Bj, err, _ = split_two_site(phi, "left", co, md, mn); tm["svd"] += time.perf_counter() - t0
        maxerr = max(maxerr, err)
        x[j] = Aj; x[j + 1] = Bj
        t0 = time.perf_counter()
        LA[j + 1] = env_left_update(LA[j], x[j], A[j])
        if b is not None:
            LB[j + 1] = benv_left_update(LB[j], x[j], b[j])
        tm["env"] += time.perf_counter() - t0
    rec = SweepRecord(sweep=1, maxlinkdim=max(c.shape[2] for c in x[:-1]), maxtruncerr=maxerr,
                      seconds=time.perf_counter() - t_start)
    return x, [rec]


# ------------------------------------------------------------------ front-ends
def linsolve(A, b, x0, a0=0.0, a1=1.0, nsweeps=1, maxdim=None, mindim=1, cutoff=1e-16, tol=1e-5, krylovdim=20,
             maxiter=10, ishermitian=False, normalize=False, fixed_matvecs=None, stats=None, timers=None, bond_window=None):
    """`linsolve(A::MPO, b::MPS, x0::MPS, a0, a1; nsweeps, maxdim, cutoff, tol, krylovdim, maxiter)`:
    body restated from examples/airy_equation/src/linsolve.jl:24-34 (defaults tol=1e-5,
    krylovdim=20, maxiter=10): ProjMPO_MPS(A, [b]); local solver = KrylovKit GMRES on
    (a0 + a1 P) phi = proj(b) starting from the current two-site tensor."""
    st = stats if stats is not None else {}
    st.setdefault("matvecs", 0); st.setdefault("gmres_resid", 0.0)

    def solver(env, phi):
        def mv(v):
            st["matvecs"] += 1
            return matvec2(env["L"], env["W1"], env["W2"], env["R"], v)
        beff = project_rhs(env["Lb"], env["b1"], env["b2"], env["Rb"])
        xloc, info = gmres_krylovkit(mv, beff, phi, a0, a1, tol, krylovdim, maxiter, fixed_matvecs)
        st["gmres_resid"] = max(st["gmres_resid"], float(info["normres"]))
        return xloc

    return two_site_sweeps(A, x0, solver, nsweeps, maxdim, mindim, cutoff, b=b, normalize=normalize, timers=timers,
                           bond_window=bond_window)


def linsolve_pseudoinverse(A, b, x0, a0=0.0, a1=1.0, nsweeps=1, maxdim=None, mindim=1, cutoff=1e-16, **_):
    """src/linsolve.jl:54-97: dense local solve T x = b via QR (`qr!(T_mat) \\ b_vec`, :88-89).
    NOTE the reference ignores a0/a1 in this variant (:67-92); restated as such."""
    def solver(env, phi):
        T = effective_operator_dense(env["L"], env["W1"], env["W2"], env["R"])
        beff = project_rhs(env["Lb"], env["b1"], env["b2"], env["Rb"]).reshape(-1)
        Q, Rr = np.linalg.qr(T)
        xv = sla.solve_triangular(Rr, Q.conj().T @ beff, lower=False)
        return xv.reshape(phi.shape)

    return two_site_sweeps(A, x0, solver, nsweeps, maxdim, mindim, cutoff, b=b)


def b_linsolve(A, b, x0, a0=0.0, a1=1.0, nsweeps=1, maxdim=None, mindim=1, cutoff=1e-16, **_):
    """src/linsolve.jl:263-307 (live path = `b_linsolve_exact_solver`, :279-300,305):
    Petrov-Galerkin environments with bra = b; RHS = product of b's two cores (`bvec`, :255-261);
    dense SVD least squares on the rectangular (chi_b^2 d^2) x (chi_x^2 d^2) matrix."""
    def solver(env, phi):
        T = effective_operator_dense(env["L"], env["W1"], env["W2"], env["R"])   # rows = bra (b) frame
        bT = np.tensordot(env["bra1"], env["bra2"], axes=([2], [0])).reshape(-1)
        xv, *_ = np.linalg.lstsq(T, bT, rcond=None)
        return xv.reshape(phi.shape)

    return two_site_sweeps(A, x0, solver, nsweeps, maxdim, mindim, cutoff, bra=b)


def dmrg_target(A, psi0, target_eigenvalue, nsweeps=1, maxdim=None, mindim=1, cutoff=1e-16, normalize=False):
    """src/dmrg_target.jl:1-12: dense effective operator, Hermitian `eigen`, pick the eigenvalue
    nearest `target_eigenvalue` (:6), return that eigenvector (:7)."""
    def solver(env, phi):
        H = effective_operator_dense(env["L"], env["W1"], env["W2"], env["R"])
        H = 0.5 * (H + H.conj().T)
        D, U = np.linalg.eigh(H)
        j = int(np.argmin(np.abs(D - target_eigenvalue)))
        return U[:, j].reshape(phi.shape)

    return two_site_sweeps(A, psi0, solver, nsweeps, maxdim, mindim, cutoff, normalize=normalize)


def eigsolve_target(matvec, lam, x0, linsolve_kwargs=None, eigsolve_kwargs=None):
    """src/eigsolve_target.jl:2-9: eigsolve(x -> linsolve(A, x, x, -lam)) -> eigenvector of A nearest lam."""
    lk = dict(tol=1e-12, krylovdim=30, maxiter=100)
    lk.update(linsolve_kwargs or {})
    ek = dict(tol=1e-12, krylovdim=30, maxiter=100)
    ek.update(eigsolve_kwargs or {})

    def f(x):
        xp, _ = gmres_krylovkit(matvec, x, x, -lam, 1.0, **lk)
        return xp

    val, vec, info = eigsolve_krylovkit(f, x0, which="LM", **ek)
    return vec


def dmrg_target_krylov(A, psi0, target_eigenvalue, nsweeps=1, maxdim=None, mindim=1, cutoff=1e-16, mode="lanczos",
                       tol=1e-12, krylovdim=30, maxiter=3, stats=None):
    """The scalable variant the reference sketches in comments (src/dmrg_target.jl:14-21):
    local Krylov eigen solve instead of dense `eigen`.  mode "lanczos": Ritz value closest to the
    target (valid for extremal targets); mode "shiftinvert": `eigsolve_target` (src/eigsolve_target.jl)."""
    st = stats if stats is not None else {}
    st.setdefault("matvecs", 0)

    def solver(env, phi):
        def mv(v):
            st["matvecs"] += 1
            return matvec2(env["L"], env["W1"], env["W2"], env["R"], v)
        if mode == "lanczos":
            _, vec, _ = eigsolve_krylovkit(mv, phi, which="target", target=target_eigenvalue, tol=tol,
                                           krylovdim=krylovdim, maxiter=maxiter)
            return vec
        return eigsolve_target(mv, target_eigenvalue, phi)

    return two_site_sweeps(A, psi0, solver, nsweeps, maxdim, mindim, cutoff, normalize=True)

import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import itensorqtt_jl_b200 as lib
n, chi, nk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = O.laplacian_mpo(n, 1.0)
lam = -((2 * np.pi * (2.0 ** nk) / 2.0 ** n) ** 2)
b = O.boundary_value_mps(n, 1 / np.sqrt(2), 1 / np.sqrt(2))
x0 = O.random_mps(n, chi, 2)
kw = dict(nsweeps=1, maxdim=chi, mindim=chi, cutoff=0.0, fixed_matvecs=30)
xg = lib.linsolve(A, b, x0, -lam, 1.0, **kw)
np.save(sys.argv[4], np.array(xg, dtype=object), allow_pickle=True)

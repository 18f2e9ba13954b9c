"""CPU oracle for the ITensorQTT sweep-engine hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy *restatement* of the reference algorithm (ITensor/ITensorQTT.jl plus the
ITensorMPS / ITensors / KrylovKit routines it calls).  It exists to check the CUDA path; it is never
the thing shipped or measured.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product package
(`itensorqtt.jl_b200/`) must never import anything from here.

Parity pinning status:
  * operator builders, grid/bit maps, TT-SVD conversion, MPO.MPS apply (DFT, prolongation,
    retraction, Fourier interpolation, repeat): PINNED against every assertion of the reference's
    own test-suite (`/root/reference/test/runtests.jl`), restated in `tests/test_oracle_kat.py`.
  * the two-site sweep engine (`linsolve`, `linsolve_pseudoinverse`, `b_linsolve`, `dmrg_target`,
    `eigsolve_target`): the reference has NO test and the Julia toolchain / ITensorMPS / KrylovKit
    sources are absent from this container, so against Julia this part is "PARITY UNPINNED".
    It is anchored instead on the substitutes the reference's examples use: dense solves of the
    same system (`examples/airy_equation/src/airy_utils.jl:41-57`), analytic solutions
    (`airy_utils.jl:8`, `examples/helmholtz_equation/helmholtz_equation.jl:48`) and the residual
    metric `sqeuclidean_normalized` (`src/mps_utils.jl:109-116`).

Third-party dependencies of the reference whose algorithms are restated here (not vendored under
/root/reference; pinned only by compat ranges in `/root/reference/Project.toml:13-18`):
ITensorMPS 0.2, ITensors 0.3-0.6 (NDTensors), KrylovKit 0.5-0.7.

Conventions (bit-exact contract):
  * MPS core j : ndarray (chi_{j-1}, 2, chi_j)            index order (l, s, r)
  * MPO core j : ndarray (w_{j-1}, w_j, 2, 2)             index order (l, r, s_in, s_out)
    [= `itensor(T, l_j, l_{j+1}, s_j, s_j')`, `/root/reference/src/laplacian.jl:32`]
  * site 1 is the most significant bit; index value 0 <-> bit 0.
"""
from .core import *      # noqa: F401,F403
from .builders import *  # noqa: F401,F403
from .apply import *     # noqa: F401,F403
from .sweep import *     # noqa: F401,F403

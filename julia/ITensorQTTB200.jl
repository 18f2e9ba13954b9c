# ITensorQTTB200.jl -- the reference-side binding of libqtt_b200.so (include/qtt_b200.h).
#
# Drop-in for ONE hot path of ITensorQTT: `linsolve(A, b, x0, a0, a1; nsweeps, maxdim, cutoff, ...)`,
# `dmrg_target`, `apply_dft_mpo` / `apply_idft_mpo`, `prolongate`, `retract` keep their signatures
# (reference examples/airy_equation/src/linsolve.jl:24-34, src/dmrg_target.jl:1-12, src/dft.jl:121-133,
# src/prolongation.jl:29-45,70-90); the two-site sweep / MPO.MPS apply run on a B200 through `ccall`.
# Only dense column-major arrays and shapes cross the boundary: Index identity, tags, prime levels and the
# `reverse` / output-bit relabelling of src/dft.jl:76-79,124 stay on the Julia side.
#
# NOTE: Julia is not installed in the build container, so this file is written against the C header and the
# ITensors API but has not been executed there; the same ABI is exercised from Python `ctypes` by the test-suite.
module ITensorQTTB200

using ITensors
using ITensorQTT: dft_mpo, idft_mpo, prolongation_mpo, siteinds_per_dimension

const libqtt = get(ENV, "LIBQTT_B200", "libqtt_b200.so")

# ---- mirror of the C structs (field order and types as in include/qtt_b200.h) ------------------------------
struct QttSweepParams
  nsweeps::Int32
  maxdim::Ptr{Int32}
  mindim::Ptr{Int32}
  cutoff::Ptr{Float64}
  solver::Int32
  a0::Float64
  a1::Float64
  tol::Float64
  krylovdim::Int32
  maxiter::Int32
  fixed_matvecs::Int32
  normalize::Int32
  outputlevel::Int32
  target::Float64
  svd_max_sweeps::Int32
  x0_canonical::Int32
end

struct QttSweepInfo
  maxlinkdim::Int32
  svd_sweeps_max::Int32
  maxtruncerr::Float64
  seconds::Float64
  solver_resid::Float64
  matvecs::Int64
  flops_matvec::Float64
  t_matvec::Float64
  t_env::Float64
  t_svd::Float64
  t_krylov::Float64
  energy::Float64
  device_ms::Float64
end

const QTT_SOLVER_GMRES = Int32(0)
const QTT_SOLVER_LANCZOS = Int32(1)
const QTT_SOLVER_SHIFT_INVERT = Int32(2)   # `eigsolve_target` (reference src/eigsolve_target.jl:2-9) as the local solver
const QTT_SOLVER_DENSE_QR = Int32(3)       # `linsolve_pseudoinverse` (reference src/linsolve.jl:54-97): dense operator, qr! \ b
const QTT_SOLVER_DENSE_EIGEN = Int32(4)    # `dmrg_target` as written (reference src/dmrg_target.jl:1-12): dense Hermitian eigen
const QTT_SOLVER_PG_LSQ = Int32(5)         # `b_linsolve` (reference src/linsolve.jl:175-307): Petrov-Galerkin sweep, dense SVD least squares

struct QttError <: Exception
  code::Int
  msg::String
end

mutable struct Context
  handle::Ptr{Cvoid}
  function Context(device::Integer=0)
    h = ccall((:qtt_ctx_create, libqtt), Ptr{Cvoid}, (Cint,), device)
    h == C_NULL && throw(QttError(-2, unsafe_string(ccall((:qtt_global_error, libqtt), Cstring, ()))))
    ctx = new(h)
    finalizer(c -> ccall((:qtt_ctx_destroy, libqtt), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
    return ctx
  end
end

const default_ctx = Ref{Union{Nothing,Context}}(nothing)
context() = something(default_ctx[], (default_ctx[] = Context(0)))

function check(ctx::Context, rc::Integer)
  rc == 0 && return nothing
  throw(QttError(rc, unsafe_string(ccall((:qtt_last_error, libqtt), Cstring, (Ptr{Cvoid},), ctx.handle))))
end

# process_sweeps semantics of tdvp: scalars or per-sweep vectors, last value repeated
sweepvec(::Type{T}, v::Number, n) where {T} = fill(T(v), n)
sweepvec(::Type{T}, v::AbstractVector, n) where {T} = T[v[min(i, length(v))] for i in 1:n]

# ---- chains: ITensors -> dense column-major cores ------------------------------------------------------------
# MPS core j as array(x[j], l_{j-1}, s_j, l_j) with size-1 dummy links at the ends
function mps_cores(x::MPS)
  n = length(x)
  cores = Vector{Array}(undef, n)
  chi = ones(Int64, n + 1)
  for j in 1:n
    s = siteind(x, j)
    l = j > 1 ? linkind(x, j - 1) : nothing
    r = j < n ? linkind(x, j) : nothing
    is = filter(!isnothing, (l, s, r))
    a = array(x[j], is...)
    dl = isnothing(l) ? 1 : dim(l)
    dr = isnothing(r) ? 1 : dim(r)
    cores[j] = reshape(a, dl, 2, dr)
    chi[j] = dl
    chi[j + 1] = dr
  end
  return cores, chi
end

# The input (plev 0) and output (plev 1) site index of MPO tensor j, whatever their ids: `dft_mpo` relabels the outputs
# (`reverse_output_bits`, reference src/dft.jl:76-79,98), so tensor j of a DFT MPO carries (s_j, s_{n+1-j}'), and
# `idft_mpo = swapprime(dag(dft_mpo))` carries (s_{n+1-j}, s_j').
function mpo_site_inds(A::MPO, j::Int)
  links = Index[]
  j > 1 && push!(links, linkind(A, j - 1))
  j < length(A) && push!(links, linkind(A, j))
  sites = [i for i in inds(A[j]) if !(i in links)]
  length(sites) == 2 || error("MPO tensor $j: expected two site indices, found $(length(sites))")
  sin = only(filter(i -> plev(i) == 0, sites))
  sout = only(filter(i -> plev(i) == 1, sites))
  return sin, sout
end

# MPO core j as array(A[j], l, r, s_in, s_out)  (index order of reference src/laplacian.jl:32)
function mpo_cores(A::MPO)
  n = length(A)
  cores = Vector{Array}(undef, n)
  w = ones(Int64, n + 1)
  for j in 1:n
    s, sp = mpo_site_inds(A, j)
    l = j > 1 ? linkind(A, j - 1) : nothing
    r = j < n ? linkind(A, j) : nothing
    is = filter(!isnothing, (l, r, s, sp))
    a = array(A[j], is...)
    dl = isnothing(l) ? 1 : dim(l)
    dr = isnothing(r) ? 1 : dim(r)
    cores[j] = reshape(a, dl, dr, 2, 2)
    w[j] = dl
    w[j + 1] = dr
  end
  return cores, w
end

dtype_code(cores) = any(c -> eltype(c) <: Complex, cores) ? Cint(1) : Cint(0)

function upload(ctx::Context, x::MPS)
  cores, chi = mps_cores(x)
  T = dtype_code(cores) == 1 ? ComplexF64 : Float64
  cores = [convert(Array{T,3}, c) for c in cores]
  out = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve cores begin
    ptrs = [Ptr{Cvoid}(pointer(c)) for c in cores]
    check(ctx, ccall((:qtt_mps_upload, libqtt), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Ptr{Cvoid}}),
                     ctx.handle, length(cores), chi, ptrs, dtype_code(cores), out))
  end
  return out[]
end

function upload(ctx::Context, A::MPO)
  cores, w = mpo_cores(A)
  T = dtype_code(cores) == 1 ? ComplexF64 : Float64
  cores = [convert(Array{T,4}, c) for c in cores]
  out = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve cores begin
    ptrs = [Ptr{Cvoid}(pointer(c)) for c in cores]
    check(ctx, ccall((:qtt_mpo_upload, libqtt), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Ptr{Cvoid}}),
                     ctx.handle, length(cores), w, ptrs, dtype_code(cores), out))
  end
  return out[]
end

# device chain -> MPS on the given site indices with fresh link indices
function download(ctx::Context, h::Ptr{Cvoid}, sites::Vector{<:Index})
  n = ccall((:qtt_mps_length, libqtt), Cint, (Ptr{Cvoid},), h)
  chi = Vector{Int64}(undef, n + 1)
  check(ctx, ccall((:qtt_mps_dims, libqtt), Cint, (Ptr{Cvoid}, Ptr{Int64}), h, chi))
  T = ccall((:qtt_mps_dtype, libqtt), Cint, (Ptr{Cvoid},), h) == 1 ? ComplexF64 : Float64
  cores = [Array{T,3}(undef, chi[j], 2, chi[j + 1]) for j in 1:n]
  GC.@preserve cores begin
    ptrs = [Ptr{Cvoid}(pointer(c)) for c in cores]
    check(ctx, ccall((:qtt_mps_download, libqtt), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), h, ptrs))
  end
  links = [Index(chi[j + 1], "Link,l=$j") for j in 1:(n - 1)]
  ts = Vector{ITensor}(undef, n)
  for j in 1:n
    if n == 1
      ts[j] = itensor(reshape(cores[j], 2), sites[j])
    elseif j == 1
      ts[j] = itensor(reshape(cores[j], 2, chi[2]), sites[j], links[1])
    elseif j == n
      ts[j] = itensor(reshape(cores[j], chi[n], 2), links[n - 1], sites[j])
    else
      ts[j] = itensor(cores[j], links[j - 1], sites[j], links[j])
    end
  end
  return MPS(ts)
end

free_mps(h) = ccall((:qtt_mps_free, libqtt), Cvoid, (Ptr{Cvoid},), h)
free_mpo(h) = ccall((:qtt_mpo_free, libqtt), Cvoid, (Ptr{Cvoid},), h)

# ---- the drop-in entry points ---------------------------------------------------------------------------------
"""
    linsolve(A::MPO, b::MPS, x0::MPS, a0=0, a1=1; nsweeps, maxdim, mindim, cutoff, tol, krylovdim, maxiter, ishermitian)

Same signature and defaults as the reference (examples/airy_equation/src/linsolve.jl:24-34): solves (a0 + a1 A) x = b
by two-site sweeps with a local GMRES, on the GPU.
"""
function linsolve(A::MPO, b::MPS, x0::MPS, a0::Number=0, a1::Number=1; nsweeps, maxdim=typemax(Int32), mindim=1,
                  cutoff=1e-16, tol=1e-5, krylovdim=20, maxiter=10, ishermitian=false, normalize=false, outputlevel=0,
                  ctx::Context=context(), local_solver=QTT_SOLVER_GMRES)
  (a0 isa Real && a1 isa Real) || error("libqtt_b200: complex a0/a1 are not supported by this build")
  hA = upload(ctx, A); hb = upload(ctx, b); hx = upload(ctx, x0)
  md = sweepvec(Int32, min.(maxdim, typemax(Int32)), nsweeps)
  mn = sweepvec(Int32, mindim, nsweeps)
  co = sweepvec(Float64, cutoff, nsweeps)
  infos = Vector{QttSweepInfo}(undef, nsweeps)
  try
    GC.@preserve md mn co begin
      p = Ref(QttSweepParams(nsweeps, pointer(md), pointer(mn), pointer(co), local_solver, a0, a1, tol, krylovdim, maxiter,
                             0, normalize, outputlevel, 0.0, 0, 0))
      check(ctx, ccall((:qtt_linsolve, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{QttSweepParams}, Ptr{QttSweepInfo}),
                       ctx.handle, hA, hb, hx, p, infos))
    end
    return download(ctx, hx, siteinds(x0))
  finally
    free_mps(hx); free_mps(hb); free_mpo(hA)
  end
end

"""
    linsolve_pseudoinverse(A::MPO, b::MPS, x0::MPS, a0=0, a1=1; nsweeps, maxdim, mindim, cutoff, kwargs...)

Reference src/linsolve.jl:54-97: the same sweep with the dense local solve `qr!(T_mat) \\ b_vec` (:88-89) on the GPU
(blocked Householder QR on the DMMA GEMM core).  As in the reference body, `a0`, `a1` and the Krylov keywords are
accepted and ignored (:67-92).  Local dimension 4 chi_l chi_r <= 8192; larger bonds raise a QttError.
"""
linsolve_pseudoinverse(A::MPO, b::MPS, x0::MPS, a0::Number=0, a1::Number=1; tol=1e-5, krylovdim=20, maxiter=10,
                       ishermitian=false, reverse_step=false, nsite=2, kwargs...) =
  linsolve(A, b, x0, 0, 1; local_solver=QTT_SOLVER_DENSE_QR, kwargs...)

"""
    b_linsolve(A::MPO, b::MPS, x0::MPS, a0=0, a1=1; nsweeps, maxdim, mindim, cutoff, kwargs...)

Reference src/linsolve.jl:263-307 (live solver `b_linsolve_exact_solver`): Petrov-Galerkin sweep with `b` as the bra of the
operator environments (`ProjMPOLinsolve`, :175-253) and the dense SVD least-squares local solve `svd(A) \\ b`.  `a0`, `a1` are
accepted and ignored, as in the reference body.
"""
b_linsolve(A::MPO, b::MPS, x0::MPS, a0::Number=0, a1::Number=1; reverse_step=false, kwargs...) =
  linsolve(A, b, x0, 0, 1; local_solver=QTT_SOLVER_PG_LSQ, kwargs...)

"""
    dmrg_target(A::MPO, psi0::MPS; target_eigenvalue, nsweeps, maxdim, cutoff, ...)

Reference src/dmrg_target.jl:1-12.  `local_solver = QTT_SOLVER_DENSE_EIGEN` is the reference body itself (dense
effective operator, eigenvector of the eigenvalue nearest the target; 4 chi_l chi_r <= 8192); the default is the Krylov
local solver sketched at :14-21, which scales to chi = 512 where the dense operator would be 8.8 TB.
"""
function dmrg_target(A::MPO, psi0::MPS; target_eigenvalue, nsweeps, maxdim=typemax(Int32), mindim=1, cutoff=1e-16, tol=1e-12,
                     krylovdim=30, maxiter=1, outputlevel=0, local_solver=QTT_SOLVER_LANCZOS, ctx::Context=context())
  hA = upload(ctx, A); hx = upload(ctx, psi0)
  md = sweepvec(Int32, min.(maxdim, typemax(Int32)), nsweeps)
  mn = sweepvec(Int32, mindim, nsweeps)
  co = sweepvec(Float64, cutoff, nsweeps)
  infos = Vector{QttSweepInfo}(undef, nsweeps)
  try
    GC.@preserve md mn co begin
      p = Ref(QttSweepParams(nsweeps, pointer(md), pointer(mn), pointer(co), local_solver, 0.0, 1.0, tol, krylovdim, maxiter,
                             0, 1, outputlevel, target_eigenvalue, 0, 0))
      check(ctx, ccall((:qtt_dmrg_target, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{QttSweepParams}, Ptr{QttSweepInfo}),
                       ctx.handle, hA, hx, p, infos))
    end
    return infos[end].energy, download(ctx, hx, siteinds(psi0))
  finally
    free_mps(hx); free_mpo(hA)
  end
end

# `apply(A, psi; cutoff, maxdim, mindim)` (density-matrix algorithm); like ITensors `apply`, the output MPS carries
# `noprime` of A's OUTPUT site indices (for a DFT MPO: s_{n+1-j} on tensor j)
function apply_mpo(A::MPO, psi::MPS; cutoff=1e-13, maxdim=0, mindim=1, ctx::Context=context())
  for j in 1:length(A)
    mpo_site_inds(A, j)[1] == siteind(psi, j) || error("apply_mpo: input site index of MPO tensor $j is not the site index of MPS tensor $j")
  end
  sout = [noprime(mpo_site_inds(A, j)[2]) for j in 1:length(A)]
  hA = upload(ctx, A); hp = upload(ctx, psi)
  out = Ref{Ptr{Cvoid}}(C_NULL)
  try
    check(ctx, ccall((:qtt_apply, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int32, Int32, Ptr{Ptr{Cvoid}}),
                     ctx.handle, hA, hp, cutoff, maxdim, mindim, out))
    return download(ctx, out[], sout)
  finally
    out[] != C_NULL && free_mps(out[]); free_mps(hp); free_mpo(hA)
  end
end

# reference src/dft.jl:121-133: the chain reversal stays here.  The reference rebuilds the DFT MPO on every call (:123);
# its dense cores only depend on (n, cutoff), so they are cached here and re-labelled with the caller's site indices.
const _dft_cache = Dict{Tuple{Int,Float64,Bool},MPO}()
function cached_dft_mpo(s; cutoff, inverse=false)
  key = (length(s), Float64(cutoff), inverse)
  F = get!(_dft_cache, key) do
    inverse ? idft_mpo(s; cutoff) : dft_mpo(s; cutoff)
  end
  # the site indices the cached MPO was built on, in chain order: dft_mpo has s_j unprimed on tensor j, idft_mpo has s_j primed
  s0 = inverse ? [noprime(mpo_site_inds(F, j)[2]) for j in 1:length(F)] : [mpo_site_inds(F, j)[1] for j in 1:length(F)]
  return s0 == s ? F : replace_siteinds(F, s0, s)
end
# every site index of every tensor is mapped by id and keeps its prime level (tensor j does NOT carry (s_j, s_j'))
function replace_siteinds(F::MPO, old, new)
  function relabel(T)
    for i in inds(T)
      k = findfirst(==(noprime(i)), old)
      isnothing(k) || (T = replaceind(T, i, setprime(new[k], plev(i))))
    end
    return T
  end
  return MPO([relabel(F[j]) for j in 1:length(F)])
end

# reference src/dft.jl:121-126: reverse(apply(dft_mpo(s), psi; kwargs...)) -- the OUTPUT chain is reversed
function apply_dft_mpo(psi::MPS; kwargs...)
  s = siteinds(psi)
  F = cached_dft_mpo(s; cutoff=1e-15)
  phi = apply_mpo(F, psi; kwargs...)
  return MPS(reverse([phi[j] for j in 1:length(phi)]))
end

# reference src/dft.jl:128-133: apply(idft_mpo(s), reverse(psi); kwargs...) -- the INPUT chain is reversed, the result is not
function apply_idft_mpo(psi::MPS; kwargs...)
  s = siteinds(psi)
  F = cached_dft_mpo(s; cutoff=1e-15, inverse=true)
  rpsi = MPS(reverse([psi[j] for j in 1:length(psi)]))
  return apply_mpo(F, rpsi; kwargs...)
end

# `+(x, y; cutoff, maxdim)` of ITensorMPS (reference src/fourier_interpolation.jl:8, src/function_to_mps.jl:123) on the device:
# the direct sum is assembled in HBM and compressed by the density-matrix pass of `qtt_apply`
function add_mps(x::MPS, y::MPS; alpha=1.0, beta=1.0, cutoff=1e-15, maxdim=0, mindim=1, ctx::Context=context())
  hx = upload(ctx, x); hy = upload(ctx, y)
  out = Ref{Ptr{Cvoid}}(C_NULL)
  try
    check(ctx, ccall((:qtt_mps_add, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Int32, Int32, Ptr{Ptr{Cvoid}}),
                     ctx.handle, hx, hy, alpha, beta, cutoff, maxdim, mindim, out))
    return download(ctx, out[], siteinds(x))
  finally
    out[] != C_NULL && free_mps(out[]); free_mps(hy); free_mps(hx)
  end
end

# `mps_to_discrete_function(psi)` (reference src/function_to_mps.jl:140-145, one dimension): expanded on the device
function mps_to_discrete_function(psi::MPS; ctx::Context=context())
  hp = upload(ctx, psi)
  T = any(j -> eltype(psi[j]) <: Complex, 1:length(psi)) ? ComplexF64 : Float64
  out = Vector{T}(undef, 2^length(psi))
  try
    GC.@preserve out check(ctx, ccall((:qtt_mps_to_vec, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32), ctx.handle, hp, pointer(out), 0))
    return out
  finally
    free_mps(hp)
  end
end

# values at 0-based grid indices: `project_bits(psi, bits)` with every site projected (reference src/project_bits.jl:1-15), batched
function sample_values(psi::MPS, idx::AbstractVector{<:Integer}; ctx::Context=context())
  hp = upload(ctx, psi)
  T = any(j -> eltype(psi[j]) <: Complex, 1:length(psi)) ? ComplexF64 : Float64
  ii = Vector{Int64}(idx)
  out = Vector{T}(undef, length(ii))
  try
    GC.@preserve ii out check(ctx, ccall((:qtt_mps_sample, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cvoid}),
                                         ctx.handle, hp, pointer(ii), length(ii), pointer(out)))
    return out
  finally
    free_mps(hp)
  end
end

# residual metric of reference src/mps_utils.jl:109-116
"""
    vec_to_mps(v::AbstractVector, s::Vector{<:Index}; cutoff=1e-15, maxdim=0, mindim=1)
    function_to_mps(f, s, xstart, xstop; cutoff=1e-15, kwargs...)     # alg = "factorize"

Reference src/function_to_mps.jl:41-52,170-177 (`MPS(vec_to_tensor(v), s; cutoff)`): the TT-SVD runs on the GPU
(streaming Gram + orthogonal projection per site); a 2^30-sample vector takes ~0.2 s of device time on one B200.
The Julia vector is already in the order the C ABI wants (index j-1 = sum_k b_k 2^(n-k), site 1 = MSB), no permute.
"""
function vec_to_mps(v::AbstractVector{<:Real}, s::Vector{<:Index}; cutoff=1e-15, maxdim=0, mindim=1, ctx::Context=context())
  n = length(s)
  length(v) == 2^n || error("vec_to_mps: length(v) must be 2^length(s)")
  vv = convert(Vector{Float64}, v)
  out = Ref{Ptr{Cvoid}}(C_NULL); err = Ref{Cdouble}(0.0)
  try
    check(ctx, ccall((:qtt_vec_to_mps, libqtt), Cint,
                     (Ptr{Cvoid}, Ptr{Cdouble}, Int32, Int32, Cdouble, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Cdouble}),
                     ctx.handle, vv, n, 0, cutoff, min(maxdim, typemax(Int32)), mindim, out, err))
    return download(ctx, out[], s)
  finally
    out[] != C_NULL && free_mps(out[])
  end
end

function function_to_mps(f::Function, s::Vector{<:Index}, xstart, xstop; cutoff=1e-15, kwargs...)
  n = length(s)
  x = range(; start=xstart, stop=xstop, length=(2^n + 1))[1:(2^n)]   # qtt_xrange, src/function_to_mps.jl:10-12
  return vec_to_mps(f.(x), s; cutoff, kwargs...)
end

function sqeuclidean_normalized(A::MPO, x::MPS, b::MPS; ctx::Context=context())
  hA = upload(ctx, A); hx = upload(ctx, x); hb = upload(ctx, b)
  sq = Ref{Float64}(0.0); sqn = Ref{Float64}(0.0)
  try
    check(ctx, ccall((:qtt_residual, libqtt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                     ctx.handle, hA, hx, hb, sq, sqn))
    return sqn[]
  finally
    free_mps(hb); free_mps(hx); free_mpo(hA)
  end
end

end # module

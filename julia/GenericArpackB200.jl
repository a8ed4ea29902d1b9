# GenericArpackB200.jl -- the binding a GenericArpack.jl maintainer adds to route the Lanczos hot path
# to libgarpack_b200.so.  NOT executable in the build image (no Julia there: never parsed or run); it mirrors,
# call for call, what genericarpack.jl_b200/csrc/ga_host.cu does in C++.  It binds the UNMODIFIED GenericArpack
# v0.2.1 -- every GenericArpack name used below exists in /root/reference/src (cited at its use); nothing has to be
# added to or factored out of the reference:
#   * operator hook   abstract type ArpackOp                      src/idonow_ops.jl:50
#   * state hook      abstract type AbstractArpackState{T}        src/macros.jl:356
#     (every inner routine takes `state` positionally so a subtype can override it:
#      dgetv0! src/dgetv0.jl:504-521, dsaitr! src/dsaitr.jl:962-981, dsapps! src/dsapps.jl:577-594;
#      the reference's own override tests: test/dsapps_override.jl:8-62, test/dgetv0_override.jl:6-47)
module GenericArpackB200

using GenericArpack, LinearAlgebra, SparseArrays
import GenericArpack: ArpackOp, ArpackNormalSVDOp, AbstractArpackState, ArpackState, AitrState, Getv0State, Saup2State,
                      dgetv0!, dsaitr!, dsapps!, dorm2r!, norm2, opx!, arpack_mode, bmat,
                      is_arpack_mode_valid_for_op, symeigs, eigenvecs_to_singvecs!, _uv_size,
                      _adjust_nev_which_for_svd, _adjust_eigenvalues_to_singular_values!,
                      dsaupd!, simple_dseupd!, shift, ArpackEigen, ArpackException   # src/dsaupd.jl:921, src/dseupd.jl:986, src/idonow_ops.jl:54, src/interface.jl:170-203,125-127

const lib = "libgarpack_b200.so"
const GA_F64, GA_C64, GA_DD, GA_F32 = Cint(0), Cint(1), Cint(2), Cint(3)
ga_dtype(::Type{Float64}) = GA_F64
ga_dtype(::Type{ComplexF64}) = GA_C64
ga_dtype(::Type{Float32}) = GA_F32      # Float32 vectors; the state stays B200ArpackState{Float64}: eigs(Float32, Float64, op, k)
# DoubleFloats.Double64 is two adjacent Float64 (hi, lo): pass pointers as-is with GA_DD.

check(rc, ctx) = rc < -99 ? error("garpack_b200: ", unsafe_string(ccall((:ga_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))) : rc

# ---- operator: an ArpackOp whose OP*x lives on the GPU (src/idonow_ops.jl:115-123 for the CPU one)
mutable struct B200ArpackOp{T} <: ArpackOp
  ctx::Ptr{Cvoid}       # ga_ctx*
  n::Int
  ncv::Int
end
function B200ArpackOp(A::SparseMatrixCSC{T}, ncv::Int; device::Int=0) where T
  n = size(A, 1)
  At = SparseMatrixCSC(transpose(A))            # CSC of A' == CSR of A (full storage, both triangles)
  rowptr = Int64.(At.colptr .- 1); col = Int32.(At.rowval .- 1)
  ctxref = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:ga_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint),
             ctxref, ga_dtype(T), n, n, 0, ncv, device, 0, 1)
  rc == 0 || error("ga_create failed ($rc): no CUDA device? there is no CPU fallback")
  check(ccall((:ga_set_csr, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}), ctxref[], rowptr, col, At.nzval), ctxref[])
  op = B200ArpackOp{T}(ctxref[], n, ncv)
  finalizer(o -> ccall((:ga_destroy, lib), Cvoid, (Ptr{Cvoid},), o.ctx), op)
  return op
end
Base.size(op::B200ArpackOp) = op.n
arpack_mode(::B200ArpackOp) = 1
bmat(::B200ArpackOp) = Val(:I)
is_arpack_mode_valid_for_op(mode::Int, ::B200ArpackOp) = mode == 1
# opx! is never called on host vectors in the overridden dsaitr!; provided for Matrix(T, op) etc.
opx!(y, op::B200ArpackOp, x) = (check(ccall((:ga_opx, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), op.ctx, x, y), op.ctx); y)

# ---- state: the dispatch hook.  Same fields as ArpackState (src/macros.jl:358-366) + the context.
Base.@kwdef mutable struct B200ArpackState{T} <: AbstractArpackState{T}
  aitr::AitrState{T} = AitrState{T}()
  getv0::Getv0State{T} = Getv0State{T}()
  saup2::Saup2State{T} = Saup2State{T}()
  aupd_nev0::Base.RefValue{Int} = Ref{Int}(0)
  aupd_np::Base.RefValue{Int} = Ref{Int}(0)
  aupd_mxiter::Base.RefValue{Int} = Ref{Int}(0)
  aup2_rnorm::Base.RefValue{T} = Ref{T}(zero(T))
  ctx::Ptr{Cvoid} = C_NULL
end

# ---- device-resident stand-ins for V / resid / workd: sizes for the check macros (src/macros.jl:208-231), no host data.
# Host code outside the overridden routines touches the big arrays in exactly the places SURVEY.md 8(b) lists; each of
# them gets ONE method below, keyed on which array and which range is addressed, and anything else is an error (scalar
# indexing included) -- if a future reference version touches the arrays differently the binding fails loudly instead
# of computing on stale data.
struct DeviceVector{T} <: AbstractVector{T}; ctx::Ptr{Cvoid}; n::Int; len::Int; kind::Symbol; end    # kind = :resid (len n) | :workd (len 3n)
struct DeviceMatrix{T} <: AbstractMatrix{T}; ctx::Ptr{Cvoid}; n::Int; ncv::Int; end                  # the Lanczos basis V
struct DeviceVecView{T} <: AbstractVector{T}; parent::DeviceVector{T}; r::UnitRange{Int}; end        # @view(x[a:b])
struct DeviceMatView{T} <: AbstractMatrix{T}; parent::DeviceMatrix{T}; rows::UnitRange{Int}; cols::UnitRange{Int}; end
Base.size(v::DeviceVector) = (v.len,); Base.size(V::DeviceMatrix) = (V.n, V.ncv)
Base.size(v::DeviceVecView) = (length(v.r),); Base.size(V::DeviceMatView) = (length(V.rows), length(V.cols))
const _DeviceArray = Union{DeviceVector,DeviceMatrix,DeviceVecView,DeviceMatView}
Base.getindex(::_DeviceArray, i...) = error("GenericArpackB200: device-resident array, no host indexing (unexpected host access to V/resid/workd)")
Base.setindex!(::_DeviceArray, v, i...) = error("GenericArpackB200: device-resident array, no host indexing")
Base.view(v::DeviceVector, r::UnitRange{Int}) = DeviceVecView(v, r)                                   # what @view(x[a:b]) lowers to
Base.view(V::DeviceMatrix, rows::UnitRange{Int}, cols::UnitRange{Int}) = DeviceMatView(V, rows, cols)
_is(v::DeviceVecView, kind::Symbol, r::UnitRange{Int}) = v.parent.kind === kind && v.r == r

# src/dsaup2.jl:1397  copyto!(@view(workd[1:n]), @view(resid[1:n]))   and   :1386 (bmat = :G)  workd[n+1:2n] <- resid:
# the library's next Lanczos step reads resid itself (ga_apply_shifts leaves it ready), nothing to move.
function Base.copyto!(d::DeviceVecView{T}, s::DeviceVecView{T}) where T
  n = s.parent.n
  (_is(s, :resid, 1:n) && (_is(d, :workd, 1:n) || _is(d, :workd, n+1:2n))) ||
    error("GenericArpackB200: unexpected device copy $(d.parent.kind)[$(d.r)] <- $(s.parent.kind)[$(s.r)]")
  return d
end
# src/interface.jl:260  copyto!(resid, v0)
function Base.copyto!(d::DeviceVector{T}, v0::AbstractVector{T}) where T
  (d.kind === :resid && length(v0) == d.n) || error("GenericArpackB200: v0 must have n entries")
  h = Vector{T}(v0)
  check(ccall((:ga_upload_resid, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), d.ctx, h), d.ctx); d
end
# src/dsaup2.jl:1414  rnorm = norm2(TR, @view(resid[1:n]));  src/dseupd.jl:1180 (bmat = :G only)  bnorm2 = norm2(TR, @view(workd[1:n]))
function norm2(::Type{TR}, v::DeviceVecView) where TR
  n = v.parent.n; out = Ref{TR}(zero(TR))
  if _is(v, :resid, 1:n)
    check(ccall((:ga_norm2_resid, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), v.parent.ctx, out), v.parent.ctx)
  elseif _is(v, :workd, 1:n)       # ||B resid||: only feeds the Ritz-vector purification of modes 3-5 (src/dseupd.jl:1526-1529), unused in mode 2
    check(ccall((:ga_bnorm_resid, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), v.parent.ctx, out), v.parent.ctx)
  else
    error("GenericArpackB200: unexpected device norm of $(v.parent.kind)[$(v.r)]")
  end
  return out[]
end

# ---- dgetv0! (src/dgetv0.jl:504-753)
function dgetv0!(ido::Ref{Int}, ::Val{:I}, itry::Int, initv::Bool, n::Int, j::Int, V::DeviceMatrix{T}, ldv::Int,
                 resid::DeviceVector{T}, rnorm::Ref{TR}, ipntr, workd, state::B200ArpackState{TR};
                 stats=nothing, debug=nothing, idonow=nothing) where {T,TR}
  iseed = collect(Int64, state.getv0.iseed[])
  ierr = ccall((:ga_getv0, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Cint, Cint, Ptr{Cvoid}), state.ctx, iseed, initv, j, rnorm)
  state.getv0.iseed[] = Tuple(iseed)
  ido[] = 99
  return Int(check(ierr, state.ctx))
end

# ---- dsaitr! (src/dsaitr.jl:962-1535): np steps on the device, H(k+1:k+np,:) comes back
function dsaitr!(ido::Ref{Int}, ::Val{:I}, n::Int, k::Int, np::Int, mode::Int, resid::DeviceVector{T}, rnorm::Ref{TR},
                 V::DeviceMatrix{T}, ldv::Int, H::AbstractMatrix{TR}, ldh::Int, ipntr, workd,
                 state::B200ArpackState{TR}; stats=nothing, debug=nothing, idonow=nothing) where {T,TR}
  iseed = collect(Int64, state.getv0.iseed[])
  info = ccall((:ga_lanczos, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Int64}),
               state.ctx, k, np, H, ldh, rnorm, iseed)     # H is a reshaped view into workl: column-major ldh x 2
  state.getv0.iseed[] = Tuple(iseed)
  if stats !== nothing   # same counters the reference maintains (src/dsaitr.jl:1113,1335,1430)
    s = ga_stats(state.ctx); stats.nopx = s.nopx; stats.nrorth = s.nrorth; stats.nitref = s.nitref; stats.nrstrt = s.nrstrt
    stats.taitr = s.taitr
  end
  ido[] = 99
  return Int(check(info, state.ctx))
end

# ---- dsapps! (src/dsapps.jl:577-869).  The np implicit QR shifts on the tridiagonal H (:609-805) are host work on
# kplusp x kplusp data and stay in the reference: its own GENERIC method is called on zero-row stand-ins (n = 0: a
# 1 x kplusp matrix satisfies the length checks :597-607 with ldv = 0, and every mul!/copyto!/axpy! of the tail :813-854
# runs on empty views), which leaves Q and the updated H exactly as the full call would.  The n-row tail is then one
# library call.  `state = nothing` is what the reference's own callers pass outside dsaup2! (test/dsapps_simple.jl).
function dsapps!(n::Int, kev::Int, np::Int, shift::AbstractVecOrMat{TR}, V::DeviceMatrix{T}, ldv::Int,
                 H::AbstractMatrix{TR}, ldh::Int, resid::DeviceVector{T}, Q::AbstractMatrix{TR}, ldq::Int, workd::DeviceVector{T},
                 state::B200ArpackState{TR}; stats=nothing, debug=nothing) where {T,TR}
  kplusp = kev + np
  dsapps!(0, kev, np, shift, zeros(T, 1, kplusp), 0, H, ldh, T[], Q, ldq, T[], nothing; stats, debug)   # host chase only
  beta = Ref(H[kev+1, 1]); rn = Ref(zero(TR))
  check(ccall((:ga_apply_shifts, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
              state.ctx, kev, np, Q, ldq, beta, rn), state.ctx)   # V*Q, resid update, ||resid|| (src/dsapps.jl:813-854, dsaup2.jl:1397-1414)
  return nothing
end

# ---- dorm2r!(:R,:N) on V inside simple_dseupd! (src/dseupd.jl:1431-1432: C = @view(V[1:n,1:ncv]), work = @view(workd[n+1:2n]))
function dorm2r!(::Val{:R}, ::Val{:N}, m::Integer, n::Integer, k::Integer, A::AbstractMatrix{TR}, tau::AbstractVector{TR},
                 C::DeviceMatView{T}, work::DeviceVecView{T}) where {T,TR}
  (C.rows == 1:C.parent.n && C.cols == 1:C.parent.ncv) || error("GenericArpackB200: dorm2r! on an unexpected block of V")
  Qh = Matrix{TR}(I, n, n)
  GenericArpack.dorm2r!(Val(:R), Val(:N), n, n, k, A, tau, Qh, Vector{TR}(undef, n))   # Qh = H(1)...H(k), ncv x ncv on the host (src/arpack-blas-qr.jl:409-476)
  check(ccall((:ga_ritz_vectors, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), C.parent.ctx, k, Qh, n, C_NULL, 0), C.parent.ctx)
  return C
end
# src/dseupd.jl:1434  copyto!(@view(Z[1:n, 1:nconv]), @view(V[1:n, 1:nconv]))
function Base.copyto!(Z::SubArray{T,2,Matrix{T}}, V::DeviceMatView{T}) where T
  (V.rows == 1:V.parent.n && first(V.cols) == 1 && size(Z) == size(V)) || error("GenericArpackB200: unexpected device -> host copy of V")
  check(ccall((:ga_download_v, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64), V.parent.ctx, 0, size(Z, 2), pointer(Z), stride(Z, 2)), V.parent.ctx)
  return Z
end

struct ga_stats_t; nopx::Int64; nbx::Int64; nrorth::Int64; nitref::Int64; nrstrt::Int64
  taupd::Float64; taup2::Float64; taitr::Float64; teigt::Float64; tgets::Float64; tapps::Float64; tconv::Float64
  tmvopx::Float64; tmvbx::Float64; tgetv0::Float64; titref::Float64; trvec::Float64; end
ga_stats(ctx) = (r = Ref{ga_stats_t}(); ccall((:ga_get_stats, lib), Cint, (Ptr{Cvoid}, Ptr{ga_stats_t}), ctx, r); r[])

# ---- the drop-in: a more specific symeigs method.  Its body is the reference's own (src/interface.jl:215-290) with the
# three big allocations (:245-246,249) replaced by device handles; dsaupd! and simple_dseupd! are the UNCHANGED
# reference routines, which reach the GPU through the dgetv0!/dsaitr!/dsapps!/dorm2r! methods above.
function _b200_symeigs(::Type{TV}, ::Type{TF}, op, nev::Integer;
    which::Symbol=:LM, maxiter::Int=max(300, ceil(Int, sqrt(size(op)))), bmat=bmat(op), ncv::Int=op.ncv,
    mode=arpack_mode(op), v0=nothing, stats=nothing, debug=nothing, tol::TF=zero(TF), ritzvec::Bool=true,
    failonmaxiter::Bool=true) where {TV,TF}
  n = size(op)
  ncv == op.ncv || throw(ArgumentError("ncv=$ncv differs from the device context's ncv=$(op.ncv)"))
  ncv <= n || throw(ArgumentError("ncv=$ncv must be at most n=$n"))                          # :231
  nev < ncv || throw(ArgumentError("nev=$nev must be strictly less than ncv=$ncv"))           # :232
  is_arpack_mode_valid_for_op(mode, op) || @warn("mode $mode is not valid for op with type $(typeof(op))")   # :239-241
  state = B200ArpackState{TF}(ctx=op.ctx)
  V = DeviceMatrix{TV}(op.ctx, n, ncv)                                                         # :245
  resid = DeviceVector{TV}(op.ctx, n, n, :resid)                                               # :246
  iparam = zeros(Int, 11); ipntr = zeros(Int, 11)                                              # :247-248
  workd = DeviceVector{TV}(op.ctx, n, 3n, :workd)                                              # :249
  lworkl = ncv*ncv + 8*ncv; workl = Vector{TF}(undef, lworkl)                                  # :250-251
  iparam[1] = 1; iparam[4] = 1; iparam[3] = maxiter; iparam[7] = mode                          # :253-256
  info_initv = 0
  if v0 !== nothing; copyto!(resid, v0); info_initv = 1; end                                   # :258-262
  ido = Ref{Int}(0)
  info = dsaupd!(ido, bmat, n, which, nev, tol, resid, ncv, V, size(V, 1), iparam, ipntr, workd, workl, lworkl, info_initv;
                 state, stats, debug, idonow=op)[1]                                            # :264-267
  if info == 1 && failonmaxiter                                                                # :269-273
    throw(ArpackException("symmetric aupd hit maxiter=$maxiter with $(iparam[5]) < $nev eigenvalues converged"))
  elseif (info != 0 && info != 1) || ido[] != 99
    throw(ArpackException("symmetric aupd gave error code info=$info with ido=$(ido[])"))
  end
  nconv = iparam[5]; nvectors = ritzvec ? min(nev, nconv) : 0                                  # :276-277
  vectors = Matrix{TV}(undef, n, nvectors); values = Vector{TF}(undef, nconv); select = Vector{Int}(undef, ncv)
  ierr = simple_dseupd!(ritzvec, select, values, vectors, shift(TF, op), bmat, n, which, nev, tol, resid, ncv, V,
                        iparam, ipntr, workd, workl; debug, stats, state)                      # :281-283
  ierr == 0 || throw(ArpackException("symmetric eupd gave error code ierr=$ierr"))              # :285-287
  return ArpackEigen(which, ipntr, iparam, V, workd, workl, resid, values, vectors, bmat, op, state)   # :289
end
symeigs(::Type{TV}, ::Type{TF}, op::B200ArpackOp{TV}, nev::Integer; kwargs...) where {TV,TF} = _b200_symeigs(TV, TF, op, nev; kwargs...)

# ---- svds: ArpackNormalOp on the device (src/idonow_ops.jl:216-333).  `svds(B200NormalOp(A, ncv), k)` runs the
# reference's own svds (src/interface.jl:312-337): symeigs on A'A / AA' through the methods above, then
# eigenvecs_to_singvecs!, which here is ONE library call (A*v_i, norms, scaling and orth! on the device).
mutable struct B200NormalOp{T} <: ArpackNormalSVDOp
  ctx::Ptr{Cvoid}; m::Int; n::Int; ncv::Int
end
function B200NormalOp(A::SparseMatrixCSC{T}, ncv::Int; device::Int=0) where T
  m, n = size(A)
  At = SparseMatrixCSC(transpose(A)); rowptr = Int64.(At.colptr .- 1); col = Int32.(At.rowval .- 1)   # CSR of A
  ctxref = Ref{Ptr{Cvoid}}(C_NULL); sz = min(m, n)
  rc = ccall((:ga_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint),
             ctxref, ga_dtype(T), sz, sz, 0, ncv, device, 0, 1)
  rc == 0 || error("ga_create failed ($rc)")
  check(ccall((:ga_set_normal_csr, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}),
              ctxref[], m, n, rowptr, col, At.nzval), ctxref[])
  op = B200NormalOp{T}(ctxref[], m, n, ncv)
  finalizer(o -> ccall((:ga_destroy, lib), Cvoid, (Ptr{Cvoid},), o.ctx), op)
  return op
end
Base.size(op::B200NormalOp) = min(op.m, op.n)
arpack_mode(::B200NormalOp) = 1
bmat(::B200NormalOp) = Val(:I)
_uv_size(op::B200NormalOp) = (op.m, op.n)
# _adjust_nev_which_for_svd / _adjust_eigenvalues_to_singular_values! are inherited from ArpackNormalSVDOp (:240-260)

symeigs(::Type{TV}, ::Type{TF}, op::B200NormalOp{TV}, nev::Integer; kwargs...) where {TV,TF} = _b200_symeigs(TV, TF, op, nev; kwargs...)

# eigenvecs_to_singvecs!(U, V, OP, Z) (src/idonow_ops.jl:295-333), called by the reference's svds (src/interface.jl:333)
# with the eigenvectors already re-sorted on the host (:321-328).  Z goes back into the first k basis columns (one
# n x k upload per solve -- no assumption about how the host permuted it), then ONE library call forms the other
# side's vectors on the device: A*v_i (or A'*u_i), norms, scaling and orth!.  The reference's check_norm thresholds
# (:299-307) are applied to the norms the library returns.
function eigenvecs_to_singvecs!(U::AbstractMatrix{T}, V::AbstractMatrix{T}, op::B200NormalOp{T}, Z::AbstractMatrix{T}) where T
  k = size(Z, 2); left = op.m > op.n
  W = left ? U : V; copyto!(left ? V : U, Z)
  Zh = Matrix{T}(Z)
  check(ccall((:ga_upload_v, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64), op.ctx, 0, k, Zh, size(Zh, 1)), op.ctx)
  perm = Int32.(0:k-1)
  norms = Vector{real(T)}(undef, k)
  rc = ccall((:ga_svd_vectors, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
             op.ctx, k, perm, W, size(W, 1), norms)
  rc == -104 && throw(ArpackException("encountered σ ≈ 0 when computing $(left ? "U" : "V") subspace"))
  check(rc, op.ctx)
  any(norms .<= eps(real(T))/2) && throw(ArpackException("encountered σ ≈ $(minimum(norms)) when computing $(left ? "U" : "V") subspace"))
  any(norms .<= GenericArpack._eps23(real(T))) && @warn "$(left ? "U" : "V") subspace may be inaccurate due to small singular value"
  return nothing
end

# ---- generalised problem A x = lambda B x, mode 2 / bmat = :G: the device counterpart of
# ArpackSymmetricGeneralizedOp(A, S, B) (src/idonow_ops.jl:462-492), reached by eigs(Symmetric(A), Symmetric(B), k)
# (src/interface.jl:143-163).  The factorisation object S of the reference becomes the library's device CG with B
# (or the caller's own device solver through ga_set_solve_callback).
mutable struct B200GeneralizedOp{T} <: ArpackOp
  ctx::Ptr{Cvoid}; n::Int; ncv::Int
end
function B200GeneralizedOp(A::SparseMatrixCSC{T}, B::SparseMatrixCSC{T}, ncv::Int; device::Int=0,
                           cg_tol::Union{Nothing,Float64}=nothing, cg_maxit::Int=0) where T
  n = size(A, 1)
  At = SparseMatrixCSC(transpose(A)); Bt = SparseMatrixCSC(transpose(B))        # CSR of A and of B
  ctxref = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:ga_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint),
             ctxref, ga_dtype(T), n, n, 0, ncv, device, 0, 1)
  rc == 0 || error("ga_create failed ($rc)")
  tol = cg_tol === nothing ? C_NULL : Ref(cg_tol)
  rc = ccall((:ga_set_generalized_csr, lib), Cint,
             (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Float64}, Cint),
             ctxref[], Int64.(At.colptr .- 1), Int32.(At.rowval .- 1), At.nzval,
             Int64.(Bt.colptr .- 1), Int32.(Bt.rowval .- 1), Bt.nzval, tol, cg_maxit)
  rc == -105 && throw(ArgumentError("B is not positive definite"))               # GA_ERR_SOLVE
  check(rc, ctxref[])
  op = B200GeneralizedOp{T}(ctxref[], n, ncv)
  finalizer(o -> ccall((:ga_destroy, lib), Cvoid, (Ptr{Cvoid},), o.ctx), op)
  return op
end
Base.size(op::B200GeneralizedOp) = op.n
arpack_mode(::B200GeneralizedOp) = 2
bmat(::B200GeneralizedOp) = Val(:G)
is_arpack_mode_valid_for_op(mode::Int, ::B200GeneralizedOp) = mode == 2
GenericArpack.bx!(y, op::B200GeneralizedOp, x) =
  (check(ccall((:ga_bx, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), op.ctx, x, y), op.ctx); y)
opx!(y, op::B200GeneralizedOp, x) = (check(ccall((:ga_opx, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), op.ctx, x, y), op.ctx); y)
# The B-norm branches (src/dsaitr.jl:1195-1204,1298-1300,1405-1407; src/dgetv0.jl:625-627,696-698) live inside the
# library once ga_set_generalized_csr has been called, so the :G methods are the :I methods above with another Val:
for f in (:dgetv0!, :dsaitr!)
  @eval $f(ido::Ref{Int}, ::Val{:G}, args...; kw...) = $f(ido, Val(:I), args...; kw...)
end
# dsaup2!'s own :G touches outside the overridden routines (SURVEY.md 8b): the ido = 2 exchange that forms B*resid in
# workd (src/dsaup2.jl:1383-1390, `_i_do_now_bx!`) and rnorm = sqrt(abs(dot(resid, workd))) (:1410-1411) collapse into one
# library call, ga_bnorm_resid; simple_dseupd!'s bnorm2 = norm2(workd[1:n]) (src/dseupd.jl:1178-1180) only feeds the
# mode 3-5 Ritz-vector purification and is not used for mode 2.
GenericArpack._i_do_now_bx!(idonow::B200GeneralizedOp, ipntr, workd::DeviceVector, n) = nothing    # src/idonow_ops.jl:86-88, called at src/dsaup2.jl:1394
function LinearAlgebra.dot(r::DeviceVecView{T}, w::DeviceVecView{T}) where T                             # src/dsaup2.jl:1410
  n = r.parent.n
  (_is(r, :resid, 1:n) && _is(w, :workd, 1:n)) || error("GenericArpackB200: unexpected device dot product")
  out = Ref{real(T)}(zero(real(T)))
  check(ccall((:ga_bnorm_resid, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), r.parent.ctx, out), r.parent.ctx)
  return out[] * out[]                                         # the caller applies sqrt(abs(.))  (:1411)
end
symeigs(::Type{TV}, ::Type{TF}, op::B200GeneralizedOp{TV}, nev::Integer; kwargs...) where {TV,TF} = _b200_symeigs(TV, TF, op, nev; kwargs...)

# ---- function operators (src/idonow_ops.jl:360-368): the user's opx! as a callback on DEVICE vectors.  `F(y, x, stream)`
# receives CuPtr-backed views (e.g. CUDA.unsafe_wrap(CuArray, ...)) and must launch its kernels on `stream`.
mutable struct B200FunctionOp{T} <: ArpackOp
  ctx::Ptr{Cvoid}; n::Int; ncv::Int; F::Function; cfun::Ptr{Cvoid}
end
function _op_thunk(user::Ptr{Cvoid}, x::Ptr{Cvoid}, y::Ptr{Cvoid}, nx::Int64, ny::Int64, stream::Ptr{Cvoid})::Cint
  op = unsafe_pointer_to_objref(user)::B200FunctionOp
  try op.F(y, x, nx, ny, stream); return Cint(0) catch; return Cint(1) end
end
function B200FunctionOp(::Type{T}, F::Function, n::Int, ncv::Int; device::Int=0) where T
  ctxref = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:ga_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint),
             ctxref, ga_dtype(T), n, n, 0, ncv, device, 0, 1)
  rc == 0 || error("ga_create failed ($rc)")
  cfun = @cfunction(_op_thunk, Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}))
  op = B200FunctionOp{T}(ctxref[], n, ncv, F, cfun)
  check(ccall((:ga_set_op_callback, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Any), op.ctx, cfun, op), op.ctx)
  finalizer(o -> ccall((:ga_destroy, lib), Cvoid, (Ptr{Cvoid},), o.ctx), op)
  return op
end
Base.size(op::B200FunctionOp) = op.n
arpack_mode(::B200FunctionOp) = 1
bmat(::B200FunctionOp) = Val(:I)
is_arpack_mode_valid_for_op(mode::Int, ::B200FunctionOp) = mode == 1
symeigs(::Type{TV}, ::Type{TF}, op::B200FunctionOp{TV}, nev::Integer; kwargs...) where {TV,TF} = _b200_symeigs(TV, TF, op, nev; kwargs...)

end # module

# PDENLPCuda.jl — the reference-side binding of libpdenlp_cuda.so (include/pdenlp_cuda.h).
#
# NOT RUNNABLE IN THIS REPO'S CONTAINER (no Julia toolchain); it is the stub a PDENLPModels.jl
# maintainer would add.  The GridapPDENLPModel constructor and the NLPModels API surface stay
# unchanged: `attach_cuda!(nlp)` flattens the Gridap objects once, after the stock constructor
# (src/additional_constructors.jl:130-266), and the callback *bodies* of
# src/GridapPDENLPModel.jl:245-661 become one `ccall` each.  @lencheck and Counters stay in Julia.
module PDENLPCuda

using Gridap, NLPModels, PDENLPModels
import NLPModels: obj, grad!, cons_nln!, jac_nln_coord!, hess_coord!, hprod!, jprod_nln!, jtprod_nln!,
                  increment!, @lencheck

const LIB = get(ENV, "PDENLP_CUDA_LIB", "libpdenlp_cuda.so")

# integrand ids of include/pdenlp_cuda.h
@enum Integrand MEMBRANE=1 POISSON_BOLTZMANN=2 BURGER_LIN=3 BURGER_NL=4 KAPPA_POISSON=5 PENALIZED_POISSON=6 BASIC_UNCONSTRAINED=7 CONTROL_SIR=8 CELL_INCREASE=9 DYNAMIC_SIR=10 TORE_BRACHISTOCHRONE=11 POISSON_PARAM=12 SIS=13 KAPPA_POISSON2=14

struct Desc   # mirrors struct pdenlp_desc field by field
  abi_version::Int32; integrand::Int32; dim::Int32; ny::Int32; nu::Int32; nq::Int32; nparam::Int32
  memspace::Int32; device::Int32; nfields_y::Int32
  ncells::Int64; nfree_y::Int64; nfree_u::Int64; ndir_y::Int64; ndir_u::Int64
  cell_dofs_y::Ptr{Int32}; cell_dofs_u::Ptr{Int32}; dir_y::Ptr{Float64}; dir_u::Ptr{Float64}
  phi_y::Ptr{Float64}; grad_y::Ptr{Float64}; phi_u::Ptr{Float64}; wdet::Ptr{Float64}
  ncoef::Int32; nfields_u::Int32; coef::Ptr{Float64}
  params::NTuple{8, Float64}
  field_nfree::Ptr{Int64}; field_ndir::Ptr{Int64}   # multi-field spaces only, else C_NULL
  cell_jinvT::Ptr{Float64}; cell_detj::Ptr{Float64}  # per-cell geometry (ABI v3), C_NULL = identical affine cells
  geom_nq::Int32; flags::Int32   # flags: 1 = deterministic accumulation (coloured, no atomics), 2 = no templates (every cell evaluated)
end

struct Info
  nvar::Int64; ncon::Int64; nnzj::Int64; nnzj_k::Int64; nnzj_y::Int64; nnzj_u::Int64
  nnzh::Int64; nnzh_obj::Int64; nnzh_k_obj::Int64; nnzh_fe::Int64; nnzh_k_con::Int64; nnzh_fe_obj::Int64
end

check(rc) = rc == 0 || error("libpdenlp_cuda ($rc): " * unsafe_string(ccall((:pdenlp_last_error, LIB), Cstring, ())))

# nlp => handle.  Weak keys: the table must not keep the model alive, or its finalizer (which frees the device model,
# its dof maps and staging buffers) would never run.
mutable struct Handle
  ptr::Ptr{Cvoid}
  sharded::Bool      # pdenlp_sharded* (several GPUs, host vectors) instead of pdenlp_model*
  memspace::Int32
end
const HANDLES = WeakKeyDict{Any, Handle}()

function release!(h::Handle)
  h.ptr == C_NULL && return
  h.sharded ? ccall((:pdenlp_sharded_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr) : ccall((:pdenlp_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr)
  h.ptr = C_NULL
end
"""    detach_cuda!(nlp): free the device model now (otherwise it is freed when `nlp` is collected)."""
function detach_cuda!(nlp)
  haskey(HANDLES, nlp) && (release!(HANDLES[nlp]); delete!(HANDLES, nlp))
  return nlp
end

# ---- device-resident vectors: the `S` parameter of the model (src/GridapPDENLPModel.jl:65-73, 219-220) --------------
# A model attached with `memspace = 0` takes its vectors in device memory: x, λ, v, vals stay on the GPU between the
# callbacks of a solver iteration, so a step runs at the kernels' rate (1.6 ms at config 2) instead of the PCIe rate of
# the host-vector path (130-155 ms).  DeviceVector needs no CUDA.jl: it wraps pdenlp_malloc / pdenlp_memcpy_*.
mutable struct DeviceVector <: AbstractVector{Float64}
  ptr::Ptr{Float64}
  len::Int
  function DeviceVector(n::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pdenlp_malloc, LIB), Cint, (Ref{Ptr{Cvoid}}, Int64), p, 8 * n))
    v = new(convert(Ptr{Float64}, p[]), n)
    finalizer(w -> (w.ptr != C_NULL && ccall((:pdenlp_free, LIB), Cint, (Ptr{Cvoid},), w.ptr); w.ptr = C_NULL), v)
    return v
  end
end
DeviceVector(h::AbstractVector{<:Real}) = copyto!(DeviceVector(length(h)), h)
Base.size(v::DeviceVector) = (v.len,)
Base.IndexStyle(::Type{DeviceVector}) = IndexLinear()
Base.similar(v::DeviceVector) = DeviceVector(v.len)
Base.similar(v::DeviceVector, ::Type{Float64}, dims::Dims{1}) = DeviceVector(dims[1])
Base.unsafe_convert(::Type{Ptr{Float64}}, v::DeviceVector) = v.ptr
function Base.copyto!(d::DeviceVector, h::AbstractVector{<:Real})
  length(h) == d.len || throw(DimensionMismatch("DeviceVector of length $(d.len), source of length $(length(h))"))
  hh = h isa Vector{Float64} ? h : collect(Float64, h)
  GC.@preserve hh check(ccall((:pdenlp_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), d.ptr, pointer(hh), 8 * d.len))
  return d
end
function Base.copyto!(h::Vector{Float64}, d::DeviceVector)
  length(h) == d.len || throw(DimensionMismatch("DeviceVector of length $(d.len), destination of length $(length(h))"))
  GC.@preserve h check(ccall((:pdenlp_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), pointer(h), d.ptr, 8 * d.len))
  return h
end
Base.Array(d::DeviceVector) = copyto!(Vector{Float64}(undef, d.len), d)
Base.collect(d::DeviceVector) = Array(d)
# scalar indexing is a device round trip: for inspection only (solvers use broadcasting-free vector kernels or Array(v))
function Base.getindex(d::DeviceVector, i::Int)
  @boundscheck checkbounds(d, i)
  r = Ref{Float64}()
  check(ccall((:pdenlp_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), r, d.ptr + 8 * (i - 1), 8))
  return r[]
end
function Base.setindex!(d::DeviceVector, x, i::Int)
  @boundscheck checkbounds(d, i)
  r = Ref{Float64}(x)
  check(ccall((:pdenlp_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), d.ptr + 8 * (i - 1), r, 8))
  return d
end
Base.fill!(d::DeviceVector, x) = copyto!(d, fill(Float64(x), d.len))

"""
    attach_cuda!(nlp, integrand; params, target, source, memspace = 1, device = -1, devices = nothing)

Flatten the Gridap side of `nlp` (cell→dof maps, Dirichlet values, reference-FE tables at the
quadrature points, weights·detJ or per-cell Jacobians, coefficient fields sampled at the quadrature points) and create
the device model.  `memspace = 1` (PDENLP_MEM_HOST): the callbacks take ordinary `Vector{Float64}`; `memspace = 0`:
they take `DeviceVector`s.  `devices = [0, 1, ...]`: one cell slab per GPU behind one handle (pdenlp_sharded_create;
host vectors).
"""
function attach_cuda!(nlp::GridapPDENLPModel, integrand::Integrand; params = Float64[], target = x -> 0.0,
                      source = x -> 0.0, degree::Int = 1, memspace::Integer = 1, device::Integer = -1,
                      devices::Union{Nothing, AbstractVector{<:Integer}} = nothing,
                      deterministic::Bool = false, templates::Bool = true)
  Ypde, Ycon = nlp.pdemeta.Ypde, nlp.pdemeta.Ycon
  trian = get_triangulation(Ypde)
  dΩ = Measure(trian, degree)
  quad = get_cell_quadrature(dΩ)   # Gridap.CellData.CellQuadrature
  qpts = get_cell_points(quad)
  # single-field spaces of Ypde / Ycon (a MultiFieldFESpace contributes its fields, each with its OWN dof ids:
  # the library applies the consecutive multi-field offsets itself, see include/pdenlp_cuda.h)
  fields(sp) = sp isa MultiFieldFESpace ? collect(sp.spaces) : [sp]
  function flat(fs)   # -> cell ids [nfields*nloc] x nc (field-major), Dirichlet values, per-field sizes, nloc
    ids = [get_cell_dof_ids(f) for f in fs]
    ncl = length(ids[1]); nl = length(ids[1][1])
    cell = Matrix{Int32}(undef, nl * length(fs), ncl)   # column-major == C row-major [nc][nfields*nloc]
    for (k, idk) in enumerate(ids), (c, v) in enumerate(idk); cell[((k - 1) * nl + 1):(k * nl), c] .= v; end
    dir = reduce(vcat, [collect(Float64, get_dirichlet_values(f)) for f in fs])
    return cell, dir, Int64[num_free_dofs(f) for f in fs], Int64[length(get_dirichlet_values(f)) for f in fs], nl
  end
  yf = fields(Ypde)
  celly, diry, nfree_yf, ndir_yf, ny = flat(yf)
  nc = size(celly, 2)
  has_u = !(Ycon isa PDENLPModels.VoidFESpace)
  uf = has_u ? fields(Ycon) : []
  cellu, diru, nfree_uf, ndir_uf, nu = has_u ? flat(uf) : (Matrix{Int32}(undef, 0, nc), Float64[], Int64[], Int64[], 0)
  field_nfree = vcat(nfree_yf, nfree_uf); field_ndir = vcat(ndir_yf, ndir_uf)
  # tables from cell 1 (all fields of a space share the reference element)
  sfy = get_cell_shapefuns(yf[1])
  φ = evaluate(sfy, qpts)[1]                    # nq × ny
  ∇φ = evaluate(∇(sfy), qpts)[1]                # nq × ny of VectorValue (PHYSICAL gradients on cell 1)
  nq, D = size(φ, 1), length(∇φ[1, 1])
  phi_y = permutedims(Float64.(φ))              # ny × nq column-major == [nq][ny]
  phi_u = has_u ? permutedims(Float64.(evaluate(get_cell_shapefuns(uf[1]), qpts)[1])) : zeros(0, nq)
  w = get_cell_weights(quad)[1]; cmap = get_cell_map(trian)
  # Geometry.  Gridap's gradient of the cell map is G[i,j] = ∂x_j/∂ξ_i (the TRANSPOSE of the usual Jacobian), so
  # ∇ₓφ = inv(G)·∇_ξφ and ∇_ξφ = G·∇ₓφ.  G is evaluated at every quadrature point of every cell; the uniform fast path
  # (physical gradients folded into the tables) is taken only when ALL cells carry the same constant G — anything else
  # (a CartesianDiscreteModel with `map`, an unstructured / simplicial mesh) goes through cell_jinvT / cell_detj.
  Gs = evaluate(∇(cmap), get_cell_coordinates(quad))     # per cell: nq TensorValues
  G1 = Gs[1][1]
  tolG = 1e-13 * maximum(abs, Tuple(G1))
  affine = all(all(maximum(abs, Tuple(Gc[q] - Gc[1])) <= tolG for q in 1:nq) for Gc in Gs)
  uniform = affine && all(maximum(abs, Tuple(Gc[1] - G1)) <= tolG for Gc in Gs)
  if uniform
    grad_y = [∇φ[q, a][d] for d in 1:D, a in 1:ny, q in 1:nq]   # [nq][ny][dim], physical
    wdet = Float64.(w .* [abs(det(Gs[1][q])) for q in 1:nq])
    jinvT = Float64[]; detj = Float64[]; geom_nq = 0
  else
    # reference gradients from cell 1's physical ones: ∇_ξφ = G·∇ₓφ
    grad_y = [(Gs[1][q] ⋅ ∇φ[q, a])[d] for d in 1:D, a in 1:ny, q in 1:nq]
    wdet = Float64.(w)
    geom_nq = affine ? 1 : nq
    # cell_jinvT[c][g][d][e] = inv(G)[d,e]  (C row-major [nc][geom_nq][dim][dim] == Julia (e, d, g, c))
    jinvT = Array{Float64}(undef, D, D, geom_nq, nc); detj = Array{Float64}(undef, geom_nq, nc)
    for c in 1:nc, g in 1:geom_nq
      Gi = inv(Gs[c][g])
      for d in 1:D, e in 1:D; jinvT[e, d, g, c] = Gi[d, e]; end
      detj[g, c] = abs(det(Gs[c][g]))
    end
  end
  xq = evaluate(cmap, get_cell_coordinates(quad))   # physical quadrature points per cell
  coef = Array{Float64}(undef, nq, nc, 2)           # [2][nc][nq]
  for c in 1:nc, q in 1:nq; coef[q, c, 1] = target(xq[c][q]); coef[q, c, 2] = source(xq[c][q]); end
  par = ntuple(i -> i <= length(params) ? Float64(params[i]) : 0.0, 8)
  sharded = devices !== nothing
  devs = sharded ? Int32.(devices) : Int32[]
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve celly cellu diry diru phi_y grad_y phi_u wdet coef field_nfree field_ndir jinvT detj devs begin
    d = Desc(3, Int32(integrand), D, ny, nu, nq, nlp.pdemeta.nparam, sharded ? 1 : memspace, device, length(yf),
             nc, num_free_dofs(Ypde), has_u ? num_free_dofs(Ycon) : 0, length(diry), length(diru),
             pointer(celly), pointer(cellu), pointer(diry), pointer(diru),
             pointer(phi_y), pointer(grad_y), pointer(phi_u), pointer(wdet), 2, length(uf), pointer(coef), par,
             pointer(field_nfree), pointer(field_ndir),
             geom_nq == 0 ? Ptr{Float64}(C_NULL) : pointer(jinvT), geom_nq == 0 ? Ptr{Float64}(C_NULL) : pointer(detj), geom_nq,
             Int32((deterministic ? 1 : 0) | (templates ? 0 : 2)))   # PDENLP_FLAG_DETERMINISTIC | PDENLP_FLAG_NO_TEMPLATES
    if sharded
      check(ccall((:pdenlp_sharded_create, LIB), Cint, (Ref{Desc}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), d, length(devs), devs, h))
    else
      check(ccall((:pdenlp_create, LIB), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), d, h))
    end
  end
  info = Ref{Info}()
  check(sharded ? ccall((:pdenlp_sharded_get_info, LIB), Cint, (Ptr{Cvoid}, Ref{Info}), h[], info) :
                  ccall((:pdenlp_get_info, LIB), Cint, (Ptr{Cvoid}, Ref{Info}), h[], info))
  # Gridap's structure stays authoritative; the device structure only cross-checks it
  @assert info[].nnzj == nlp.meta.nnzj && info[].nnzh == nlp.meta.nnzh && info[].nnzh_obj == nlp.pdemeta.nnzh_obj
  if sharded || memspace == 1   # (a device-memspace model returns its structure in device memory)
    rows = Vector{Int64}(undef, nlp.meta.nnzh); cols = similar(rows)
    check(sharded ? ccall((:pdenlp_sharded_hess_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], rows, cols) :
                    ccall((:pdenlp_hess_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], rows, cols))
    @assert rows == nlp.pdemeta.Hrows && cols == nlp.pdemeta.Hcols "COO order differs from hess_structure!"
  end
  hd = Handle(h[], sharded, sharded ? 1 : memspace)
  finalizer(release!, hd)        # the handle frees the device model when it is collected — with the model, see HANDLES
  HANDLES[nlp] = hd
  return nlp
end

handle(nlp) = HANDLES[nlp].ptr
# Entry point of a callback: `pdenlp_<name>` for one device, `pdenlp_sharded_<name>` for the multi-GPU handle (same
# argument lists).  ccall needs the symbol as a constant, hence the macro.
macro pcall(name, argtypes, nlp, args...)
  one = QuoteNode(Symbol("pdenlp_", name)); many = QuoteNode(Symbol("pdenlp_sharded_", name))
  quote
    local hd = HANDLES[$(esc(nlp))]
    check(hd.sharded ? ccall(($many, LIB), Cint, $(esc(argtypes)), hd.ptr, $(map(esc, args)...)) :
                       ccall(($one, LIB), Cint, $(esc(argtypes)), hd.ptr, $(map(esc, args)...)))
  end
end
# callbacks arrive with contiguous whole-vector views (SURVEY §8b); anything else is staged.  DeviceVectors pass through.
dense(v::DeviceVector) = v
dense(v::StridedVector{Float64}) = stride(v, 1) == 1 ? v : collect(v)
dense(v) = collect(Float64, v)

# ---- the callback bodies: @lencheck + Counters exactly as in the reference, then one ccall ----
function obj(nlp::GridapPDENLPModel, x::AbstractVector)                       # src/GridapPDENLPModel.jl:245
  @lencheck nlp.meta.nvar x
  increment!(nlp, :neval_obj)
  f = Ref{Float64}()
  @pcall(obj, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), nlp, dense(x), f)
  return f[]
end

function grad!(nlp::GridapPDENLPModel, x::AbstractVector, g::AbstractVector)   # :257
  @lencheck nlp.meta.nvar x g
  increment!(nlp, :neval_grad)
  out = dense(g)
  @pcall(grad, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), out)
  out === g || copyto!(g, out)
  return g
end

# optional override of the NLPModels default (obj then grad!): one sweep over the cells for both
function objgrad!(nlp::GridapPDENLPModel, x::AbstractVector, g::AbstractVector)
  @lencheck nlp.meta.nvar x g
  HANDLES[nlp].sharded && return obj(nlp, x), grad!(nlp, x, g)   # (no fused entry point on the multi-GPU handle)
  increment!(nlp, :neval_obj); increment!(nlp, :neval_grad)
  out = dense(g); f = Ref{Float64}(0.0)
  @pcall(objgrad, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}), nlp, dense(x), f, out)
  out === g || copyto!(g, out)
  return f[], g
end

function cons_nln!(nlp::GridapPDENLPModel, x::AbstractVector, c::AbstractVector)   # :433
  @lencheck nlp.meta.nvar x
  @lencheck nlp.meta.nnln c
  increment!(nlp, :neval_cons_nln)
  out = dense(c)
  @pcall(cons, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), out)
  out === c || copyto!(c, out)
  return c
end

function jac_nln_coord!(nlp::GridapPDENLPModel, x::AbstractVector, vals::AbstractVector)   # :567
  @lencheck nlp.meta.nvar x
  @lencheck nlp.meta.nln_nnzj vals
  increment!(nlp, :neval_jac_nln)
  out = dense(vals)
  @pcall(jac_coord, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), out)
  out === vals || copyto!(vals, out)
  return vals
end

function hess_coord!(nlp::GridapPDENLPModel, x::AbstractVector{T}, vals::AbstractVector; obj_weight::Real = one(T)) where {T}   # :268
  @lencheck nlp.meta.nvar x
  @lencheck nlp.meta.nnzh vals
  increment!(nlp, :neval_hess)
  out = dense(vals)
  @pcall(hess_coord, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}), nlp, dense(x), C_NULL, obj_weight, out)
  out === vals || copyto!(vals, out)
  return vals
end

function hess_coord!(nlp::GridapPDENLPModel, x::AbstractVector{T}, λ::AbstractVector, vals::AbstractVector;
                     obj_weight::Real = one(T)) where {T}   # :341
  @lencheck nlp.meta.nvar x
  @lencheck nlp.meta.ncon λ
  @lencheck nlp.meta.nnzh vals
  increment!(nlp, :neval_hess)
  out = dense(vals)
  @pcall(hess_coord, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}), nlp, dense(x), dense(λ), obj_weight, out)
  out === vals || copyto!(vals, out)
  return vals
end

function jprod_nln!(nlp::GridapPDENLPModel, x::AbstractVector, v::AbstractVector, Jv::AbstractVector)   # :473
  @lencheck nlp.meta.nvar x v
  @lencheck nlp.meta.nnln Jv
  increment!(nlp, :neval_jprod_nln)      # net effect of the reference's increment/decrement pair (:483-484)
  out = dense(Jv)
  @pcall(jprod, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), dense(v), out)
  out === Jv || copyto!(Jv, out)
  return Jv
end

function jtprod_nln!(nlp::GridapPDENLPModel, x::AbstractVector, v::AbstractVector, Jtv::AbstractVector)   # :505
  @lencheck nlp.meta.nvar x Jtv
  @lencheck nlp.meta.nnln v
  increment!(nlp, :neval_jtprod_nln)
  out = dense(Jtv)
  @pcall(jtprod, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), dense(v), out)
  out === Jtv || copyto!(Jtv, out)
  return Jtv
end

function hprod!(nlp::GridapPDENLPModel, x::AbstractVector{T}, v::AbstractVector, Hv::AbstractVector; obj_weight::Real = one(T)) where {T}   # :303
  @lencheck nlp.meta.nvar x v Hv
  increment!(nlp, :neval_hprod)
  out = dense(Hv)
  @pcall(hprod, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), C_NULL, obj_weight, dense(v), out)
  out === Hv || copyto!(Hv, out)
  return Hv
end

function hprod!(nlp::GridapPDENLPModel, x::AbstractVector{T}, λ::AbstractVector, v::AbstractVector, Hv::AbstractVector;
                obj_weight::Real = one(T)) where {T}   # :413
  @lencheck nlp.meta.nvar x v Hv
  @lencheck nlp.meta.ncon λ
  increment!(nlp, :neval_hprod)
  out = dense(Hv)
  @pcall(hprod, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}), nlp, dense(x), dense(λ), obj_weight, dense(v), out)
  out === Hv || copyto!(Hv, out)
  return Hv
end

end # module

# export_golden.jl — run on a machine WITH Julia 1.6+, Gridap 0.15.5 and PDENLPModels 0.3.5 to dump
# true reference vectors with the same keys as tests/golden/make_golden.py (one .npz per case, via NPZ.jl).
# Not runnable in this repo's container.  Inputs are read from the oracle-made fixtures so that x, λ, v
# are bit-identical on both sides.
using Gridap, PDENLPModels, NLPModels, NPZ, BenchmarkTools

include(joinpath(@__DIR__, "..", "..", "reference", "test", "problems", "burger1d.jl"))   # adjust the path
# README membrane (res-function variant, README.md:58-98)
function membrane(n)
  model = CartesianDiscreteModel((-1, 1, -1, 1), (n, n))
  Xpde = TestFESpace(model, ReferenceFE(lagrangian, Float64, 2); conformity = :H1, dirichlet_tags = "boundary")
  Ypde = TrialFESpace(Xpde, x -> 0.0)
  Xcon = TestFESpace(model, ReferenceFE(lagrangian, Float64, 1); conformity = :H1)
  Ycon = TrialFESpace(Xcon)
  trian = Triangulation(model); dΩ = Measure(trian, 1)
  yd(x) = -x[1]^2; α = 1e-2; ω = π - 1 / 8; h(x) = -sin(ω * x[1]) * sin(ω * x[2])
  f(y, u) = ∫(0.5 * (yd - y) * (yd - y) + 0.5 * α * u * u) * dΩ
  res(y, u, v) = ∫(∇(v) ⊙ ∇(y) - v * u - v * h) * dΩ
  GridapPDENLPModel(zeros(num_free_dofs(Ypde) + num_free_dofs(Ycon)), f, trian, Ypde, Ycon, Xpde, Xcon, res)
end

function dump(name, nlp, fixture)
  g = npzread(fixture)
  x, λ, v, ow = g["x"], g["lam"], g["v"], g["obj_weight"]
  jr, jc = jac_structure(nlp); hr, hc = hess_structure(nlp)
  out = Dict(
    "x" => x, "lam" => λ, "v" => v, "obj_weight" => ow, "obj" => obj(nlp, x), "grad" => grad(nlp, x), "cons" => cons(nlp, x),
    "jac_rows" => Int64.(jr), "jac_cols" => Int64.(jc), "jac_vals" => jac_coord(nlp, x),
    "hess_rows" => Int64.(hr), "hess_cols" => Int64.(hc), "hess_vals" => hess_coord(nlp, x, λ, obj_weight = ow),
    "hess_obj_vals" => hess_coord(nlp, x, obj_weight = ow), "jprod" => jprod(nlp, x, v), "jtprod" => jtprod(nlp, x, λ),
    "hprod" => hprod(nlp, x, λ, v, obj_weight = ow), "nnzh_obj" => nlp.pdemeta.nnzh_obj,
    "t_jac_coord" => @belapsed(jac_coord($nlp, $x)), "t_hess_coord" => @belapsed(hess_coord($nlp, $x, $λ)),
  )
  npzwrite(name * "_reference.npz", out)
end

dump("membrane_n3", membrane(3), joinpath(@__DIR__, "..", "tests", "golden", "membrane_n3.npz"))
dump("burger1d_n6", burger1d(n = 6), joinpath(@__DIR__, "..", "tests", "golden", "burger1d_n6.npz"))

# export_golden.jl — ONE command that turns "parity unpinned" into "pinned by the reference":
#
#     julia --project=<env with Gridap 0.15.5, PDENLPModels 0.3.5, NLPModels, NPZ> julia/export_golden.jl /path/to/PDENLPModels.jl
#
# Run on any machine WITH Julia (this repo's container has none).  For each of the 15 cases of
# tests/golden/make_golden.py it builds the problem with the REFERENCE's own constructor (its test/problems/*.jl files,
# or the README / docs formulation where the case is not a test problem), reads x, λ, v, obj_weight from the committed
# oracle-made fixture tests/golden/<case>.npz (so both sides evaluate at bit-identical points) and writes
# tests/golden/ref/<case>.npz with the SAME keys, from the stock NLPModels API of the unmodified package.
# tests/test_golden.py prefers tests/golden/ref/*.npz when present and reports which provenance it used.
using Gridap, PDENLPModels, NLPModels, NPZ, LinearAlgebra

const REF = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "..", "reference")
for f in ("burger1d", "burger1d_param", "poissonmixed", "poissonmixed2", "poissonparam", "penalizedpoisson", "basicunconstrained",
          "controlsir", "cellincrease", "dynamicsir", "torebrachistochrone", "sis")
  include(joinpath(REF, "test", "problems", f * ".jl"))
end

# README membrane, res-function variant (README.md:58-98; docstring src/GridapPDENLPModel.jl:164-204)
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

# docs/src/poisson-boltzman.md:30-67: Q1 state, L2-conforming Q1 control, one quadrature point
function poissonboltzman(n)
  model = CartesianDiscreteModel((-1, 1, -1, 1), (n, n))
  Xpde = TestFESpace(model, ReferenceFE(lagrangian, Float64, 1); conformity = :H1, dirichlet_tags = "boundary")
  Ypde = TrialFESpace(Xpde, x -> 0.0)
  Xcon = TestFESpace(model, ReferenceFE(lagrangian, Float64, 1); conformity = :L2)
  Ycon = TrialFESpace(Xcon)
  trian = Triangulation(model); dΩ = Measure(trian, 1)
  yd(x) = min(x[1] - 0.25, 0.75 - x[1], x[2] - 0.25, 0.75 - x[2]) >= 0.0 ? 10.0 : 5.0
  α = 1e-4; ω = π - 1 / 8; h(x) = -sin(ω * x[1]) * sin(ω * x[2])
  f(y, u) = ∫(0.5 * (yd - y) * (yd - y) + 0.5 * α * u * u) * dΩ
  res(y, u, v) = ∫(∇(v) ⊙ ∇(y) + (sinh ∘ y) * v - u * v - v * h) * dΩ
  GridapPDENLPModel(zeros(num_free_dofs(Ypde) + num_free_dofs(Ycon)), f, trian, Ypde, Ycon, Xpde, Xcon, res)
end

# BASELINE config 5: test/problems/poissonmixed.jl on the unit cube with a Q1 distributed control (res ... - v*u)
function inversepoisson3d(n)
  model = CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (n, n, n))
  sol(x) = sin(2 * pi * x[1]) * x[2] * x[3]
  f(x) = (2 * pi^2) * sin(2 * pi * x[1]) * x[2] * x[3]
  reffe = ReferenceFE(lagrangian, Float64, 1)
  V0 = TestFESpace(model, reffe; conformity = :H1, dirichlet_tags = "boundary")
  Ug = TrialFESpace(V0, x -> 0.0)
  Xcon = TestFESpace(model, reffe; conformity = :H1); Ycon = TrialFESpace(Xcon)
  trian = Triangulation(model); dΩ = Measure(trian, 2)
  res(k, y, u, v) = ∫(k[1] * ∇(v) ⊙ ∇(y) - v * f * k[2] - v * u)dΩ
  fk(k, y, u) = ∫(0.5 * (sol - y) * (sol - y) + 0.5 * (k[1] - 1.0) * (k[1] - 1.0) + 0.5 * (k[2] - 1.0) * (k[2] - 1.0))dΩ
  nrj = MixedEnergyFETerm(fk, trian, 2, true, Ug, Ycon)
  GridapPDENLPModel(zeros(2 + num_free_dofs(Ug) + num_free_dofs(Ycon)), nrj, Ug, Ycon, V0, Xcon, res)
end

const CASES = [   # names and sizes of tests/golden/make_golden.py
  ("membrane_n3", () -> membrane(3)), ("poissonboltzman_l2_n4", () -> poissonboltzman(4)),
  ("burger1d_n6", () -> burger1d(n = 6)), ("burger1d_param_n6", () -> burger1d_param(n = 6)),
  ("poissonmixed_n3", () -> poissonmixed(n = 3)), ("inversepoisson3d_n2", () -> inversepoisson3d(2)),
  ("penalizedpoisson_n4", () -> penalizedpoisson(n = 4)), ("basicunconstrained_n3", () -> basicunconstrained(n = 3)),
  ("controlsir_n6", () -> controlsir(n = 6)), ("cellincrease_n5", () -> cellincrease(n = 5)),
  ("dynamicsir_n6", () -> dynamicsir(n = 6)), ("torebrachistochrone_n4", () -> torebrachistochrone(n = 4)),
  ("poissonmixed2_n3", () -> poissonmixed2(n = 3)), ("poissonparam_n3", () -> poissonparam(n = 3)), ("sis_n6", () -> sis(n = 6)),
]

function dump(name, nlp, fixture, outdir)
  g = npzread(fixture)
  x, v, ow = g["x"], g["v"], Float64(g["obj_weight"])
  @assert length(x) == nlp.meta.nvar "$name: nvar $(nlp.meta.nvar) != fixture $(length(x)) — the host layer of the repo builds a different space"
  hr, hc = hess_structure(nlp)
  out = Dict{String, Any}("x" => x, "v" => v, "obj_weight" => ow, "obj" => obj(nlp, x), "grad" => grad(nlp, x),
                          "nnzh_obj" => Int64(nlp.pdemeta.nnzh_obj), "provenance" => "PDENLPModels.jl (reference)")
  if nlp.meta.ncon == 0   # unconstrained: the keys of make_golden.py's unconstrained branch
    no = nlp.pdemeta.nnzh_obj
    out["unconstrained"] = true
    out["hess_rows"] = Int64.(hr[1:no]); out["hess_cols"] = Int64.(hc[1:no])
    out["hess_obj_vals"] = hess_coord(nlp, x, obj_weight = ow)[1:no]
    out["hprod"] = hprod(nlp, x, v, obj_weight = ow)
  else
    λ = g["lam"]
    jr, jc = jac_structure(nlp)
    out["lam"] = λ; out["cons"] = cons(nlp, x)
    out["jac_rows"] = Int64.(jr); out["jac_cols"] = Int64.(jc); out["jac_vals"] = jac_coord(nlp, x)
    out["hess_rows"] = Int64.(hr); out["hess_cols"] = Int64.(hc)
    out["hess_vals"] = hess_coord(nlp, x, λ, obj_weight = ow); out["hess_obj_vals"] = hess_coord(nlp, x, obj_weight = ow)
    out["jprod"] = jprod(nlp, x, v); out["jtprod"] = jtprod(nlp, x, λ); out["hprod"] = hprod(nlp, x, λ, v, obj_weight = ow)
  end
  npzwrite(joinpath(outdir, name * ".npz"), out)
  println(name, ": nvar ", nlp.meta.nvar, " ncon ", nlp.meta.ncon, " nnzj ", nlp.meta.nnzj, " nnzh ", nlp.meta.nnzh)
end

function main()
  golden = joinpath(@__DIR__, "..", "tests", "golden")
  outdir = joinpath(golden, "ref")
  mkpath(outdir)
  for (name, mk) in CASES
    dump(name, mk(), joinpath(golden, name * ".npz"), outdir)
  end
  println("wrote ", length(CASES), " reference fixtures to ", outdir, "; now run: python -m pytest tests/test_golden.py -q -rA")
end
main()

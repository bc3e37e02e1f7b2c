"""Problem catalogue: the closed, named set of weak-form integrands
 and the BASELINE.json configurations built on them.

Each ``ProblemSpec`` is what the user of the reference would have written with
Gridap (mesh, spaces, quadrature degree, f, res) — here reduced to the named
integrand id plus the coefficient functions, which the flattening layer samples
at the quadrature points.  Formulas follow the reference files cited per
problem; paths are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from .cartesian import CartesianDiscreteModel, LagrangianFESpace

# integrand ids — keep in sync with include/pdenlp_cuda.h
MEMBRANE = 1
POISSON_BOLTZMANN = 2
BURGER_LIN = 3
BURGER_NL = 4
KAPPA_POISSON = 5
PENALIZED_POISSON = 6
BASIC_UNCONSTRAINED = 7
CONTROL_SIR = 8
CELL_INCREASE = 9
DYNAMIC_SIR = 10
TORE_BRACHISTOCHRONE = 11
POISSON_PARAM = 12
SIS = 13
KAPPA_POISSON2 = 14

INTEGRAND_NAMES = {
    MEMBRANE: "membrane",
    POISSON_BOLTZMANN: "poisson_boltzmann",
    BURGER_LIN: "burger_lin",
    BURGER_NL: "burger_nl",
    KAPPA_POISSON: "kappa_poisson",
    PENALIZED_POISSON: "penalized_poisson",
    BASIC_UNCONSTRAINED: "basic_unconstrained",
    CONTROL_SIR: "control_sir",
    CELL_INCREASE: "cell_increase",
    DYNAMIC_SIR: "dynamic_sir",
    TORE_BRACHISTOCHRONE: "tore_brachistochrone",
    POISSON_PARAM: "poisson_param",
    SIS: "sis",
    KAPPA_POISSON2: "kappa_poisson2",
}


@dataclass
class ProblemSpec:
    name: str
    integrand: int
    model: CartesianDiscreteModel
    Ypde: LagrangianFESpace
    Ycon: Optional[LagrangianFESpace]
    degree: int
    nparam: int = 0
    params: Sequence[float] = ()
    # coefficient fields: callables X (npts, dim) -> (npts,)
    target: Optional[Callable] = None  # yd / sol
    source: Optional[Callable] = None  # h / f
    # True: the constraint is an AffineFEOperator (linear API: cons_lin!, jac_lin_coord!, ...;
    # Jacobian structure = findnz(A); empty constraint-Hessian structure)
    affine: bool = False
    # True: no constraint at all (the reference's unconstrained constructor,
    # src/additional_constructors.jl:1-66): ncon = 0, nnzj = 0, nnzh = nnzh_obj
    unconstrained: bool = False
    # True: the objective is a NoFETerm, f(kappa) without any integral (src/additional_obj_terms.jl:59-106):
    # no FE block in the objective Hessian, kappa-block p x p lower triangle, zero FE gradient
    no_fe_obj: bool = False

    @property
    def nvar(self) -> int:
        return self.nparam + self.Ypde.nfree + (self.Ycon.nfree if self.Ycon is not None else 0)

    @property
    def ncon(self) -> int:
        return 0 if self.unconstrained else self.Ypde.nfree


@dataclass
class MFProblemSpec:
    """Multi-field variant: Ypde / Ycon are MultiFieldFESpaces of scalar
    Lagrangian fields of one element type each; x = [y_1 .. y_nfy ; u_1 .. u_nfu] (consecutive
    multi-field numbering), the test space has one field per state field."""
    name: str
    integrand: int
    model: CartesianDiscreteModel
    Yfields: Sequence[LagrangianFESpace]
    Ufields: Sequence[LagrangianFESpace]
    degree: int
    params: Sequence[float] = ()
    target: Optional[Callable] = None
    source: Optional[Callable] = None
    unconstrained: bool = False
    nparam: int = 0
    affine: bool = False
    no_fe_obj: bool = False
    is_multifield: bool = True

    @property
    def nvar(self) -> int:
        return sum(f.nfree for f in self.Yfields) + sum(f.nfree for f in self.Ufields)

    @property
    def ncon(self) -> int:
        return 0 if self.unconstrained else sum(f.nfree for f in self.Yfields)


_OMEGA = math.pi - 1.0 / 8.0


def _h_membrane(X):
    return -np.sin(_OMEGA * X[:, 0]) * np.sin(_OMEGA * X[:, 1])


def controlelasticmembrane(n: int = 100, affine: bool = False, degree: int = 1, state_order: int = 2, dim: int = 2,
                           mesh_map: Optional[Callable] = None) -> ProblemSpec:
    """README 'Control elastic membrane' (README.md:58-98; docstring
    src/GridapPDENLPModel.jl:164-204): Q2 state with Dirichlet boundary, Q1
    control, degree 1, res-function (FEOperatorFromWeakForm) variant.
    ``affine=True`` is the test-suite variant built on an AffineFEOperator
    (test/problems/controlelasticmembrane.jl:48-53): same PDE, linear API.
    ``degree`` is the user's `Measure(trian, degree)` (1 in the README; 2 or 3 give 2x2 Gauss points),
    ``state_order`` the order of the state element (2 in the README); ``dim = 3`` is the same problem on (-1,1)^3
    (yd and h keep depending on x1, x2 only).  ``mesh_map``: Gridap's ``CartesianDiscreteModel(...; map = f)`` —
    the vertices are mapped, the cells become general Q1 images (per-cell geometry path)."""
    model = CartesianDiscreteModel((-1, 1) * dim, (n,) * dim, map=mesh_map)
    Ypde = LagrangianFESpace(model, state_order, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, "H1")
    return ProblemSpec(
        name="Control elastic membrane" if not affine else "controlelasticmembrane1",
        integrand=MEMBRANE, model=model, Ypde=Ypde, Ycon=Ycon, degree=degree,
        params=(1e-2,),
        target=lambda X: -X[:, 0] ** 2,
        source=_h_membrane,
        affine=affine,
    )


def poissonboltzman2d(n: int = 100, control_conformity: str = "L2", degree: int = 1, dim: int = 2,
                      state_order: int = 1, mesh_map: Optional[Callable] = None) -> ProblemSpec:
    """docs/src/poisson-boltzman.md:30-67: Q1 state, Q1 control (L2 in the
    doc), degree 1, sinh reaction term (``dim = 3``: the same on (-1,1)^3)."""
    model = CartesianDiscreteModel((-1, 1) * dim, (n,) * dim, map=mesh_map)
    Ypde = LagrangianFESpace(model, state_order, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, control_conformity)

    def yd(X):
        m = np.minimum(np.minimum(X[:, 0] - 0.25, 0.75 - X[:, 0]), np.minimum(X[:, 1] - 0.25, 0.75 - X[:, 1]))
        return np.where(m >= 0.0, 10.0, 5.0)

    return ProblemSpec(
        name="poissonboltzman2d",
        integrand=POISSON_BOLTZMANN, model=model, Ypde=Ypde, Ycon=Ycon, degree=degree,
        params=(1e-4,), target=yd, source=_h_membrane,
    )


def _burger_spaces(n, mesh_map=None):
    model = CartesianDiscreteModel((0, 1), n, map=mesh_map)
    # dirichlet_tags = ["diri0", "diri1"] = entities [1], [2]; values 0 and -1
    Ypde = LagrangianFESpace(model, 1, "H1", [1, 2], {1: 0.0, 2: -1.0})
    Ycon = LagrangianFESpace(model, 1, "H1")
    return model, Ypde, Ycon


def burger1d(n: int = 512, degree: int = 1, mesh_map: Optional[Callable] = None) -> ProblemSpec:
    """test/problems/burger1d.jl:1-52 (the test-suite variant: dt(y,v) = y' v
    is linear, src/util_functions.jl:109-110)."""
    model, Ypde, Ycon = _burger_spaces(n, mesh_map)
    nu = 0.08
    return ProblemSpec(
        name="burger1d",
        integrand=BURGER_LIN, model=model, Ypde=Ypde, Ycon=Ycon, degree=degree,
        params=(1e-2, nu),
        target=lambda X: -X[:, 0] ** 2,
        source=lambda X: 2 * (nu + X[:, 0] ** 3),
    )


def burger1d_param(n: int = 512, degree: int = 1, mesh_map: Optional[Callable] = None) -> ProblemSpec:
    """test/problems/burger1d_param.jl:1-54: nonlinear convection v*y*y',
    one algebraic unknown."""
    model, Ypde, Ycon = _burger_spaces(n, mesh_map)
    nu = 0.08
    return ProblemSpec(
        name="burger1d_param",
        integrand=BURGER_NL, model=model, Ypde=Ypde, Ycon=Ycon, degree=degree,
        nparam=1, params=(1e-2, nu, 0.08),
        target=lambda X: -X[:, 0] ** 2,
        source=lambda X: 2 * (nu + X[:, 0] ** 3),
    )


def poissonmixed(n: int = 3, dim: int = 2, control: bool = False, mesh_map: Optional[Callable] = None) -> ProblemSpec:
    """test/problems/poissonmixed.jl:14-48 (2 algebraic unknowns, degree 2);
    ``dim=3, control=True`` is BASELINE config 5 (inverse Poisson with kappa +
    3-D Q1 Poisson control)."""
    dom = (0, 1) * dim
    model = CartesianDiscreteModel(dom, (n,) * dim, map=mesh_map)
    Ypde = LagrangianFESpace(model, 1, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, "H1") if control else None

    def sol(X):
        s = np.sin(2 * np.pi * X[:, 0]) * X[:, 1]
        return s * X[:, 2] if dim == 3 else s

    def f(X):
        s = (2 * np.pi ** 2) * np.sin(2 * np.pi * X[:, 0]) * X[:, 1]
        return s * X[:, 2] if dim == 3 else s

    return ProblemSpec(
        name="poissonmixed" if not control else "inversePoissonproblem",
        integrand=KAPPA_POISSON, model=model, Ypde=Ypde, Ycon=Ycon, degree=2,
        nparam=2, params=(), target=sol, source=f,
    )


def poissonparam(n: int = 3) -> ProblemSpec:
    """test/problems/poissonparam.jl:14-46: one algebraic unknown, objective NoFETerm(0.5*|k - 1|^2),
    res = k1*grad(v).grad(y) - v*f, Dirichlet data = the manufactured solution (Q1, degree 2)."""
    model = CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    sol = lambda X: np.sin(2 * np.pi * X[:, 0]) * X[:, 1]
    Ypde = LagrangianFESpace(model, 1, "H1", "boundary", sol)
    return ProblemSpec(
        name="poissonparam", integrand=POISSON_PARAM, model=model, Ypde=Ypde, Ycon=None, degree=2, nparam=1,
        source=lambda X: (2 * np.pi ** 2) * np.sin(2 * np.pi * X[:, 0]) * X[:, 1], no_fe_obj=True,
    )


def poissonmixed2(n: int = 3) -> ProblemSpec:
    """test/problems/poissonmixed2.jl:14-49 as evidently intended: f = k1*(sol - y)^2 + 0.5*(k2 - 1)^2 with a plain
    MixedEnergyFETerm (inde = false: the objective kappa-block has cross rows with y), res as poissonmixed.
    (The file passes the test space V0 in the Ycon slot of MixedEnergyFETerm; it is on the reference's
    local-only list.)"""
    model = CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    Ypde = LagrangianFESpace(model, 1, "H1", "boundary", lambda X: np.sin(2 * np.pi * X[:, 0]) * X[:, 1])
    return ProblemSpec(
        name="poissonmixed2", integrand=KAPPA_POISSON2, model=model, Ypde=Ypde, Ycon=None, degree=2, nparam=2,
        target=lambda X: np.sin(2 * np.pi * X[:, 0]) * X[:, 1],
        source=lambda X: (2 * np.pi ** 2) * np.sin(2 * np.pi * X[:, 0]) * X[:, 1],
    )


def penalizedpoisson(n: int = 16) -> ProblemSpec:
    """test/problems/penalizedpoisson.jl:14-42: min 0.5*int |grad y|^2 - w*y, y = 0 on the boundary
    (Q1, degree 2, unconstrained; w = 1)."""
    model = CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    Ypde = LagrangianFESpace(model, 1, "H1", "boundary", 0.0)
    return ProblemSpec(
        name="penalizedpoisson", integrand=PENALIZED_POISSON, model=model, Ypde=Ypde, Ycon=None, degree=2,
        source=lambda X: np.ones(X.shape[0]), unconstrained=True,
    )


def basicunconstrained(n: int = 16) -> ProblemSpec:
    """test/problems/basicunconstrained.jl:1-30: min 0.5*int (ubis-u)^2 + y^2 over the two fields
    [y (Q1, Dirichlet 0), u (Q1)], degree 2, unconstrained."""
    model = CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    Ypde = LagrangianFESpace(model, 1, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, "H1")
    return ProblemSpec(
        name="basicunconstrained", integrand=BASIC_UNCONSTRAINED, model=model, Ypde=Ypde, Ycon=Ycon, degree=2,
        target=lambda X: X[:, 0] ** 2 + X[:, 1] ** 2, unconstrained=True,
    )


def _ode_state_fields(model, x0):
    # VI, VS: Q1, Dirichlet at t = 0 ("diri0" = entity 1) with the initial values x0
    return [LagrangianFESpace(model, 1, "H1", [1], float(v)) for v in x0]


def controlsir(n: int = 10, x0=(1.0, 2.0), a: float = 0.2, b: float = 0.1, mu: float = 0.1, T: float = 1.0) -> MFProblemSpec:
    """test/problems/controlsir.jl:1-38: SIR ODE system as constraints on the two fields (I, S),
    f = 0.5*I^2, no control; params = (a, b, mu)."""
    model = CartesianDiscreteModel((0, T), n)
    return MFProblemSpec(name="controlsir", integrand=CONTROL_SIR, model=model, Yfields=_ode_state_fields(model, x0),
                         Ufields=[], degree=1, params=(a, b, mu))


def sis(n: int = 10, x0=(1.0, 2.0), a: float = 0.2, b: float = 0.7, T: float = 1.0) -> MFProblemSpec:
    """test/problems/sis.jl:1-35: SIS ODE system on the two fields (I, S), objective NoFETerm() = 0, no control;
    params = (a, b)."""
    model = CartesianDiscreteModel((0, T), n)
    return MFProblemSpec(name="sis", integrand=SIS, model=model, Yfields=_ode_state_fields(model, x0),
                         Ufields=[], degree=1, params=(a, b), no_fe_obj=True)


def cellincrease(n: int = 10, x0=(0.6, 0.1), T: float = 7.0) -> MFProblemSpec:
    """test/problems/cellincrease.jl:1-46: two state fields (cf, pf), one L2 control; f = kp*pf;
    params = (kp, kr) = (1.01, 2.03)."""
    model = CartesianDiscreteModel((0, T), n)
    return MFProblemSpec(name="cellincrease", integrand=CELL_INCREASE, model=model, Yfields=_ode_state_fields(model, x0),
                         Ufields=[LagrangianFESpace(model, 1, "L2")], degree=1, params=(1.01, 2.03))


def dynamicsir(n: int = 10, x0=(1.0, 2.0), T: float = 1.0) -> MFProblemSpec:
    """test/problems/dynamicsir.jl:1-64: state (I, S), controls (bf, cf) in H1;
    f = 0.5*(bf*w0 - w1)^2 + 0.5*(cf*w0 - w2)^2 with w0 = 1 + t, w1 = 1, w2 = 2; params = (w1, w2)."""
    model = CartesianDiscreteModel((0, T), n)
    return MFProblemSpec(name="dynamicsir", integrand=DYNAMIC_SIR, model=model, Yfields=_ode_state_fields(model, x0),
                         Ufields=[LagrangianFESpace(model, 1, "H1"), LagrangianFESpace(model, 1, "H1")], degree=1,
                         params=(1.0, 2.0), target=lambda X: 1.0 + X[:, 0])


def torebrachistochrone(n: int = 3, a: float = 1.0, c: float = 3.0) -> MFProblemSpec:
    """test/problems/torebrachistochrone.jl:1-67: geodesic on a torus, fields (phi, theta) fixed at both
    ends (0 and pi), unconstrained; f = a^2 phi'^2 + (c + a cos(phi))^2 theta'^2; params = (a, c)."""
    model = CartesianDiscreteModel((0, 1), n)
    fields = [LagrangianFESpace(model, 1, "H1", [1, 2], {1: 0.0, 2: math.pi}) for _ in range(2)]
    return MFProblemSpec(name="torebrachistochrone", integrand=TORE_BRACHISTOCHRONE, model=model, Yfields=fields,
                         Ufields=[], degree=1, params=(a, c), unconstrained=True)


BASELINE_CONFIGS = {
    "membrane_n100": lambda: controlelasticmembrane(100),
    "membrane_n2048": lambda: controlelasticmembrane(2048),
    "poissonboltzman2d_n1024": lambda: poissonboltzman2d(1024),
    "burger1d_1M": lambda: burger1d_param(1_000_000),
    "kappa_poisson3d_128": lambda: poissonmixed(128, dim=3, control=True),
}

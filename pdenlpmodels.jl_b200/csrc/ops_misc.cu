// ops_misc.cu — one of the translation units that instantiate the closed set of (integrand, element) kernels
// (split only so that they compile in parallel; see ops.cuh).
#include "ops.cuh"

namespace pdenlp {
namespace host {

Ops* select_ops_part3(const pdenlp_desc& d) {
  PDENLP_SELECT_PROLOGUE(d);
  CASE(PDENLP_KAPPA_POISSON2, 2, 4, 0, 4, 2, KappaPoisson2<2>)
  CASE(PDENLP_PENALIZED_POISSON, 2, 4, 0, 4, 0, PenalizedPoisson<2>)
  CASE(PDENLP_BASIC_UNCONSTRAINED, 2, 4, 4, 4, 0, BasicUnconstrained<2>)
  CASE(PDENLP_POISSON_PARAM, 2, 4, 0, 4, 1, PoissonParam<2>)   // NoFETerm objective
  // 1-D: Burgers with 1 or 2 points
  CASE(PDENLP_BURGER_LIN, 1, 2, 2, 1, 0, BurgerLin)
  CASE(PDENLP_BURGER_NL, 1, 2, 2, 1, 1, BurgerNl)
  CASE(PDENLP_BURGER_LIN, 1, 2, 2, 2, 0, BurgerLin)
  CASE(PDENLP_BURGER_NL, 1, 2, 2, 2, 1, BurgerNl)
  // multi-field ODE problems: 1-D, P1 state fields, one quadrature point
  CASE(PDENLP_CONTROL_SIR, 1, 2, 0, 1, 0, ControlSir)
  CASE(PDENLP_CELL_INCREASE, 1, 2, 2, 1, 0, CellIncrease)
  CASE(PDENLP_DYNAMIC_SIR, 1, 2, 2, 1, 0, DynamicSir)
  CASE(PDENLP_TORE_BRACHISTOCHRONE, 1, 2, 0, 1, 0, ToreBrachistochrone)
  CASE(PDENLP_SIS, 1, 2, 0, 1, 0, Sis)                          // NoFETerm objective
  // test-only integrand with deliberately wrong shortcut traits: pdenlp_create must fail its self-check
  CASE(PDENLP_SELFTEST_WRONG_TRAIT, 2, 4, 4, 1, 0, WrongTraitPB<2>)
  CASE(PDENLP_SELFTEST_WRONG_TRAIT, 2, 4, 4, 4, 0, WrongTraitPB<2>)
  return nullptr;
}

}  // namespace host
}  // namespace pdenlp

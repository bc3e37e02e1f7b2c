// ops_membrane4.cu — one of the translation units that instantiate the closed set of (integrand, element) kernels
// (split only so that they compile in parallel; see ops.cuh).
#include "ops.cuh"

namespace pdenlp {
namespace host {

Ops* select_ops_part1(const pdenlp_desc& d) {
  PDENLP_SELECT_PROLOGUE(d);
  // the membrane problem with the user's quadrature degree raised to 2 or 3 (2 Gauss points per direction)
  CASE(PDENLP_MEMBRANE, 2, 9, 4, 4, 0, Membrane<2>)
  CASE(PDENLP_MEMBRANE, 2, 4, 4, 4, 0, Membrane<2>)
  // Poisson-Boltzmann with a Q2 state
  CASE(PDENLP_POISSON_BOLTZMANN, 2, 9, 4, 1, 0, PoissonBoltzmann<2>)
  CASE(PDENLP_POISSON_BOLTZMANN, 2, 9, 4, 4, 0, PoissonBoltzmann<2>)
  return nullptr;
}

}  // namespace host
}  // namespace pdenlp

// ops_membrane.cu — one of the translation units that instantiate the closed set of (integrand, element) kernels
// (split only so that they compile in parallel; see ops.cuh).
#include "ops.cuh"

namespace pdenlp {
namespace host {

Ops* select_ops_part0(const pdenlp_desc& d) {
  PDENLP_SELECT_PROLOGUE(d);
  CASE(PDENLP_MEMBRANE, 2, 9, 4, 1, 0, Membrane<2>)
  CASE(PDENLP_MEMBRANE, 2, 4, 4, 1, 0, Membrane<2>)
  CASE(PDENLP_POISSON_BOLTZMANN, 2, 4, 4, 1, 0, PoissonBoltzmann<2>)
  CASE(PDENLP_POISSON_BOLTZMANN, 2, 4, 4, 4, 0, PoissonBoltzmann<2>)
  return nullptr;
}

}  // namespace host
}  // namespace pdenlp

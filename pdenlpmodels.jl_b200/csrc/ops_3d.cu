// ops_3d.cu — one of the translation units that instantiate the closed set of (integrand, element) kernels
// (split only so that they compile in parallel; see ops.cuh).
#include "ops.cuh"

namespace pdenlp {
namespace host {

Ops* select_ops_part4(const pdenlp_desc& d) {
  PDENLP_SELECT_PROLOGUE(d);
  // membrane / Poisson-Boltzmann on 3-D Q1/Q1 meshes, one or 2x2x2 quadrature points
  CASE(PDENLP_MEMBRANE, 3, 8, 8, 1, 0, Membrane<3>)
  CASE(PDENLP_MEMBRANE, 3, 8, 8, 8, 0, Membrane<3>)
  CASE(PDENLP_POISSON_BOLTZMANN, 3, 8, 8, 1, 0, PoissonBoltzmann<3>)
  return nullptr;
}

}  // namespace host
}  // namespace pdenlp

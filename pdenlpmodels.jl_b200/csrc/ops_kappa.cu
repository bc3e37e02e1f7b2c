// ops_kappa.cu — one of the translation units that instantiate the closed set of (integrand, element) kernels
// (split only so that they compile in parallel; see ops.cuh).
#include "ops.cuh"

namespace pdenlp {
namespace host {

Ops* select_ops_part2(const pdenlp_desc& d) {
  PDENLP_SELECT_PROLOGUE(d);
  using KP2 = KappaPoisson<2, false>;
  using KP2U = KappaPoisson<2, true>;
  using KP3U = KappaPoisson<3, true>;
  CASE(PDENLP_KAPPA_POISSON, 2, 4, 0, 4, 2, KP2)
  CASE(PDENLP_KAPPA_POISSON, 2, 4, 4, 4, 2, KP2U)
  CASE(PDENLP_KAPPA_POISSON, 3, 8, 8, 8, 2, KP3U)
  return nullptr;
}

}  // namespace host
}  // namespace pdenlp

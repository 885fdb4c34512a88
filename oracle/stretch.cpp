// stretch.cpp -- stretched-mesh pieces of the oracle: stretching (src/stretching.f90),
// matrice_refinement (src/poisson.f90:1814-2249), inversion5_v1/v2 (src/tools.f90:1225-1498).
// (TEST INFRASTRUCTURE)  -- restated in a later step; uniform meshes (istret=0) never get here.
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {
void stretching(Stretch &, int, int, int, int, bool) { throw std::runtime_error("oracle: stretching not restated yet"); }
void Poisson::matrice_refinement() { throw std::runtime_error("oracle: matrice_refinement not restated yet"); }
void inversion5_v1(const std::vector<cplx> &, cplx *, int, int, int) { throw std::runtime_error("oracle: inversion5_v1 not restated yet"); }
void inversion5_v2(std::vector<cplx> &, cplx *, int, int, int) { throw std::runtime_error("oracle: inversion5_v2 not restated yet"); }
}  // namespace x3do

// kernels_fluid.cu -- WCA and square-well fluids on shared-memory cell lists.
#include "make_set.cuh"
#include "sys_cell_fluid.cuh"
namespace sadmc {
KernelSet kernels_cell_fluid(bool square_well, const DevParams& P) {
  return square_well ? make_set<CellFluidSys<true>, true>(P) : make_set<CellFluidSys<false>, true>(P);
}
} // namespace sadmc

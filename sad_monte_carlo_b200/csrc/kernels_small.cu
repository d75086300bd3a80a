// kernels_small.cu -- one thread per walker systems: Ising and the analytic test systems.
#include "make_set.cuh"
#include "sys_fake.cuh"
#include "sys_ising.cuh"
namespace sadmc {
KernelSet kernels_ising(const DevParams& P) { return make_set<IsingSys, true>(P); }
KernelSet kernels_fake(const DevParams& P) { return make_set<FakeSys, true>(P); }
KernelSet kernels_two_wells(const DevParams& P) { return make_set<TwoWellsSys, true>(P); }
KernelSet kernels_erfinv(const DevParams& P) { return make_set<ErfInvSys, true>(P); }
} // namespace sadmc

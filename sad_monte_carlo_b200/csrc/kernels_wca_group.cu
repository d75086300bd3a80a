// kernels_wca_group.cu -- WCA fluid, 4 / 8 / 16 lanes per walker (sys_wca_group.cuh), reference and fast arithmetic.
#include "make_set.cuh"
#include "sys_wca_group.cuh"
namespace sadmc {
bool kernels_wca_group(int G, bool fast, const DevParams& P, KernelSet* out) {
  switch (G) {
    case 4: *out = fast ? make_set<WcaGroupSys<4, true>>(P) : make_set<WcaGroupSys<4, false>>(P); return true;
    case 8: *out = fast ? make_set<WcaGroupSys<8, true>>(P) : make_set<WcaGroupSys<8, false>>(P); return true;
    case 16: *out = fast ? make_set<WcaGroupSys<16, true>>(P) : make_set<WcaGroupSys<16, false>>(P); return true;
  }
  return false;
}
} // namespace sadmc

// sys_limits.hpp -- compile-time capacities of the register-resident test systems (shared by the
// kernels and by the host-side parameter checks in engine.cu).
#pragma once
namespace sadmc {
constexpr int FAKE_MAX_DIM = 16;
constexpr int TW_MAX_DIM = 48;
constexpr int ERFINV_MAX_DIM = 32;
} // namespace sadmc

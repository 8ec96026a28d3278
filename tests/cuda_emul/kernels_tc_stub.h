// Stand-in for clonealign_b200/csrc/kernels_tc.cuh under the CPU emulation (tests only): the tcgen05 / TMA kernels
// cannot be emulated, so the tensor path reports itself unavailable and AUTO resolves to the CUDA-core kernels.
#pragma once
#include <stdexcept>
#include "common.cuh"

namespace ca {
constexpr bool kTcAvailable = false;
struct TcPlan {
  bool ok = false;
  int fsplit = 1, nsplit = 1;
};
inline void tc_plan_create(TcPlan&, int, int64_t, int64_t, int, int64_t, int, int, __nv_bfloat16*, __nv_bfloat16*, __half*) {
  throw std::runtime_error("the tcgen05 path is not available under the CPU emulation");
}
inline void tc_plan_destroy(TcPlan& p) { p.ok = false; }
inline void tc_launch_fwd(const TcPlan&, const float*, const float*, const float*, float*, cudaStream_t) {
  throw std::runtime_error("the tcgen05 path is not available under the CPU emulation");
}
inline void tc_launch_bwd(const TcPlan&, const float*, const float*, const float*, float*, cudaStream_t) {
  throw std::runtime_error("the tcgen05 path is not available under the CPU emulation");
}
}  // namespace ca

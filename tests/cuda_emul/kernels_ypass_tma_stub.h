// Emulation counterpart of clonealign_b200/csrc/kernels_ypass_tma.cuh (TEST INFRASTRUCTURE ONLY): the tensor-copy Y pass is not emulated.
#pragma once
#include <stdexcept>
#include <vector>
namespace ca {
constexpr bool kY7Available = false;
struct Y7Plan { bool ok = false; int RB = 0; std::vector<int> tiles, offs; const int* d_tiles = nullptr; const int* d_offs = nullptr; };
inline void y7_plan_tiles(Y7Plan&, int64_t, int64_t, int, int) {}
inline void y7_plan_create(Y7Plan&, const void*, int64_t, int64_t) { throw std::runtime_error("the tensor-copy Y pass is not available under the CPU emulation"); }
inline void y7_launch(const Y7Plan&, unsigned, cudaStream_t, int64_t, int64_t, int, int, int, int, const float*, const float*, float*, float*) {
  throw std::runtime_error("the tensor-copy Y pass is not available under the CPU emulation");
}
inline cudaError_t y7_set_attributes() { return cudaSuccess; }
}  // namespace ca

// strip_s16.cuh -- placeholder while the packed s16x2 kernel is being brought up: constants only.
#pragma once
#include "strip_common.cuh"
#include "strip_s32.cuh"
namespace b200 {
constexpr int kR16 = 8;
constexpr int kSH16 = 64 * kR16;
#define B200_NO_S16 1
template <int R, bool SW, bool TRACK>
__global__ void strip_kernel_s16(const StripParams p) {}
}  // namespace b200

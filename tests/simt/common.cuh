// tests/simt/common.cuh — what csrc/common.cuh gives a kernel source, for the CPU lane emulation: the real header's host part
// (parameter blocks, format ids; AURORA_REAL_COMMON is its path) plus the device helpers on top of simt.hpp.
#pragma once
#include "cuda_runtime.h"
#include AURORA_REAL_COMMON
#include "simt.hpp"

namespace aurora {
inline int lane_id() { return simt_lane(); }
}  // namespace aurora

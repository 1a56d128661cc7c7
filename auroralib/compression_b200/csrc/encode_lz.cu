// encode_lz.cu — encoder kernels (placeholder until the match-search kernels land in this round).
#include "common.cuh"

namespace aurora {
cudaError_t launch_encode_lz(const EncodeParams&, int, cudaStream_t) { return cudaErrorNotSupported; }
size_t encode_scratch_per_warp(int) { return 0; }
int encode_resident_warps(int sm_count) { return sm_count * 8; }
}  // namespace aurora

// tests/simt/stage.cuh — csrc/stage.cuh for the CPU lane emulation: the shared-memory accessors on the emulated warp's array,
// then the REAL header (AURORA_REAL_STAGE: constants, the staged input stream, the reciprocal table, the warp scan) with its
// own PTX accessors switched off (AURORA_SIMT).
#pragma once
#include "common.cuh"

namespace aurora {
inline uint32_t lds_u8(uint32_t a) { return *simt_smem(a, 1); }
inline uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    memcpy(&v, simt_smem(a, 2), 2);
    return v;
}
inline uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    memcpy(&v, simt_smem(a, 4), 4);
    return v;
}
inline uint2 lds_u64(uint32_t a) {
    uint2 v;
    memcpy(&v, simt_smem(a, 8), 8);
    return v;
}
inline uint4 lds_u128(uint32_t a) {
    uint4 v;
    memcpy(&v, simt_smem(a, 16), 16);
    return v;
}
inline void sts_u8(uint32_t a, uint32_t v) { *simt_smem(a, 1) = uint8_t(v); }
inline void sts_u16(uint32_t a, uint32_t v) {
    const uint16_t x = uint16_t(v);
    memcpy(simt_smem(a, 2), &x, 2);
}
inline void sts_u32(uint32_t a, uint32_t v) { memcpy(simt_smem(a, 4), &v, 4); }
inline void sts_u64(uint32_t a, uint32_t x, uint32_t y) {
    const uint32_t v[2] = {x, y};
    memcpy(simt_smem(a, 8), v, 8);
}
inline void sts_u128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    const uint32_t v[4] = {x, y, z, w};
    memcpy(simt_smem(a, 16), v, 16);
}
}  // namespace aurora

#define AURORA_SIMT 1
#include AURORA_REAL_STAGE

// tests/simt/stage.cuh — the shared-memory accessors and warp primitives of csrc/stage.cuh for the CPU lane emulation.
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
inline void sts_u8(uint32_t a, uint32_t v) { *simt_smem(a, 1) = uint8_t(v); }
inline void sts_u16(uint32_t a, uint32_t v) {
    const uint16_t x = uint16_t(v);
    memcpy(simt_smem(a, 2), &x, 2);
}
inline void sts_u32(uint32_t a, uint32_t v) { memcpy(simt_smem(a, 4), &v, 4); }
inline void sts_u128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    const uint32_t v[4] = {x, y, z, w};
    memcpy(simt_smem(a, 16), v, 16);
}
// ceil(2^20 / d) (csrc/stage.cuh: the constant-memory reciprocal table of the self-overlapping copies)
struct RcpTable {
    uint32_t v[512];
    constexpr RcpTable() : v() {
        for (uint32_t d = 1; d < 512; ++d) v[d] = ((1u << 20) + d - 1) / d;
    }
};
static const RcpTable c_rcp = RcpTable();

// the same code as csrc/stage.cuh
inline uint32_t warp_incl_scan(uint32_t v) {
    const int lane = lane_id();
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
}  // namespace aurora

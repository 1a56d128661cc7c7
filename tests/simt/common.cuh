// tests/simt/common.cuh — what csrc/common.cuh gives a kernel source, for the CPU lane emulation: the real header's host part
// (parameter blocks, format ids; AURORA_REAL_COMMON is its path) plus its device helpers on top of simt.hpp: shared addresses
// are byte offsets into the emulated warp's shared-memory array, a TMA bulk copy is a checked memcpy that completes at once and
// signals its mbarrier, which the waiting lanes really wait for (phases, parities, transaction bytes).
#pragma once
#include "cuda_runtime.h"
#include AURORA_REAL_COMMON
#include "simt.hpp"

namespace aurora {
inline int lane_id() { return simt_lane(); }
inline uint32_t smem_u32(const void* p) {
    const simt::Warp* w = simt::current();
    const uint8_t* b = static_cast<const uint8_t*>(p);
    if (b < w->smem || b > w->smem + w->smem_size) simt::fail("smem_u32 of a pointer outside the warp's shared memory");
    return uint32_t(b - w->smem);
}
inline void mbar_init(uint64_t* bar, uint32_t count) { simt::mbar_init_impl(bar, count); }
inline void mbar_arrive(uint64_t* bar) { simt::mbar_arrive_impl(bar, 0); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { simt::mbar_arrive_impl(bar, int64_t(bytes)); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return simt::mbar_test(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { simt::mbar_wait_impl(bar, parity); }
inline void fence_proxy_async() {}
inline void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    const simt::Warp* w = simt::current();
    const uint8_t* g = static_cast<const uint8_t*>(gmem_src);
    uint8_t* d = static_cast<uint8_t*>(smem_dst);
    if (g < w->g_lo || g + bytes > w->g_hi) simt::fail("TMA bulk copy reads outside the batch's source bytes");
    if (d < w->smem || d + bytes > w->smem + w->smem_size) simt::fail("TMA bulk copy writes outside the warp's shared memory");
    if (bytes % 16 || (reinterpret_cast<uintptr_t>(g) % 16) || ((d - w->smem) % 16)) simt::fail("TMA bulk copy: size and addresses must be multiples of 16");
    memcpy(d, g, bytes);
    simt::mbar_complete_tx(bar, int64_t(bytes));   // the copy completes at once; the waiters still have to wait for it
}
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
// the named barriers of the flag-LZ decode kernel's warp pairs (PTX bar.sync / bar.arrive id, 64)
inline void bar_sync(uint32_t id) { simt::named_barrier(id, true); }
inline void bar_arrive(uint32_t id) { simt::named_barrier(id, false); }
}  // namespace aurora

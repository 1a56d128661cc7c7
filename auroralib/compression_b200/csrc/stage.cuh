// stage.cuh — TMA-staged input sub-streams and small warp primitives shared by the decode kernels.
#pragma once
#include "common.cuh"

namespace aurora {

constexpr int kInChunk = 1024;       // TMA bulk chunk
constexpr int kInRing = 2 * kInChunk;
constexpr int kInMask = kInRing - 1;
constexpr int kLookahead = 192;      // input bytes one parse step may touch past its cursor
constexpr int kInMirror = 640;       // copy of slot 0's head behind the ring: any window of <= 640 bytes is contiguous
constexpr int kInStage = kInRing + kInMirror;   // shared-memory bytes of one staged sub-stream

#ifndef AURORA_SIMT   // (tests/simt supplies these on its CPU lane emulation)
// ---- shared-memory accessors on 32-bit shared addresses (explicit program order for the ring traffic)
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// TMA-staged input sub-stream: a 2 x 1 KiB shared-memory ring filled by cp.async.bulk (1-D TMA) with one
// mbarrier per slot.  All members are warp-uniform registers.  Loads are numbered by a counter that runs
// over the lifetime of the warp (slot = k & 1, mbarrier parity = (k >> 1) & 1), so phases stay
// consistent across streams; stream chunk c maps to load kbase + (c - cbase).  Access is sequential
// with occasional forward jumps (skipped frames), which re-base the mapping.
// Positions passed to ensure()/at() are relative to the pointer given to begin().
// ---------------------------------------------------------------------------------------------
struct InStream {
    uint8_t* ring;
    uint64_t* bar;
    const uint8_t* gbase;   // 16-byte aligned global address of stream chunk 0
    uint32_t glimit;        // loadable bytes from gbase (multiple of 16)
    uint32_t nchunks;       // ceil(glimit / kInChunk)
    uint32_t skew;          // relative byte 0 sits at gbase + skew
    uint32_t kbase, cbase;  // load number / stream chunk of the current mapping
    uint32_t issued, ready; // running load counters
    uint32_t rbias;         // ring index of relative byte 0 under the current mapping
    uint32_t ready_end;     // bytes (from gbase) below this are staged and waited for
    uint32_t issue_trig;    // cursor (from gbase) at which the next chunk has to be issued

    __device__ __forceinline__ void init(uint8_t* r, uint64_t* b) {
        ring = r;
        bar = b;
        issued = ready = 0;
        if (lane_id() == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
        }
    }
    __device__ __forceinline__ void drain_inflight() {
        while (ready < issued) {
            mbar_wait(&bar[ready & 1], (ready >> 1) & 1);
            ready++;
        }
    }
    __device__ __forceinline__ void rebase(uint32_t chunk) {
        drain_inflight();
        kbase = issued;
        cbase = chunk;
        rbias = skew + ((kbase - cbase) & 1u) * kInChunk;
        ready_end = 0;
        issue_trig = 0;
    }
    __device__ __forceinline__ void begin(const uint8_t* src_base, uint64_t src_limit, const uint8_t* p) {
        skew = uint32_t(reinterpret_cast<uintptr_t>(p) & 15);
        gbase = p - skew;
        const uint64_t lim = uint64_t((src_base + src_limit) - gbase);
        glimit = lim > 0xFFFF0000ull ? 0xFFFF0000u : uint32_t(lim);   // (the chunk count below must not wrap)
        nchunks = (glimit + kInChunk - 1) / kInChunk;
        rebase(0);
    }
    // make relative bytes [pos, pos + span) readable (span <= kInChunk); pos never moves backwards
    __device__ __forceinline__ void ensure(uint32_t pos, uint32_t span = kLookahead) {
        // fast path (two compares): the bytes are staged and the cursor has not reached the next issue point
        if (pos + skew + span < ready_end && pos + skew < issue_trig) return;
        ensure_slow(pos, span);
    }
    __device__ __forceinline__ void ensure_slow(uint32_t pos, uint32_t span) {
        const uint32_t lo = (pos + skew) / kInChunk;
        const uint32_t hi = (pos + skew + span) / kInChunk;
        uint32_t cend = cbase + (issued - kbase);   // next stream chunk to issue
        if (lo > cend) {
            rebase(lo);
            cend = lo;
        }
        const uint32_t want = min(lo + 2, nchunks);
        if (cend < want) {
            __syncwarp();
            if (lane_id() == 0) {
                fence_proxy_async();
                for (uint32_t c = cend; c < want; c++) {
                    const uint32_t k = kbase + (c - cbase);
                    const uint32_t bytes = min(uint32_t(kInChunk), glimit - c * kInChunk);
                    const uint32_t mir = (k & 1) ? 0u : min(bytes, uint32_t(kInMirror));
                    mbar_expect_tx(&bar[k & 1], bytes + mir);
                    tma_bulk_g2s(ring + (k & 1) * kInChunk, gbase + size_t(c) * kInChunk, bytes, &bar[k & 1]);
                    if (mir) tma_bulk_g2s(ring + kInRing, gbase + size_t(c) * kInChunk, mir, &bar[k & 1]);
                }
            }
            issued += want - cend;
        }
        const uint32_t want_c = min(hi + 1, nchunks);
        if (want_c > cbase) {
            const uint32_t want_k = kbase + (want_c - cbase);
            while (int32_t(want_k - ready) > 0) {
                mbar_wait(&bar[ready & 1], (ready >> 1) & 1);
                ready++;
            }
        }
        // limits of the fast path under the current mapping
        const uint32_t rchunks = cbase + (ready - kbase);                  // stream chunks [cbase, rchunks) are complete
        ready_end = rchunks >= nchunks ? 0xFFFFFFFFu : rchunks * kInChunk;  // nothing to wait for past the last chunk
        const uint32_t nend = cbase + (issued - kbase);                    // next stream chunk to issue
        issue_trig = nend >= nchunks ? 0xFFFFFFFFu : (nend - 1) * kInChunk;
    }
    __device__ __forceinline__ uint32_t at(uint32_t pos) const {
        return lds_u8(smem_u32(ring) + ((pos + rbias) & kInMask));
    }
    // pointer to relative byte `pos`, contiguous for kInMirror bytes (after ensure(pos, <= kInMirror))
    __device__ __forceinline__ const uint8_t* window(uint32_t pos) const { return ring + ((pos + rbias) & kInMask); }
};

// ceil(2^20 / d): i mod d for the self-overlapping copy without an integer division (exact for i, d < 512)
struct RcpTable {
    uint32_t v[512];
    constexpr RcpTable() : v() {
        for (uint32_t d = 1; d < 512; ++d) v[d] = ((1u << 20) + d - 1) / d;
    }
};
#ifndef AURORA_SIMT
static __constant__ RcpTable c_rcp = RcpTable();
#else
static const RcpTable c_rcp = RcpTable();
#endif

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace aurora

// common.cuh — shared device/host definitions for libaurora_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/aurora_cuda.h"

namespace aurora {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

// LZSS parameters resolved on the host from aurora_lz_props (LzProperties.cs)
struct LzssParams {
    int windows_bits, length_bits, min_length, max_distance, windows_start, initial_fill;
};

// One batch on one device.  All pointers are device pointers.
struct DecodeParams {
    const uint8_t* src_base;
    uint64_t src_limit;          // readable bytes from src_base, a multiple of 16 (TMA bulk granularity)
    const uint64_t* src_off;
    const uint64_t* src_len;
    uint8_t* dst_base;
    const uint64_t* dst_off;
    const uint64_t* dst_cap;
    uint64_t* out_len;
    uint64_t* consumed;
    int32_t* status;
    const uint32_t* order;       // optional: stream indices, largest first (nullptr = identity)
    unsigned int* ticket;        // global work counter (zeroed by the launcher)
    uint32_t n;
    int format;
    int byte_order;              // aurora_endian
    int size_only;               // parse without copying (decoded_size_batch size_scan)
    int lz4_verify;              // LZ4.HashAlgorithm set: verify XXH32 checksums
    int headerless;              // LZ10/LZ11/LZSS DecompressHeaderless(source, destination, size): the stream is the bare token
                                 // body and dst_cap[i] carries (size << 32) | min(capacity, 2^32 - 1)  (wrapper formats)
    LzssParams lzss;
};

struct EncodeParams {
    const uint8_t* src_base;
    uint64_t src_limit;
    const uint64_t* src_off;
    const uint64_t* src_len;
    uint8_t* dst_base;
    const uint64_t* dst_off;
    const uint64_t* dst_cap;
    uint64_t* out_len;
    int32_t* status;
    unsigned int* ticket;
    uint32_t n;
    int format;
    int byte_order;
    // CompressionSettings -> LzChainMatchFinder parameters (LzChainMatchFinder.cs:108-119)
    int max_chain, lazy_threshold, hash_bits, chain_bits, min_length, max_length, min_distance, max_distance, no_self_overlap,
        use_min_table;
    int finder_choice;   // 0 automatic, 1 one lane per window position (encode_lz_par.cu), 2 sequential replay (finder.cuh)
    uint32_t yaz0_alignment;
    uint32_t lz4_block_size;
    LzssParams lzss;
    uint8_t* scratch;            // per-resident-warp match-finder tables
    uint64_t scratch_per_warp;
};

// kernel launchers (one translation unit per kernel family); all return cudaGetLastError()
cudaError_t launch_decode_flaglz(const DecodeParams& p, int sm_count, cudaStream_t st);
cudaError_t launch_decode_bytelz(const DecodeParams& p, int sm_count, cudaStream_t st);
cudaError_t launch_decode_blz(const DecodeParams& p, int sm_count, cudaStream_t st);
// in-place byte reversal of bytes [0, min(len[i], cap[i])) of every stream (BLZ encode; decode_blz.cu); cap may be null
cudaError_t launch_reverse_bytes(uint8_t* base, const uint64_t* d_off, const uint64_t* d_len, const uint64_t* d_cap, uint32_t n,
                                 cudaStream_t st);
cudaError_t launch_encode_lz(const EncodeParams& p, int warps, cudaStream_t st);
cudaError_t launch_encode_bytelz(const EncodeParams& p, int warps, cudaStream_t st);
bool encode_lz_par_supported(const EncodeParams& p);   // encode_lz_par.cu: one lane per window position
cudaError_t launch_encode_lz_par(const EncodeParams& p, int sm_count, cudaStream_t st);
size_t encode_scratch_per_warp(int format, int hash_bits, int chain_bits, uint64_t max_src_len);
int encode_resident_warps(int sm_count);
cudaError_t launch_size_order(const uint64_t* d_size, uint32_t n, uint32_t* d_hist64, uint32_t* d_order, cudaStream_t st);
// LZ00's per-byte LCG keystream over bytes [off[i] + skip, off[i] + min(len[i], cap[i])) of every stream (keystream.cu); cap may be null
cudaError_t launch_lcg_xor(uint8_t* base, const uint64_t* d_off, const uint64_t* d_len, const uint64_t* d_cap, const uint32_t* d_key,
                           uint32_t skip, uint32_t n, cudaStream_t st);
cudaError_t launch_scan(const uint8_t* image, uint64_t len, uint8_t* match, int format, cudaStream_t st);
cudaError_t launch_is_match(const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len, uint8_t* match, uint32_t n,
                            int format, cudaStream_t st);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (global -> shared), the sm_90+/sm_100a async staging path.
// SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK (see /opt/skills/guides/B200_PROFILING.md).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {   // release.cta: orders the arriving thread's earlier writes
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint expires, so a
    // waiting warp does not burn issue slots polling
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
// (Measured and rejected, round 1: __nanosleep back-off in this loop; the polls are a third of the issued instructions
// of the pipelined flag-LZ kernel, but the hand-off latency it adds costs more than the issue slots it frees.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

}  // namespace aurora

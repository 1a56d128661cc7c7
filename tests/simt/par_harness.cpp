// tests/simt/par_harness.cpp — runs the DEVICE part of csrc/encode_lz_par.cu (everything above its "// ---- kernel" line, cut
// out of the real file by tests/test_simt_kernels.py and included below) on the CPU lane emulation of simt.hpp.
// TEST INFRASTRUCTURE: the product never loads this.  One emulated warp encodes the streams of a batch one after the other
// with the same tables, as a resident warp of the kernel does.
#include <vector>

#include "common.cuh"
#include "stage.cuh"
#include PAR_DEVICE_INC   // opens `namespace aurora { namespace {` and leaves both open

template <int K, bool M>
static int run_batch(const EncodeParams& P) {
    simt::Warp w;
    std::vector<uint8_t> smem(size_t(kTablesPerWarp<M>), 0xCD);
    w.smem = smem.data();
    w.smem_size = smem.size();
    w.g_lo = P.src_base;
    w.g_hi = P.src_base + P.src_limit;
    simt::run_warp(w, [&](int) {
        ParState S;
        par_setup(S, P, 0u, P.scratch);
        for (uint32_t t = 0; t < P.n; t++) encode_stream_par<K, M>(P, t, S);
    });
    return 0;
}

template <int K>
static int run_kind(const EncodeParams& P) {
    return (P.use_min_table && P.min_length < 4) ? run_batch<K, true>(P) : run_batch<K, false>(P);
}

}  // namespace
}  // namespace aurora

extern "C" int simt_encode_lz_par(int format, int byte_order, const int* finder /* max_chain, lazy, hash_bits, chain_bits, min_length,
                                  max_length, min_distance, max_distance, no_self_overlap, use_min_table */,
                                  uint32_t yaz0_alignment, const int* lzss /* windows_bits, length_bits, min_length, max_distance,
                                  windows_start */, const uint8_t* src_base, uint64_t src_limit, const uint64_t* src_off,
                                  const uint64_t* src_len, uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* dst_cap,
                                  uint64_t* out_len, int32_t* status, uint32_t n, uint8_t* scratch, uint64_t scratch_per_warp) {
    using namespace aurora;
    EncodeParams P{};
    P.src_base = src_base;
    P.src_limit = src_limit;
    P.src_off = src_off;
    P.src_len = src_len;
    P.dst_base = dst_base;
    P.dst_off = dst_off;
    P.dst_cap = dst_cap;
    P.out_len = out_len;
    P.status = status;
    P.n = n;
    P.format = format;
    P.byte_order = byte_order;
    P.max_chain = finder[0];
    P.lazy_threshold = finder[1];
    P.hash_bits = finder[2];
    P.chain_bits = finder[3];
    P.min_length = finder[4];
    P.max_length = finder[5];
    P.min_distance = finder[6];
    P.max_distance = finder[7];
    P.no_self_overlap = finder[8];
    P.use_min_table = finder[9];
    P.yaz0_alignment = yaz0_alignment;
    P.lzss = LzssParams{lzss[0], lzss[1], lzss[2], lzss[3], lzss[4], 0};
    P.scratch = scratch;
    P.scratch_per_warp = scratch_per_warp;
    switch (format) {
        case AURORA_FMT_LZ10:
        case AURORA_FMT_BLZ: return run_kind<P_LZ10>(P);
        case AURORA_FMT_YAZ0:
        case AURORA_FMT_YAZ1: return run_kind<P_YAZ0>(P);
        case AURORA_FMT_LZSS: return run_kind<P_LZSS>(P);
        case AURORA_FMT_MIO0: return run_kind<P_MIO0>(P);
        case AURORA_FMT_YAY0: return run_kind<P_YAY0>(P);
        case AURORA_FMT_LZ11:
        case AURORA_FMT_LZ40:
        case AURORA_FMT_LZ60: return run_kind<P_LZ11>(P);
        case AURORA_FMT_LZHUDSON: return run_kind<P_HUDSON>(P);
        case AURORA_FMT_SMSR00: return run_kind<P_SMSR>(P);
        default: return -1;
    }
}

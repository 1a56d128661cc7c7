// tests/simt/seq_harness.cpp — runs the DEVICE part of csrc/encode_lz.cu or csrc/encode_bytelz.cu (with csrc/finder.cuh: the
// sequential replay of the reference's match finder; everything above the file's "// ---- kernel" line, cut out of the real
// file by tests/test_simt_kernels.py) on the CPU lane emulation of simt.hpp.  TEST INFRASTRUCTURE: the product never loads this.
#include <vector>

#include "common.cuh"
#include "stage.cuh"
#include SEQ_DEVICE_INC   // opens `namespace aurora { namespace {` and leaves both open

static int run_batch(const EncodeParams& P) {
    simt::Warp w;
    w.smem = nullptr;
    w.smem_size = 0;
    w.g_lo = P.src_base;
    w.g_hi = P.src_base + P.src_limit;
#ifdef SEQ_BYTELZ
    for (uint32_t i = 0; i < 256; i++) {   // what launch_encode_bytelz copies into the constant bank
        uint32_t c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        c_crc32c[i] = c;
    }
#endif
    int rc = 0;
    simt::run_warp(w, [&](int) {
        Finder f;
        finder_setup(f, P, P.scratch);
        for (uint32_t t = 0; t < P.n; t++) {
#ifdef SEQ_BYTELZ
            encode_stream(P, t, f);
#else
            switch (P.format) {   // launch_encode_lz
                case AURORA_FMT_LZ10:
                case AURORA_FMT_BLZ: encode_stream<E_LZ10>(P, t, f); break;
                case AURORA_FMT_LZ11:
                case AURORA_FMT_LZ40:
                case AURORA_FMT_LZ60: encode_stream<E_LZ11>(P, t, f); break;
                case AURORA_FMT_YAZ0:
                case AURORA_FMT_YAZ1:
                case AURORA_FMT_LZHUDSON: encode_stream<E_YAZ0>(P, t, f); break;
                case AURORA_FMT_LZSS: encode_stream<E_LZSS>(P, t, f); break;
                case AURORA_FMT_MIO0:
                case AURORA_FMT_SMSR00: encode_stream<E_MIO0>(P, t, f); break;
                case AURORA_FMT_YAY0: encode_stream<E_YAY0>(P, t, f); break;
                default: rc = -1;
            }
#endif
        }
    });
    return rc;
}

}  // namespace
}  // namespace aurora

extern "C" int simt_encode_seq(int format, int byte_order, const int* finder /* max_chain, lazy, hash_bits, chain_bits, min_length,
                               max_length, min_distance, max_distance, no_self_overlap, use_min_table */,
                               uint32_t yaz0_alignment, uint32_t lz4_block_size, const int* lzss /* windows_bits, length_bits,
                               min_length, max_distance, windows_start */, const uint8_t* src_base, uint64_t src_limit,
                               const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst_base, const uint64_t* dst_off,
                               const uint64_t* dst_cap, uint64_t* out_len, int32_t* status, uint32_t n, uint8_t* scratch,
                               uint64_t scratch_per_warp) {
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
    P.lz4_block_size = lz4_block_size;
    P.lzss = LzssParams{lzss[0], lzss[1], lzss[2], lzss[3], lzss[4], 0};
    P.scratch = scratch;
    P.scratch_per_warp = scratch_per_warp;
    return run_batch(P);
}

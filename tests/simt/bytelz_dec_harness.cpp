// tests/simt/bytelz_dec_harness.cpp — runs the DEVICE part of csrc/decode_bytelz.cu (LZ4 block / legacy / frame, Snappy block /
// framed, LZO, PRS; everything above the file's "// ---- kernel" line, cut out of the real file by tests/test_simt_kernels.py)
// on the CPU lane emulation of simt.hpp, with the REAL staged input stream of csrc/stage.cuh over an emulated TMA.
// TEST INFRASTRUCTURE: the product never loads this.
#include <vector>

#include "common.cuh"
#include "stage.cuh"
#include DEC_DEVICE_INC   // opens `namespace aurora { namespace {` and leaves both open

template <int K>
static int run_batch(const DecodeParams& P) {
    simt::Warp w;
    // the warp's slice as the kernel carves it: the output ring first (its address must be a multiple of its size), then the
    // staged input and its two mbarriers
    std::vector<uint8_t> smem(size_t(kSmemPerWarp) + 64, 0xCD);
    w.smem = smem.data();
    w.smem_size = size_t(kSmemPerWarp);
    w.g_lo = P.src_base;
    w.g_hi = P.src_base + P.src_limit;
    simt::run_warp(w, [&](int) {
        uint8_t* wbase = simt::current()->smem + kORing;
        InStream in;
        in.init(wbase, reinterpret_cast<uint64_t*>(wbase + kInStage));
        __syncwarp();
        for (uint32_t t = 0; t < P.n; t++) decode_stream<K>(P, t, in, 0u);
        in.drain_inflight();
    });
    return 0;
}

}  // namespace
}  // namespace aurora

extern "C" int simt_decode_bytelz(int format, int byte_order, int lz4_verify, int size_only, const uint8_t* src_base, uint64_t src_limit,
                                  const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst_base, const uint64_t* dst_off,
                                  const uint64_t* dst_cap, uint64_t* out_len, uint64_t* consumed, int32_t* status, uint32_t n) {
    using namespace aurora;
    DecodeParams P{};
    P.src_base = src_base;
    P.src_limit = src_limit;
    P.src_off = src_off;
    P.src_len = src_len;
    P.dst_base = dst_base;
    P.dst_off = dst_off;
    P.dst_cap = dst_cap;
    P.out_len = out_len;
    P.consumed = consumed;
    P.status = status;
    P.n = n;
    P.format = format;
    P.byte_order = byte_order;
    P.lz4_verify = lz4_verify;
    P.size_only = size_only;
    switch (format) {   // launch_decode_bytelz
        case AURORA_FMT_LZ4:
        case AURORA_FMT_LZ4_LEGACY: return run_batch<B_LZ4>(P);
        case AURORA_FMT_LZ4_BLOCK: return run_batch<B_LZ4_BLOCK>(P);
        case AURORA_FMT_SNAPPY: return run_batch<B_SNAPPY>(P);
        case AURORA_FMT_SNAPPY_BLOCK: return run_batch<B_SNAPPY_BLOCK>(P);
        case AURORA_FMT_LZO: return run_batch<B_LZO>(P);
        case AURORA_FMT_PRS: return run_batch<B_PRS>(P);
        default: return -1;
    }
}

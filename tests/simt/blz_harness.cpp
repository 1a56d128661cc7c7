// tests/simt/blz_harness.cpp — runs the DEVICE part of csrc/decode_blz.cu (everything above its "// ---- kernel" line, cut out
// of the real file by tests/test_simt_kernels.py) on the CPU lane emulation of simt.hpp.  TEST INFRASTRUCTURE.
#include "common.cuh"
#include "stage.cuh"
#include BLZ_DEVICE_INC   // opens `namespace aurora { namespace {` and leaves both open

static void run_batch(const DecodeParams& P) {
    simt::Warp w;
    w.g_lo = P.src_base;
    w.g_hi = P.src_base + P.src_limit;
    simt::run_warp(w, [&](int) {
        for (uint32_t t = 0; t < P.n; t++) blz_decode_stream(P, t);
    });
}

}  // namespace
}  // namespace aurora

extern "C" int simt_decode_blz(const uint8_t* src_base, uint64_t src_limit, const uint64_t* src_off, const uint64_t* src_len,
                               uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t* out_len,
                               uint64_t* consumed, int32_t* status, uint32_t n, int size_only) {
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
    P.format = AURORA_FMT_BLZ;
    P.size_only = size_only;
    run_batch(P);
    return 0;
}

// tests/simt/flaglz_dec_harness.cpp — runs the DEVICE part of csrc/decode_flaglz.cu (the headline kernel: LZ10, LZ11 / LZ40 / LZ60,
// Yaz0 / Yaz1, LZSS, MIO0, Yay0, LZHudson, SMSR00; everything above the file's "// ---- kernel" line, cut out of the real file by
// tests/test_simt_kernels.py) on the CPU lane emulation of simt.hpp: ONE stream slot = a parser warp and a resolver warp (64
// fibers) that hand batches over through two emulated named barriers, the real staged input streams over an emulated TMA.
// The slot is carved and the two roles are started exactly as decode_flaglz_kernel does (csrc/decode_flaglz.cu, "// ---- kernel").
// TEST INFRASTRUCTURE: the product never loads this.
#include <vector>

#include "common.cuh"
#include "stage.cuh"
#include DEC_DEVICE_INC   // opens `namespace aurora { namespace {` and leaves both open

template <int K>
static int run_batch(const DecodeParams& P) {
    using T = Traits<K>;
    simt::Warp w;
    // one slot: the 8 KiB ring first (its shared address must be a multiple of its size), then the slot's aux block
    const size_t bytes = size_t(kRing) + size_t(T::kAuxBytes);
    std::vector<uint8_t> smem(bytes + 64, 0xCD);
    w.smem = smem.data();
    w.smem_size = bytes;
    w.g_lo = P.src_base;
    w.g_hi = P.src_base + P.src_limit;
    unsigned ticket = 0;
    DecodeParams Q = P;
    Q.ticket = &ticket;
    simt::run_block(w, 2, [&](int thread) {
        const int role = thread >> 5;   // warp 0 parses, warp 1 resolves
        uint8_t* ring = simt::current()->smem;
        uint8_t* aux = ring + kRing;
        uint8_t* qptr = aux + T::kStreams * kInStage;
        const uint32_t qbase = smem_u32(qptr), gaddr = qbase + T::kQueueBytes, mail = gaddr + 128;
        uint64_t* bars = reinterpret_cast<uint64_t*>(qptr + T::kQueueBytes + 128 + 48);
        const uint32_t bar_f = 0, bar_e = 1;
        if (role == 0) {
            InStream in[T::kStreams];
            for (int s = 0; s < T::kStreams; s++) in[s].init(aux + s * kInStage, bars + 2 * s);
            fence_proxy_async();
            __syncwarp();
            SlotSink sink;
            sink.rbase = smem_u32(ring);
            sink.qbase = qbase;
            sink.mail = mail;
            sink.bar_f = bar_f;
            sink.bar_e = bar_e;
            sink.init();
            for (;;) {
                uint32_t t = 0;
                if (lane_id() == 0) t = atomicAdd(Q.ticket, 1u);
                t = __shfl_sync(kFull, t, 0);
                if (t >= Q.n) break;
                const uint32_t idx = Q.order ? Q.order[t] : t;
                decode_stream<K>(Q, idx, in, ring, sink, gaddr);
            }
            sink.exit();
            for (int s = 0; s < T::kStreams; s++) in[s].drain_inflight();
        } else {
            resolver_role(ring, qbase, mail, bar_f, bar_e);
        }
    });
    return 0;
}

}  // namespace
}  // namespace aurora

extern "C" int simt_decode_flaglz(int format, int byte_order, int headerless, const int* lzss /* windows_bits, length_bits, min_length,
                                  max_distance, windows_start, initial_fill */, const uint8_t* src_base, uint64_t src_limit,
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
    P.headerless = headerless;
    P.lzss = LzssParams{lzss[0], lzss[1], lzss[2], lzss[3], lzss[4], lzss[5]};
    switch (format) {   // launch_decode_flaglz
        case AURORA_FMT_LZ10: return run_batch<K_LZ10>(P);
        case AURORA_FMT_LZ11: return run_batch<K_LZ11>(P);
        case AURORA_FMT_YAZ0:
        case AURORA_FMT_YAZ1: return run_batch<K_YAZ0>(P);
        case AURORA_FMT_LZSS: return run_batch<K_LZSS>(P);
        case AURORA_FMT_MIO0: return run_batch<K_MIO0>(P);
        case AURORA_FMT_YAY0: return run_batch<K_YAY0>(P);
        case AURORA_FMT_LZHUDSON: return run_batch<K_HUDSON>(P);
        case AURORA_FMT_LZ40:
        case AURORA_FMT_LZ60: return run_batch<K_LZ40>(P);
        case AURORA_FMT_SMSR00: return run_batch<K_SMSR>(P);
        default: return -1;
    }
}

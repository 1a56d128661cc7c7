// decode_blz.cu — batched decoder for Nintendo BLZ ("bottom LZ"), the one format of the family that is parsed and written
// BACKWARDS from the end of the stream.  One compressed stream per warp.
//
// Reference semantics restated on the device (src/AuroraLib.Compression.Nintendo/Nintendo/BLZ.cs):
//   Decompress :45-69            footer at Length - 8: u24 LE compressed size, u8 footer-and-padding size (>= 8), i32 LE
//                                (decoded size - compressed size); the codes are the compressed size minus the footer and
//                                padding, read from Length - compressed size; the decoded buffer is written to the
//                                destination only when the whole decode succeeded
//   DecompressHeaderless :101-141  src runs from the end of the codes down to 0, dst from the decoded size down to 0: flag
//                                byte (MSB first, 0 = literal), literal byte, or a u16 (first byte read = high byte)
//                                (length - 3) << 12 | (distance - 3) copying destination[dst - 1] = destination[dst - 1 +
//                                distance] while dst > 0; DecompressedSizeException when dst != 0 at the end
//
// Design: the token walk is a literal transcription — all 32 lanes step through the same scalars (the parse never looks at
// decoded bytes, so the statuses are independent of the data written), the literal bits that follow each other inside one
// flag byte are one warp step (65 -> 174 GB/s on the C2 corpus against one literal per step), a match is one
// warp-cooperative step (length <= 18) with the periodic form when distance < length.  Everything lives in global memory
// (codes through L1, output written once and re-read by back-references through L1/L2): there is no staging ring to run
// backwards, and occupancy (64 warps per SM) hides the dependent-load latency of the walk.  When the decoded size exceeds
// the destination the walk still runs (stores dropped) so that a corrupt stream reports its own error, like the reference,
// before the failed destination write.  Algorithmic bytes per stream = compressed + decoded.
#include "stage.cuh"

namespace aurora {

namespace {

constexpr int kBlzWarpsPerBlock = 16;

__device__ void blz_decode_stream(const DecodeParams& P, uint32_t idx) {
    const uint32_t lane = lane_id();
    const uint8_t* p = P.src_base + P.src_off[idx];
    const uint64_t len64 = P.src_len[idx];
    uint8_t* out = P.dst_base + P.dst_off[idx];
    const uint64_t cap = P.dst_cap[idx];
    int status = AURORA_OK;
    uint64_t consumed = len64, out_len = 0;
    if (len64 < 8) {
        status = AURORA_END_OF_STREAM;   // Position = Length - 8 / the footer reads
    } else {
        const uint8_t* f = p + len64 - 8;
        const uint32_t csize = uint32_t(f[0]) | (uint32_t(f[1]) << 8) | (uint32_t(f[2]) << 16);
        const uint32_t hdr = f[3];
        const uint32_t delta = uint32_t(f[4]) | (uint32_t(f[5]) << 8) | (uint32_t(f[6]) << 16) | (uint32_t(f[7]) << 24);
        const int32_t m = int32_t(delta + csize);
        const int32_t n = int32_t(csize) - int32_t(hdr);
        if (hdr < 8 || uint64_t(csize) > len64 || n < 0 || m < 0) {
            status = AURORA_INVALID_DATA;   // "Invalid BLZ header." / Position < 0 / ArrayPool.Rent(negative)
        } else {
            consumed = len64 - hdr;         // source.Read(inBuffer, 0, codeSize)
            const uint8_t* in = p + (len64 - csize);
            const bool store = uint64_t(m) <= cap && !P.size_only;
            int32_t src = n, d = m;
            uint32_t flags = 0, mask = 0;
            while (src > 0) {
                if ((mask >>= 1) == 0) {
                    flags = in[--src];
                    mask = 0x80;
                }
                if ((flags & mask) == 0) {
                    // the literal bits that follow in the SAME flag byte govern the next code bytes: one warp step for the run.
                    // The first literal fails on dst == 0 (checked first) or src == 0; before every further one the loop
                    // condition ends the walk cleanly when the codes are used up, and dst == 0 is the error that remains.
                    if (d == 0) { status = AURORA_INVALID_DATA; break; }
                    if (src == 0) { status = AURORA_END_OF_STREAM; break; }
                    const uint32_t below = flags & ((mask << 1) - 1u);            // the flag bits from `mask` downwards
                    const int32_t run = below ? int32_t(__clz(below)) - int32_t(__clz(mask)) : 32 - int32_t(__clz(mask));
                    const int32_t k = min(run, min(d, src));
                    if (store && int32_t(lane) < k) out[d - 1 - int32_t(lane)] = in[src - 1 - int32_t(lane)];
                    d -= k;
                    src -= k;
                    if (k < run) {
                        if (src != 0) status = AURORA_INVALID_DATA;   // dst == 0 with codes left
                        break;
                    }
                    mask >>= (k - 1);
                } else {
                    if (src < 2) { status = AURORA_END_OF_STREAM; break; }
                    const uint32_t info = (uint32_t(in[src - 1]) << 8) | in[src - 2];
                    src -= 2;
                    const int32_t dist = int32_t(info & 0xFFF) + 3;
                    const int32_t L = min(int32_t(info >> 12) + 3, d);
                    if (L > 0) {
                        if (d - 1 + dist >= m) { status = AURORA_INVALID_DATA; break; }   // reads past the end of the buffer
                        if (store) {
                            __syncwarp();   // literals and earlier matches of this warp are visible
                            if (int32_t(lane) < L) {
                                // destination[d-1-i] = destination[d-1-i+dist]; for i >= dist that byte was written by this
                                // very match: the copy is periodic with period dist
                                const int32_t i = int32_t(lane);
                                const int32_t k = dist < L ? i - int32_t((uint32_t(i) * c_rcp.v[dist]) >> 20) * dist : i;
                                out[d - 1 - i] = out[d - 1 + dist - k];
                            }
                            __syncwarp();
                        }
                        d -= L;
                    }
                }
            }
            if (status == AURORA_OK && d != 0) status = AURORA_SIZE_MISMATCH;
            if (status == AURORA_OK) {
                if (uint64_t(m) > cap && !P.size_only) status = AURORA_DST_TOO_SMALL;   // destination.Write refuses the whole buffer
                else out_len = uint64_t(m);
            }
        }
    }
    if (lane == 0) {
        P.out_len[idx] = out_len;
        P.consumed[idx] = consumed;
        P.status[idx] = status;
    }
    __syncwarp();
}

// ---- kernel
__global__ void __launch_bounds__(kBlzWarpsPerBlock * 32) decode_blz_kernel(const DecodeParams P) {
    for (;;) {
        uint32_t t = 0;
        if (lane_id() == 0) t = atomicAdd(P.ticket, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= P.n) break;
        blz_decode_stream(P, P.order ? P.order[t] : t);
    }
}

// in-place byte reversal of bytes [0, min(len[i], cap[i])) of every stream (the encoder's match finder "only works in one
// direction", BLZ.cs:152-156: the source is reversed, encoded forwards, and the code stream reversed again)
__global__ void __launch_bounds__(256) reverse_bytes_kernel(uint8_t* base, const uint64_t* off, const uint64_t* len, const uint64_t* cap,
                                                           uint32_t n) {
    for (uint32_t s = blockIdx.x; s < n; s += gridDim.x) {
        uint64_t m = len[s];
        if (cap && cap[s] < m) m = cap[s];
        uint8_t* p = base + off[s];
        for (uint64_t i = uint64_t(blockIdx.y) * blockDim.x + threadIdx.x; i < m / 2; i += uint64_t(gridDim.y) * blockDim.x) {
            const uint8_t a = p[i], b = p[m - 1 - i];
            p[i] = b;
            p[m - 1 - i] = a;
        }
    }
}

}  // namespace

cudaError_t launch_reverse_bytes(uint8_t* base, const uint64_t* d_off, const uint64_t* d_len, const uint64_t* d_cap, uint32_t n,
                                 cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const dim3 grid(n < 4736u ? n : 4736u, n >= 1184u ? 1u : n >= 148u ? 8u : 32u);
    reverse_bytes_kernel<<<grid, 256, 0, st>>>(base, d_off, d_len, d_cap, n);
    return cudaGetLastError();
}

cudaError_t launch_decode_blz(const DecodeParams& p, int sm_count, cudaStream_t st) {
    int blocks = sm_count * 4;   // 64 resident warps per SM
    const int needed = int((p.n + kBlzWarpsPerBlock - 1) / kBlzWarpsPerBlock);
    if (needed < blocks) blocks = needed > 0 ? needed : 1;
    decode_blz_kernel<<<blocks, kBlzWarpsPerBlock * 32, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace aurora

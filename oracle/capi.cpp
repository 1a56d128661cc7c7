// capi.cpp — CPU ORACLE (test infrastructure, NOT product code).
// C ABI of liboracle.so: the same batch shapes as include/aurora_cuda.h (prefix ora_, no context),
// a per-format dispatcher that maps the reference's exceptions to status codes, and a std::thread
// batch driver that stands in for the north-star's Parallel.ForEach CPU baseline (no .NET here).
#include <atomic>
#include <thread>

#include "../include/aurora_cuda.h"
#include "oracle_core.hpp"

namespace ora {

uint32_t nintendo_decoded_size(int fmt, Src& s, const CodecOpts& o);
bool lz1x_validate(Src& source, bool lz11);
int prs_get_byte_order(Src& stream);

static CodecOpts to_opts(const aurora_codec_opts* c) {
    CodecOpts o;
    if (!c) return o;
    if (c->byte_order == AURORA_ENDIAN_LITTLE || c->byte_order == AURORA_ENDIAN_BIG) {
        o.byteOrder = c->byte_order == AURORA_ENDIAN_BIG ? Endian::Big : Endian::Little;
        o.byteOrderDefault = false;
    }
    o.settings.Quality = c->quality < 0 ? 8 : c->quality;
    o.settings.MaxWindowBits = c->max_window_bits;
    o.settings.Strategy = c->strategy;
    o.vramMode = c->vram_mode;
    if (c->lzss.windows_bits != 0) {
        o.lzss.WindowsBits = c->lzss.windows_bits;
        o.lzss.LengthBits = c->lzss.length_bits;
        o.lzss.MinLength = c->lzss.min_length;
        o.lzss.MaxLength = c->lzss.max_length;
        o.lzss.MaxDistance = c->lzss.max_distance;
        o.lzss.MinDistance = c->lzss.min_distance;
        o.lzss.WindowsStart = c->lzss.windows_start;
    }
    o.lzssInitialFill = c->lzss_initial_fill;
    o.lz4BlockSize = c->lz4_block_size ? c->lz4_block_size : 0x400000;
    o.lz4Verify = c->lz4_verify != 0;
    o.yaz0Alignment = c->yaz0_alignment;
    if (c->struct_size >= sizeof(aurora_codec_opts)) {
        o.lz77Type = c->lz77_type ? int(c->lz77_type) : 0x10;
        o.lz77ChunkSize = c->lz77_chunk_size ? int(c->lz77_chunk_size) : 0x1000;
        o.level5Type = c->level5_type ? int(c->level5_type) : 1;
        o.lz00Key = c->lz00_key;
        o.ecdPlainSize = c->ecd_plain_size ? int(c->ecd_plain_size) : 4;
    }
    return o;
}

DecodeResult decode_one(int fmt, const CodecOpts& o, const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap, bool size_only) {
    DecodeResult r;
    Src s(src, n);
    Sink d(dst, cap);
    d.size_only = size_only;
    try {
        switch (fmt) {
            case FMT_YAZ0: yaz0_decode(s, d, o, "Yaz0"); break;
            case FMT_YAZ1: yaz0_decode(s, d, o, "Yaz1"); break;
            case FMT_YAY0: yay0_decode(s, d, o); break;
            case FMT_MIO0: mio0_decode(s, d, o); break;
            case FMT_LZ10: lz10_decode(s, d); break;
            case FMT_LZ11: lz11_decode(s, d); break;
            case FMT_LZSS: lzss_decode(s, d, o); break;
            case FMT_LZ4:
            case FMT_LZ4_LEGACY: lz4_decode(s, d, o); break;
            case FMT_LZ4_BLOCK: lz4_block_decode(s, d); break;
            case FMT_LZO: lzo_decode(s, d); break;
            case FMT_SNAPPY: snappy_decode(s, d); break;
            case FMT_SNAPPY_BLOCK: snappy_block_decode(s, d); break;
            case FMT_PRS: prs_decode(s, d); break;
            case FMT_LZHUDSON: lzhudson_decode(s, d); break;
            case FMT_LZ40: lz40_decode(s, d, 0x40); break;
            case FMT_LZ60: lz40_decode(s, d, 0x60); break;
            case FMT_SMSR00: smsr00_decode(s, d); break;
            case FMT_BLZ: blz_decode(s, d); break;
            default:
                if (!is_wrapper_format(fmt)) fail(INVALID_ARGUMENT);
                wrapper_decode(fmt, s, d, o);
        }
        if (d.pos > d.cap) r.status = DST_TOO_SMALL;
    } catch (const Error& e) {
        r.status = e.status;
    }
    r.out_len = d.pos;
    r.consumed = std::min<int64_t>(std::max<int64_t>(s.pos, 0), n);
    return r;
}

int encode_one(int fmt, const CodecOpts& o, const uint8_t* src, int64_t n64, std::vector<uint8_t>& out) {
    OutBuf b;
    if (n64 > 0x7FFFFFFF) return INVALID_ARGUMENT;
    int n = int(n64);
    try {
        switch (fmt) {
            case FMT_YAZ0: yaz0_encode(src, n, b, o, "Yaz0"); break;
            case FMT_YAZ1: yaz0_encode(src, n, b, o, "Yaz1"); break;
            case FMT_YAY0: yay0_encode(src, n, b, o); break;
            case FMT_MIO0: mio0_encode(src, n, b, o); break;
            case FMT_LZ10: lz10_encode(src, n, b, o); break;
            case FMT_LZ11: lz11_encode(src, n, b, o); break;
            case FMT_LZSS: lzss_encode(src, n, b, o); break;
            case FMT_LZ4: lz4_encode(src, n, b, o, false); break;
            case FMT_LZ4_LEGACY: lz4_encode(src, n, b, o, true); break;
            case FMT_LZ4_BLOCK: lz4_block_encode(src, n, b, o); break;
            case FMT_LZO: lzo_encode(src, n, b, o); break;
            case FMT_SNAPPY: snappy_encode(src, n, b, o); break;
            case FMT_SNAPPY_BLOCK: snappy_block_encode(src, n, b, o); break;
            case FMT_PRS: prs_encode(src, n, b, o); break;
            case FMT_LZHUDSON: lzhudson_encode(src, n, b, o); break;
            case FMT_LZ40: lz40_encode(src, n, b, o, 0x40); break;
            case FMT_LZ60: lz40_encode(src, n, b, o, 0x60); break;
            case FMT_SMSR00: smsr00_encode(src, n, b, o); break;
            case FMT_BLZ: blz_encode(src, n, b, o); break;
            default:
                if (!is_wrapper_format(fmt)) return INVALID_ARGUMENT;
                wrapper_encode(fmt, src, n, b, o, o.lz77Type, o.lz77ChunkSize, o.level5Type);
        }
    } catch (const Error& e) {
        return e.status;
    }
    out.swap(b.v);
    return OK;
}

// IProvidesDecompressedSize.GetDecompressedSize (a Peek: the stream position is restored)
uint32_t decoded_size(int fmt, Src& s, const CodecOpts& o) {
    int64_t p0 = s.pos;
    struct R { Src& s; int64_t p; ~R() { s.pos = p; } } r{s, p0};
    switch (fmt) {
        case FMT_YAZ0: case FMT_YAZ1: case FMT_YAY0: case FMT_MIO0: case FMT_LZ10: case FMT_LZ11: case FMT_LZ40: case FMT_LZ60:
        case FMT_SMSR00: case FMT_BLZ:
            return nintendo_decoded_size(fmt, s, o);
        case FMT_LZSS:   // LZSS.cs:45-50
            s.MatchThrow("LZSS", 4);
            return s.ReadUInt32(Endian::Big);
        case FMT_LZHUDSON: return s.ReadUInt32(Endian::Big);   // LZHudson.cs:37-38
    }
    if (is_wrapper_format(fmt)) return wrapper_decoded_size(fmt, s);
    fail(NOT_SUPPORTED);
}

// IsMatch(Stream) (SURVEY.md Appendix A "IsMatch rules")
bool is_match(int fmt, Src& s, const CodecOpts&) {
    int64_t p0 = s.pos;
    struct R { Src& s; int64_t p; ~R() { s.pos = p; } } r{s, p0};
    auto magic16 = [&](const char* m) { return s.pos + 0x10 < s.len && std::memcmp(s.p + s.pos, m, 4) == 0; };
    switch (fmt) {
        case FMT_YAZ0: return magic16("Yaz0");
        case FMT_YAZ1: return magic16("Yaz1");
        case FMT_YAY0: return magic16("Yay0");
        case FMT_MIO0: return magic16("MIO0");
        case FMT_LZSS: return magic16("LZSS");
        case FMT_LZ4_LEGACY: return magic16("\x02\x21\x4C\x18");
        case FMT_LZ4: {   // LZ4.cs:46-47
            if (!(s.pos + 0x10 < s.len)) return false;
            uint32_t v = s.ReadUInt32();
            return v == 0x184C2102u || v == 0x184D2204u || (v >= 0x184D2A50u && v <= 0x184D2A5Fu);
        }
        case FMT_SNAPPY: {   // Snappy.cs:35-36
            static const uint8_t id[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};
            return s.pos + 0x10 < s.len && std::memcmp(s.p + s.pos, id, 10) == 0;
        }
        case FMT_LZ10: return s.pos + 0x8 < s.len && lz1x_validate(s, false);   // LZ10.cs:40-41
        case FMT_LZ11: return s.pos + 0x8 < s.len && lz1x_validate(s, true);    // LZ11.cs:39-40
        case FMT_LZO: {   // LZO.cs:31-39 (no file name given)
            int flag = s.PeekByte();
            return (flag > 11 && flag < 0x20) || (flag != -1 && flag < 0x10);
        }
        case FMT_PRS: return s.pos + 0x4 < s.len && prs_get_byte_order(s) >= 0;   // PRS.cs:31-32
        case FMT_LZHUDSON: return s.pos + 0x8 < s.len && s.ReadUInt32() != 0;     // LZHudson.cs:31-32 (no file name given)
        case FMT_SMSR00: return s.pos + 0x10 < s.len && s.Match("SMSR00", 6);   // SMSR00.cs:36-37
        case FMT_BLZ: {   // BLZ.cs:31-32: the footer's compressed size is the whole stream, footer size >= 8
            if (s.len < 8) return false;
            s.pos = s.len - 8;
            return int64_t(s.ReadUInt24()) == s.len && s.ReadByte() >= 8;
        }
        case FMT_LZ40: case FMT_LZ60:   // LZ40.cs:41-43, LZ60.cs:31-33: identifier and a non-zero size ("recognition is inaccurate!")
            return s.pos + 0x8 < s.len && s.ReadByte() == (fmt == FMT_LZ40 ? 0x40 : 0x60) && (s.ReadUInt24() != 0 || s.ReadUInt32() != 0);
        // wrapper formats (GCLZ.cs:30-31, CXLZ.cs:31-32, COMP.cs:30-31, 3DS-LZ.cs:29-30, LZ77.cs:46-47, LZOn.cs:29-30,
        // Level5LZSS.cs:29-30); Level5.IsMatch needs zlib and the file name: not restated
        case FMT_GCLZ: return s.pos + 0x8 < s.len && s.Match("GCLZ", 4) && s.pos + 0x8 < s.len && lz1x_validate(s, false);
        case FMT_CXLZ: return s.pos + 0x8 < s.len && s.Match("CXLZ", 4) && s.pos + 0x8 < s.len && lz1x_validate(s, false);
        case FMT_COMP: return s.pos + 0x8 < s.len && s.Match("COMP", 4) && s.pos + 0x8 < s.len && lz1x_validate(s, true);
        case FMT_LZ_3DS: return s.pos + 0x10 < s.len && s.Match("3DS-LZ\r\n", 8);
        case FMT_LZON: {
            static const uint8_t id[8] = {'L', 'Z', 'O', 'n', 0x00, 0x2F, 0xF1, 0x71};
            return s.pos + 0x10 < s.len && s.Match(id, 8);
        }
        case FMT_LEVEL5_LZSS: return s.pos + 0x10 < s.len && s.Match("SSZL", 4) && s.ReadUInt32() == 0;
        // the LZSS-property family: identifier only (AKLZ.cs:31-32, LZ01.cs:33-34, FCMP.cs:31-32, IECP.cs:30-31, MDB4.cs:28-29);
        // LZSega / GCZ have no identifier (length heuristic / file extension): not restated
        case FMT_AKLZ: {
            static const uint8_t id[12] = {'A', 'K', 'L', 'Z', '~', '?', 'Q', 'd', '=', 0xCC, 0xCC, 0xCD};
            return s.pos + 0x10 < s.len && s.Match(id, 12);
        }
        case FMT_SDPC: return s.pos + 0x10 < s.len && s.Match("SDPC", 4) && s.ReadUInt32() != 0;   // SDPC.cs:31-32
        case FMT_ECD: {   // ECD.cs:37-38 + GetDecompressedSizeStatic :43-51
            if (!(s.pos + 0x10 < s.len && s.Match("ECD", 3))) return false;
            s.pos += 5;
            if (uint64_t(s.ReadUInt32(Endian::Big)) + 0x10 > uint64_t(s.len)) return false;
            return s.ReadUInt32(Endian::Big) != 0;
        }
        case FMT_LZ00: return s.pos + 0x40 < s.len && s.Match("LZ00", 4);   // LZ00.cs:37-38
        case FMT_LZ01: return s.pos + 0x10 < s.len && s.Match("LZ01", 4);
        case FMT_FCMP: return s.pos + 0x10 < s.len && s.Match("FCMP", 4);
        case FMT_IECP: return s.pos + 0x10 < s.len && s.Match("IECP", 4);
        case FMT_MDB4: return s.pos + 0x10 < s.len && s.Match("MDB4", 4);
        case FMT_LZ77: {
            if (!(s.pos + 0x8 < s.len && s.Match("LZ77", 4))) return false;
            uint8_t t = s.ReadUInt8();
            return t == 0x10 || t == 0x11 || t == 0x24 || t == 0x28 || t == 0x30 || t == 0xF7;
        }
    }
    fail(NOT_SUPPORTED);
}

template <typename F>
static void parallel_for(size_t n, int threads, F&& f) {
    if (threads <= 0) threads = int(std::thread::hardware_concurrency());
    if (threads < 1) threads = 1;
    if (size_t(threads) > n) threads = int(n ? n : 1);
    if (threads == 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& th : pool) th.join();
}

}  // namespace ora

using namespace ora;

extern "C" {

int ora_abi_version(void) { return AURORA_ABI_VERSION; }
int ora_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

int ora_decode_batch(int format, const aurora_codec_opts* opts, size_t n, const uint8_t* src_base,
                     const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst_base,
                     const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t* out_len,
                     uint64_t* consumed, int32_t* status, int threads) {
    CodecOpts o = to_opts(opts);
    parallel_for(n, threads, [&](size_t i) {
        DecodeResult r = decode_one(format, o, src_base + src_off[i], int64_t(src_len[i]), dst_base + dst_off[i], int64_t(dst_cap[i]));
        if (out_len) out_len[i] = uint64_t(r.out_len);
        if (consumed) consumed[i] = uint64_t(r.consumed);
        if (status) status[i] = r.status;
    });
    return OK;
}

int ora_encode_batch(int format, const aurora_codec_opts* opts, size_t n, const uint8_t* src_base,
                     const uint64_t* src_off, const uint64_t* src_len, uint8_t* dst_base,
                     const uint64_t* dst_off, const uint64_t* dst_cap, uint64_t* out_len,
                     int32_t* status, int threads) {
    CodecOpts o = to_opts(opts);
    parallel_for(n, threads, [&](size_t i) {
        std::vector<uint8_t> out;
        int st = encode_one(format, o, src_base + src_off[i], int64_t(src_len[i]), out);
        if (st == OK) {
            if (out.size() > dst_cap[i]) st = DST_TOO_SMALL;
            else if (!out.empty()) std::memcpy(dst_base + dst_off[i], out.data(), out.size());
        }
        if (out_len) out_len[i] = out.size();
        if (status) status[i] = st;
    });
    return OK;
}

int ora_decoded_size_batch(int format, const aurora_codec_opts* opts, size_t n, const uint8_t* src_base,
                           const uint64_t* src_off, const uint64_t* src_len, int size_scan,
                           uint64_t* out_size, int32_t* status) {
    CodecOpts o = to_opts(opts);
    for (size_t i = 0; i < n; i++) {
        int st = OK;
        uint64_t sz = 0;
        try {
            Src s(src_base + src_off[i], int64_t(src_len[i]));
            sz = decoded_size(format, s, o);
        } catch (const Error& e) {
            st = e.status;
            if (st == NOT_SUPPORTED && size_scan) {   // size-only pre-pass: decode into a zero-capacity sink
                DecodeResult r = decode_one(format, o, src_base + src_off[i], int64_t(src_len[i]), nullptr, 0, true);
                st = (r.status == DST_TOO_SMALL) ? OK : r.status;
                sz = uint64_t(r.out_len);
            }
        }
        out_size[i] = sz;
        if (status) status[i] = st;
    }
    return OK;
}

int ora_is_match_batch(int format, const aurora_codec_opts* opts, size_t n, const uint8_t* src_base,
                       const uint64_t* src_off, const uint64_t* src_len, uint8_t* match) {
    CodecOpts o = to_opts(opts);
    for (size_t i = 0; i < n; i++) {
        bool m = false;
        try {
            Src s(src_base + src_off[i], int64_t(src_len[i]));
            m = is_match(format, s, o);
        } catch (const Error&) {
            m = false;
        }
        match[i] = m ? 1 : 0;
    }
    return OK;
}

uint64_t ora_xxh64(const uint8_t* p, size_t n, uint64_t seed) { return xxh64(p, n, seed); }
uint32_t ora_xxh32(const uint8_t* p, size_t n, uint32_t seed) { return xxh32(p, n, seed); }
uint32_t ora_crc32c(const uint8_t* p, size_t n) { return crc32c(p, n); }

void ora_lz_props_window(aurora_lz_props* out, int32_t windows_size, int32_t max_length, int32_t min_length,
                         int32_t windows_start, int32_t min_distance) {
    LzProps p = LzProps::Window(windows_size, max_length, min_length, windows_start, min_distance);
    *out = aurora_lz_props{p.WindowsBits, p.LengthBits, p.MinLength, p.MaxLength, p.MaxDistance, p.MinDistance, p.WindowsStart, 0};
}
void ora_lz_props_bits(aurora_lz_props* out, int32_t distance_bits, int32_t length_bits, int32_t threshold) {
    LzProps p = LzProps::Bits(distance_bits, length_bits, threshold);
    *out = aurora_lz_props{p.WindowsBits, p.LengthBits, p.MinLength, p.MaxLength, p.MaxDistance, p.MinDistance, p.WindowsStart, 0};
}

}  // extern "C"

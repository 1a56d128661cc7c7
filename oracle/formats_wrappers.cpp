// formats_wrappers.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of the wrapper formats of src/AuroraLib.Compression.Nintendo that put a header around a core this
// repository already has:
//   Nintendo/GCLZ.cs:39-53      "GCLZ" + LZ10            Sega/CXLZ.cs:41-55       "CXLZ" + LZ10
//   Sega/COMP.cs:39-53          "COMP" + LZ11            Nintendo/3DS-LZ.cs:37-50 "3DS-LZ\r\n" + LZ10
//   Nintendo/LZ77.cs:59-158     "LZ77" + type byte + u24 size: LZ10 / LZ11 / ChunkLZ10 (independent LZ10 chunks)
//   Level5/Level5.cs:63-148     u32 (type | size << 3): OnlySave / LZ10 headerless
//   Nintendo/LZOn.cs:41-80      "LZOn" 00 2F F1 71 + BE size + BE compressed size + LZO headerless
//   Level5/Level5LZSS.cs:41-72  "SSZL" + u32 + compressed size + size + LZSS headerless, Lzss0Properties
// and the LZSS-property family (a fixed header + LZSS.DecompressHeaderless / CompressHeaderless):
//   src/AuroraLib.Compression.Sega/Sega/AKLZ.cs:43-56, LZ01.cs:47-82, LZSega.cs:49-67
//   src/AuroraLib.Compression-Extended/Marvelous/FCMP.cs:43-59, IECP.cs:42-55, Konami/GCZ.cs:40-51, Specialized/MDB4.cs:41-80
// and two LZSS wrappers with extra work around the core:
//   src/AuroraLib.Compression-Extended/Specialized/ECD.cs:56-121   "ECD" + flag + plain size + compressed size + size (BE),
//                                     plain bytes + LZSS(0x400, 0x42, 3, 0x3BE) headerless, or the stored payload
//   src/AuroraLib.Compression.Sega/Sega/LZ00.cs:50-203             64-byte header, LZSS (Lzss0) body under StreamTransformer:
//                                     every byte is XORed with a value derived from a 32-bit LCG key that steps per byte
// The Huffman / RLE / zlib sub-types of LZ77 and Level5 are outside the LZ hot path: NOT_SUPPORTED here and on the GPU.
#include "oracle_core.hpp"

namespace ora {

void lz10_headerless(Src& source, Sink& destination, uint32_t decomLength);
void lz11_headerless(Src& source, Sink& destination, uint32_t decomLength);
void lzss_headerless(Src& source, Sink& destination, uint32_t decomLength, const LzProps& lz, uint8_t initialFill);

static const uint8_t kLzonMagic[8] = {'L', 'Z', 'O', 'n', 0x00, 0x2F, 0xF1, 0x71};
static const LzProps kLzss0 = LzProps::Window(0x1000, 0xF + 3, 3, 0xFEE);   // LZSS.cs:34

static void prefixed(Src& s, Sink& d, const char* magic, int n, bool lz11) {
    s.MatchThrow(magic, n);
    if (lz11) lz11_decode(s, d);
    else lz10_decode(s, d);
}

// LZ77.cs:108-157
static void lz77_decode(Src& source, Sink& destination) {
    source.MatchThrow("LZ77", 4);
    uint8_t type = source.ReadUInt8();
    uint32_t decompressedSize = source.ReadUInt24();
    if (decompressedSize == 0) decompressedSize = source.ReadUInt32();
    switch (type) {
        case 0x10: lz10_headerless(source, destination, decompressedSize); break;
        case 0x11: lz11_headerless(source, destination, decompressedSize); break;
        case 0xF7: {
            int64_t destinationEndPosition = destination.pos + decompressedSize;
            std::vector<uint16_t> segmentEndOffsets;
            do {
                segmentEndOffsets.push_back(source.ReadUInt16(Endian::Little));
            } while (int64_t(segmentEndOffsets.back()) + source.pos != source.len);
            int64_t headerEndOffset = source.pos;
            for (size_t i = 0; i < segmentEndOffsets.size(); i++) {
                lz10_decode(source, destination);
                source.pos = segmentEndOffsets[i] + headerEndOffset;
            }
            if (destination.pos > destinationEndPosition)
                fail(SIZE_MISMATCH, decompressedSize, destination.pos - (destinationEndPosition - decompressedSize));
            break;
        }
        default: fail(NOT_SUPPORTED);   // HUF20 / RLE30 sub-types and undefined values
    }
}

// Level5.cs:63-113
static void level5_decode(Src& source, Sink& destination) {
    uint32_t typeAndSize = source.ReadUInt32();
    if (source.pos >= source.len) fail(END_OF_STREAM);    // Peek<byte>() at the end of the stream
    if (source.PeekByte() == 0x78) fail(NOT_SUPPORTED);   // zlib payload
    uint32_t type = typeAndSize & 7, decompressedSize = typeAndSize >> 3;
    switch (type) {
        case 0: {   // OnlySave
            source.need(decompressedSize);
            destination.Write(source.p + source.pos, decompressedSize);
            source.pos += decompressedSize;
            break;
        }
        case 1: lz10_headerless(source, destination, decompressedSize); break;
        default: fail(NOT_SUPPORTED);
    }
}

// LZOn.cs:41-61
static void lzon_decode(Src& source, Sink& destination) {
    source.MatchThrow(kLzonMagic, 8);
    uint32_t decompressedSize = source.ReadUInt32(Endian::Big);
    (void)source.ReadUInt32(Endian::Big);
    int64_t start = destination.pos;
    lzo_decode(source, destination);
    if (destination.pos - start != int64_t(decompressedSize)) fail(SIZE_MISMATCH, decompressedSize, destination.pos - start);
}

// AuroraLib.Compression-Extended/Specialized/SDPC.cs:43-56
static void sdpc_decode(Src& source, Sink& destination) {
    source.MatchThrow("SDPC", 4);
    uint32_t decompressedSize = source.ReadUInt32();
    int64_t endPosition = destination.pos + decompressedSize;
    destination.SetLength(endPosition);
    lzo_decode(source, destination);
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decompressedSize, destination.pos - (endPosition - decompressedSize));
}

// Level5LZSS.cs:41-57
static void sszl_decode(Src& source, Sink& destination, const CodecOpts& o) {
    source.MatchThrow("SSZL", 4);
    (void)source.ReadUInt32();
    (void)source.ReadUInt32();
    uint32_t decompressedSize = source.ReadUInt32();
    lzss_headerless(source, destination, decompressedSize, kLzss0, uint8_t(o.lzssInitialFill));
}

bool is_wrapper_format(int fmt) { return fmt >= FMT_GCLZ && fmt <= FMT_LZ00; }

// ---- ECD (ECD.cs)
static const LzProps kEcdProps = LzProps::Window(0x400, 0x42, 3, 0x3BE);   // ECD.cs:18

static void ecd_decode(Src& source, Sink& destination) {   // ECD.cs:56-86
    source.MatchThrow("ECD", 3);
    bool isCompressed = source.ReadByte() == 1;
    uint32_t plainSize = source.ReadUInt32(Endian::Big);
    (void)source.ReadUInt32(Endian::Big);   // compressed size: only traced on mismatch
    uint32_t decompressedSize = source.ReadUInt32(Endian::Big);
    if (isCompressed) {
        // destination.WriteByte((byte)source.ReadByte()): past the end of the source ReadByte() is -1, i.e. 0xFF is written.
        // A fixed-size destination refuses the first byte past its capacity (NotSupportedException).
        for (uint32_t i = 0; i < plainSize; i++) {
            if (destination.pos >= destination.cap && !destination.size_only) fail(DST_TOO_SMALL);
            destination.WriteByte(uint8_t(source.ReadByte()));
        }
        lzss_headerless(source, destination, decompressedSize - plainSize, kEcdProps, 0);
    } else {
        // source.CopyTo(destination): everything that is left
        const int64_t rest = source.len - source.pos;
        if (rest > 0) {
            destination.Write(source.p + source.pos, rest);
            source.pos = source.len;
        }
    }
}

static void ecd_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o) {   // ECD.cs:88-121
    bool isCompressed = o.settings.Quality != 0 && n > 0x10;
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t start = out.size();
        const int plainSize = isCompressed ? o.ecdPlainSize : 0;
        if (plainSize > n) fail(INVALID_ARGUMENT);   // source.Slice(0, plainSize): ArgumentOutOfRangeException
        out.Write(reinterpret_cast<const uint8_t*>("ECD"), 3);
        out.WriteByte(isCompressed ? 1 : 0);
        out.WriteU32(uint32_t(plainSize), Endian::Big);
        out.WriteU32(isCompressed ? 0u : uint32_t(n), Endian::Big);
        out.WriteU32(uint32_t(n), Endian::Big);
        if (!isCompressed) {
            out.Write(src, size_t(n));
            return;
        }
        out.Write(src, size_t(plainSize));
        CodecOpts oo = o;
        oo.lzss = kEcdProps;
        OutBuf core;
        lzss_encode(src + plainSize, n - plainSize, core, oo);   // "LZSS" header (16 bytes) + body
        out.Write(core.v.data() + 0x10, core.size() - 0x10);
        const uint32_t compressedSize = uint32_t(out.size() - start - 0x10);
        out.PatchU32(start + 8, compressedSize, Endian::Big);
        if (compressedSize < uint32_t(n)) return;
        out.v.resize(start);   // ineffective: store (Compress(source, destination, CompressionLevel.NoCompression))
        isCompressed = false;
    }
}

// ---- LZ00 (LZ00.cs)
// StreamTransformer.GenerateNextKey (LZ00.cs:125-131): the shifts and subtractions multiply by 1103515245
static inline uint32_t lz00_next_key(uint32_t key) {
    uint32_t x = (((((((key << 1) + key) << 5) - key) << 5) + key) << 7) - key;
    x = (x << 6) - x;
    x = (x << 4) - x;
    return (x << 2) - x + 12345u;
}
static inline uint8_t lz00_transform(uint32_t& key, uint8_t value) {   // LZ00.cs:133-138
    key = lz00_next_key(key);
    const uint32_t t = (key >> 16) & 0x7FFF;
    return uint8_t(value ^ (((t << 8) - t) >> 15));
}

static void lz00_decode(Src& source, Sink& destination, const CodecOpts& o) {   // LZ00.cs:50-72
    source.MatchThrow("LZ00", 4);
    (void)source.ReadUInt32();   // sourceLength: only traced on mismatch
    source.pos += 8;
    source.need(32);             // Name = ReadString(32)
    source.pos += 32;
    uint32_t decompressedSize = source.ReadUInt32();
    uint32_t key = source.ReadUInt32();
    source.pos += 8;
    // the LZSS decoder reads the transformed stream strictly byte by byte, front to back: transforming the rest of
    // the source up front is the same thing
    const int64_t at = std::min(source.pos, source.len);
    std::vector<uint8_t> plain(source.p + at, source.p + source.len);
    for (auto& b : plain) b = lz00_transform(key, b);
    Src inner(plain.data(), int64_t(plain.size()));
    struct Sync { Src& outer; Src& in; int64_t base; ~Sync() { outer.pos = base + in.pos; } } sync{source, inner, source.pos};
    lzss_headerless(inner, destination, decompressedSize, kLzss0, uint8_t(o.lzssInitialFill));
}

static void lz00_encode(const uint8_t* src, int n, OutBuf& out, const CodecOpts& o) {   // LZ00.cs:83-110
    CodecOpts oo = o;
    oo.lzss = kLzss0;
    OutBuf core;
    lzss_encode(src, n, core, oo);
    const size_t blen = core.size() - 0x10;
    out.Write(reinterpret_cast<const uint8_t*>("LZ00"), 4);
    out.WriteU32(uint32_t(64 + blen));
    out.WriteU32(0);
    out.WriteU32(0);
    uint8_t name[32] = {'T', 'e', 'm', 'p', '.', 'd', 'a', 't'};   // LZ00.Name default, zero padded (WriteString(Name, 32, 0))
    out.Write(name, 32);
    out.WriteU32(uint32_t(n));
    out.WriteU32(o.lz00Key);
    out.WriteU32(0);
    out.WriteU32(0);
    uint32_t key = o.lz00Key;
    for (size_t i = 0; i < blen; i++) out.WriteByte(lz00_transform(key, core.v[0x10 + i]));
}

static const uint8_t kAklzMagic[12] = {'A', 'K', 'L', 'Z', '~', '?', 'Q', 'd', '=', 0xCC, 0xCC, 0xCD};   // Identifier("AKLZ~?Qd=ÌÌÍ")
static const LzProps kLzssDefault = LzProps::Bits(12, 4, 2);   // LZSS.cs:33

static void lzss_family_decode(int fmt, Src& source, Sink& destination, const CodecOpts& o) {
    uint32_t decompressedSize = 0;
    const LzProps* lz = &kLzss0;
    switch (fmt) {
        case FMT_AKLZ:   // AKLZ.cs:43-48
            source.MatchThrow(kAklzMagic, 12);
            decompressedSize = source.ReadUInt32(Endian::Big);
            lz = &kLzssDefault;
            break;
        case FMT_LZ01:   // LZ01.cs:47-62
            source.MatchThrow("LZ01", 4);
            (void)source.ReadUInt32();
            decompressedSize = source.ReadUInt32();
            (void)source.ReadUInt32();
            break;
        case FMT_FCMP:   // FCMP.cs:43-49
            source.MatchThrow("FCMP", 4);
            decompressedSize = source.ReadUInt32();
            (void)source.ReadUInt32();
            break;
        case FMT_IECP:   // IECP.cs:42-47
            source.MatchThrow("IECP", 4);
            decompressedSize = source.ReadUInt32();
            break;
        case FMT_MDB4:   // MDB4.cs:41-58
            source.MatchThrow("MDB4", 4);
            (void)source.ReadUInt32();
            decompressedSize = source.ReadUInt32();
            (void)source.ReadUInt32();
            source.pos += 16;   // Skip(4 * 4): a seek, not a read
            lz = &kLzssDefault;
            break;
        case FMT_LZSEGA:   // LZSega.cs:49-54
            (void)source.ReadUInt32();
            decompressedSize = source.ReadUInt32();
            lz = &kLzssDefault;
            break;
        case FMT_GCZ:   // GCZ.cs:40-44
            decompressedSize = source.ReadUInt32();
            break;
    }
    lzss_headerless(source, destination, decompressedSize, *lz, uint8_t(o.lzssInitialFill));
}

static void lzss_family_encode(int fmt, const uint8_t* src, int n, OutBuf& out, const CodecOpts& o) {
    CodecOpts oo = o;
    oo.lzss = (fmt == FMT_AKLZ || fmt == FMT_MDB4 || fmt == FMT_LZSEGA) ? kLzssDefault : kLzss0;
    OutBuf core;
    lzss_encode(src, n, core, oo);   // "LZSS" header (16 bytes) + body; CompressHeaderless is the body
    const uint8_t* body = core.v.data() + 0x10;
    const size_t blen = core.size() - 0x10;
    switch (fmt) {
        case FMT_AKLZ: out.Write(kAklzMagic, 12); out.WriteU32(uint32_t(n), Endian::Big); break;
        case FMT_LZ01:
            out.Write(reinterpret_cast<const uint8_t*>("LZ01"), 4);
            out.WriteU32(uint32_t(16 + blen));   // the whole file (LZ01.cs:80-81)
            out.WriteU32(uint32_t(n));
            out.WriteU32(0);
            break;
        case FMT_FCMP:
            out.Write(reinterpret_cast<const uint8_t*>("FCMP"), 4);
            out.WriteU32(uint32_t(n));
            out.WriteU32(305397760u);
            break;
        case FMT_IECP: out.Write(reinterpret_cast<const uint8_t*>("IECP"), 4); out.WriteU32(uint32_t(n)); break;
        case FMT_MDB4:
            out.Write(reinterpret_cast<const uint8_t*>("MDB4"), 4);
            out.WriteU32(uint32_t(n) + 1);
            out.WriteU32(uint32_t(n));
            out.WriteU32(uint32_t(32 + blen - 0x10));   // Position - start - 0x10 (MDB4.cs:78)
            for (int i = 0; i < 4; i++) out.WriteU32(0);
            break;
        case FMT_LZSEGA: out.WriteU32(uint32_t(blen)); out.WriteU32(uint32_t(n)); break;
        case FMT_GCZ: out.WriteU32(uint32_t(n)); break;
    }
    out.Write(body, blen);
}

void wrapper_decode(int fmt, Src& s, Sink& d, const CodecOpts& o) {
    switch (fmt) {
        case FMT_GCLZ: prefixed(s, d, "GCLZ", 4, false); break;
        case FMT_CXLZ: prefixed(s, d, "CXLZ", 4, false); break;
        case FMT_COMP: prefixed(s, d, "COMP", 4, true); break;
        case FMT_LZ_3DS: prefixed(s, d, "3DS-LZ\r\n", 8, false); break;
        case FMT_LZ77: lz77_decode(s, d); break;
        case FMT_LEVEL5: level5_decode(s, d); break;
        case FMT_LZON: lzon_decode(s, d); break;
        case FMT_LEVEL5_LZSS: sszl_decode(s, d, o); break;
        case FMT_SDPC: sdpc_decode(s, d); break;
        case FMT_ECD: ecd_decode(s, d); break;
        case FMT_LZ00: lz00_decode(s, d, o); break;
        default:
            if (fmt < FMT_AKLZ || fmt > FMT_GCZ) fail(INVALID_ARGUMENT);
            lzss_family_decode(fmt, s, d, o);
    }
}

// ---------------------------------------------------------------------------------------------- encoders
static void strip_lz1x_header(const OutBuf& core, OutBuf& out) {   // CompressHeaderless == Compress minus the 4/8-byte header
    size_t h = (core.v.size() >= 4 && (core.v[1] | core.v[2] | core.v[3]) == 0) ? 8 : 4;
    out.Write(core.v.data() + h, core.v.size() - h);
}

void wrapper_encode(int fmt, const uint8_t* src, int n, OutBuf& out, const CodecOpts& o, int lz77_type, int chunk_size, int level5_type) {
    if (fmt >= FMT_AKLZ && fmt <= FMT_GCZ) return lzss_family_encode(fmt, src, n, out, o);
    if (fmt == FMT_ECD) return ecd_encode(src, n, out, o);
    if (fmt == FMT_LZ00) return lz00_encode(src, n, out, o);
    switch (fmt) {
        case FMT_GCLZ: out.Write(reinterpret_cast<const uint8_t*>("GCLZ"), 4); lz10_encode(src, n, out, o); break;
        case FMT_CXLZ: out.Write(reinterpret_cast<const uint8_t*>("CXLZ"), 4); lz10_encode(src, n, out, o); break;
        case FMT_COMP: out.Write(reinterpret_cast<const uint8_t*>("COMP"), 4); lz11_encode(src, n, out, o); break;
        case FMT_LZ_3DS: out.Write(reinterpret_cast<const uint8_t*>("3DS-LZ\r\n"), 8); lz10_encode(src, n, out, o); break;
        case FMT_LZ77: {   // LZ77.cs:59-105
            out.Write(reinterpret_cast<const uint8_t*>("LZ77"), 4);
            if (lz77_type == 0x11) { lz11_encode(src, n, out, o); break; }
            if (lz77_type == 0x10 || (lz77_type == 0xF7 && chunk_size >= n)) { lz10_encode(src, n, out, o); break; }
            if (lz77_type != 0xF7) fail(NOT_SUPPORTED);
            out.WriteU32(0xF7u | (uint32_t(n) << 8));
            int segments = (n + chunk_size - 1) / chunk_size;
            size_t table = out.size();
            for (int i = 0; i < segments; i++) out.WriteU16(0, Endian::Little);
            size_t headerEnd = out.size();
            for (int i = 0; i < segments; i++) {
                int start = i * chunk_size, size = std::min(chunk_size, n - start);
                lz10_encode(src + start, size, out, o);
                size_t end = out.size() - headerEnd;
                if (end > 0xFFFF) fail(INVALID_ARGUMENT);   // ArgumentOutOfRangeException: chunks too large
                out.v[table + 2 * i] = uint8_t(end);
                out.v[table + 2 * i + 1] = uint8_t(end >> 8);
            }
            break;
        }
        case FMT_LEVEL5: {   // Level5.cs:116-148
            int type = o.settings.Quality == 0 ? 0 : level5_type;
            out.WriteU32(uint32_t(type) | (uint32_t(n) << 3));
            if (type == 0) out.Write(src, size_t(n));
            else if (type == 1) {
                OutBuf core;
                lz10_encode(src, n, core, o);
                strip_lz1x_header(core, out);
            } else fail(NOT_SUPPORTED);
            break;
        }
        case FMT_SDPC:   // SDPC.cs:59-64
            out.Write(reinterpret_cast<const uint8_t*>("SDPC"), 4);
            out.WriteU32(uint32_t(n));
            lzo_encode(src, n, out, o);
            break;
        case FMT_LZON: {   // LZOn.cs:64-79
            out.Write(kLzonMagic, 8);
            out.WriteU32(uint32_t(n), Endian::Big);
            size_t at = out.size();
            out.WriteU32(0);
            lzo_encode(src, n, out, o);
            out.PatchU32(at, uint32_t(out.size() - 0x10), Endian::Big);
            break;
        }
        case FMT_LEVEL5_LZSS: {   // Level5LZSS.cs:60-72
            CodecOpts oo = o;
            oo.lzss = kLzss0;
            OutBuf core;
            lzss_encode(src, n, core, oo);
            out.Write(reinterpret_cast<const uint8_t*>("SSZL"), 4);
            out.WriteU32(0);
            out.WriteU32(uint32_t(core.size() - 0x10));
            out.WriteU32(uint32_t(n));
            out.Write(core.v.data() + 0x10, core.size() - 0x10);
            break;
        }
        default: fail(INVALID_ARGUMENT);
    }
}

// GetDecompressedSize of the wrappers (a Peek)
uint32_t wrapper_decoded_size(int fmt, Src& s) {
    switch (fmt) {
        case FMT_GCLZ: s.MatchThrow("GCLZ", 4); break;
        case FMT_CXLZ: s.MatchThrow("CXLZ", 4); break;
        case FMT_COMP: s.MatchThrow("COMP", 4); break;
        case FMT_LZ_3DS: s.MatchThrow("3DS-LZ\r\n", 8); break;
        case FMT_LZ77: {
            s.MatchThrow("LZ77", 4);
            s.need(1);
            s.pos += 1;
            uint32_t v = s.ReadUInt24();
            return v ? v : s.ReadUInt32();
        }
        case FMT_LEVEL5: {
            uint32_t v = s.ReadUInt32();
            return s.PeekByte() == 0x78 ? v : v >> 3;
        }
        case FMT_LZON: s.MatchThrow(kLzonMagic, 8); return s.ReadUInt32(Endian::Big);
        case FMT_LEVEL5_LZSS: s.MatchThrow("SSZL", 4); s.need(8); s.pos += 8; return s.ReadUInt32();
        case FMT_AKLZ: s.MatchThrow(kAklzMagic, 12); return s.ReadUInt32(Endian::Big);        // AKLZ.cs:35-40
        case FMT_LZ01: s.MatchThrow("LZ01", 4); s.pos += 4; return s.ReadUInt32();             // LZ01.cs:37-43
        case FMT_FCMP: s.MatchThrow("FCMP", 4); return s.ReadUInt32();                         // FCMP.cs:35-40
        case FMT_IECP: s.MatchThrow("IECP", 4); return s.ReadUInt32();                         // IECP.cs:34-39
        case FMT_MDB4: s.MatchThrow("MDB4", 4); (void)s.ReadUInt32(); return s.ReadUInt32();   // MDB4.cs:32-38
        case FMT_LZSEGA: s.pos = 4; return s.ReadUInt32();                                     // LZSega.cs:41-46 (Position = +4)
        case FMT_GCZ: return s.ReadUInt32();                                                   // GCZ.cs:37
        case FMT_SDPC: s.MatchThrow("SDPC", 4); return s.ReadUInt32();                         // SDPC.cs:35-40
        case FMT_ECD: {                                                                        // ECD.cs:43-51
            s.MatchThrow("ECD", 3);
            s.pos += 5;
            if (uint64_t(s.ReadUInt32(Endian::Big)) + 0x10 > uint64_t(s.len)) return 0;
            return s.ReadUInt32(Endian::Big);
        }
        case FMT_LZ00: s.MatchThrow("LZ00", 4); s.pos += 4 + 8 + 32; return s.ReadUInt32();     // LZ00.cs:41-47
        default: fail(INVALID_ARGUMENT);
    }
    // the prefixed LZ10 / LZ11 streams: type byte + u24 (LZ10.cs:47-57)
    uint8_t id = s.ReadUInt8();
    if (id != (fmt == FMT_COMP ? 0x11 : 0x10)) fail(INVALID_IDENTIFIER);
    uint32_t v = s.ReadUInt24();
    return v ? v : s.ReadUInt32();
}

}  // namespace ora

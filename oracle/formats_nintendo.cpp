// formats_nintendo.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of src/AuroraLib.Compression.Nintendo/Nintendo/{LZ10,LZ11,Yaz0,Yay0,MIO0}.cs.
#include "oracle_core.hpp"

namespace ora {

static const LzProps kLz10 = LzProps::Window(0x1000, 18, 3);            // LZ10.cs:25, MIO0.cs:28
static const LzProps kLz10Vram = LzProps::Window(0x1000, 18, 3, 0, 2);  // LZ10.cs:30
static const LzProps kLz11 = LzProps::Window(0x1000, 0x4000, 3);        // LZ11.cs:25
static const LzProps kLz11Vram = LzProps::Window(0x1000, 0x4000, 3, 0, 2);
static const LzProps kYay0 = LzProps::Window(0x1000, 0xff + 0x12, 3);   // Yay0.cs:27

// ------------------------------------------------------------------ LZ10
// LZ10.cs:47-57
static uint32_t lz1x_size(Src& s, uint8_t id) {
    uint8_t identifier = s.ReadUInt8();
    if (identifier != id) fail(INVALID_IDENTIFIER);
    uint32_t size = s.ReadUInt24();
    if (size == 0) size = s.ReadUInt32();
    return size;
}

// LZ10.cs:82-111
void lz10_headerless(Src& source, Sink& destination, uint32_t decomLength) {
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    {
        LzWindows buffer(&destination, kLz10.WindowsBits);
        FlagReader flag(&source, Endian::Big);
        while (destination.pos + buffer.Position() < endPosition) {
            if (flag.Readbit()) {
                uint8_t b1 = source.ReadUInt8();
                uint8_t b2 = source.ReadUInt8();
                int distance = ((b1 & 0xf) << 8 | b2) + 1;
                int length = (b1 >> 4) + 3;
                buffer.BackCopy(distance, length);
            } else {
                buffer.WriteByte(source.ReadUInt8());
            }
        }
    }
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

void lz10_decode(Src& s, Sink& d) { lz10_headerless(s, d, lz1x_size(s, 0x10)); }

// LZ10.cs:67-80 header, :113-137 body
static void lz1x_header(OutBuf& out, uint8_t id, int n) {
    if (n <= 0xFFFFFF) {
        out.WriteU32(uint32_t(id) | (uint32_t(n) << 8));
    } else {
        out.WriteU32(id);
        out.WriteU32(uint32_t(n));
    }
}

void lz10_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    lz1x_header(destination, 0x10, n);
    bool vram = o.vramMode < 0 ? true : o.vramMode != 0;   // LZ10.cs:33 default true
    int sourcePointer = 0;
    MatchFinder mf(vram ? kLz10Vram : kLz10, o.settings);
    FlagWriter flag(&destination, Endian::Big);
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        while (plain != 0) {
            plain--;
            flag.Buffer.WriteByte(source[sourcePointer++]);
            flag.WriteBit(false);
        }
        if (match.Length == 0) break;
        flag.Buffer.WriteU16(uint16_t((match.Length - 3) << 12 | ((match.Distance - 1) & 0xFFF)), Endian::Big);
        sourcePointer += match.Length;
        flag.WriteBit(true);
    }
    flag.Dispose();
}

// ------------------------------------------------------------------ LZ11
// LZ11.cs:83-133
void lz11_headerless(Src& source, Sink& destination, uint32_t decomLength) {
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    {
        LzWindows buffer(&destination, kLz11.WindowsBits);
        FlagReader flag(&source, Endian::Big);
        while (destination.pos + buffer.Position() < endPosition) {
            if (flag.Readbit()) {
                int distance, length;
                uint8_t b1 = source.ReadUInt8();
                uint8_t b2 = source.ReadUInt8();
                if (b1 >> 4 == 0) {
                    uint8_t b3 = source.ReadUInt8();
                    distance = ((b2 & 0xf) << 8 | b3) + 1;
                    length = ((b1 & 0xf) << 4 | b2 >> 4) + 17;
                } else if (b1 >> 4 == 1) {
                    uint8_t b3 = source.ReadUInt8();
                    uint8_t b4 = source.ReadUInt8();
                    distance = ((b3 & 0xf) << 8 | b4) + 1;
                    length = ((b1 & 0xf) << 12 | b2 << 4 | b3 >> 4) + 273;
                } else {
                    distance = ((b1 & 0xf) << 8 | b2) + 1;
                    length = (b1 >> 4) + 1;
                }
                buffer.BackCopy(distance, length);
            } else {
                buffer.WriteByte(source.ReadUInt8());
            }
        }
    }
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

void lz11_decode(Src& s, Sink& d) { lz11_headerless(s, d, lz1x_size(s, 0x11)); }

// LZ11.cs:65-81, :135-171
void lz11_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    lz1x_header(destination, 0x11, n);
    bool vram = o.vramMode < 0 ? false : o.vramMode != 0;  // LZ11.cs:29 default false
    int sourcePointer = 0;
    MatchFinder mf(vram ? kLz11Vram : kLz11, o.settings);
    FlagWriter flag(&destination, Endian::Big);
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        while (plain != 0) {
            plain--;
            flag.Buffer.WriteByte(source[sourcePointer++]);
            flag.WriteBit(false);
        }
        if (match.Length == 0) break;
        if (match.Length <= 16) {
            flag.Buffer.WriteU16(uint16_t((match.Length - 1) << 12 | ((match.Distance - 1) & 0xFFF)), Endian::Big);
        } else if (match.Length <= 272) {
            flag.Buffer.WriteByte(uint8_t(((match.Length - 17) & 0xFF) >> 4));
            flag.Buffer.WriteU16(uint16_t((match.Length - 17) << 12 | ((match.Distance - 1) & 0xFFF)), Endian::Big);
        } else {
            flag.Buffer.WriteU32(uint32_t(0x10000000 | ((match.Length - 273) & 0xFFFF) << 12 | ((match.Distance - 1) & 0xFFF)), Endian::Big);
        }
        sourcePointer += match.Length;
        flag.WriteBit(true);
    }
    flag.Dispose();
}

// ------------------------------------------------------------------ Yay0 core (shared with Yaz0)
// Yay0.cs:110-144.  The three sources may be one and the same stream (Yaz0) or three (Yay0).
static void yay0_core(FlagReader& flag, Src& compressedSource, Src& uncompressedSource, Sink& destination, uint32_t decomLength) {
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    {
        LzWindows buffer(&destination, kYay0.WindowsBits);
        while (destination.pos + buffer.Position() < endPosition) {
            if (flag.Readbit()) {
                buffer.WriteByte(uncompressedSource.ReadUInt8());
            } else {
                uint8_t b1 = compressedSource.ReadUInt8();
                uint8_t b2 = compressedSource.ReadUInt8();
                int distance = (((b1 & 0x0F) << 8) | b2) + 0x1;
                int length = b1 >> 4;
                if (length == 0)
                    length = uncompressedSource.ReadByte() + 0x12;   // BCL ReadByte: -1 at EOF -> 0x11 (Yay0.cs:131)
                else
                    length += 2;
                buffer.BackCopy(distance, length);
            }
        }
    }
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

// Yay0.cs:152-184
static void yay0_core_encode(const uint8_t* source, int n, OutBuf& compressedData, OutBuf& uncompressedData, FlagWriter& flag, const Settings& settings) {
    int sourcePointer = 0;
    MatchFinder mf(kYay0, settings);
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        while (plain != 0) {
            plain--;
            uncompressedData.WriteByte(source[sourcePointer++]);
            flag.WriteBit(true);
        }
        if (match.Length == 0) return;
        if (match.Length < 18) {
            compressedData.WriteU16(uint16_t((match.Distance - 0x1) | ((match.Length - 0x2) << 12)), Endian::Big);
        } else {
            compressedData.WriteU16(uint16_t((match.Distance - 0x1) & 0xFFF), Endian::Big);
            uncompressedData.WriteByte(uint8_t(match.Length - 0x12));
        }
        sourcePointer += match.Length;
        flag.WriteBit(false);
    }
}

// ------------------------------------------------------------------ Yaz0 / Yaz1
static uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

// Yaz0.cs:58-79 (+ :91-92).  Any exception in the first attempt -> retry with byte-swapped size.
void yaz0_decode(Src& source, Sink& destination, const CodecOpts& o, const char* magic) {
    source.MatchThrow(magic, 4);
    Endian order = o.byteOrderDefault ? Endian::Big : o.byteOrder;
    uint32_t decompressedSize = source.ReadUInt32(order);
    (void)source.ReadUInt32(order);   // MemoryAlignment
    (void)source.ReadUInt32(order);
    int64_t sourceDataStartPosition = source.pos;
    int64_t destinationStartPosition = destination.pos;
    try {
        FlagReader flag(&source, Endian::Big);
        yay0_core(flag, source, source, destination, decompressedSize);
    } catch (const Error&) {
        source.pos = sourceDataStartPosition;
        destination.pos = destinationStartPosition;
        decompressedSize = bswap32(decompressedSize);
        FlagReader flag(&source, Endian::Big);
        yay0_core(flag, source, source, destination, decompressedSize);
    }
}

// Yaz0.cs:82-98
void yaz0_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o, const char* magic) {
    Endian order = o.byteOrderDefault ? Endian::Big : o.byteOrder;
    destination.Write(reinterpret_cast<const uint8_t*>(magic), 4);
    destination.WriteU32(uint32_t(n), order);
    destination.WriteU32(o.yaz0Alignment, order);
    destination.WriteU32(0);
    FlagWriter flag(&destination, Endian::Big);
    yay0_core_encode(source, n, flag.Buffer, flag.Buffer, flag, o.settings);
    flag.Dispose();
}

// ------------------------------------------------------------------ LZ40 / LZ60
// AuroraLib.Compression.Nintendo/Nintendo/LZ40.cs:47-52 (header as LZ10 with identifier 0x40), :73-124 (body), :126-168
// (encoder); LZ60.cs:53-71 is the same codec under identifier 0x60.
static void lz40_headerless(Src& source, Sink& destination, uint32_t decomLength) {
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    {
        LzWindows buffer(&destination, kLz11.WindowsBits);   // LzProperties(0x1000, 0x4000, 3): 12 window bits
        int flag = 0, flagbits = 0;
        while (destination.pos + buffer.Position() < endPosition) {
            if (flagbits == 0) {
                flag = uint8_t(-source.ReadByte());   // BCL ReadByte: -1 at the end -> flag 0x01, the token read then throws
                flagbits = 8;
            }
            if ((flag & 0x80) != 0) {
                int distance = source.ReadUInt16(Endian::Little);
                int length = distance & 0xF;
                distance >>= 4;
                if (length <= 1) {
                    if (length == 0) length = source.ReadUInt8() + 16;
                    else length = source.ReadUInt16(Endian::Little) + 272;
                }
                buffer.BackCopy(distance, length);
            } else {
                buffer.WriteByte(source.ReadUInt8());
            }
            flag <<= 1;
            flagbits--;
        }
    }
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

void lz40_decode(Src& s, Sink& d, uint8_t id) { lz40_headerless(s, d, lz1x_size(s, id)); }

void lz40_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o, uint8_t id) {
    lz1x_header(destination, id, n);
    bool vram = o.vramMode < 0 ? false : o.vramMode != 0;   // LZ40.cs:35 default false
    int sourcePointer = 0;
    MatchFinder mf(vram ? kLz11Vram : kLz11, o.settings);
    FlagWriter flag(&destination, Endian::Big);
    flag.negate8 = true;
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        while (plain != 0) {
            plain--;
            flag.Buffer.WriteByte(source[sourcePointer++]);
            flag.WriteBit(false);
        }
        if (match.Length == 0) break;
        // (ushort)(match.Distance << 4 | ...): a distance of 0x1000 wraps to 0, which BackCopy resolves one window back
        if (match.Length < 16) {
            flag.Buffer.WriteU16(uint16_t(match.Distance << 4 | match.Length), Endian::Little);
        } else if (match.Length < 272) {
            flag.Buffer.WriteU16(uint16_t(match.Distance << 4), Endian::Little);
            flag.Buffer.WriteByte(uint8_t(match.Length - 16));
        } else {
            flag.Buffer.WriteU16(uint16_t(match.Distance << 4 | 1), Endian::Little);
            flag.Buffer.WriteU16(uint16_t(match.Length - 272), Endian::Little);
        }
        sourcePointer += match.Length;
        flag.WriteBit(true);
    }
    flag.Dispose();
}

// ------------------------------------------------------------------ LZHudson
// AuroraLib.Compression.Nintendo/HudsonSoft/LZHudson.cs:41-59: u32 BE size, then the Yay0 token core with all three
// sub-streams = the source and a FlagReader over 4-byte big-endian flag words, MSB first.
void lzhudson_decode(Src& source, Sink& destination) {
    uint32_t decompressedSize = source.ReadUInt32(Endian::Big);
    FlagReader flag(&source, Endian::Big, 4, Endian::Big);
    yay0_core(flag, source, source, destination, decompressedSize);
}

void lzhudson_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    destination.WriteU32(uint32_t(n), Endian::Big);
    FlagWriter flag(&destination, Endian::Big, 4, Endian::Big);
    yay0_core_encode(source, n, flag.Buffer, flag.Buffer, flag, o.settings);
    flag.Dispose();
}

// ------------------------------------------------------------------ DetectByteOrder<uint>(3)
// AuroraLib.Core 1.7.0 (not in the tree): "parity unpinned".  Restated as the plausibility rule of
// an order is plausible when 0x10 <= compOff <= litOff <= total length.
static Endian detect_order3(const Src& s, const CodecOpts& o) {
    if (!o.byteOrderDefault) return o.byteOrder;   // explicit override
    if (s.pos + 12 > s.len) return Endian::Big;
    auto rd = [&](int i, bool be) {
        const uint8_t* q = s.p + s.pos + 4 * i;
        return be ? (uint32_t(q[0]) << 24 | uint32_t(q[1]) << 16 | uint32_t(q[2]) << 8 | q[3])
                  : (uint32_t(q[0]) | uint32_t(q[1]) << 8 | uint32_t(q[2]) << 16 | uint32_t(q[3]) << 24);
    };
    uint64_t total = uint64_t(s.len - (s.pos - 4));
    auto plausible = [&](bool be) {
        uint32_t c = rd(1, be), l = rd(2, be);
        return c >= 0x10 && c <= l && l <= total;
    };
    if (plausible(true)) return Endian::Big;
    if (plausible(false)) return Endian::Little;
    return Endian::Big;
}

// ------------------------------------------------------------------ Yay0
// Yay0.cs:50-60, :80-108
void yay0_decode(Src& source, Sink& destination, const CodecOpts& o) {
    const int flagDataStart = 0x10;
    uint32_t startPosition = uint32_t(source.pos);
    source.MatchThrow("Yay0", 4);
    Endian endian = detect_order3(source, o);
    uint32_t uncompressedSize = source.ReadUInt32(endian);
    uint32_t compressedDataPointer = source.ReadUInt32(endian) + startPosition;
    uint32_t uncompressedDataPointer = source.ReadUInt32(endian) + startPosition;
    int cdp = int(compressedDataPointer) - flagDataStart, udp = int(uncompressedDataPointer) - flagDataStart;

    int64_t dataLength = source.len - source.pos;
    const uint8_t* data = source.p + source.pos;
    source.pos = source.len;   // ReadExactly(data, 0, dataLength)
    if (cdp < 0 || cdp > dataLength || udp < 0 || udp > dataLength) fail(INVALID_DATA);   // Span.Slice throws ArgumentOutOfRange
    Src flagSource(data, dataLength);
    Src compressedSource(data + cdp, dataLength - cdp);
    Src uncompressedSource(data + udp, dataLength - udp);
    FlagReader flag(&flagSource, Endian::Big);
    auto rewind = [&]() {
        int64_t read = std::max<int64_t>(cdp + compressedSource.pos, udp + uncompressedSource.pos);
        source.pos -= dataLength - read;
    };
    try {
        yay0_core(flag, compressedSource, uncompressedSource, destination, uncompressedSize);
    } catch (const Error&) {
        throw;   // the reference leaves source.Position at the end of the stream on failure
    }
    rewind();
}

// Yay0.cs:63-78
void yay0_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    Endian order = o.byteOrderDefault ? Endian::Big : o.byteOrder;
    OutBuf compressedData, uncompressedData, flagData;
    {
        FlagWriter flag(&flagData, Endian::Big);
        yay0_core_encode(source, n, compressedData, uncompressedData, flag, o.settings);
        flag.Dispose();
    }
    destination.Write(reinterpret_cast<const uint8_t*>("Yay0"), 4);
    destination.WriteU32(uint32_t(n), order);
    destination.WriteU32(uint32_t(0x10 + flagData.size()), order);
    destination.WriteU32(uint32_t(0x10 + flagData.size() + compressedData.size()), order);
    destination.Write(flagData.v.data(), flagData.size());
    destination.Write(compressedData.v.data(), compressedData.size());
    destination.Write(uncompressedData.v.data(), uncompressedData.size());
}

// ------------------------------------------------------------------ MIO0
// MIO0.cs:51-61, :83-149
void mio0_decode(Src& source, Sink& destination, const CodecOpts& o) {
    const int flagDataStart = 0x10;
    uint32_t startPosition = uint32_t(source.pos);
    source.MatchThrow("MIO0", 4);
    Endian endian = detect_order3(source, o);
    uint32_t decomLength = source.ReadUInt32(endian);
    int compressedDataPointer = int(source.ReadUInt32(endian) + startPosition) - flagDataStart;
    int uncompressedDataPointer = int(source.ReadUInt32(endian) + startPosition) - flagDataStart;

    int64_t bufferLength = source.len - source.pos;
    const uint8_t* src = source.p + source.pos;
    source.pos = source.len;
    // The reference indexes the span lazily (IndexOutOfRangeException at the first bad access); the oracle
    // and the GPU path reject sub-stream pointers outside the blob up front, like Yay0's Span.Slice does.
    if (compressedDataPointer < 0 || compressedDataPointer > bufferLength || uncompressedDataPointer < 0 || uncompressedDataPointer > bufferLength)
        fail(INVALID_DATA);
    auto at = [&](int i) -> uint8_t {
        if (i < 0 || i >= bufferLength) fail(END_OF_STREAM);   // IndexOutOfRangeException on the span: input exhausted
        return src[i];
    };

    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    int64_t total;
    {
        LzWindows buffer(&destination, kLz10.WindowsBits);
        int flagDataPointer = 0;
        int maskBitCounter = 0, currentMask = 0;
        while (destination.pos + buffer.Position() < endPosition) {
            if (maskBitCounter == 0) {
                currentMask = at(flagDataPointer++);
                maskBitCounter = 8;
            }
            if ((currentMask & 0x80) == 0x80) {
                buffer.WriteByte(at(uncompressedDataPointer++));
            } else {
                uint8_t b1 = at(compressedDataPointer++);
                uint8_t b2 = at(compressedDataPointer++);
                int distance = (((b1 & 0x0F) << 8) | b2) + 0x1;
                int length = (b1 >> 4) + 3;
                buffer.BackCopy(distance, length);
            }
            currentMask <<= 1;
            maskBitCounter--;
        }
        total = destination.pos + buffer.Position();
        if (total > endPosition) {
            buffer.Dispose();
            fail(SIZE_MISMATCH, decomLength, total - (endPosition - decomLength));
        }
    }
    int64_t read = std::max(compressedDataPointer, uncompressedDataPointer);
    source.pos -= bufferLength - read;
}

// MIO0.cs:64-81, :159-184
void mio0_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    Endian order = o.byteOrderDefault ? Endian::Big : o.byteOrder;
    OutBuf compressedData, uncompressedData, flagData;
    {
        FlagWriter flag(&flagData, Endian::Big);
        int sourcePointer = 0;
        MatchFinder mf(kLz10, o.settings);
        while (true) {
            LzMatch match = mf.FindNextBestMatch(source, n);
            int plain = match.Offset - sourcePointer;
            while (plain != 0) {
                plain--;
                uncompressedData.WriteByte(source[sourcePointer++]);
                flag.WriteBit(true);
            }
            if (match.Length == 0) break;
            compressedData.WriteU16(uint16_t((match.Distance - 0x1) | ((match.Length - 0x3) << 12)), Endian::Big);
            sourcePointer += match.Length;
            flag.WriteBit(false);
        }
        flag.Dispose();
    }
    destination.Write(reinterpret_cast<const uint8_t*>("MIO0"), 4);
    destination.WriteU32(uint32_t(n), order);
    destination.WriteU32(uint32_t(0x10 + flagData.size()), order);
    destination.WriteU32(uint32_t(0x10 + flagData.size() + compressedData.size()), order);
    destination.Write(flagData.v.data(), flagData.size());
    destination.Write(compressedData.v.data(), compressedData.size());
    destination.Write(uncompressedData.v.data(), uncompressedData.size());
}

// ------------------------------------------------------------------ SMSR00
// AuroraLib.Compression.Nintendo/Nintendo/SMSR00.cs:49-57 (header), :77-141 (body), :60-75 + :143-147 (encoder).
// Header 0x10 bytes: "SMSR00", 2 skipped, u32 BE size, u32 BE pointer to the literal section.  The code section between
// the header and that pointer interleaves 16-bit big-endian mask words (MSB first, 1 = literal) with the 16-bit big-endian
// match codes of their 0 bits (MIO0's code layout); literals are read from the stream behind the code section.
void smsr00_decode(Src& source, Sink& destination) {
    source.MatchThrow("SMSR00", 6);
    source.pos += 2;   // Skip(2): a seek
    uint32_t decomLength = source.ReadUInt32(Endian::Big);
    uint32_t uncompressedDataPointer = source.ReadUInt32(Endian::Big);
    int codesLength = int(int64_t(uncompressedDataPointer) - source.pos);
    if (codesLength < 0) fail(INVALID_DATA);   // ArrayPool.Rent(negative): ArgumentOutOfRangeException
    source.need(codesLength);                  // ReadExactly
    const uint8_t* codes = source.p + source.pos;
    source.pos += codesLength;
    const int ncodes = codesLength / 2;        // MemoryMarshal.Cast<byte, ushort>: an odd trailing byte is not addressable
    auto code = [&](int i) -> uint16_t {
        if (i >= ncodes) { source.pos = source.len; fail(END_OF_STREAM); }   // IndexOutOfRangeException on the span: input exhausted
        return uint16_t(codes[2 * i] << 8 | codes[2 * i + 1]);               // ReverseEndianness of the little-endian cast
    };
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    {
        LzWindows buffer(&destination, kLz10.WindowsBits);   // LzProperties(0x1000, 18, 3)
        int codePointer = 0, maskBitCounter = 0;
        uint16_t currentMask = 0;
        while (destination.pos + buffer.Position() < endPosition) {
            if (maskBitCounter == 0) {
                currentMask = code(codePointer++);
                maskBitCounter = 16;
            }
            if ((currentMask & 0x8000) == 0x8000) {
                buffer.WriteByte(source.ReadUInt8());
            } else {
                uint16_t data = code(codePointer++);
                buffer.BackCopy((data & 0x0FFF) + 1, (data >> 12) + 3);
            }
            currentMask = uint16_t(currentMask << 1);
            maskBitCounter--;
        }
    }
    if (destination.pos > endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

void smsr00_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    OutBuf uncompressedData, codeData;
    {
        FlagWriter flag(&codeData, Endian::Big, 2, Endian::Big);
        int sourcePointer = 0;
        MatchFinder mf(kLz10, o.settings);   // MIO0.CompressHeaderless with MIO0's LzProperties(0x1000, 18, 3)
        while (true) {
            LzMatch match = mf.FindNextBestMatch(source, n);
            int plain = match.Offset - sourcePointer;
            while (plain != 0) {
                plain--;
                uncompressedData.WriteByte(source[sourcePointer++]);
                flag.WriteBit(true);
            }
            if (match.Length == 0) break;
            flag.Buffer.WriteU16(uint16_t((match.Distance - 0x1) | ((match.Length - 0x3) << 12)), Endian::Big);
            sourcePointer += match.Length;
            flag.WriteBit(false);
        }
        flag.Dispose();
    }
    destination.Write(reinterpret_cast<const uint8_t*>("SMSR00"), 6);
    destination.WriteU16(0, Endian::Little);
    destination.WriteU32(uint32_t(n), Endian::Big);
    destination.WriteU32(uint32_t(0x10 + codeData.size()), Endian::Big);
    destination.Write(codeData.v.data(), codeData.size());
    destination.Write(uncompressedData.v.data(), uncompressedData.size());
}

// ------------------------------------------------------------------ BLZ
// AuroraLib.Compression.Nintendo/Nintendo/BLZ.cs: the stream is parsed from its END.  Footer (last 8 bytes): u24 LE compressed
// size (codes + padding + footer), u8 footer-and-padding size (>= 8), i32 LE (decoded size - compressed size).  The codes are
// read backwards (:104-141): flag byte (MSB first, 0 = literal), literals and big-endian-when-reversed u16 codes
// (length - 3) << 12 | (distance - 3), and the output is written backwards.  Decompress (:45-69) decodes into a rented
// buffer and writes it to the destination only when the decode succeeded.
static const LzProps kBlz = LzProps::Window(0x1000, 18, 3, 0, 3);   // BLZ.cs:24

void blz_decode(Src& source, Sink& destination) {
    const int64_t length = source.len;
    if (length < 8) { source.pos = length; fail(END_OF_STREAM); }   // Position = Length - 8 is negative / the footer reads run out
    source.pos = length - 8;
    const uint32_t compressedSize = source.ReadUInt24();
    const uint8_t headerAndPaddingSize = source.ReadUInt8();
    const int32_t decompressedSize = int32_t(uint32_t(source.ReadUInt32()) + compressedSize);
    const int32_t codeSize = int32_t(compressedSize) - headerAndPaddingSize;
    if (headerAndPaddingSize < 8) fail(INVALID_DATA);                // "Invalid BLZ header."
    if (int64_t(compressedSize) > length) fail(INVALID_DATA);        // Position = negative: ArgumentOutOfRangeException
    if (codeSize < 0 || decompressedSize < 0) fail(INVALID_DATA);    // ArrayPool.Rent(negative)
    source.pos = length - compressedSize;
    const uint8_t* in = source.p + source.pos;
    source.pos += codeSize;                                          // source.Read(inBuffer, 0, codeSize)
    std::vector<uint8_t> out(size_t(decompressedSize) + 1);
    {
        int src = codeSize, dst = decompressedSize;
        int flags = 0, mask = 0;
        while (src > 0) {
            if ((mask >>= 1) == 0) {
                flags = in[--src];
                mask = 0x80;
            }
            if ((flags & mask) == 0) {
                if (dst == 0) fail(INVALID_DATA);     // destination[--dst]: IndexOutOfRangeException (evaluated before the source read)
                if (src == 0) fail(END_OF_STREAM);    // source[--src]
                out[size_t(--dst)] = in[--src];
            } else {
                if (src < 2) fail(END_OF_STREAM);
                int info = (in[src - 1] << 8) | in[src - 2];
                src -= 2;
                int distance = (info & 0x0FFF) + 3;
                int len = ((info >> 12) & 0xF) + 3;
                for (int i = 0; i < len && dst > 0; i++) {
                    if (dst - 1 + distance >= decompressedSize) fail(INVALID_DATA);   // reads past the end of the buffer
                    out[size_t(dst - 1)] = out[size_t(dst - 1 + distance)];
                    dst--;
                }
            }
        }
        if (dst != 0) fail(SIZE_MISMATCH, uint32_t(decompressedSize), int64_t(decompressedSize) - dst);
    }
    // destination.Write(outBuffer, 0, decompressedSize): a fixed-size destination refuses the whole write
    if (destination.pos + decompressedSize > destination.cap && !destination.size_only) fail(DST_TOO_SMALL);
    destination.Write(out.data(), decompressedSize);
}

void blz_encode(const uint8_t* sourceIn, int n, OutBuf& destinationStream, const CodecOpts& o) {
    // CompressHeaderless (:143-215): the match finder runs over the REVERSED source, tokens are written from the end of a buffer
    std::vector<uint8_t> destination(size_t(n) + size_t(n) / 5 + 64);
    std::vector<uint8_t> source(sourceIn, sourceIn + n);
    std::reverse(source.begin(), source.end());
    int src = 0, dst = int(destination.size()) - 2, flag = 1, flagPos = int(destination.size()) - 1;
    MatchFinder mf(kBlz, o.settings);
    auto WriteFlag = [&]() {
        if ((flag & 0x100) != 0) {
            destination[size_t(flagPos)] = uint8_t(flag);
            flag = 1;
            flagPos = dst--;
        }
    };
    while (src < n) {
        LzMatch match = mf.FindNextBestMatch(source.data(), n);
        int plain = match.Offset - src;
        while (plain != 0) {
            flag <<= 1;
            destination[size_t(dst--)] = source[size_t(src++)];
            WriteFlag();
            plain--;
        }
        if (match.Length == 0) break;
        flag <<= 1;
        int lzCode = uint16_t((match.Length - 3) << 12 | ((match.Distance - 3) & 0xFFF));
        destination[size_t(dst--)] = uint8_t(lzCode >> 8);
        destination[size_t(dst--)] = uint8_t(lzCode);
        src += match.Length;
        flag |= 1;
        WriteFlag();
    }
    if (flag != 1) {
        while ((flag & 0x100) == 0) flag <<= 1;   // == flag << (8 - Log2(flag)): left-align the partial flag byte
        destination[size_t(flagPos)] = uint8_t(flag);
    } else {
        dst++;   // no flag written, one step back
    }
    const int compressedSize = int(destination.size()) - dst - 1;
    // Compress (:71-99): the codes, 0xFF padding to 16 bytes, then the footer
    destinationStream.Write(destination.data() + destination.size() - size_t(compressedSize), size_t(compressedSize));
    int headerSize = 8;
    int totalSize = compressedSize + headerSize;
    const int padding = (16 - (totalSize % 16)) % 16;
    totalSize += padding;
    headerSize += padding;
    for (int i = 0; i < padding; i++) destinationStream.WriteByte(0xFF);
    destinationStream.WriteU24(uint32_t(totalSize), Endian::Little);
    destinationStream.WriteByte(uint8_t(headerSize));
    destinationStream.WriteU32(uint32_t(n - totalSize), Endian::Little);
}

// ------------------------------------------------------------------ sizes / IsMatch helpers
uint32_t nintendo_decoded_size(int fmt, Src& s, const CodecOpts& o) {
    switch (fmt) {
        case FMT_LZ10: return lz1x_size(s, 0x10);
        case FMT_LZ11: return lz1x_size(s, 0x11);
        case FMT_LZ40: return lz1x_size(s, 0x40);   // LZ40.cs:47-52
        case FMT_LZ60: return lz1x_size(s, 0x60);   // LZ60.cs:37-47
        case FMT_SMSR00: s.MatchThrow("SMSR00", 6); s.pos += 2; return s.ReadUInt32(Endian::Big);   // SMSR00.cs:40-46
        case FMT_BLZ: {   // BLZ.cs:35-43
            if (s.len < 8) { s.pos = s.len; fail(END_OF_STREAM); }
            s.pos = s.len - 8;
            uint32_t compressedSize = s.ReadUInt24();
            if (s.ReadUInt8() < 8) fail(INVALID_DATA);
            return s.ReadUInt32() + compressedSize;
        }
        case FMT_YAZ0:
        case FMT_YAZ1: {   // Yaz0.cs:50-55
            s.MatchThrow(fmt == FMT_YAZ0 ? "Yaz0" : "Yaz1", 4);
            return s.ReadUInt32(o.byteOrderDefault ? Endian::Big : o.byteOrder);
        }
        case FMT_YAY0: {   // Yay0.cs:41-47: detects the order, then reads big-endian regardless
            s.MatchThrow("Yay0", 4);
            (void)detect_order3(s, o);
            return s.ReadUInt32(Endian::Big);
        }
        case FMT_MIO0: {   // MIO0.cs:42-48
            s.MatchThrow("MIO0", 4);
            Endian e = detect_order3(s, o);
            return s.ReadUInt32(e);
        }
    }
    fail(NOT_SUPPORTED);
}

// LZ10.cs:139-175 / LZ11.cs:173-223 (Validate)
bool lz1x_validate(Src& source, bool lz11) {
    if (source.ReadByte() != (lz11 ? 0x11 : 0x10)) return false;
    uint32_t decompressedSize = source.ReadUInt24();
    if (decompressedSize == 0) decompressedSize = source.ReadUInt32();
    if (decompressedSize == 0) return false;
    int i = 3;
    int64_t Buffer = 0;
    FlagReader flag(&source, Endian::Big);
    while (source.pos < source.len) {
        if (flag.Readbit()) {
            int distance, length;
            uint8_t b1 = source.ReadUInt8();
            uint8_t b2 = source.ReadUInt8();
            if (lz11 && (b1 >> 4) == 0) {
                uint8_t b3 = source.ReadUInt8();
                distance = ((b2 & 0xf) << 8 | b3) + 1;
                length = ((b1 & 0xf) << 4 | b2 >> 4) + 17;
            } else if (lz11 && (b1 >> 4) == 1) {
                uint8_t b3 = source.ReadUInt8();
                uint8_t b4 = source.ReadUInt8();
                distance = ((b3 & 0xf) << 8 | b4) + 1;
                length = ((b1 & 0xf) << 12 | b2 << 4 | b3 >> 4) + 273;
            } else if (lz11) {
                distance = ((b1 & 0xf) << 8 | b2) + 1;
                length = (b1 >> 4) + 1;
            } else {
                distance = ((b1 & 0xf) << 8 | b2) + 1;
                length = (b1 >> 4) + 3;
            }
            if (distance > Buffer) return false;
            if (i == 0) return true;
            i--;
            Buffer += length;
        } else {
            source.pos++;
            Buffer++;
        }
    }
    return Buffer == int64_t(decompressedSize);
}

}  // namespace ora

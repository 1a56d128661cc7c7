// formats_common.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of src/AuroraLib.Compression/Formats/Common/{LZSS,LZ4,LZ4.Frame,LZ4.FrameDescriptor,
// LZ4Legacy,LZO,Snappy}.cs.
#include "oracle_core.hpp"

namespace ora {

// ================================================================== LZSS
// LZSS.cs:91-130
void lzss_headerless(Src& source, Sink& destination, uint32_t decomLength, const LzProps& lz, uint8_t initialFill) {
    int64_t endPosition = destination.pos + decomLength;
    destination.SetLength(endPosition);
    FlagReader flag(&source, Endian::Little);
    {
        LzWindows buffer(&destination, lz.WindowsBits, initialFill);
        int f = lz.GetLengthBitsFlag();
        int n = lz.GetWindowsFlag();
        while (destination.pos + buffer.Position() < endPosition) {
            if (flag.Readbit()) {
                buffer.WriteByte(source.ReadUInt8());
            } else {
                uint8_t b1 = source.ReadUInt8();
                uint8_t b2 = source.ReadUInt8();
                int offset = (b2 >> lz.LengthBits << 8) | b1;
                int length = (b2 & f) + lz.MinLength;
                offset = (lz.MaxDistance + offset - lz.WindowsStart) & n;
                buffer.OffsetCopy(offset, length);
            }
        }
    }
    if (destination.pos != endPosition) fail(SIZE_MISMATCH, decomLength, destination.pos - (endPosition - decomLength));
}

// LZSS.cs:53-69
void lzss_decode(Src& source, Sink& destination, const CodecOpts& o) {
    source.MatchThrow("LZSS", 4);
    uint32_t decompressedSize = source.ReadUInt32(Endian::Big);
    (void)source.ReadUInt32(Endian::Big);   // compressedSize (only traced on mismatch)
    (void)source.ReadUInt32(Endian::Big);
    lzss_headerless(source, destination, decompressedSize, o.lzss, uint8_t(o.lzssInitialFill));
}

// LZSS.cs:72-89, :132-160
void lzss_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    const LzProps& lz = o.lzss;
    size_t start = destination.size();
    destination.Write(reinterpret_cast<const uint8_t*>("LZSS"), 4);
    destination.WriteU32(uint32_t(n), Endian::Big);
    destination.WriteU32(0);
    destination.WriteU32(0);
    {
        int sourcePointer = 0;
        MatchFinder mf(lz, o.settings);
        FlagWriter flag(&destination, Endian::Little);
        int nn = lz.GetWindowsFlag();
        int f = lz.GetLengthBitsFlag();
        while (true) {
            LzMatch match = mf.FindNextBestMatch(source, n);
            int plain = match.Offset - sourcePointer;
            while (plain != 0) {
                plain--;
                flag.Buffer.WriteByte(source[sourcePointer++]);
                flag.WriteBit(true);
            }
            if (match.Length == 0) break;
            int offset = ((lz.WindowsStart + sourcePointer - match.Distance) & nn);
            flag.Buffer.WriteU16(uint16_t((offset & 0xFF) | (offset & 0xFF00) << lz.LengthBits | ((match.Length - lz.MinLength) & f) << 8), Endian::Little);
            flag.WriteBit(false);
            sourcePointer += match.Length;
        }
        flag.Dispose();
    }
    destination.PatchU32(start + 8, uint32_t(destination.size() - start - 0x10), Endian::Big);
}

// ================================================================== LZ4
static const LzProps kLz4 = LzProps::Window(0xFFFF, 0x7FFFFFFF, 4);   // LZ4.cs:29

static bool lz4_is_frame_magic(uint32_t v) {   // Enum.IsDefined(typeof(FrameTypes), v)  (LZ4.Frame.cs:52-95)
    return v == 0x184C2102u || v == 0x184D2204u || (v >= 0x184D2A50u && v <= 0x184D2A5Fu);
}

// LZ4.cs:241-252
static void lz4_read_ext(const uint8_t* src, int64_t n, int64_t& length, int64_t& sp) {
    if (length == 0xF) {
        int b;
        do {
            if (sp >= n) fail(END_OF_STREAM);   // IndexOutOfRangeException on the span
            b = src[sp++];
            length += b;
        } while (b == 255);
    }
}

// LZ4.cs:176-200
static void lz4_block(const uint8_t* src, int64_t n, LzWindows& buffer) {
    int64_t sp = 0;
    while (sp < n) {
        int token = src[sp++];
        int64_t plainLength = token >> 4;
        lz4_read_ext(src, n, plainLength, sp);
        if (sp + plainLength > n) fail(END_OF_STREAM);   // Slice -> ArgumentOutOfRangeException
        buffer.Write(src + sp, int(plainLength));
        sp += plainLength;
        if (sp >= n) break;
        int64_t matchLength = token & 0xF;
        if (sp + 2 > n) fail(END_OF_STREAM);
        int matchDistance = src[sp] | src[sp + 1] << 8;
        sp += 2;
        lz4_read_ext(src, n, matchLength, sp);
        buffer.BackCopy(matchDistance, int(matchLength + 4));
    }
}

// LZ4.cs:162-175 : a fresh window per call
static void lz4_block_stream(Src& source, Sink& destination, uint32_t compressedBlockSize) {
    LzWindows windows(&destination, kLz4.WindowsBits);
    // source.Read may return fewer bytes; the reference then decodes pool garbage.  Treated as truncation.
    if (source.pos + int64_t(compressedBlockSize) > source.len) { source.pos = source.len; fail(END_OF_STREAM); }
    const uint8_t* blk = source.p + source.pos;
    source.pos += compressedBlockSize;
    lz4_block(blk, compressedBlockSize, windows);
}

void lz4_block_decode(Src& s, Sink& d) { lz4_block_stream(s, d, uint32_t(s.len - s.pos)); }

// LZ4.cs:96-111
static uint32_t lz4_read_legacy(Src& source, Sink& destination) {
    uint32_t blockSize = source.ReadUInt32();
    do {
        lz4_block_stream(source, destination, blockSize);
        int b = source.ReadByte();
        if (int8_t(b) == -1) return 0;   // EOF, or the 0xFF end flag the encoder appends (LZ4.cs:133)
        source.pos--;
        blockSize = source.ReadUInt32();
    } while (!lz4_is_frame_magic(blockSize));
    return blockSize;
}

// LZ4.Frame.cs:107-174 with LZ4.FrameDescriptor.cs:18-26
static void lz4_frame(Src& source, Sink& destination, const CodecOpts& o) {
    int64_t destStartPos = destination.pos;
    uint8_t FLG = source.ReadUInt8();
    uint8_t BD = source.ReadUInt8();
    int64_t blockMaxSize;
    switch ((BD & 0x70) >> 4) {
        case 4: blockMaxSize = 0x10000; break;
        case 5: blockMaxSize = 0x40000; break;
        case 6: blockMaxSize = 0x100000; break;
        case 7: blockMaxSize = 0x400000; break;
        default: fail(INVALID_DATA);   // ArgumentOutOfRangeException
    }
    uint64_t contentSize = (FLG & 8) ? source.ReadUInt64LE() : 0;
    if (FLG & 1) (void)source.ReadUInt32();
    (void)source.ReadUInt8();   // header checksum: read, never verified
    if (FLG & 1) fail(NOT_SUPPORTED);
    {
        LzWindows windows(&destination, kLz4.WindowsBits);
        while (true) {
            uint32_t blockSize = source.ReadUInt32();
            if (blockSize == 0) break;
            bool isUncompressed = (blockSize & 0x80000000u) != 0;
            int64_t sz = blockSize & 0x7FFFFFFFu;
            if (sz > blockMaxSize) fail(INVALID_DATA);   // buffer.AsSpan(0, sz) on the rented array
            if (source.pos + sz > source.len) { source.pos = source.len; fail(END_OF_STREAM); }
            const uint8_t* blk = source.p + source.pos;
            source.pos += sz;
            if (FLG & 16) {
                uint32_t checksum = source.ReadUInt32();
                if (o.lz4Verify && checksum != xxh32(blk, size_t(sz))) fail(INVALID_DATA);
            }
            if (isUncompressed) windows.Write(blk, int(sz));
            else lz4_block(blk, sz, windows);
        }
    }
    if ((FLG & 8) && uint64_t(destination.pos) != uint64_t(destStartPos) + contentSize)
        fail(SIZE_MISMATCH, (long long)contentSize, destination.pos + destStartPos);
    if (FLG & 4) {
        uint32_t checksum = source.ReadUInt32();
        if (o.lz4Verify) {
            if (destination.overflowed()) fail(DST_TOO_SMALL);
            if (checksum != xxh32(destination.p + destStartPos, size_t(destination.pos - destStartPos))) fail(INVALID_DATA);
        }
    }
}

// LZ4.cs:50-94
void lz4_decode(Src& source, Sink& destination, const CodecOpts& o) {
    while (source.pos < source.len) {
        uint32_t magic = source.ReadUInt32();
    SwitchStart:
        if (magic == 0x184C2102u) {
            uint32_t blockSize = lz4_read_legacy(source, destination);
            if (blockSize == 0) return;
            magic = blockSize;
            goto SwitchStart;
        } else if (magic == 0x184D2204u) {
            lz4_frame(source, destination, o);
        } else if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {
            uint32_t blockSize = source.ReadUInt32();
            source.pos += blockSize;
        } else {
            source.pos -= 4;
            return;
        }
    }
}

// LZ4.cs:254-268
static void lz4_write_ext(OutBuf& out, int length) {
    length -= 0xF;
    if (length >= 0) {
        int byteToWrite;
        do {
            byteToWrite = std::min(length, 0xFF);
            out.WriteByte(uint8_t(byteToWrite));
            length -= byteToWrite;
        } while (byteToWrite == 0xFF);
    }
}

// LZ4.cs:202-238
void lz4_block_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    if (n < 5) fail(INVALID_ARGUMENT);   // source.Slice(0, Length - 5) throws ArgumentOutOfRangeException
    int sourcePointer = 0, plainLength, token;
    MatchFinder mf(kLz4, o.settings);
    const int encodeLen = n - 5;
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, encodeLen);
        plainLength = match.Offset - sourcePointer;
        token = (plainLength > 0xF ? 0xF : plainLength) << 4;
        if (match.Length != 0) {
            token |= (match.Length - 4 > 0xF ? 0xF : match.Length - 4);
        } else {
            plainLength = n - sourcePointer;
            token = (plainLength > 0xF ? 0xF : plainLength) << 4;
        }
        destination.WriteByte(uint8_t(token));
        lz4_write_ext(destination, plainLength);
        destination.Write(source + sourcePointer, size_t(plainLength));
        sourcePointer += plainLength;
        if (sourcePointer >= n) break;
        destination.WriteU16(uint16_t(match.Distance), Endian::Little);
        lz4_write_ext(destination, match.Length - 4);
        sourcePointer += match.Length;
    }
}

// LZ4.cs:114-160 (legacy) and LZ4.Frame.cs:176-227 (v1 frame; flags are wiped to IsVersion1, :184)
void lz4_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o, bool legacy) {
    if (legacy) {
        destination.WriteU32(0x184C2102u);
        int sourcePointer = 0;
        while (sourcePointer != n) {
            size_t blockStart = destination.size();
            destination.WriteU32(0);
            int blockLen = std::min(0x400000 * 2, n - sourcePointer);
            lz4_block_encode(source + sourcePointer, blockLen, destination, o);
            sourcePointer += blockLen;
            destination.PatchU32(blockStart, uint32_t(destination.size() - blockStart - 4), Endian::Little);
        }
        destination.WriteByte(0xFF);
        return;
    }
    destination.WriteU32(0x184D2204u);
    uint32_t blockSize = o.lz4BlockSize ? o.lz4BlockSize : 0x400000;
    uint8_t desc[3];
    desc[0] = 0x40;   // Flags &= IsVersion1
    switch (blockSize) {
        case 0x10000: desc[1] = 0x40; break;
        case 0x40000: desc[1] = 0x50; break;
        case 0x100000: desc[1] = 0x60; break;
        case 0x400000: desc[1] = 0x70; break;
        default: fail(INVALID_ARGUMENT);
    }
    desc[2] = uint8_t((xxh32(desc, 2) >> 8) & 0xFF);
    destination.Write(desc, 3);
    int sourcePointer = 0;
    while (sourcePointer != n) {
        OutBuf buffer;
        int blockLen = std::min(int(blockSize), n - sourcePointer);
        lz4_block_encode(source + sourcePointer, blockLen, buffer, o);
        if (buffer.size() >= blockSize) {
            destination.WriteU32(uint32_t(blockLen) | 0x80000000u);
            destination.Write(source + sourcePointer, size_t(blockLen));
        } else {
            destination.WriteU32(uint32_t(buffer.size()));
            destination.Write(buffer.v.data(), buffer.size());
        }
        sourcePointer += blockLen;
    }
    destination.WriteU32(0);
}

// ================================================================== LZO
static const LzProps kLzo = LzProps::Window(0xBFFF, 0x7FFFFFFF, 3);   // LZO.cs:24

// LZO.cs:252-262
static int lzo_read_ext(Src& source) {
    int b, length = 0;
    while ((b = source.ReadByte()) == 0) length += 255;
    if (b == -1) fail(END_OF_STREAM);
    return length + b;
}

// LZO.cs:49-139.  Every ReadByte() that returns -1 in the middle of a token feeds garbage into the
// reference's copy and then ends in EndOfStreamException (:138); the oracle raises it at once.
void lzo_decode(Src& source, Sink& destination) {
    int flag, length, distance, plain = 0;
    LzWindows buffer(&destination, kLzo.WindowsBits);
    auto rb = [&]() {
        int b = source.ReadByte();
        if (b < 0) fail(END_OF_STREAM);
        return b;
    };
    flag = rb();
    if (flag > 17) {
        length = flag - 17;
        buffer.CopyFrom(source, length);
        flag = rb();
    }
    do {
        int flagcode = flag >> 4;
        if (flagcode == 0) {
            if (plain == 0) {
                length = 3 + flag;
                if (length == 3) length = 18 + lzo_read_ext(source);
                plain = 4;
                buffer.CopyFrom(source, length);
                continue;
            } else if (plain <= 3) {
                distance = rb();
                distance = (distance << 2) + (flag >> 2) + 1;
                length = 2;
            } else {
                distance = rb();
                distance = (distance << 2) + (flag >> 2) + (2048 + 1);
                length = 3;
            }
        } else if (flagcode == 1) {
            length = 2 + (flag & 0x7);
            if (length == 2) length = 9 + lzo_read_ext(source);
            distance = 16384 + ((flag & 0x8) << 11);
            flag = rb();
            distance |= (rb() << 6 | flag >> 2);
            if (distance == 16384) return;
        } else if (flagcode <= 3) {
            length = 2 + (flag & 0x1f);
            if (length == 2) length = 33 + lzo_read_ext(source);
            flag = rb();
            distance = rb();
            distance = (distance << 6 | flag >> 2) + 1;
        } else if (flagcode <= 7) {
            length = 3 + ((flag >> 5) & 0x1);
            distance = rb();
            distance = (distance << 3) + ((flag >> 2) & 0x7) + 1;
        } else {
            length = 5 + ((flag >> 5) & 0x3);
            distance = rb();
            distance = (distance << 3) + ((flag & 0x1c) >> 2) + 1;
        }
        plain = flag & 0x3;
        buffer.BackCopy(distance, length);
        buffer.CopyFrom(source, plain);
    } while ((flag = source.ReadByte()) != -1);
    fail(END_OF_STREAM);
}

// LZO.cs:263-271
static void lzo_write_ext(OutBuf& destination, int value) {
    while (value > 255) {
        destination.WriteByte(0);
        value -= 255;
    }
    destination.WriteByte(uint8_t(value));
}

// LZO.cs:141-250
void lzo_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    if (n < 0x10) {
        destination.WriteByte(uint8_t(17 + n));
        destination.Write(source, size_t(n));
        destination.WriteByte(0x11);
        destination.WriteByte(0x0);
        destination.WriteByte(0x0);
        return;
    }
    int sourcePointer = 0;
    MatchFinder mf(kLzo, o.settings);
    LzMatch match = mf.FindNextBestMatch(source, n);
    LzMatch next = mf.FindNextBestMatch(source, n);
    while (sourcePointer != n) {
        int plain = match.Offset - sourcePointer;
        if (plain != 0) {
            if (plain < 4) {
                int dif = 4 - plain;
                match = LzMatch{match.Offset + dif, match.Distance, match.Length - dif};
                plain = 4;
            }
            if (plain > 18) {
                destination.WriteByte(0);
                lzo_write_ext(destination, plain - 18);
            } else {
                destination.WriteByte(uint8_t(plain - 3));
            }
            if (sourcePointer + plain > n) fail(INVALID_ARGUMENT);   // Slice -> ArgumentOutOfRangeException
            destination.Write(source + sourcePointer, size_t(plain));
            sourcePointer += plain;
        }
        if (match.Length >= kLzo.MinLength) {
            sourcePointer += match.Length;
            plain = next.Offset - sourcePointer;
            if (plain > 3) plain = 0;
            if (match.Length <= 8 && match.Distance <= 2048) {
                uint8_t flag = uint8_t(plain | (((match.Distance - 1) & 0x7) << 2));
                if (match.Length <= 4) destination.WriteByte(uint8_t(flag | 0x40 | ((match.Length - 3) << 5)));
                else destination.WriteByte(uint8_t(flag | 0x80 | ((match.Length - 5) << 5)));
                destination.WriteByte(uint8_t((match.Distance - 1) >> 3));
            } else if (match.Distance <= 16384) {
                if (match.Length > 33) {
                    destination.WriteByte(0x20);
                    lzo_write_ext(destination, match.Length - 33);
                } else {
                    destination.WriteByte(uint8_t(0x20 | (match.Length - 2)));
                }
                destination.WriteByte(uint8_t(plain | (match.Distance - 1) << 2));
                destination.WriteByte(uint8_t((match.Distance - 1) >> 6));
            } else {
                const int hFlag = 0x4000;
                int distance = match.Distance - hFlag;
                uint8_t flag = uint8_t(0x10 | ((distance & hFlag) >> 11));
                if (match.Length > 9) {
                    destination.WriteByte(flag);
                    lzo_write_ext(destination, match.Length - 9);
                } else {
                    destination.WriteByte(uint8_t(flag | (match.Length - 2)));
                }
                destination.WriteByte(uint8_t(plain | distance << 2));
                destination.WriteByte(uint8_t(distance >> 6));
            }
            if (plain < 0 || sourcePointer + plain > n) fail(INVALID_ARGUMENT);
            destination.Write(source + sourcePointer, size_t(plain));
            sourcePointer += plain;
        }
        match = next;
        next = mf.FindNextBestMatch(source, n);
    }
    destination.WriteByte(0x11);
    destination.WriteByte(0x0);
    destination.WriteByte(0x0);
}

// ================================================================== Snappy
static const LzProps kSnappy = LzProps::Window(0x8000, 63 + 1, 4);   // Snappy.cs:28
static const uint8_t kSnappyId[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};

// Snappy.cs:109-122
static uint32_t snappy_read_size(Src& source) {
    uint32_t result = 0;
    int shift = 0;
    int b = -1;
    while ((b & 0x80) != 0) {
        b = source.ReadUInt8();
        result |= shift < 32 ? uint32_t(b & 0x7F) << shift : 0;   // C#: shift count is taken mod 32; streams never get there
        shift += 7;
    }
    return result;
}

// Snappy.cs:205-250
void snappy_block_decode(Src& source, Sink& destination) {
    int64_t distance, length;
    uint32_t decompressedSize = snappy_read_size(source);
    int64_t endPosition = destination.pos + decompressedSize;
    destination.SetLength(endPosition);
    LzWindows buffer(&destination, kSnappy.WindowsBits + 1);
    while (destination.pos + buffer.Position() < endPosition) {
        int tag = source.ReadByte();
        if (tag < 0) fail(END_OF_STREAM);   // the reference reads a 4-byte offset next and throws there
        int type = tag & 0x3;
        length = tag >> 2;
        if (type == 0) {
            if (length >= 60) {
                int lenBytes = int(length) - 59;
                length = 0;
                for (int i = 0; i < lenBytes; i++) {
                    int b = source.ReadByte();
                    if (b < 0) fail(END_OF_STREAM);
                    length |= int64_t(uint32_t(b) << (8 * i));
                }
                length = int32_t(uint32_t(length));
            }
            if (length + 1 < 0) fail(INVALID_DATA);   // ReadExactly(count < 0) -> ArgumentOutOfRangeException
            buffer.CopyFrom(source, int(length + 1));
            continue;
        } else if (type == 1) {
            length = (length & 0x7) + 3;
            distance = ((tag >> 5) << 8) | source.ReadUInt8();
        } else if (type == 2) {
            distance = source.ReadUInt16(Endian::Little);
        } else {
            distance = int32_t(source.ReadUInt32(Endian::Little));
        }
        if (distance < 0 || distance > buffer.Length()) fail(INVALID_DATA);   // aliases in the reference's 64 KiB ring
        buffer.BackCopy(int(distance), int(length + 1));
    }
}

// Snappy.cs:39-68
void snappy_decode(Src& source, Sink& destination) {
    source.MatchThrow(kSnappyId, 10);
    while (source.pos < source.len) {
        int chunkType = source.ReadByte();
        int chunkLength = int(source.ReadUInt24(Endian::Little));
        if (chunkType == 0) {
            (void)source.ReadUInt32(Endian::Little);
            snappy_block_decode(source, destination);
        } else if (chunkType == 1) {
            (void)source.ReadUInt32(Endian::Little);
            int64_t n = chunkLength - 4;
            if (n < 0) fail(INVALID_DATA);
            n = std::min<int64_t>(n, source.len - source.pos);   // SubStream.CopyTo copies what is there
            destination.Write(source.p + source.pos, n);
            source.pos += n;
        } else {
            if (chunkType >= 0x02 && chunkType <= 0x7F) fail(INVALID_DATA);
            source.pos += chunkLength;
        }
    }
}

// Snappy.cs:130-203
static void snappy_headerless_encode(const uint8_t* source, int n, OutBuf& destination, MatchFinder& mf) {
    int v = n;
    while (v >= 0x80) {
        destination.WriteByte(uint8_t(v | 0x80));
        v >>= 7;
    }
    destination.WriteByte(uint8_t(v));
    int sourcePointer = 0;
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        if (plain > 0) {
            if (plain <= 60) {
                destination.WriteByte(uint8_t((plain - 1) << 2));
            } else {
                int len = plain - 1;
                if (len <= 0xFF) {
                    destination.WriteByte(uint8_t(60 << 2));
                    destination.WriteByte(uint8_t(len));
                } else if (len <= 0xFFFF) {
                    destination.WriteByte(uint8_t(61 << 2));
                    destination.WriteU16(uint16_t(len), Endian::Little);
                } else if (len <= 0xFFFFFF) {
                    destination.WriteByte(uint8_t(62 << 2));
                    destination.WriteU24(uint32_t(len), Endian::Little);
                } else {
                    destination.WriteByte(uint8_t(63 << 2));
                    destination.WriteU32(uint32_t(len));
                }
            }
            destination.Write(source + sourcePointer, size_t(plain));
            sourcePointer += plain;
        }
        if (match.Length == 0) return;
        sourcePointer += match.Length;
        if (match.Distance < 2048 && match.Length >= 4 && match.Length <= 11) {
            uint8_t tag = uint8_t(1 | ((match.Length - 4) << 2) | ((match.Distance >> 8) << 5));
            destination.WriteByte(tag);
            destination.WriteByte(uint8_t(match.Distance));
        } else {
            uint8_t tag = uint8_t(2 | ((match.Length - 1) << 2));
            destination.WriteByte(tag);
            destination.WriteU16(uint16_t(match.Distance), Endian::Little);
        }
    }
}

void snappy_block_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    MatchFinder mf(kSnappy, o.settings);
    snappy_headerless_encode(source, n, destination, mf);
}

// Snappy.cs:71-107 (CRC32C of the raw chunk, masked :252)
void snappy_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    destination.Write(kSnappyId, 10);
    MatchFinder mf(kSnappy, o.settings);
    int pos = 0;
    while (pos < n) {
        OutBuf buffer;
        int chunkSize = std::min(0x10000, n - pos);
        snappy_headerless_encode(source + pos, chunkSize, buffer, mf);
        mf.Reset();
        uint32_t crc = crc32c(source + pos, size_t(chunkSize));
        crc = ((crc >> 15) | (crc << 17)) + 0xa282ead8u;
        if (int(buffer.size()) >= chunkSize) {
            destination.WriteByte(1);
            destination.WriteU24(uint32_t(chunkSize + 4), Endian::Little);
            destination.WriteU32(crc);
            destination.Write(source + pos, size_t(chunkSize));
        } else {
            destination.WriteByte(0);
            destination.WriteU24(uint32_t(buffer.size() + 4), Endian::Little);
            destination.WriteU32(crc);
            destination.Write(buffer.v.data(), buffer.size());
        }
        pos += chunkSize;
    }
}

}  // namespace ora

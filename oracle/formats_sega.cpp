// formats_sega.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of src/AuroraLib.Compression.Sega/Sega/PRS.cs.
#include "oracle_core.hpp"

namespace ora {

static const LzProps kPrs = LzProps::Window(0x1FFF, 0x100, 2);   // PRS.cs:21

// PRS.cs:172-218.  Exceptions (EndOfStream inside a token) propagate to the caller like in the reference.
static bool prs_validate(Src& stream, Endian order) {
    int i = 3;
    int64_t startPos = stream.pos;
    FlagReader flag(&stream, order);
    int64_t Buffer = 0;
    struct Restore {
        Src& s; int64_t p;
        ~Restore() { s.pos = p; }
    } restore{stream, startPos};
    while (stream.pos < stream.len) {
        if (flag.Readbit()) {
            stream.pos++;
            Buffer++;
        } else {
            int distance, length;
            if (flag.Readbit()) {
                distance = stream.ReadUInt16(order);
                if (distance == 0) return true;
                length = distance & 7;
                distance = 0x2000 - (distance >> 3);
                length = length == 0 ? stream.ReadUInt8() + 1 : length + 2;
            } else {
                length = flag.ReadInt(2, true) + 2;
                distance = 0x100 - stream.ReadUInt8();
            }
            if (distance > Buffer) return false;
            if (i == 0) return true;
            i--;
            Buffer += length;
        }
    }
    return false;
}

// PRS.cs:161-170: 0 = Little, 1 = Big, -1 = null
int prs_get_byte_order(Src& stream) {
    int flag = stream.PeekByte();
    if (flag < 0) fail(END_OF_STREAM);
    if (flag > 12 && (flag & 0x1) == 1 && prs_validate(stream, Endian::Little)) return 0;
    if ((flag & 128) == 128 && prs_validate(stream, Endian::Big)) return 1;
    return -1;
}

// PRS.cs:59-102
static void prs_headerless(Src& source, Sink& destination, Endian order) {
    LzWindows buffer(&destination, kPrs.WindowsBits);
    FlagReader flag(&source, order);
    while (source.pos < source.len) {
        if (flag.Readbit()) {
            buffer.WriteByte(source.ReadUInt8());
        } else {
            int distance, length;
            if (flag.Readbit()) {
                distance = source.ReadUInt16(order);
                if (distance == 0) return;
                length = distance & 7;
                distance = 0x2000 - (distance >> 3);
                if (length == 0) length = source.ReadUInt8() + 1;
                else length += 2;
            } else {
                length = flag.ReadInt(2, true) + 2;
                distance = 0x100 - source.ReadUInt8();
            }
            buffer.BackCopy(distance, length);
        }
    }
    fail(END_OF_STREAM);
}

// PRS.cs:42-57
void prs_decode(Src& source, Sink& destination) {
    Endian detected = prs_get_byte_order(source) == 1 ? Endian::Big : Endian::Little;
    int64_t sourcePos = source.pos;
    int64_t destinationPos = destination.pos;
    try {
        prs_headerless(source, destination, detected);
    } catch (const Error&) {
        source.pos = sourcePos;
        destination.pos = destinationPos;
        prs_headerless(source, destination, detected == Endian::Big ? Endian::Little : Endian::Big);
    }
}

// PRS.cs:104-159
void prs_encode(const uint8_t* source, int n, OutBuf& destination, const CodecOpts& o) {
    Endian order = o.byteOrderDefault ? Endian::Big : o.byteOrder;   // PRS.cs:24 FormatByteOrder default Big
    int sourcePointer = 0;
    MatchFinder mf(kPrs, o.settings);
    FlagWriter flag(&destination, order);
    while (true) {
        LzMatch match = mf.FindNextBestMatch(source, n);
        int plain = match.Offset - sourcePointer;
        while (plain != 0) {
            plain--;
            flag.Buffer.WriteByte(source[sourcePointer++]);
            flag.WriteBit(true);
        }
        if (match.Length == 0) break;
        if (match.Length == 2 && match.Distance > 0x100) continue;
        sourcePointer += match.Length;
        int distance = match.Distance * -1;
        int length = match.Length;
        flag.WriteBit(false);
        if ((distance >= -0x100) && (length <= 5)) {
            flag.WriteBit(false);
            flag.WriteInt(length - 2, 2, true);
            flag.Buffer.WriteByte(uint8_t(distance));
            flag.FlushIfNecessary();
        } else {
            if (length > 9) {
                flag.Buffer.WriteU16(uint16_t(distance << 3), order);
                flag.Buffer.WriteByte(uint8_t(length - 1));
            } else {
                flag.Buffer.WriteU16(uint16_t((distance << 3) | (length - 2)), order);
            }
            flag.WriteBit(true);
        }
    }
    flag.WriteBit(false);
    flag.Buffer.WriteByte(0);
    flag.Buffer.WriteByte(0);
    flag.WriteBit(true);
    flag.Dispose();
}

}  // namespace ora

// GPU-backed drop-ins for the reference's format classes: same interfaces (ICompressionAlgorithm,
// IProvidesDecompressedSize, IEndianDependentFormat), work done by libaurora_cuda.so as a 1-element batch.
// UNTESTED here (no .NET toolchain in the build image).  See INTEGRATION.md.
using System;
using System.IO;
using AuroraLib.Compression.Exceptions;
using AuroraLib.Compression.Interfaces;
using AuroraLib.Core;
using AuroraLib.Core.Exceptions;
using AuroraLib.Core.Format;

namespace AuroraLib.Compression.Cuda
{
    public abstract unsafe class GpuCodec : ICompressionAlgorithm
    {
        internal static readonly Lazy<IntPtr> Context = new Lazy<IntPtr>(() =>
        {
            IntPtr ctx = Native.aurora_init(0);
            if (ctx == IntPtr.Zero) throw new PlatformNotSupportedException("aurora_init failed: no B200 visible (there is no CPU fallback)");
            return ctx;
        });

        protected abstract AuroraFormat Format { get; }
        protected abstract ICompressionAlgorithm Managed { get; }   // only for Info / IsMatch metadata
        public IFormatInfo Info => Managed.Info;
        public virtual bool IsMatch(Stream stream, ReadOnlySpan<char> fileNameAndExtension = default) => Managed.IsMatch(stream, fileNameAndExtension);

        protected virtual void FillOptions(ref AuroraCodecOpts o, CompressionSettings settings) { }

        internal static void ThrowFor(AuroraStatus st, long expected, long actual)
        {
            switch (st)
            {
                case AuroraStatus.Ok: return;
                case AuroraStatus.EndOfStream: throw new EndOfStreamException();
                case AuroraStatus.InvalidIdentifier: throw new InvalidIdentifierException();
                case AuroraStatus.SizeMismatch: throw new DecompressedSizeException(expected, actual);
                case AuroraStatus.DstTooSmall: throw new NotSupportedException("destination is not expandable");
                case AuroraStatus.InvalidData: throw new InvalidDataException();
                case AuroraStatus.NotSupported: throw new NotSupportedException();
                case AuroraStatus.InvalidArgument: throw new ArgumentException();
                default: throw new InvalidOperationException("CUDA error: " + System.Runtime.InteropServices.Marshal.PtrToStringAnsi(Native.aurora_last_error_string(Context.Value)));
            }
        }

        /// <inheritdoc/>
        public void Decompress(Stream source, Stream destination)
        {
            long start = source.Position;
            byte[] src = new byte[source.Length - start];
            source.ReadExactly(src, 0, src.Length);
            AuroraCodecOpts o;
            Native.aurora_codec_opts_init(&o);
            FillOptions(ref o, default);
            ulong off = 0, len = (ulong)src.Length, size = 0, outLen = 0, consumed = 0, dOff = 0;
            int st = 0;
            fixed (byte* ps = src)
            {
                Native.aurora_decoded_size_batch(Context.Value, (int)Format, &o, (UIntPtr)1, ps, &off, &len, 1, &size, &st);
                ulong cap = st == 0 ? size : 0;
                byte[] dst = new byte[checked((int)Math.Max(cap, 1UL))];   // (ulong, ulong) overload; a managed array is int-indexed
                fixed (byte* pd = dst)
                {
                    int rc = Native.aurora_decode_batch(Context.Value, (int)Format, &o, (UIntPtr)1, ps, &off, &len, pd, &dOff, &cap, &outLen, &consumed, &st);
                    if (rc != 0) ThrowFor((AuroraStatus)rc, 0, 0);
                }
                source.Position = start + (long)consumed;
                destination.Write(dst, 0, (int)Math.Min(outLen, cap));
                ThrowFor((AuroraStatus)st, (long)size, (long)outLen);
            }
        }

        /// <inheritdoc/>
        public void Compress(ReadOnlySpan<byte> source, Stream destination, CompressionSettings settings = default)
        {
            AuroraCodecOpts o;
            Native.aurora_codec_opts_init(&o);
            o.Quality = settings.Quality;
            o.MaxWindowBits = settings.MaxWindowBits;
            o.Strategy = (int)settings.Strategy;
            FillOptions(ref o, settings);
            ulong off = 0, len = (ulong)source.Length, dOff = 0, outLen = 0;
            ulong cap = Native.aurora_encode_bound((int)Format, len);
            byte[] dst = new byte[cap];
            int st = 0;
            fixed (byte* ps = source)
            fixed (byte* pd = dst)
            {
                int rc = Native.aurora_encode_batch(Context.Value, (int)Format, &o, (UIntPtr)1, ps, &off, &len, pd, &dOff, &cap, &outLen, &st);
                if (rc != 0) ThrowFor((AuroraStatus)rc, 0, 0);
            }
            ThrowFor((AuroraStatus)st, 0, 0);
            destination.Write(dst, 0, (int)outLen);
        }

        protected virtual long PeekBytes => 16;   // header bytes GetDecompressedSize looks at

        protected uint PeekSize(Stream source)
        {
            long start = source.Position;
            byte[] head = new byte[Math.Min(PeekBytes, source.Length - start)];
            source.ReadExactly(head, 0, head.Length);
            source.Position = start;
            AuroraCodecOpts o;
            Native.aurora_codec_opts_init(&o);
            FillOptions(ref o, default);
            ulong off = 0, len = (ulong)head.Length, size = 0;
            int st = 0;
            fixed (byte* ph = head)
                Native.aurora_decoded_size_batch(Context.Value, (int)Format, &o, (UIntPtr)1, ph, &off, &len, 0, &size, &st);
            ThrowFor((AuroraStatus)st, 0, 0);
            return (uint)size;
        }
    }

    public sealed class GpuYaz0 : GpuCodec, IEndianDependentFormat, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.Yaz0 _managed = new Formats.Nintendo.Yaz0();
        protected override AuroraFormat Format => AuroraFormat.Yaz0;
        protected override ICompressionAlgorithm Managed => _managed;
        public Endian FormatByteOrder { get; set; } = Endian.Big;
        public uint MemoryAlignment { get; set; } = 0;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.ByteOrder = FormatByteOrder == Endian.Big ? 1 : 0; o.Yaz0Alignment = MemoryAlignment; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ10 : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.LZ10 _managed = new Formats.Nintendo.LZ10();
        protected override AuroraFormat Format => AuroraFormat.LZ10;
        protected override ICompressionAlgorithm Managed => _managed;
        public bool GbaVramCompatibilityMode { get; set; } = true;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.VramMode = GbaVramCompatibilityMode ? 1 : 0; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    // ---- wrapper formats (AuroraLib.Compression.Nintendo): the header is resolved by libaurora_cuda.so on the host, the core
    // runs on the device; ChunkLZ10 chunks decode as independent streams of one batch
    public sealed class GpuGCLZ : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.GCLZ _managed = new Formats.Nintendo.GCLZ();
        protected override AuroraFormat Format => AuroraFormat.GCLZ;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuCXLZ : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.CXLZ _managed = new Formats.Nintendo.CXLZ();
        protected override AuroraFormat Format => AuroraFormat.CXLZ;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuCOMP : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.COMP _managed = new Formats.Nintendo.COMP();
        protected override AuroraFormat Format => AuroraFormat.COMP;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ_3DS : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.LZ_3DS _managed = new Formats.Nintendo.LZ_3DS();
        protected override AuroraFormat Format => AuroraFormat.LZ_3DS;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ77 : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.LZ77 _managed = new Formats.Nintendo.LZ77();
        protected override AuroraFormat Format => AuroraFormat.LZ77;
        protected override ICompressionAlgorithm Managed => _managed;
        public Formats.Nintendo.LZ77.CompressionType Type { get; set; } = Formats.Nintendo.LZ77.CompressionType.LZ10;
        public uint ChunkSize { get; set; } = 0x1000;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.Lz77Type = (uint)Type; o.Lz77ChunkSize = ChunkSize; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLevel5 : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Level5.Level5 _managed = new Formats.Level5.Level5();
        protected override AuroraFormat Format => AuroraFormat.Level5;
        protected override ICompressionAlgorithm Managed => _managed;
        public Formats.Level5.Level5.CompressionType Type { get; set; } = Formats.Level5.Level5.CompressionType.LZ10;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.Level5Type = (uint)Type; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZOn : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.LZOn _managed = new Formats.Nintendo.LZOn();
        protected override AuroraFormat Format => AuroraFormat.LZOn;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLevel5LZSS : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Level5.Level5LZSS _managed = new Formats.Level5.Level5LZSS();
        protected override AuroraFormat Format => AuroraFormat.Level5LZSS;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuAKLZ : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Sega.AKLZ _managed = new Formats.Sega.AKLZ();
        protected override AuroraFormat Format => AuroraFormat.AKLZ;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ01 : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Sega.LZ01 _managed = new Formats.Sega.LZ01();
        protected override AuroraFormat Format => AuroraFormat.LZ01;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuFCMP : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Marvelous.FCMP _managed = new Formats.Marvelous.FCMP();
        protected override AuroraFormat Format => AuroraFormat.FCMP;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuIECP : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Marvelous.IECP _managed = new Formats.Marvelous.IECP();
        protected override AuroraFormat Format => AuroraFormat.IECP;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuMDB4 : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Specialized.MDB4 _managed = new Formats.Specialized.MDB4();
        protected override AuroraFormat Format => AuroraFormat.MDB4;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZSega : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Sega.LZSega _managed = new Formats.Sega.LZSega();
        protected override AuroraFormat Format => AuroraFormat.LZSega;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuGCZ : GpuCodec, IProvidesDecompressedSize   // header + LZSS.DecompressHeaderless (wrappers.cu, kFamily)
    {
        private readonly Formats.Konami.GCZ _managed = new Formats.Konami.GCZ();
        protected override AuroraFormat Format => AuroraFormat.GCZ;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuSDPC : GpuCodec, IProvidesDecompressedSize   // "SDPC" + size + LZO.DecompressHeaderless
    {
        private readonly Formats.Specialized.SDPC _managed = new Formats.Specialized.SDPC();
        protected override AuroraFormat Format => AuroraFormat.SDPC;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZHudson : GpuCodec, IProvidesDecompressedSize   // u32 BE size + Yay0 tokens under 4-byte flag words
    {
        private readonly Formats.HudsonSoft.LZHudson _managed = new Formats.HudsonSoft.LZHudson();
        protected override AuroraFormat Format => AuroraFormat.LZHudson;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuBLZ : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.BLZ _managed = new Formats.Nintendo.BLZ();
        protected override AuroraFormat Format => AuroraFormat.BLZ;
        protected override ICompressionAlgorithm Managed => _managed;
        protected override long PeekBytes => long.MaxValue;   // the footer sits at Length - 8
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuSMSR00 : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.SMSR00 _managed = new Formats.Nintendo.SMSR00();
        protected override AuroraFormat Format => AuroraFormat.SMSR00;
        protected override ICompressionAlgorithm Managed => _managed;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ40 : GpuCodec, IProvidesDecompressedSize
    {
        private readonly Formats.Nintendo.LZ40 _managed = new Formats.Nintendo.LZ40();
        protected override AuroraFormat Format => AuroraFormat.LZ40;
        protected override ICompressionAlgorithm Managed => _managed;
        public bool GbaVramCompatibilityMode { get; set; } = false;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.VramMode = GbaVramCompatibilityMode ? 1 : 0; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ60 : GpuCodec, IProvidesDecompressedSize   // the LZ40 codec under identifier 0x60
    {
        private readonly Formats.Nintendo.LZ60 _managed = new Formats.Nintendo.LZ60();
        protected override AuroraFormat Format => AuroraFormat.LZ60;
        protected override ICompressionAlgorithm Managed => _managed;
        public bool GbaVramCompatibilityMode { get; set; } = false;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.VramMode = GbaVramCompatibilityMode ? 1 : 0; }
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuECD : GpuCodec, IProvidesDecompressedSize   // plain bytes + LZSS(0x400, 0x42, 3, 0x3BE), or stored
    {
        private readonly Formats.Specialized.ECD _managed = new Formats.Specialized.ECD();
        protected override AuroraFormat Format => AuroraFormat.ECD;
        protected override ICompressionAlgorithm Managed => _managed;
        public byte PlainSize { get; set; } = 4;
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s) { o.EcdPlainSize = PlainSize; }
        protected override long PeekBytes => long.MaxValue;   // the compressed size is compared with the stream length
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ00 : GpuCodec, IProvidesDecompressedSize   // LZSS (Lzss0) under the StreamTransformer keystream (device pass)
    {
        private readonly Formats.Sega.LZ00 _managed = new Formats.Sega.LZ00();
        protected override AuroraFormat Format => AuroraFormat.LZ00;
        protected override ICompressionAlgorithm Managed => _managed;
        /// <summary>Key written by <see cref="Compress"/>; null = the Unix time of the call, like the reference.</summary>
        public uint? Key { get; set; }
        protected override void FillOptions(ref AuroraCodecOpts o, CompressionSettings s)
            => o.Lz00Key = Key ?? (uint)DateTimeOffset.UtcNow.ToUnixTimeSeconds();
        protected override long PeekBytes => 64;
        public uint GetDecompressedSize(Stream source) => PeekSize(source);
    }

    public sealed class GpuLZ4 : GpuCodec
    {
        private readonly Formats.Common.LZ4 _managed = new Formats.Common.LZ4();
        protected override AuroraFormat Format => AuroraFormat.LZ4;
        protected override ICompressionAlgorithm Managed => _managed;
    }
    // GpuYaz1, GpuYay0, GpuMIO0, GpuLZ11, GpuLZSS, GpuLZ4Legacy, GpuLZO, GpuSnappy, GpuPRS follow the same pattern
    // (one AuroraFormat value each; Yay0/MIO0/PRS expose FormatByteOrder, LZSS takes LzProperties -> AuroraLzProps).

    /// <summary>The new batch entry point: many independent blobs at once, sharded over all GPUs of the box.</summary>
    public static unsafe class BatchCodec
    {
        public static AuroraStatus[] DecompressBatch(AuroraFormat format, ReadOnlyMemory<byte>[] sources, Memory<byte>[] destinations, out long[] written)
        {
            int n = sources.Length;
            ulong total = 0, dtotal = 0;
            var sOff = new ulong[n]; var sLen = new ulong[n]; var dOff = new ulong[n]; var dCap = new ulong[n];
            for (int i = 0; i < n; i++)
            {
                sOff[i] = total; sLen[i] = (ulong)sources[i].Length; total += (sLen[i] + 15) & ~15UL;
                dOff[i] = dtotal; dCap[i] = (ulong)destinations[i].Length; dtotal += (dCap[i] + 15) & ~15UL;
            }
            IntPtr ps = Native.aurora_pinned_alloc((UIntPtr)(total + 16)), pd = Native.aurora_pinned_alloc((UIntPtr)(dtotal + 16));
            try
            {
                for (int i = 0; i < n; i++) sources[i].Span.CopyTo(new Span<byte>((byte*)ps + sOff[i], (int)sLen[i]));
                var outLen = new ulong[n]; var consumed = new ulong[n]; var status = new int[n];
                AuroraCodecOpts o;
                Native.aurora_codec_opts_init(&o);
                fixed (ulong* a = sOff, b = sLen, c = dOff, d = dCap, e = outLen, f = consumed)
                fixed (int* g = status)
                {
                    int rc = Native.aurora_decode_batch(GpuCodec.Context.Value, (int)format, &o, (UIntPtr)n, (byte*)ps, a, b, (byte*)pd, c, d, e, f, g);
                    if (rc != 0) GpuCodec.ThrowFor((AuroraStatus)rc, 0, 0);
                }
                written = new long[n];
                var result = new AuroraStatus[n];
                for (int i = 0; i < n; i++)
                {
                    written[i] = (long)outLen[i];
                    result[i] = (AuroraStatus)status[i];
                    new Span<byte>((byte*)pd + dOff[i], (int)Math.Min(outLen[i], dCap[i])).CopyTo(destinations[i].Span);
                }
                return result;
            }
            finally { Native.aurora_pinned_free(ps); Native.aurora_pinned_free(pd); }
        }
    }
}

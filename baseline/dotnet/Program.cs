// CPU baseline on the reference's own runtime: decode (or encode) a batch of independent streams with
// Parallel.ForEach, one AuroraLib.Compression codec instance per worker, and print ONE JSON line in bench.py's format.
//
//   dotnet run -c Release -- <format> <packed.bin> <index.bin> [--threads N] [--steps K] [--warmup W] [--encode]
//
// packed.bin : the compressed (decode) or raw (encode) streams back to back
// index.bin  : per stream three little-endian u64: offset, length, decoded size — what `python bench.py --dump-batch DIR`
//              writes for the C2 sample of the reference arm
// The harness mirrors /root/reference/Benchmarks/Benchmarks/TestAllAlgorithms.cs:26-69 (a MemoryStream per stream,
// ICompressionAlgorithm.Decompress(source, destination)); it is auto-skipped by bench.py when no `dotnet` is on PATH.
using System.Diagnostics;
using System.Text.Json;
using AuroraLib.Compression.Algorithms;
using AuroraLib.Compression.Interfaces;

static ICompressionAlgorithm Make(string format) => format.ToUpperInvariant() switch
{
    "LZ10" => new LZ10(), "LZ11" => new LZ11(), "YAZ0" => new Yaz0(), "YAY0" => new Yay0(), "MIO0" => new MIO0(),
    "LZSS" => new LZSS(), "LZ4" => new LZ4(), "LZO" => new LZO(), "SNAPPY" => new Snappy(), "PRS" => new PRS(),
    _ => throw new ArgumentException($"unknown format {format}")
};

string format = args[0];
byte[] packed = File.ReadAllBytes(args[1]);
byte[] index = File.ReadAllBytes(args[2]);
int threads = Environment.ProcessorCount, steps = 3, warmup = 1;
bool encode = false;
for (int i = 3; i < args.Length; i++)
{
    if (args[i] == "--threads") threads = int.Parse(args[++i]);
    else if (args[i] == "--steps") steps = int.Parse(args[++i]);
    else if (args[i] == "--warmup") warmup = int.Parse(args[++i]);
    else if (args[i] == "--encode") encode = true;
}
int n = index.Length / 24;
var off = new long[n]; var len = new long[n]; var size = new long[n];
for (int i = 0; i < n; i++)
{
    off[i] = BitConverter.ToInt64(index, 24 * i);
    len[i] = BitConverter.ToInt64(index, 24 * i + 8);
    size[i] = BitConverter.ToInt64(index, 24 * i + 16);
}
long outBytes = 0;
var po = new ParallelOptions { MaxDegreeOfParallelism = threads };
double best = double.MaxValue, total = 0;
for (int it = 0; it < warmup + steps; it++)
{
    long produced = 0;
    var sw = Stopwatch.StartNew();
    Parallel.ForEach(Enumerable.Range(0, n), po, () => Make(format), (i, _, codec) =>
    {
        using var src = new MemoryStream(packed, (int)off[i], (int)len[i], writable: false);
        using var dst = new MemoryStream((int)Math.Max(size[i], 16));
        if (encode) codec.Compress(new ReadOnlySpan<byte>(packed, (int)off[i], (int)len[i]), dst);
        else codec.Decompress(src, dst);
        Interlocked.Add(ref produced, encode ? len[i] : dst.Length);
        return codec;
    }, _ => { });
    sw.Stop();
    if (it >= warmup) { best = Math.Min(best, sw.Elapsed.TotalSeconds); total += sw.Elapsed.TotalSeconds; }
    outBytes = produced;
}
double gbs = outBytes / (total / steps) / 1e9;
Console.WriteLine(JsonSerializer.Serialize(new
{
    impl = "reference", metric = $"batched {format} {(encode ? "encode" : "decode")} throughput", value = Math.Round(gbs, 3), unit = "GB/s",
    steps, warmup, ms_per_step = Math.Round(1e3 * total / steps, 3), higher_is_better = true,
    cpu_baseline = new { value = Math.Round(gbs, 3), unit = "GB/s", cores = threads, kind = "reference", sample = $"{n} streams, {outBytes} bytes per step, Parallel.ForEach" },
}));

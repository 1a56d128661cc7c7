"""Developer tool (GPU box): find the streams of a C4-like batch that fail, save their raw / compressed bytes for offline analysis."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from auroralib.compression_b200 import BatchCodec, _abi as A
dev = torch.device("cuda", 0); ts = torch.cuda.Stream(device=dev)
codec = BatchCodec(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
fmt = {"lzo": A.FMT_LZO, "lz4": A.FMT_LZ4_BLOCK, "snappy": A.FMT_SNAPPY_BLOCK}[sys.argv[2] if len(sys.argv) > 2 else "lzo"]
raw, r_off, r_len = bench.ragged_corpus(n, 4 << 10, 64 << 10, "TMX", 0xC4000 + fmt, dev, group=4096)
packed, p_off, c_len, tot, ms = bench.gpu_encode(codec, fmt, raw, r_off, r_len, A.make_opts(quality=8), dev, ts)
d_dst = torch.zeros(raw.numel(), dtype=torch.uint8, device=dev); torch.cuda.synchronize()
olen = torch.zeros(n, dtype=torch.int64, device=dev); cons = torch.zeros(n, dtype=torch.int64, device=dev)
st = torch.full((n,), -1, dtype=torch.int32, device=dev)
codec.decode_device(fmt, packed, p_off, c_len, d_dst, r_off, r_len, olen, cons, st, A.make_opts(), device=0, stream=ts.cuda_stream)
ts.synchronize()
bad = (st != 0).nonzero().flatten().tolist()
print("bad", len(bad), bad[:20])
for k, i in enumerate(bad[:4]):
    po, cl, ro, rl = int(p_off[i]), int(c_len[i]), int(r_off[i]), int(r_len[i])
    print(i, "group", i // 4096, "class", "TMX"[(i // 4096) % 3], "r_len", rl, "c_len", cl, "status", int(st[i]), "olen", int(olen[i]), "cons", int(cons[i]), "dst align", ro % 16)
    np.save(f"gpurun_out/dbg_raw_{k}.npy", raw[ro:ro + rl].cpu().numpy())
    np.save(f"gpurun_out/dbg_comp_{k}.npy", packed[po:po + cl].cpu().numpy())
    # the same stream alone, 16-byte aligned destination
    one = torch.zeros(rl + 64, dtype=torch.uint8, device=dev)
    z = torch.zeros(1, dtype=torch.int64, device=dev)
    o1 = torch.zeros(1, dtype=torch.int64, device=dev); c1 = torch.zeros(1, dtype=torch.int64, device=dev); s1 = torch.zeros(1, dtype=torch.int32, device=dev)
    codec.decode_device(fmt, packed, p_off[i:i+1].contiguous(), c_len[i:i+1].contiguous(), one, z, r_len[i:i+1].contiguous(), o1, c1, s1, A.make_opts(), device=0, stream=ts.cuda_stream)
    ts.synchronize()
    print("   alone aligned: status", int(s1[0]), "olen", int(o1[0]), "equal", bool(torch.equal(one[:rl], raw[ro:ro+rl])))
    capbig = torch.tensor([rl + 48], dtype=torch.int64, device=dev)
    codec.decode_device(fmt, packed, p_off[i:i+1].contiguous(), c_len[i:i+1].contiguous(), one, z, capbig, o1, c1, s1, A.make_opts(), device=0, stream=ts.cuda_stream)
    ts.synchronize()
    print("   alone cap+48: status", int(s1[0]), "olen", int(o1[0]), "equal", bool(torch.equal(one[:rl], raw[ro:ro+rl])))

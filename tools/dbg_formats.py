import sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from auroralib.compression_b200 import BatchCodec, _abi as A, corpus
from oracle import oracle as O
codec=BatchCodec(1)
raw,_=corpus.generate_mix(600,65536,device='cuda')
raws=[raw[i].cpu().numpy().tobytes() for i in range(600)]
for fmt in (A.FMT_MIO0,A.FMT_YAY0,A.FMT_LZ11,A.FMT_LZ4_BLOCK,A.FMT_SNAPPY_BLOCK,A.FMT_LZO):
    comps,st=O.encode_batch(fmt,raws,A.make_opts(quality=8))
    outs,ol,cons,gst=codec.decode_batch(fmt,comps,[65536]*600)
    ref,rl,rc,rst=O.decode_batch(fmt,comps,[65536]*600)
    bad=[i for i in range(600) if gst[i]!=rst[i] or outs[i]!=ref[i] or ol[i]!=rl[i] or cons[i]!=rc[i]]
    print(A.FORMAT_NAMES[fmt],'bad',len(bad),'gpu nonzero',int((gst!=0).sum()),'ref nonzero',int((rst!=0).sum()))
    for i in bad[:3]:
        fd=next((k for k in range(min(len(outs[i]),len(ref[i]))) if outs[i][k]!=ref[i][k]),None)
        print('  #',i,'st',gst[i],rst[i],'ol',ol[i],rl[i],'cons',cons[i],rc[i],'len',len(comps[i]),'firstdiff',fd, 'hdr', comps[i][:16].hex())

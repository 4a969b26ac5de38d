import sys, zlib, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import htslib_b200 as H
from tools import synth
corpus = synth.bam_bgzf_corpus(0.5e9)
comp, clen, ulen = corpus["comp"], corpus["clen"], corpus["ulen"]
nb=len(clen); dev=torch.device("cuda:0")
in_off = np.concatenate([[0], np.cumsum(clen.astype(np.int64))[:-1]]).astype(np.uint64)
out_off = np.concatenate([[0], np.cumsum(ulen.astype(np.int64))[:-1]]).astype(np.uint64)
t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
d_in = torch.zeros(comp.size+64, dtype=torch.uint8, device=dev); d_in[:comp.size].copy_(torch.from_numpy(comp.copy()))
d_out = torch.zeros(int(ulen.sum())+64, dtype=torch.uint8, device=dev)
a,b,c,d = t(in_off), t(clen), t(out_off), t(ulen)
d_len = torch.zeros(nb, dtype=torch.int32, device=dev); d_st = torch.zeros(nb, dtype=torch.int32, device=dev)
ctx = H.Context(0)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ctx.bgzf_inflate_dev(d_in,a,b,d_out,c,d,d_len,d_st,s.cuda_stream); torch.cuda.synchronize()
st=d_st.cpu().numpy(); out=d_out.cpu().numpy()
bad=np.nonzero(st)[0]
print("failing", len(bad), bad[:10], st[bad][:10])
for i in bad[:4]:
    blk=comp[int(in_off[i]):int(in_off[i])+int(clen[i])].tobytes()
    want=zlib.decompress(blk[18:-8], -15)
    got=out[int(out_off[i]):int(out_off[i])+int(ulen[i])].tobytes()
    k=0
    while k<len(want) and got[k]==want[k]: k+=1
    print(i, 'clen',len(blk),'ulen',len(want),'first diff',k, 'align in',int(in_off[i])%16,'out',int(out_off[i])%16)
    np.save('gpurun_out/fail_block_%d.npy'%i, np.frombuffer(blk,dtype=np.uint8))

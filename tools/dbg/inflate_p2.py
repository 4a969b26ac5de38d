import sys, zlib, numpy as np, torch, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import htslib_b200 as H
blk=np.load(sys.argv[1])
L=H.lib()
ctx=H.Context(0)
dev=torch.device("cuda:0")
n=1
IN=int(sys.argv[3]) if len(sys.argv)>3 else 5
in_off=np.array([IN],dtype=np.uint64); in_len=np.array([blk.size],dtype=np.uint32)
d_in=torch.zeros(blk.size+64,dtype=torch.uint8,device=dev); d_in[IN:IN+blk.size].copy_(torch.from_numpy(blk))
out_off=np.array([int(sys.argv[2]) if len(sys.argv)>2 else 0],dtype=np.uint64); cap=np.array([65536],dtype=np.uint32)
t=lambda a: torch.from_numpy(a.view(np.int64) if a.dtype==np.uint64 else a.view(np.int32)).to(dev)
d_out=torch.zeros(70000,dtype=torch.uint8,device=dev); d_len=torch.zeros(1,dtype=torch.int32,device=dev); d_st=torch.zeros(1,dtype=torch.int32,device=dev)
L.hgpu_debug_p2(0, None)
s=torch.cuda.Stream()
with torch.cuda.stream(s):
    ctx.bgzf_inflate_dev(d_in,t(in_off),t(in_len),d_out,t(out_off),t(cap),d_len,d_st,s.cuda_stream); torch.cuda.synchronize()
buf=(C.c_uint*(6*256))()
print('status',int(d_st[0]),'len',int(d_len[0]), 'dbg rc', L.hgpu_debug_p2(0, buf))
a=np.array(buf[:]).reshape(256,6)
np.save('gpurun_out/p2_dbg_%s.npy' % sys.argv[4],a)
for r in range(256):
    if a[r,5] or a[r,0]: print(r, a[r].tolist())

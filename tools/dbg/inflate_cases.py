import sys, random, zlib
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import htslib_b200 as H
from _libs import BGZF_EOF, bgzf_block
from test_gpu_bgzf import gpu_blocks
ctx=H.Context(0)
rng = random.Random(3)
p = bytes(rng.choice(b"abcdefgh ") for _ in range(30000))
blocks=[]
c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED); blocks.append(('fixed',bgzf_block(p, raw_deflate=c.compress(p) + c.flush())))
c = zlib.compressobj(6, zlib.DEFLATED, -15, 8); raw=b""
for i in range(0, len(p), 1000): raw += c.compress(p[i:i + 1000]) + c.flush(zlib.Z_FULL_FLUSH)
raw += c.flush(); blocks.append(('fullflush',bgzf_block(p, raw_deflate=raw)))
c = zlib.compressobj(9, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY); blocks.append(('huffonly',bgzf_block(p, raw_deflate=c.compress(p) + c.flush())))
c = zlib.compressobj(6, zlib.DEFLATED, -15, 1); blocks.append(('memlevel1',bgzf_block(p, raw_deflate=c.compress(p) + c.flush())))
for name,b in blocks:
    (st,data),=gpu_blocks(ctx,[b])
    k=0
    while k<min(len(data),len(p)) and data[k]==p[k]: k+=1
    print(name,'clen',len(b),'status',st,'len',len(data),'first diff',k)

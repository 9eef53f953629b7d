import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennerf_b200 import _lib, ops, synthetic as S
DEV='cuda'
n=int(sys.argv[1]) if len(sys.argv)>1 else 148*128*2+77
g=S.gen(43)
w,hw,hb=S.decoder_weights(g,32,15,512,5,64,32,alpha=0.8)
xyz=S.query_points(n,(96,96,48),0.04,g)[0]
feat=torch.randn(n,32,generator=g)
dw2=ops.DecoderWeights(w,hw,hb,n_blocks=5,d_geo=32,device=DEV)
b,tb=ops.decode(dw2,xyz.to(DEV),feat.to(DEV),'fp16')
old=_lib.set_option('GNB_TC_PAIR',1)
dw1=ops.DecoderWeights(w,hw,hb,n_blocks=5,d_geo=32,device=DEV)
a,ta=ops.decode(dw1,xyz.to(DEV),feat.to(DEV),'fp16')
torch.cuda.synchronize()
err=((a-b).abs().max(dim=1).values/b.abs().max()).cpu()
tiles=(n+127)//128
bad=[]
for t in range(tiles):
    e=err[t*128:(t+1)*128].max().item()
    if e>2e-3: bad.append((t,round(e,3)))
print('tiles',tiles,'bad tiles',len(bad)); print(bad[:40])
# within a bad tile, which rows
if bad:
    t=bad[0][0]; print('rows of tile',t,[round(x,3) for x in err[t*128:(t+1)*128].tolist()][:128])
a2,ta2=ops.decode(dw1,xyz.to(DEV),feat.to(DEV),'fp16')
print('deterministic', torch.equal(a,a2))

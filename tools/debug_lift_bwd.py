import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gennerf_b200 import synthetic as S, autograd as ag
from oracle import gennerf_oracle as O
VS=0.04; ORIGIN=torch.tensor([0,0,0]).view(1,3); DEV='cuda'
wl=S.WORKLOADS['cfg5']; g=S.gen(1005)
P=S.projections(wl['T'],wl['H'],wl['W'],wl['voxel_dim'],VS,g).unsqueeze(0)
feats=S.frame_features(wl['T'],32,wl['H'],wl['W'],g)
G=torch.randn(1,32,*wl['voxel_dim'],generator=g)
fo=[f.clone().requires_grad_(True) for f in feats]
vol_o,_,_=O.encode_volume(wl['voxel_dim'],VS,ORIGIN,P,fo)
(vol_o*G).sum().backward()
for layout in ('nchw','nhwc'):
    fd=[f.to(DEV) for f in feats]
    if layout=='nhwc': fd=[f.contiguous(memory_format=torch.channels_last) for f in fd]
    fd=[f.requires_grad_(True) for f in fd]
    vol,cnt,valid=ag.backproject_frames(wl['voxel_dim'],VS,ORIGIN,P,fd)
    (vol*G.to(DEV)).sum().backward()
    for t in range(wl['T']):
        a=fd[t].grad.cpu(); b=fo[t].grad
        d=(a-b).abs()
        idx=d.flatten().argmax().item()
        print(layout,t,'max abs err',d.max().item(),'max|ref|',b.abs().max().item(),'nnz ref',(b!=0).sum().item(),'nnz gpu',(a!=0).sum().item(),'at',idx, a.flatten()[idx].item(), b.flatten()[idx].item())

"""Cycle trace of the transposed-pair decoder (cluster 0): stage issue times of the MMA warp and round times of the epilogue."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S
from gennerf_b200._lib import lib, set_option
set_option('GNB_TC_PAIR', 1)
dev = torch.device("cuda", 0)
g = S.gen(1)
w, hw, hb = S.decoder_weights(g, 32, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
n = 148 * 64 * 4
xyz = S.query_points(n, (96, 96, 48), 0.04, g)[0].to(dev)
feat = torch.randn(n, 32, device=dev)
ops.decode(dw, xyz, feat, "fp16")
buf = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
fn = lib().gnb_debug_set_trace
fn.argtypes = [C.c_void_p]
fn(buf.data_ptr())
ops.decode(dw, xyz, feat, "fp16")
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(4, 4096)
it = 1
base = t[3, it * 128].item()
names = []
J, nb = 2, 5
for j in range(J): names += [f"in{j}", f"z0_{j}"]
for i in range(nb):
    for j in range(J): names += [f"fc0_{i}({j})p{p}" for p in range(4)]
    for j in range(J):
        if i < nb - 1: names.append(f"z{i+1}_{j}")
        names += [f"fc1_{i}({j})p{p}" for p in range(4)]
names += [f"out p{p}" for p in range(4)]
print("tile", it, "stages:", len(names))
prev = base
for s, nm in enumerate(names):
    reach, bw, iss = t[3, it*128+s].item(), t[0, it*128+s].item(), t[2, it*128+s].item()
    print(f"{s:3d} {nm:12s} reached +{reach-base:7d}  (act wait {bw-reach:6d}, weight wait {iss-bw:6d})  issue +{iss-base:7d}  d={iss-prev}")
    prev = iss
for wname, off in (("keep warp 4", 0), ("remote warp 8", 2048)):
    print(wname)
    for r in range(2 * nb + 1):
        for j in range(J):
            v = [t[1, off + it*128 + r*8 + j*4 + k].item() - base for k in range(4)]
            print(f"   round {r:2d} M-tile {j}: acc seen +{v[0]:7d}  converted +{v[1]-v[0]:5d}  kh0 wait {v[2]-v[1]:5d}  stored+arrived +{v[3]-v[2]:5d}")

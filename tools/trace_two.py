"""Step-level cycle trace of the cta_group::2 query-major decoder (GNB_TC_TWO_CTA=1, tracing library): when the leader's MMA
warp reached each step, saw the own chunk / the pushed chunk / the weight stage, and when the exchange warp pushed."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S
from gennerf_b200._lib import lib
dev = torch.device("cuda", 0)
g = S.gen(1)
w, hw, hb = S.decoder_weights(g, 32, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
n = 148 * 128 * 3
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
tile = 1
st = t[3, tile * 96 * 4:(tile + 1) * 96 * 4].view(96, 4)
base = st[0, 0].item()
print("MMA warp (leader), tile 1: step  reached  own-chunk  pushed-chunk  weights   (deltas)")
for i in range(96):
    if st[i, 0] == 0: break
    a, b, c, d = [x.item() - base for x in st[i]]
    print(f"  {i:2d} +{a:7d}  own +{b-a:5d}  pushed +{c-b:5d}  weights +{d-c:5d}")
ex = t[2, 1024:].view(-1, 3)
print("exchange warp: push  a_ready seen   rfree wait   issue")
for i in range(44, 100):
    if ex[i, 0] == 0: break
    a, b, c = [x.item() - base for x in ex[i]]
    print(f"  {i:3d} +{a:7d}  rfree +{b-a:5d}  issued +{c-b:5d}")
e = t[1, 64:128].tolist()
print("epilogue warp 4, tile 1 (acc_ready seen, chunks written):")
k = 2
for r in range(11):
    print(f"  round {r:2d}: +{e[k]-base:7d} .. +{e[k+1]-base:7d}")
    k += 2

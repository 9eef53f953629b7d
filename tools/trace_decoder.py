"""Per-phase cycle trace of the tcgen05 decoder (cluster 0) -- profiling aid."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S  # noqa: E402
from gennerf_b200._lib import lib  # noqa: E402

dev = torch.device("cuda", 0)
Hd = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = S.gen(1)
w, hw, hb = S.decoder_weights(g, 32, 15, Hd, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
n = 148 * 128 * 3
xyz = S.query_points(n, (96, 96, 48), 0.04, g)[0].to(dev)
feat = torch.randn(n, 32, device=dev)
ops.decode(dw, xyz, feat, "fp16")
buf = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
L = lib()._handle
raw = C.CDLL(None)
fn = lib().gnb_debug_set_trace
fn.argtypes = [C.c_void_p]
fn(buf.data_ptr())
ops.decode(dw, xyz, feat, "fp16")
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(4, 4096)
nb = 5
for tile in range(2):
    m = t[0, tile * 64: tile * 64 + 64].tolist()
    e = t[1, tile * 64: tile * 64 + 64].tolist()
    t0 = e[0]
    print(f"--- tile {tile}: epilogue warp: prologue {e[1]-e[0]} cyc")
    names = ["lin_in"] + sum([[f"lin_z{i}", f"fc0_{i}", f"fc1_{i}"] for i in range(nb)], []) + ["lin_out"]
    print("MMA thread: in_ready at +%d" % (m[0] - t0))
    for i, nm in enumerate(names):
        wa, ww = t[2, (tile * 32 + i) * 2].item(), t[2, (tile * 32 + i) * 2 + 1].item()
        print(f"   {nm:8s} issue start +{m[1+2*i]-t0:7d}  end +{m[2+2*i]-t0:7d}  (dur {m[2+2*i]-m[1+2*i]:6d}; waiting A {wa:6d}, waiting W {ww:6d})")
    k = 2
    for r in range(2 * nb + 1):
        print(f"   E round {r:2d}: acc_ready seen +{e[k]-t0:7d}  chunks written +{e[k+1]-t0:7d}  (dur {e[k+1]-e[k]})")
        k += 2
    print(f"   final: acc_ready +{e[k]-t0:7d} done +{e[k+1]-t0:7d}")


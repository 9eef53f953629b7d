"""One pass of the bench's query phase on a smaller range (TQ_Q queries, default 1 Mi; config-4 volume + planes) for ncu:
launches bin_count / bin_reduce / bin_scan / bin_scatter / sample_binned_kernel / decoder_tc_kernel once each after a warm-up."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S
dev = "cuda"
Q = int(os.environ.get("TQ_Q", str(1 << 20)))
vd, VS = (256, 256, 96), 0.04
g = S.gen(5)
w, hw, hb = S.decoder_weights(g, 64, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
xyz = S.query_points(Q, vd, VS, g).to(dev)
vol = torch.randn(1, *vd, 32, device=dev).permute(0, 4, 1, 2, 3)
pl = {k: torch.randn(1, 256, 256, 32, device=dev).permute(0, 3, 1, 2) for k in ops.PLANES}
kw = dict(volume=vol, planes=pl, voxel_size=VS, origin=torch.zeros(1, 3), padding=0.1)
for _ in range(2):
    out, tsdf, _ = ops.query_image(dw, xyz, binned=True, **kw) if False else ops.query_image(dw, xyz, **kw)
torch.cuda.synchronize()
print("ok", float(tsdf.abs().mean()))

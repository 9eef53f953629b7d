"""Small driver for ncu: runs the lift, the sampler and (if built) the bf16 decoder on BASELINE
config-2 shapes a few times.  python tools/prof_kernels.py [lift|sample|decode|all] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
dev = torch.device("cuda", 0)
VS, C = 0.04, 32
wl = S.WORKLOADS[cfg]
g = S.gen(1002)
origin = torch.tensor([0, 0, 0]).view(1, 3)
T = wl["T"]
P = S.projections(T, wl["H"], wl["W"], wl["voxel_dim"], VS, g).unsqueeze(0)
feats = [torch.randn(1, C, wl["H"], wl["W"], device=dev) for _ in range(T)]
feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
Q = int(os.environ.get("PROF_Q", min(wl["Q"], 1 << 21)))
xyz = S.query_points(Q, wl["voxel_dim"], VS, g).to(dev)
w, hw, hb = S.decoder_weights(g, C, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, use_code=True, num_freqs=2, freq_factor=0.5, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
vol, cnt, valid = ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats_cl)
for _ in range(reps):
    if what in ("lift", "all"):
        flush.fill_(1)
        ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats)        # NCHW: transpose + lift
        flush.fill_(1)
        ops.backproject_frames(wl["voxel_dim"], VS, origin, P, feats_cl)     # NHWC: lift only
    if what in ("sample", "all"):
        flush.fill_(1)
        feat = ops.sample_features(xyz, volume=vol, voxel_size=VS, origin=origin)
    if what in ("decode", "all"):
        try:
            out, tsdf, _ = ops.query_fused(dw, xyz, volume=vol, voxel_size=VS, origin=origin, want_feat=False)
        except RuntimeError as e:
            print("no bf16 decoder:", e)
            what = "lift+sample"
torch.cuda.synchronize()
print("done", what, "valid voxel-frames", int(cnt.sum()), "of", cnt.numel() * T)

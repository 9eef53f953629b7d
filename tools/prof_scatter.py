"""Small driver for ncu: the config-3 triplane scatter (3 x 256^2 planes, C_p = 32) with all 614 400 pixels of 8 frames
240x320, unit and metric point domains.  python tools/prof_scatter.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S  # noqa: E402

g = S.gen(1003)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for domain in ("unit", "metric"):
    p = S.plane_points(614400, g, domain, voxel_dim=(96, 96, 48)).cuda()
    c = torch.randn(1, 614400, 32, generator=g).cuda()
    flush.fill_(1)
    planes, cnt = ops.scatter_mean_planes(p, c, 256, 0.1, "atomic")
torch.cuda.synchronize()
print("done", int(cnt.sum()))

"""Host time per call of the torch.library custom ops (gennerf_b200.autograd) against the plain ops.* wrappers they call:
how much of a launch-rate-bound training step is dispatcher / autograd-wrapper overhead."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import autograd as ag, ops, synthetic as S  # noqa: E402

dev = "cuda"
g = S.gen(1)
B, N, Hd, R = 1, 4096, 32, 128
p = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
c = torch.randn(B, N, Hd, generator=g).to(dev).requires_grad_(True)
vd = (64, 64, 32)
T, C, H, W = 8, 32, 120, 160
P = S.projections(T, H, W, vd, 0.04, g).unsqueeze(0)
feats = [torch.randn(1, C, H, W, generator=g).to(dev).requires_grad_(True) for _ in range(T)]
origin = torch.zeros(1, 3)
xyz = S.query_points(20000, vd, 0.04, g).to(dev)


def host(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t) / n * 1e6
    torch.cuda.synchronize()
    return dt


with torch.no_grad():
    vol = ops.backproject_frames(vd, 0.04, origin, P, [f.detach() for f in feats])[0]
volg = vol.detach().requires_grad_(True)
rows = [
    ("pool_local       custom op (grad on)", lambda: ag.pool_local(p, c, R, 0.1, "max")),
    ("pool_local       ops.* direct", lambda: ops.pool_local_fwd_keep(p, c.detach(), R, 0.1, "max")),
    ("scatter_mean     custom op (grad on)", lambda: ag.scatter_mean_planes(p, c, R, 0.1, "atomic")),
    ("scatter_mean     ops.* direct", lambda: ops.scatter_mean_planes(p, c.detach(), R, 0.1, "atomic")),
    ("sample_features  custom op (grad on)", lambda: ag.sample_features(xyz, volume=volg, voxel_size=0.04, origin=origin)),
    ("sample_features  ops.* direct", lambda: ops.sample_features(xyz, volume=vol, voxel_size=0.04, origin=origin)),
    ("backproject      custom op (grad on)", lambda: ag.backproject_frames(vd, 0.04, origin, P, feats)),
    ("backproject      ops.* direct", lambda: ops.backproject_frames(vd, 0.04, origin, P, [f.detach() for f in feats])),
]
for name, fn in rows:
    print(f"{name:40s} {host(fn):8.1f} us of host time per call")
# backward of the custom ops (autograd engine thread included: wall time of .backward() issue)
out = ag.pool_local(p, c, R, 0.1, "max")
go = torch.randn_like(out)
print(f"{'pool_local backward (custom op)':40s} {host(lambda: torch.autograd.grad(out, c, go, retain_graph=True)):8.1f} us")
f = ag.sample_features(xyz, volume=volg, voxel_size=0.04, origin=origin)
gf = torch.randn_like(f)
print(f"{'sample_features backward (custom op)':40s} {host(lambda: torch.autograd.grad(f, volg, gf, retain_graph=True)):8.1f} us")

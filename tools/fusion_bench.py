"""Timing aid: TSDF fusion (gnb_tsdf_fusion_integrate) on the config-4 grid, against its HBM roofline.
Algorithmic bytes = V*(8 read + 8 written) for the tsdf / weight volumes + T*H*W*4 of depth."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import synthetic as S  # noqa: E402
from gennerf_b200.dropin import TSDFFusion  # noqa: E402

dev = "cuda"
for vd, T, H, W in (((256, 256, 96), 32, 480, 640), ((96, 96, 48), 8, 240, 320)):
    g = S.gen(3)
    P = S.projections(T, H, W, vd, 0.04, g)
    depths = S.surface_depth_maps(T, H, W, g, mean=1.5).to(dev)
    f = TSDFFusion(vd, 0.04, (0, 0, 0), device=dev, color=False, label=False)
    f.integrate_frames(P, depths)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for _ in range(7):
        f.reset()
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f.integrate_frames(P, depths); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    m = sorted(ms)[len(ms) // 2]
    V = vd[0] * vd[1] * vd[2]
    nbytes = V * 16 + T * H * W * 4
    seen = int((f.weight_vol > 0).sum())
    print(f"grid {vd} x {T} frames {H}x{W}: {m * 1e3:.1f} us, {V * T / m / 1e6:.1f} G voxel-frames/s, "
          f"{nbytes / m / 1e6:.0f} GB/s algorithmic ({nbytes / m / 1e6 / 6550.7:.2f} of HBM peak); voxels seen {seen}")

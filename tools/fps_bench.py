"""Timing aid: farthest point sampling (gnb_farthest_point_sample), 512 samples from B clouds of N points."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, _lib  # noqa: E402

dev = "cuda"
# variants: auto = the library's choice; cluster = thread-block clusters only (GNB_FPS_GRID=-1); grid = the co-operative grid kernel
VARIANTS = [("auto", 0)] + ([("cluster", -1)] if "--variants" in sys.argv else [])
for N in (76800, 307200):
  for vname, vval in VARIANTS:
    for B in (1, 2, 4, 8, 16, 32):
        _lib.set_option("GNB_FPS_GRID", vval)
        xyz = torch.rand(B, N, 3, device=dev) * 4
        start = torch.zeros(B, dtype=torch.long, device=dev)
        ops.farthest_point_sample(xyz, 512, start)
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.farthest_point_sample(xyz, 512, start); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        m = sorted(ms)[1]
        print(f"[{vname}] N={N} B={B}: {m:.3f} ms  ({m / 512 * 1e3:.2f} us per iteration, {B * N * 512 / m / 1e6:.1f} G point-updates/s)", flush=True)

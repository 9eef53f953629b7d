"""Tuning aid: fused query time per Mi queries for volume-only / planes-only / volume+planes prologues (config-2 grid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S
dev = "cuda"
g = S.gen(1)
vd = (96, 96, 48)
n = 1 << 20
xyz = S.query_points(n, vd, 0.04, g).to(dev)
vol = torch.randn(1, *vd, 32, generator=g).to(dev).permute(0, 4, 1, 2, 3)
planes = {k: torch.randn(1, 32, 256, 256, generator=g).to(dev).contiguous(memory_format=torch.channels_last) for k in ops.PLANES}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, v, p in (("volume", vol, None), ("planes", None, planes), ("volume+planes", vol, planes)):
    d_feat = (32 if v is not None else 0) + (32 if p is not None else 0)
    w, hw, hb = S.decoder_weights(S.gen(2), d_feat, 15, 512, 5, 64, 32)
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
    for presort in (False, True):
        if presort and v is None:
            continue
        run = lambda: ops.query_fused(dw, xyz, volume=v, planes=p, voxel_size=0.04, origin=torch.zeros(1, 3), want_feat=False, presort=presort)
        run(); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        print(f"{name:14s} presort={presort!s:5s}: {sorted(ms)[2]:.3f} ms per Mi queries", flush=True)
feat = torch.randn(n, 64, device=dev)
w, hw, hb = S.decoder_weights(S.gen(2), 64, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
run = lambda: ops.decode(dw, xyz[0], feat, "fp16")
run(); torch.cuda.synchronize()
ms = []
for _ in range(5):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); b.synchronize()
    ms.append(a.elapsed_time(b))
print(f"decode only, d_feat 64: {sorted(ms)[2]:.3f} ms")

"""BASELINE config 3: scatter_mean onto 3x256^2 triplanes (C_p = 32) from N = 4 096 (reference-faithful: 8 frames x 512
sparse points) and N = 614 400 points (every pixel of 8 frames 240x320), 'unit' and 'metric' point domains (SURVEY trap T6:
metric points clamp into the border cells), atomic and deterministic modes; then the bilinear plane query + decode of
1 Mi points.  Algorithmic bytes as SURVEY 8d: N*(12 + 4*C_p) read + 3*R^2*(4*C_p + 4) written."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S  # noqa: E402

dev = "cuda"
R, Cp = 256, 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = 6550.7


def timed(fn, reps=10):
    fn()
    gr = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2]


g = S.gen(1003)
for N in (4096, 614400):
    for domain in ("unit", "metric"):
        p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48)).to(dev)
        c = torch.randn(1, N, Cp, generator=g).to(dev)
        byt = N * (12 + 4 * Cp) + 3 * R * R * (4 * Cp + 4)
        for mode in ("atomic", "deterministic"):
            m = timed(lambda: ops.scatter_mean_planes(p, c, R, 0.1, mode))
            print(f"scatter_mean N={N} {domain:6s} {mode:13s}: {m*1e3:8.1f} us  alg {byt/1e6:.1f} MB -> {byt/m/1e6:.0f} GB/s ({byt/m/1e6/PEAK:.2f} of HBM)", flush=True)
        for st in ("max", "mean"):
            m = timed(lambda: ops.pool_local(p, c, R, 0.1, st))
            print(f"pool_local   N={N} {domain:6s} {st:13s}: {m*1e3:8.1f} us", flush=True)

# plane query + decode of 1 Mi points (planes only, the experiment variant's shape C_lat = 32)
planes, _ = ops.scatter_mean_planes(S.plane_points(614400, g, "unit").to(dev), torch.randn(1, 614400, Cp, generator=g).to(dev), R, 0.1, "atomic")
pl = {k: planes[i] for i, k in enumerate(("xz", "xy", "yz"))}
Q = 1 << 20
xyz = S.plane_points(Q, g, "unit").to(dev)
m = timed(lambda: ops.sample_features(xyz, planes=pl, padding=0.1))
byt = Q * (12 + 4 * Cp) + min(3 * R * R * Cp * 4, 12 * Q * Cp * 4)
print(f"plane query  Q={Q}: {m*1e3:8.1f} us  alg {byt/1e6:.1f} MB -> {byt/m/1e6:.0f} GB/s ({byt/m/1e6/PEAK:.2f} of HBM)", flush=True)
w, hw, hb = S.decoder_weights(g, Cp, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, use_code=True, num_freqs=2, freq_factor=0.5, device=torch.device(dev))
m = timed(lambda: ops.query_fused(dw, xyz, planes=pl, padding=0.1, want_feat=False), reps=5)
print(f"plane query + decoder (fused tcgen05) Q={Q}: {m:8.3f} ms -> {Q/m/1e3:.1f} M points/s", flush=True)

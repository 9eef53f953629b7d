"""Tuning aid: sampler kernel time (1 Mi queries, BASELINE config-2 volume) + parity with the generic kernel."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gennerf_b200 import ops, synthetic as S
    dev = "cuda"
    g = S.gen(3)
    for C, Cp, R in ((32, 0, 0), (32, 32, 256), (128, 0, 0)):
        dims = (96, 96, 48)
        vol = torch.randn(1, *dims, C, device=dev).permute(0, 4, 1, 2, 3)
        planes = {k: torch.randn(1, R, R, Cp, device=dev).permute(0, 3, 1, 2) for k in ("xz", "xy", "yz")} if Cp else None
        Q = 1 << 20
        xyz = S.query_points(Q, dims, 0.04, g).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        out = ops.sample_features(xyz, volume=vol, planes=planes, voxel_size=0.04)
        gr = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr, stream=st):
                out2 = ops.sample_features(xyz, volume=vol, planes=planes, voxel_size=0.04)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        m = sorted(ms)[len(ms) // 2]
        byt = Q * (12 + 4 * (C + Cp)) + min(vol.numel() * 4, 8 * Q * C * 4) + (min(3 * R * R * Cp * 4, 12 * Q * Cp * 4) if Cp else 0)
        print(f"  C={C} Cp={Cp}: {m*1e3:.1f} us  alg {byt/1e6:.0f} MB -> {byt/m/1e6:.0f} GB/s ({byt/m/1e6/6550.7:.2f}) checksum {out.double().sum().item():.6f}", flush=True)
else:
    for mode in ("staged", "generic"):
        print(mode, flush=True)
        env = dict(os.environ)
        if mode == "generic":
            env["GNB_SAMPLE_GENERIC"] = "1"
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=200)

"""Tuning aid: sampler time (1 Mi queries, BASELINE config-2 volume; 16 Mi on the config-4 grid) for the brick-binned,
staged and generic kernels, with the algorithmic-byte HBM roofline of SURVEY 8d."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gennerf_b200 import ops, synthetic as S
    dev = "cuda"
    g = S.gen(3)
    mode = sys.argv[2]
    binned = {"binned": True, "staged": False, "generic": False}[mode]
    cases = [((96, 96, 48), 32, 0, 0, 1 << 20), ((96, 96, 48), 32, 32, 256, 1 << 20), ((96, 96, 48), 128, 0, 0, 1 << 20)]
    if mode != "generic":
        cases.append(((256, 256, 96), 32, 0, 0, 1 << 24))
        cases.append(((256, 256, 96), 32, 0, 0, 1 << 21))
    if os.environ.get("SAMPLE_WIDE"):        # the reference's default yaml: 512 spatial channels
        cases = [((96, 96, 48), 512, 0, 0, 1 << 20), ((96, 96, 48), 512, 32, 256, 1 << 20), ((96, 96, 48), 256, 0, 0, 1 << 20)]
    cases = cases[:int(os.environ.get('SAMPLE_CASES', len(cases)))]
    for dims, C, Cp, R, Q in cases:
        vol = torch.randn(1, *dims, C, device=dev).permute(0, 4, 1, 2, 3)
        planes = {k: torch.randn(1, R, R, Cp, device=dev).permute(0, 3, 1, 2) for k in ("xz", "xy", "yz")} if Cp else None
        xyz = S.query_points(Q, dims, 0.04, g).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        out = ops.sample_features(xyz, volume=vol, planes=planes, voxel_size=0.04, binned=binned)
        gr = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr, stream=st):
                out2 = ops.sample_features(xyz, volume=vol, planes=planes, voxel_size=0.04, binned=binned)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        m = sorted(ms)[len(ms) // 2]
        byt = Q * (12 + 4 * (C + Cp)) + min(vol.numel() * 4, 8 * Q * C * 4) + (min(3 * R * R * Cp * 4, 12 * Q * Cp * 4) if Cp else 0)
        print(f"  grid {dims} C={C} Cp={Cp} Q={Q}: {m*1e3:.1f} us  alg {byt/1e6:.0f} MB -> {byt/m/1e6:.0f} GB/s ({byt/m/1e6/6550.7:.2f})"
              f" checksum {out.double().sum().item():.6f}", flush=True)
        del vol, xyz, out, out2, gr
else:
    for mode in (sys.argv[1:] or ["binned", "staged", "generic"]):
        print(mode, flush=True)
        env = dict(os.environ)
        if mode == "generic":
            env["GNB_SAMPLE_GENERIC"] = "1"
        subprocess.run([sys.executable, __file__, "child", mode], env=env, timeout=300)

"""Tuning aid: lift kernel time for voxels-per-warp variants on BASELINE shapes."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gennerf_b200 import ops, synthetic as S
    dev = "cuda"
    for cfg, C in (("cfg2", 32), ("cfg2", 128), ("cfg4", 32)):
        wl = S.WORKLOADS[cfg]
        g = S.gen(1002)
        T = wl["T"]
        origin = torch.tensor([0, 0, 0]).view(1, 3)
        P = S.projections(T, wl["H"], wl["W"], wl["voxel_dim"], 0.04, g).unsqueeze(0)
        feats = [torch.randn(1, wl["H"], wl["W"], C, device=dev).permute(0, 3, 1, 2) for _ in range(T)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        vol, cnt, valid = ops.backproject_frames(wl["voxel_dim"], 0.04, origin, P, feats)
        gr = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr, stream=st):
                ops.backproject_frames(wl["voxel_dim"], 0.04, origin, P, feats)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        V = cnt.numel()
        nv = int(cnt.sum())
        byt = min(T * C * wl["H"] * wl["W"] * 4, nv * C * 4) + V * C * 4 + V * 5
        m = sorted(ms)[len(ms) // 2]
        print(f"  {cfg} C={C}: {m*1e3:.1f} us  valid vf {nv/(V*T):.3f}  alg {byt/1e6:.0f} MB -> {byt/m/1e6:.0f} GB/s ({byt/m/1e6/6550.7:.2f})", flush=True)
else:
    for n in (4, 8, 16, 32):
        print("NVW", n, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, GNB_LIFT_NVW=str(n)), timeout=200)

"""Tuning aid: tcgen05 decoder time at 1 Mi rows (config-2 shape) for the kernel variants selected by environment
variables, each in a child process.  Prints ms, TFLOP/s and the max abs difference to the default variant."""
import os
import subprocess
import sys

VARIANTS = {
    "default": {},
    "two_cta": {"GNB_TC_TWO_CTA": "1"},
    "one_cta": {"GNB_TC_TWO_CTA": "0"},
    "no_early": {"GNB_TC_NO_EARLY": "1"},
    "nocopy": {"GNB_DEBUG_NO_WCOPY": "1"},                         # timing experiment: weight stages signalled, not copied (wrong results)
    "pair_nocopy": {"GNB_TC_PAIR": "1", "GNB_DEBUG_NO_WCOPY": "1"},
    "pair": {"GNB_TC_PAIR": "1"},           # the transposed-pair kernel (decoder_tp_kernel) instead of the query-major kernel
}

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gennerf_b200 import ops, synthetic as S
    dev = "cuda"
    Hd = int(os.environ.get("TD_HIDDEN", "512"))
    n = int(os.environ.get("TD_ROWS", str(1 << 20)))
    DF = int(os.environ.get("TD_FEAT", "32"))        # latent width (544 = the reference's default yaml: streamed lin_in)
    g = S.gen(1)
    w, hw, hb = S.decoder_weights(g, DF, 15, Hd, 5, 64, 32)
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
    xyz = S.query_points(n, (96, 96, 48), 0.04, g)[0].to(dev)
    feat = torch.randn(n, DF, generator=g).to(dev)
    fused = os.environ.get("TD_FUSED") == "1"       # sample the features from a volume inside the kernel (the bench's path)
    if fused:
        vd = (96, 96, 48)
        vol = torch.randn(1, *vd, 32, generator=g).to(dev).permute(0, 4, 1, 2, 3)
        xyz = xyz.unsqueeze(0)
        run = lambda: ops.query_fused(dw, xyz, volume=vol, voxel_size=0.04, origin=torch.zeros(1, 3), want_feat=False)[:2]
    else:
        run = lambda: ops.decode(dw, xyz, feat, "fp16")
    out, tsdf = run()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    m = sorted(ms)[len(ms) // 2]
    flops = 2.0 * (DF * Hd + 5 * (15 * Hd + 2 * Hd * Hd) + Hd * 64 + 32) * n
    ck = os.environ.get("TD_CHECK")
    diff = ""
    if ck and os.path.exists(ck):
        ref = torch.load(ck)
        diff = f"  max|tsdf - default| {(tsdf.cpu() - ref['tsdf']).abs().max().item():.2e}  max|out - default| {(out.cpu() - ref['out']).abs().max().item():.2e}"
    elif ck:
        torch.save({"tsdf": tsdf.cpu(), "out": out.cpu()}, ck)
    print(f"  Hd={Hd} d_feat={DF} rows={n}: {m:.3f} ms  {flops / m / 1e9:.1f} TFLOP/s  ({flops / m / 1e9 / 1389.9:.3f} of sustained bf16 peak){diff}", flush=True)
else:
    names = sys.argv[1:] or list(VARIANTS)
    ck = "/tmp/td_check.pt"
    if os.path.exists(ck):
        os.remove(ck)
    for name in names:
        print(name, flush=True)
        env = dict(os.environ)
        env.update(VARIANTS.get(name, {}))
        if VARIANTS.get(name, {}).get("GNB_TC_TWO_CTA") == "0":
            env.pop("GNB_TC_TWO_CTA", None)
        env["TD_CHECK"] = ck
        subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300)

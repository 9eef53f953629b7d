"""Tuning aid: the query phase of the bench workload (config 4: 256x256x96 volume of 32 channels + 3x256^2x32 planes) split
into its parts, on TQ_Q queries (default 4 Mi): sampler alone (volume / planes / both), decoder alone, fused kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gennerf_b200 import ops, synthetic as S
dev = "cuda"
Q = int(os.environ.get("TQ_Q", str(1 << 22)))
vd, VS = (256, 256, 96), 0.04
g = S.gen(5)
w, hw, hb = S.decoder_weights(g, 64, 15, 512, 5, 64, 32)
dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=dev)
xyz = S.query_points(Q, vd, VS, g).to(dev)
vol = torch.randn(1, *vd, 32, device=dev).permute(0, 4, 1, 2, 3)
pl = {k: torch.randn(1, 256, 256, 32, device=dev).permute(0, 3, 1, 2) for k in ops.PLANES}
origin = torch.zeros(1, 3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2] / (Q / (1 << 20))
kw = dict(voxel_size=VS, origin=origin, padding=0.1)
feat = ops.sample_features(xyz, volume=vol, planes=pl, **kw)
print(f"queries {Q}; all times in ms per Mi queries")
print("sampler volume only      %.3f" % t(lambda: ops.sample_features(xyz, volume=vol, **kw)))
print("sampler volume (generic) %.3f" % t(lambda: ops.sample_features(xyz, volume=vol, binned=False, **kw)))
print("sampler planes only      %.3f" % t(lambda: ops.sample_features(xyz, planes=pl, **kw)))
print("sampler volume + planes  %.3f" % t(lambda: ops.sample_features(xyz, volume=vol, planes=pl, **kw)))
print("sampler both (generic)   %.3f" % t(lambda: ops.sample_features(xyz, volume=vol, planes=pl, binned=False, **kw)))
print("decoder alone (d_feat 64) %.3f" % t(lambda: ops.decode(dw, xyz, feat, "fp16")))
print("image (sampler -> operand image -> decoder) %.3f" % t(lambda: ops.query_image(dw, xyz, volume=vol, planes=pl, **kw)))
print("fused, auto              %.3f" % t(lambda: ops.query_fused(dw, xyz, volume=vol, planes=pl, want_feat=False, mode="fused", **kw)))
for ps in (False, True):
    try:
        print(f"fused, presort={ps}      %.3f" % t(lambda: ops.query_fused(dw, xyz, volume=vol, planes=pl, want_feat=False, presort=ps, mode="fused", **kw)))
    except Exception as e:
        print("fused presort", ps, "failed:", e)
print("fused volume only        %.3f" % t(lambda: ops.query_fused(dw if False else ops.DecoderWeights(*S.decoder_weights(S.gen(5), 32, 15, 512, 5, 64, 32), n_blocks=5, d_geo=32, device=dev), xyz, volume=vol, want_feat=False, **kw)))

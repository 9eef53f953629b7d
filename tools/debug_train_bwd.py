"""Debug aid: native (gnb_decode_train_bwd) vs python-chain backward of the fp16 training decoder, per tensor, with fp32 and
TF32 GEMMs, at the config-5 query count."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gennerf_b200 import train_decode  # noqa: E402
from gennerf_b200.dropin import decode_train  # noqa: E402
from test_gpu_train_decode import _setup  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def run(n, tf32, native, use_out, prec="fp16"):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    train_decode.NATIVE_BACKWARD = native
    mlp, head, code, xyz, feat, g = _setup(n, seed=91)
    gout = torch.randn(1, n, 64, generator=g).cuda()
    target = (torch.rand(1, n, 1, generator=g) * 2 - 1).cuda()
    x1, f1 = xyz.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    out, tsdf = decode_train(mlp, head, code, x1, f1, precision=prec)
    loss = (tsdf - target).abs().mean()
    if use_out:
        loss = loss + (out * gout).sum() / out.numel()
    loss.backward()
    r = {"xyz": x1.grad, "feat": f1.grad}
    r.update({k: p.grad.clone() for k, p in mlp.named_parameters()})
    r.update({"head." + k: p.grad.clone() for k, p in head.named_parameters()})
    return r


for n in (5000, 23200):
    for use_out in (False, True):
        ref = run(n, False, False, use_out, "fp32")
        for tf32 in (False, True):
            a, b = run(n, tf32, True, use_out), run(n, tf32, False, use_out)
            worst = max(((rel(a[k], b[k]), k) for k in b))
            print(f"n {n} use_out {use_out} tf32 {tf32}: native vs python worst {worst[0]:.2e} ({worst[1]}); feat {rel(a['feat'], b['feat']):.2e}; "
                  f"vs fp32 autograd: native feat {rel(a['feat'], ref['feat']):.2e} python feat {rel(b['feat'], ref['feat']):.2e} "
                  f"norms {a['feat'].norm().item():.4e} {b['feat'].norm().item():.4e} {ref['feat'].norm().item():.4e}")

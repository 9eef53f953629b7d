import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gennerf_b200 import ops, synthetic as S
from gennerf_b200.dropin import PositionalEncoding, ResnetFC, TSDFHeadSimple, decode_train
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train_decode import _setup, _run
DEV = "cuda"
n = 3000
mlp, head, code, xyz, feat, g = _setup(n)
# saved activations vs an fp32 recomputation
with torch.no_grad():
    x2 = xyz.reshape(-1, 3)
    emb = torch.sin(torch.addcmul(code._phases.to(DEV), x2.unsqueeze(1).repeat(1, 4, 1), code._freqs.to(DEV))).view(n, -1)
    z = torch.cat((x2, emb), -1)
    f2 = feat.reshape(n, -1)
    sd = dict(mlp.named_parameters())
    from gennerf_b200.torch_ops import mlp_keys
    dw = ops.DecoderWeights({k: v for k, v in mlp.state_dict().items()}, head.fc.weight, head.fc.bias, n_blocks=5, d_geo=32, use_code=2,
                            num_freqs=0, freq_factor=0.0, include_input=False, d_code=15, device=DEV)
    out, tsdf, acts = ops.decode_save(dw, z, f2)
    x = mlp.lin_in(f2)
    for i in range(5):
        x = x + mlp.alpha * mlp.lin_z[i](z)
        a = F.relu(x)
        print(i, "a err", (acts[2 * i].float() - a).abs().max().item(), a.abs().max().item())
        net = mlp.blocks[i].fc_0(a)
        h = F.relu(net)
        print(i, "h err", (acts[2 * i + 1].float() - h).abs().max().item(), h.abs().max().item())
        x = x + mlp.blocks[i].fc_1(h)
    print("final err", (acts[10].float() - F.relu(x)).abs().max().item())
target = torch.rand(1, n, 1, generator=g).to(DEV) * 2 - 1
gout = torch.randn(1, n, 64, generator=g).to(DEV)
o32, t32, g32 = _run(mlp, head, code, xyz, feat, target, gout, "fp32")
o16, t16, g16 = _run(mlp, head, code, xyz, feat, target, gout, "fp16")
for k in g32:
    a, b = g16[k].float(), g32[k].float()
    print(f"{k:32s} rel2 {((a-b).norm()/b.norm().clamp_min(1e-20)).item():.3e}  max {((a-b).abs().max()/b.abs().max().clamp_min(1e-20)).item():.3e}")
print("---- TF32 autograd (the reference's training setting) vs fp32 autograd")
torch.backends.cuda.matmul.allow_tf32 = True
torch.set_float32_matmul_precision("high")
otf, ttf, gtf = _run(mlp, head, code, xyz, feat, target, gout, "fp32")
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_float32_matmul_precision("highest")
for k in ("xyz", "feat", "mlp.lin_in.weight", "mlp.blocks.0.fc_0.weight", "mlp.blocks.4.fc_1.weight", "mlp.lin_out.weight", "head.fc.weight"):
    a, b = gtf[k].float(), g32[k].float()
    print(f"{k:32s} rel2 {((a-b).norm()/b.norm().clamp_min(1e-20)).item():.3e}")
print("tsdf tf32 vs fp32", (ttf - t32).abs().max().item(), " fp16 vs fp32", (t16 - t32).abs().max().item())

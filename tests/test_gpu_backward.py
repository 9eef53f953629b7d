"""Gradients of the sm_100a backward kernels against CPU autograd through the oracle
(BASELINE config 5).  Tolerance: 1e-4 relative to the tensor scale (fp32 atomics
sum in a nondeterministic order, as the reference's CUDA index_put_ / grid_sampler backward do)."""
import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS = 0.04
ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)


@pytest.fixture(autouse=True, params=["eager_functions", "custom_ops"])
def autograd_path(request):
    """Every test of this file runs through both spellings of the autograd formulas (gennerf_b200/autograd.py): the plain
    autograd.Functions of eager mode and the torch.library custom ops a tracer sees."""
    from gennerf_b200 import autograd as ag
    old = ag.EAGER_FUNCTIONS
    ag.EAGER_FUNCTIONS = request.param == "eager_functions"
    yield request.param
    ag.EAGER_FUNCTIONS = old


def close(a, b, rtol, what):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = b.abs().max().clamp_min(1e-30)
    err = ((a - b).abs() / torch.maximum(b.abs(), scale)).max().item()
    assert err <= rtol, f"{what}: max rel err {err:.3e} > {rtol}"


@pytest.mark.parametrize("C,layout", [(8, "nchw"), (32, "nhwc"), (3, "nchw")])
def test_lift_backward(C, layout):
    from gennerf_b200 import autograd as ag
    wl = S.WORKLOADS["small"]
    g = S.gen(51)
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
    Pb = torch.stack([P, P.flip(0)])
    feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g, B=2)
    G = torch.randn(2, C, *wl["voxel_dim"], generator=g)
    fo = [f.clone().requires_grad_(True) for f in feats]
    vol_o, _, _ = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, Pb, fo)
    (vol_o * G).sum().backward()
    fd = [f.to(DEV) for f in feats]
    if layout == "nhwc":
        fd = [f.contiguous(memory_format=torch.channels_last) for f in fd]
    fd = [f.requires_grad_(True) for f in fd]
    vol, cnt, valid = ag.backproject_frames(wl["voxel_dim"], VS, ORIGIN, Pb, fd)
    (vol * G.to(DEV)).sum().backward()
    for t in range(wl["T"]):
        assert (fo[t].grad != 0).any()
        close(fd[t].grad, fo[t].grad, 1e-4, f"grad features[{t}]")


@pytest.mark.parametrize("with_planes,with_volume", [(True, True), (False, True), (True, False)])
def test_sampler_backward(with_planes, with_volume):
    from gennerf_b200 import autograd as ag
    g = S.gen(52)
    dims, R, Cp, C = (12, 10, 6), 16, 8, 16
    vol = torch.randn(2, C, *dims, generator=g)
    planes = {k: torch.randn(2, Cp, R, R, generator=g) for k in O.PLANES}
    xyz = S.query_points(2000, dims, VS, g, B=2)
    xyz[:, :500] = (S.plane_points(500, g, "unit", B=2) * 0.9)              # inside the unit cube: plane d/dxyz is non-zero
    D = (Cp if with_planes else 0) + (C if with_volume else 0)
    G = torch.randn(2, 2000, D, generator=g)
    # oracle
    xo = xyz.clone().requires_grad_(True)
    vo = vol.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in planes.items()}
    valid = torch.ones(2, 1, *dims, dtype=torch.bool)
    ref = O.map_features(xo, vo if with_volume else None, valid if with_volume else None, po if with_planes else None, VS, 0.1)
    (ref * G).sum().backward()
    # kernels
    xd = xyz.to(DEV).requires_grad_(True)
    vd = vol.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3).requires_grad_(True)
    pd = {k: v.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for k, v in planes.items()}
    out = ag.sample_features(xd, volume=vd if with_volume else None, planes=pd if with_planes else None, voxel_size=VS,
                             origin=ORIGIN, padding=0.1)
    (out * G.to(DEV)).sum().backward()
    close(xd.grad, xo.grad, 1e-4, "grad xyz")
    if with_volume:
        close(vd.grad, vo.grad, 1e-4, "grad volume")
    if with_planes:
        for k in O.PLANES:
            close(pd[k].grad, po[k].grad, 1e-4, f"grad plane {k}")


@pytest.mark.parametrize("with_planes,with_volume,layout", [(True, False, "cl"), (False, True, "cl"), (True, True, "cl"),
                                                            (True, True, "reference")])
def test_sampler_double_backward(with_planes, with_volume, layout):
    """The eikonal / gradient losses (reference utils.py:636-649 `calculate_grad(create_graph=True)`, model.py:385-400)
    differentiate d tsdf / d xyz once more.  CUDA: sample_features -> sample_features_bwd -> gnb_sample_features_bwd2;
    checker: CPU autograd through the reference's grid_sample_2d (planes) and the written-out trilinear sum (volume)."""
    from gennerf_b200 import autograd as ag
    g = S.gen(57)
    dims, R, Cp, C, Q = (12, 10, 6), 16, 8, 16, 1500
    vol = torch.randn(2, C, *dims, generator=g)
    planes = {k: torch.randn(2, Cp, R, R, generator=g) for k in O.PLANES}
    xyz = S.query_points(Q, dims, VS, g, B=2)                                   # some outside the grid (border clip)
    xyz[:, :600] = S.plane_points(600, g, "unit", B=2) * 1.15                   # plane domain, some clamped by normalize_coordinate
    D = (Cp if with_planes else 0) + (C if with_volume else 0)
    wgt = torch.randn(D, generator=g) * 0.3
    bdir = torch.randn(3, generator=g)

    def loss_of(feat, x):
        t = torch.tanh((feat * wgt.to(feat.device)).sum(-1) + (x * bdir.to(x.device)).sum(-1))
        (gx,) = torch.autograd.grad(t.sum(), x, create_graph=True)
        return ((gx.norm(dim=-1) - 1) ** 2).mean() + t.mean(), gx

    # checker
    xo = xyz.clone().requires_grad_(True)
    vo = vol.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in planes.items()}
    ref = O.map_features_twice_differentiable(xo, vo if with_volume else None, po if with_planes else None, VS, 0.1)
    lo, gxo = loss_of(ref, xo)
    lo.backward()
    # kernels
    xd = xyz.to(DEV).requires_grad_(True)
    if layout == "cl":
        vd = vol.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3).requires_grad_(True)
        pd = {k: v.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for k, v in planes.items()}
    else:               # the reference's own layouts (B,C,nx,ny,nz) / (B,C,R,R): strided scalar gathers and reductions
        vd = vol.to(DEV).requires_grad_(True)
        pd = {k: v.to(DEV).requires_grad_(True) for k, v in planes.items()}
    out = ag.sample_features(xd, volume=vd if with_volume else None, planes=pd if with_planes else None, voxel_size=VS,
                             origin=ORIGIN, padding=0.1)
    ld, gxd = loss_of(out, xd)
    ld.backward()
    close(gxd, gxo, 1e-4, "d tsdf / d xyz")
    assert abs(ld.item() - lo.item()) <= 1e-4 * max(1.0, abs(lo.item()))
    assert xo.grad.abs().max() > 0
    close(xd.grad, xo.grad, 2e-4, "grad xyz (second order)")
    if with_volume:
        close(vd.grad, vo.grad, 2e-4, "grad volume (second order)")
    if with_planes:
        for k in O.PLANES:
            close(pd[k].grad, po[k].grad, 2e-4, f"grad plane {k} (second order)")


def test_sampler_double_backward_through_grad_volume():
    """The other half of the sampler backward's derivative: a loss on grad_volume / grad_planes (create_graph=True)
    flows back into grad_out and xyz through the sampler and its backward applied to the incoming gradients."""
    from gennerf_b200 import autograd as ag
    g = S.gen(58)
    dims, R, Cp, C, Q = (8, 6, 5), 8, 4, 8, 400
    vol = torch.randn(1, C, *dims, generator=g)
    planes = {k: torch.randn(1, Cp, R, R, generator=g) for k in O.PLANES}
    xyz = S.query_points(Q, dims, VS, g, B=1) * 0.9 + 0.01
    wv, wp = torch.randn(1, C, *dims, generator=g), torch.randn(1, Cp, R, R, generator=g)
    wgt = torch.randn(Cp + C, generator=g)

    def run(sample, dev, x, v, pl):
        t = torch.tanh((sample(x, v, pl) * wgt.to(dev)).sum(-1)).sum()
        gv, gp = torch.autograd.grad(t, [v, pl["xy"]], create_graph=True)
        loss = (gv * wv.to(dev)).sum() + (gp * wp.to(dev)).sum()
        return torch.autograd.grad(loss, [x, v, pl["xy"]], allow_unused=True)

    xo, vo = xyz.clone().requires_grad_(True), vol.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in planes.items()}
    a = run(lambda x, v, pl: O.map_features_twice_differentiable(x, v, pl, VS, 0.1), "cpu", xo, vo, po)
    xd = xyz.to(DEV).requires_grad_(True)
    vd = vol.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3).requires_grad_(True)
    pd = {k: v.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for k, v in planes.items()}
    b = run(lambda x, v, pl: ag.sample_features(x, volume=v, planes=pl, voxel_size=VS, origin=ORIGIN, padding=0.1), DEV, xd, vd, pd)
    for name, u, w in zip(("xyz", "volume", "plane xy"), b, a):
        close(u, w, 2e-4, f"grad {name} through grad_volume / grad_planes")


@pytest.mark.parametrize("scatter_type", ["max", "mean"])
def test_triplane_backward(scatter_type):
    from gennerf_b200 import autograd as ag
    g = S.gen(53)
    N, Cp, R = 3000, 8, 16
    p = S.plane_points(N, g, "unit", B=2)
    c = torch.randn(2, N, Cp, generator=g)
    Gp = torch.randn(3, 2, Cp, R, R, generator=g)
    Gl = torch.randn(2, N, Cp, generator=g)
    co = c.clone().requires_grad_(True)
    loss = sum((O.generate_plane_features(p, co, k, R, 0.1) * Gp[i]).sum() for i, k in enumerate(O.PLANES))
    loss = loss + (O.pool_local(p, co, R, 0.1, scatter_type=scatter_type) * Gl).sum()
    loss.backward()
    cd = c.to(DEV).requires_grad_(True)
    planes, _ = ag.scatter_mean_planes(p.to(DEV), cd, R, 0.1, "atomic")
    pooled = ag.pool_local(p.to(DEV), cd, R, 0.1, scatter_type)
    ((planes * Gp.to(DEV)).sum() + (pooled * Gl.to(DEV)).sum()).backward()
    close(cd.grad, co.grad, 1e-4, "grad point features")


def test_training_step_gradients(golden_dir):
    """BASELINE config 5 in miniature: encode -> query -> L1 TSDF loss -> backward; gradients w.r.t. the
    frame features, the planes and every decoder weight against CPU autograd through the oracle."""
    from test_gpu_parity import _gennerf_from_golden, load
    G = load(golden_dir, "gennerf_forward.pt")
    i = G["in"]
    T = i["projection"].shape[1]
    target = torch.rand(1, i["xyz"].shape[1], 1, generator=S.gen(54)) * 2 - 1
    # ---- oracle -------------------------------------------------------------------------------
    fo = [i["features"][t:t + 1].clone().requires_grad_(True) for t in range(T)]
    po = {k: v.clone().requires_grad_(True) for k, v in i["planes"].items()}
    wo = {k: v.clone().requires_grad_(True) for k, v in i["weights"].items()}
    hw, hb = i["head_w"].clone().requires_grad_(True), i["head_b"].clone().requires_grad_(True)
    vol_o, valid_o, _ = O.encode_volume(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"], fo)
    ref = O.gennerf_forward(i["xyz"], wo, hw, hb, volume=vol_o, valid=valid_o, planes=po, voxel_size=i["voxel_size"],
                            padding=i["padding"], num_freqs=i["num_freqs"], freq_factor=i["freq_factor"])
    (ref["tsdf"] - target).abs().mean().backward()
    # ---- drop-in model in training mode -------------------------------------------------------------
    model = _gennerf_from_golden(G, "fp16", True).train()
    fd = [f.detach().to(DEV).requires_grad_(True) for f in fo]
    image = torch.stack(fd, dim=1)
    model.cfg.encoder.use_pointnet = False
    model.encode(i["projection"], image, None, "train")
    model.cfg.encoder.use_pointnet = True
    pd = {k: v.detach().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for k, v in po.items()}
    model.c_plane = pd
    out = model(i["xyz"].to(DEV))
    (out["tsdf"] - target.to(DEV)).abs().mean().backward()
    for t in range(T):
        close(fd[t].grad, fo[t].grad, 1e-4, f"grad features[{t}]")
    for k in O.PLANES:
        close(pd[k].grad, po[k].grad, 1e-4, f"grad plane {k}")
    sd = dict(model.mlp.named_parameters())
    for k, v in wo.items():
        close(sd[k].grad, v.grad, 2e-4, f"grad mlp.{k}")
    close(model.head_geo.fc.weight.grad, hw.grad, 2e-4, "grad head weight")

"""GPU parity of the TSDF fusion path (reference src/data/tsdf.py:320-440): every volume bit-exact
against the real reference's golden vectors and against the CPU oracle, through the C ABI."""
import os

import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS = 0.04


def test_fusion_golden(golden_dir):
    from gennerf_b200.dropin import TSDFFusion
    G = torch.load(os.path.join(golden_dir, "tsdf_fusion.pt"), weights_only=False)
    i, o = G["in"], G["out"]
    T = i["projection"].shape[0]
    f = TSDFFusion(i["voxel_dim"], i["voxel_size"], i["origin"], trunc_ratio=i["trunc_ratio"], device=DEV, color=True, label=True)
    # frame by frame, the reference's call pattern
    for t in range(T):
        f.integrate(i["projection"][t], i["depth"][t].to(DEV), i["color"][t].to(DEV), i["label"][t].long().to(DEV))
        if t == 0:
            assert torch.equal(f.tsdf_vol.cpu(), o["frame0"]["tsdf_vol"])
            assert torch.equal(f.weight_vol.cpu(), o["frame0"]["weight_vol"].float())
    for name in ("tsdf_vol", "color_vol"):
        assert torch.equal(getattr(f, name).cpu(), o["all"][name]), name
    assert torch.equal(f.weight_vol.cpu(), o["all"]["weight_vol"].float())
    assert torch.equal(f.label_vol.cpu().long(), o["all"]["label_vol"].long())
    tsdf, color, label = f.get_volumes()
    assert torch.equal(tsdf.cpu().reshape(-1), o["normalised"]["tsdf"])
    assert torch.equal(color.cpu().reshape(3, -1), o["normalised"]["color"])
    assert label.dtype == torch.long and tuple(label.shape) == tuple(i["voxel_dim"])
    # all frames in one launch: same bits
    f.reset()
    assert float(f.tsdf_vol.min()) == 1.0 and float(f.weight_vol.max()) == 0.0
    f.integrate_frames(i["projection"], i["depth"].to(DEV), i["color"].to(DEV), i["label"].to(DEV))
    assert torch.equal(f.tsdf_vol.cpu(), o["all"]["tsdf_vol"]) and torch.equal(f.color_vol.cpu(), o["all"]["color_vol"])
    assert torch.equal(f.weight_vol.cpu(), o["all"]["weight_vol"].float())
    assert torch.equal(f.label_vol.cpu().long(), o["all"]["label_vol"].long())


@pytest.mark.parametrize("vd,T,H,W,trunc_ratio,color,label,origin", [
    ((96, 96, 48), 8, 120, 160, 3, True, True, (0.0, 0.0, 0.0)),
    ((50, 33, 21), 5, 48, 64, 8, False, False, (-0.3, 0.2, 0.05)),          # ragged bricks, wide truncation band
    ((64, 64, 32), 70, 30, 40, 3, True, False, (0.0, 0.0, 0.0)),            # > GNB_MAX_FRAMES: chained launches
])
def test_fusion_vs_oracle(vd, T, H, W, trunc_ratio, color, label, origin):
    from gennerf_b200.dropin import TSDFFusion
    g = S.gen(61)
    P = S.projections(T, H, W, vd, VS, g)
    depths = S.surface_depth_maps(T, H, W, g)
    colors = torch.rand(T, 3, H, W, generator=g) if color else None
    labels = torch.randint(0, 40, (T, H, W), generator=g) if label else None
    o = O.TSDFFusion(vd, VS, origin, trunc_ratio=trunc_ratio, color=color, label=label)
    for t in range(T):
        o.integrate(P[t], depths[t], colors[t] if color else None, labels[t] if label else None)
    f = TSDFFusion(vd, VS, origin, trunc_ratio=trunc_ratio, device=DEV, color=color, label=label)
    f.integrate_frames(P, depths.to(DEV), colors.to(DEV) if color else None, labels.to(DEV) if label else None)
    assert int((o.weight_vol > 0).sum()) > 500
    assert torch.equal(f.tsdf_vol.cpu(), o.tsdf_vol) and torch.equal(f.weight_vol.cpu(), o.weight_vol)
    if color:
        assert torch.equal(f.color_vol.cpu(), o.color_vol)
    if label:
        assert torch.equal(f.label_vol.cpu().long(), o.label_vol)
    # the depth-band culling (16x16 min/max tiles of the depth maps) must not change a single bit
    from gennerf_b200 import ops
    t2, w2 = torch.ones_like(f.tsdf_vol), torch.zeros_like(f.weight_vol)
    ops.tsdf_fusion_integrate(vd, VS, origin, VS * trunc_ratio, P, depths.to(DEV), t2, w2, depth_culling=False)
    assert torch.equal(t2, f.tsdf_vol) and torch.equal(w2, f.weight_vol)
    to, co, lo = o.get_volumes()
    tf, cf, lf = f.get_volumes()
    assert torch.equal(tf.cpu(), to) and (co is None or torch.equal(cf.cpu(), co)) and (lo is None or torch.equal(lf.cpu(), lo))


def test_fusion_edge_cases():
    from gennerf_b200 import ops
    from gennerf_b200.dropin import TSDFFusion
    vd = (16, 12, 20)
    f = TSDFFusion(vd, VS, (0, 0, 0), device=DEV, color=False, label=False)
    # no frames: a no-op that needs no valid pointers
    f.integrate_frames(torch.zeros(0, 3, 4), torch.zeros(0, 8, 8, device=DEV))
    assert float(f.tsdf_vol.min()) == 1.0 and float(f.weight_vol.max()) == 0.0
    # a frame without a single measurement, and a camera that looks away from the volume
    g = S.gen(5)
    P = S.projections(2, 24, 32, vd, VS, g)
    f.integrate_frames(P, torch.zeros(2, 24, 32, device=DEV))
    assert float(f.tsdf_vol.min()) == 1.0 and float(f.weight_vol.max()) == 0.0
    # a camera 50 m away that looks away from the volume: every voxel is behind it
    P_away = torch.tensor([[20.0, 0, 16, 0], [0, 20.0, 12, 0], [0, 0, 1, 50.0]]).mul(torch.tensor([[1.0], [1.0], [-1.0]])).expand(2, 3, 4)
    f.integrate_frames(P_away, torch.ones(2, 24, 32, device=DEV))
    assert float(f.tsdf_vol.min()) == 1.0 and float(f.weight_vol.max()) == 0.0
    # mirrored cameras (what used to be behind is now in front): whatever happens must match the oracle
    P_flip = P.clone()
    P_flip[:, 2, :] = -P_flip[:, 2, :]
    o = O.TSDFFusion(vd, VS, (0, 0, 0), color=False, label=False)
    for t in range(2):
        o.integrate(P_flip[t], torch.ones(24, 32))
    f.integrate_frames(P_flip, torch.ones(2, 24, 32, device=DEV))
    assert torch.equal(f.tsdf_vol.cpu(), o.tsdf_vol) and torch.equal(f.weight_vol.cpu(), o.weight_vol)
    f.reset()
    with pytest.raises(RuntimeError):
        ops.tsdf_fusion_integrate(vd, VS, (0, 0, 0), 0.12, P, torch.ones(2, 24, 32), f.tsdf_vol, f.weight_vol)   # CPU depth
    with pytest.raises(ValueError):
        ops.tsdf_fusion_integrate(vd, VS, (0, 0, 0), 0.12, P, torch.ones(2, 24, 32, device=DEV), f.tsdf_vol[:-1], f.weight_vol)


# ---- training-time ray sampler (reference src/models/utils.py:458-540) ------------------------------
def test_sample_points_on_rays_golden(golden_dir):
    from gennerf_b200 import ops
    G = torch.load(os.path.join(golden_dir, "ray_points.pt"), weights_only=False)
    i, o = G["in"], G["out"]
    args = [i["h_idxs"].long().to(DEV), i["w_idxs"].long().to(DEV), i["depths"].to(DEV), i["intrinsics"].to(DEV), i["poses"].to(DEV)]
    xyz, z = ops.sample_points_on_rays(*args, i["N"], i["M"], i["delta"], i["min_dist"], i["sigma"], gaussian_depths=i["gaussian_depths"].to(DEV))
    assert ((z.cpu() - o["z"]).abs() <= 1e-6 * o["z"].abs().clamp_min(1.0)).all()                       # vs the real reference (CPU linspace)
    assert ((xyz.cpu() - o["xyz_world"]).abs() <= 1e-6 * o["xyz_world"].abs().clamp_min(1.0)).all()
    # vs the oracle (same element formula): bit-exact
    xo, zo = O.sample_points_on_rays(i["h_idxs"].long(), i["w_idxs"].long(), i["depths"], i["intrinsics"], i["poses"], i["N"], i["M"],
                                     i["delta"], i["min_dist"], i["gaussian_depths"])
    assert torch.equal(z.cpu(), zo) and torch.equal(xyz.cpu(), xo)
    # the drop-in signature draws the gaussian depths itself: shapes, surface sample, spread
    from gennerf_b200.dropin import sample_points_on_rays
    x2, z2 = sample_points_on_rays(*args, N=i["N"], M=i["M"], delta=i["delta"], min_dist=i["min_dist"], sigma=i["sigma"])
    assert x2.shape == xyz.shape and torch.equal(z2[..., :1 + i["N"]], z[..., :1 + i["N"]])
    spread = (z2[..., 1 + i["N"]:] - z2[..., :1]).std().item()
    assert 0.08 < spread < 0.12
    # empty batch
    e = ops.sample_points_on_rays(args[0][:0], args[1][:0], args[2][:0], args[3][:0], args[4][:0], 20, 8, 0.1, 0.07, 0.1)
    assert e[0].shape == (0, 100, 29, 3)

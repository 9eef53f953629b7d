"""Pin the CPU oracle against the REAL reference, imported live from /root/reference.

These tests only run where the reference tree is mounted (the build container).  On the
GPU box the same guarantees travel as committed golden vectors (tests/golden/*.pt, made by
tests/golden/make_golden.py from the real reference; checked in test_oracle_golden.py).
Bit-for-bit equality is required everywhere: the oracle calls the same ATen CPU kernels
in the same order as the reference.
"""
import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O
from oracle import ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")]

ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)          # reference model.py:57 (int64 zeros)
VS = 0.04


@pytest.fixture(scope="module")
def ref():
    return ref_shim.ref_modules()


@pytest.mark.parametrize("name,C", [("tiny", 4), ("small", 8), ("small", 3)])
def test_backproject_and_accumulate(ref, name, C):
    wl = S.WORKLOADS[name]
    g = S.gen(11)
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
    feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g)
    vol_ref = val_ref = None
    for t in range(wl["T"]):
        v, m = ref.utils.backproject(wl["voxel_dim"], VS, ORIGIN, P[t:t + 1], feats[t])
        vo, mo = O.backproject(wl["voxel_dim"], VS, ORIGIN, P[t:t + 1], feats[t])
        assert torch.equal(v, vo) and torch.equal(m, mo) and mo.dtype == torch.bool
        # accumulate as GenNerf.encode does (model.py:122-127)
        vol_ref = v if vol_ref is None else vol_ref + v
        val_ref = m if val_ref is None else val_ref + m
    vol_o, val_o, cnt_o = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, P.unsqueeze(0), feats)
    assert torch.equal(vol_ref, vol_o) and torch.equal(val_ref, val_o)
    assert val_o.dtype == torch.bool                      # trap T2: OR, not a count
    assert torch.equal(cnt_o > 0, val_o.squeeze(1))
    # the normalised volume the decoder reads equals the SUM (trap T2)
    assert torch.equal(O.normalize_volume(vol_o, val_o), vol_o)


def test_projection_is_fma_chain(ref):
    wl = S.WORKLOADS["cfg1"]
    P = S.projections(4, wl["H"], wl["W"], wl["voxel_dim"], VS, S.gen(5))
    a = O.project_indices(wl["voxel_dim"], VS, ORIGIN, P, wl["H"], wl["W"])
    b = O.project_indices_explicit(wl["voxel_dim"], VS, ORIGIN, P, wl["H"], wl["W"])
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert torch.equal(ref.tsdf.coordinates(wl["voxel_dim"], torch.device("cpu")),
                       O.coordinates(wl["voxel_dim"]))


@pytest.mark.parametrize("C", [1, 8])
def test_trilinear(ref, C):
    g = S.gen(21)
    dims = (9, 7, 5)
    vol = torch.randn(2, C, *dims, generator=g).permute(0, 2, 3, 4, 1)     # strided view, as model.py:201
    xyz = S.query_points(513, dims, VS, g, B=2)
    a = ref.utils.trilinear_interpolation(vol, xyz, ORIGIN.squeeze(), VS)
    b = O.trilinear_interpolation(vol, xyz, ORIGIN.squeeze(), VS)
    assert torch.equal(a, b)
    c = O.trilinear_interpolation_explicit(vol, xyz, ORIGIN.squeeze(), VS)
    assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("domain", ["unit", "metric"])
def test_plane_indices_and_scatter(ref, domain):
    g = S.gen(31)
    N, Cp, R = 3001, 8, 32
    p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48), B=2)
    p[0, :7] = torch.tensor([[-0.55, 0.55, 0.0], [0.55, 0.55, 0.55], [-0.6, 0.7, -0.7], [0.0, 0.0, 0.0],
                             [0.549999, -0.549999, 0.5500001], [1e-9, -1e-9, 3.0], [-3.0, 3.0, 0.1]])
    c = torch.randn(2, N, Cp, generator=g)
    pn = ref.pointnet.LocalPoolPointnet(c_dim=Cp, dim=3, hidden_dim=8, scatter_type="max", unet=False,
                                        plane_resolution=R, plane_type=["xz", "xy", "yz"], padding=0.1, n_blocks=2)
    for plane in O.PLANES:
        xy_r = ref.utils.normalize_coordinate(p.clone(), padding=0.1, plane=plane)
        xy_o = O.normalize_coordinate(p.clone(), padding=0.1, plane=plane)
        assert torch.equal(xy_r, xy_o)
        assert torch.equal(ref.utils.coordinate2index(xy_r, R), O.coordinate2index(xy_o, R))
        assert torch.equal(pn.generate_plane_features(p, c, plane), O.generate_plane_features(p, c, plane, R, 0.1))
    coord = {k: ref.utils.normalize_coordinate(p.clone(), plane=k, padding=0.1) for k in O.PLANES}
    index = {k: ref.utils.coordinate2index(coord[k], R) for k in O.PLANES}
    assert torch.equal(pn.pool_local(coord, index, c), O.pool_local(p, c, R, 0.1, scatter_type="max"))
    pn.scatter = ref.pointnet.scatter_mean
    assert torch.equal(pn.pool_local(coord, index, c), O.pool_local(p, c, R, 0.1, scatter_type="mean"))


def test_scatter_add_is_sequential_in_point_order():
    """CPU scatter_add_ == sequential fp32 sum in point order."""
    g = S.gen(41)
    N, cells = 20000, 64
    src = torch.randn(1, 1, N, generator=g)
    idx = torch.randint(0, cells, (1, 1, N), generator=g)
    out = torch.zeros(1, 1, cells).scatter_add_(2, idx, src)
    seq = torch.zeros(cells)
    s, i = src.view(-1).tolist(), idx.view(-1).tolist()
    import numpy as np
    acc = np.zeros(cells, dtype=np.float32)
    for k in range(N):
        acc[i[k]] = np.float32(acc[i[k]] + np.float32(s[k]))
    assert torch.equal(out.view(-1), torch.from_numpy(acc))


def test_plane_query(ref):
    g = S.gen(51)
    Cp, R = 8, 16
    planes = {k: torch.randn(2, Cp, R, R, generator=g) for k in O.PLANES}
    xyz = S.plane_points(777, g, "unit", B=2) * 1.2
    GenNerf = ref_shim.ref_gennerf()
    fake = type("F", (), {})()
    fake.cfg = ref_shim.to_attr({"encoder": {"pointnet": {"padding": 0.1, "sample_mode": "bilinear"}},
                                 "loss": {"use_eikonal": False, "use_gradient": False}})
    for k in O.PLANES:
        a = GenNerf.sample_plane_feature(fake, xyz, planes[k], plane=k)
        b = O.sample_plane_feature(xyz, planes[k], k, 0.1)
        assert torch.equal(a, b)
        c = O.sample_plane_feature_explicit(xyz, planes[k], k, 0.1)
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)
        # the pure-torch variant used with eikonal/gradient losses (utils.py:1117) agrees too
        xy = ref.utils.normalize_coordinate(xyz.clone(), plane=k, padding=0.1)
        d = ref.utils.grid_sample_2d(planes[k], 2.0 * xy[:, :, None].float() - 1.0).squeeze(-1)
        assert torch.allclose(a, d, rtol=1e-5, atol=1e-6)


def test_plane_query_eikonal_variant_and_its_double_backward(ref):
    """model.py:157-158 with loss.use_eikonal: the reference's grid_sample_2d (utils.py:1117-1174).  The oracle's restatement
    equals it bit for bit -- values, first derivatives and the double backward of an eikonal-style loss."""
    g = S.gen(53)
    Cp, R = 8, 16
    xyz = S.plane_points(300, g, "unit", B=2) * 1.2
    GenNerf = ref_shim.ref_gennerf()
    fake = type("F", (), {})()
    fake.cfg = ref_shim.to_attr({"encoder": {"pointnet": {"padding": 0.1, "sample_mode": "bilinear"}},
                                 "loss": {"use_eikonal": True, "use_gradient": False}})
    wgt = torch.randn(Cp, generator=g)

    def run(sample):
        x = xyz.clone().requires_grad_(True)
        pl = {k: torch.randn(2, Cp, R, R, generator=S.gen(54 + i)).requires_grad_(True) for i, k in enumerate(O.PLANES)}
        f = sum(sample(x, pl[k], k) for k in O.PLANES)                     # (B,Cp,Q)
        t = torch.tanh((f * wgt.view(1, -1, 1)).sum(1))
        (gx,) = torch.autograd.grad(t.sum(), x, create_graph=True)
        loss = ((gx.norm(dim=-1) - 1) ** 2).mean() + t.mean()
        grads = torch.autograd.grad(loss, [x] + [pl[k] for k in O.PLANES])
        return [f.detach(), gx.detach()] + list(grads)

    a = run(lambda x, c, k: GenNerf.sample_plane_feature(fake, x, c, plane=k))
    b = run(lambda x, c, k: O.sample_plane_feature_eikonal(x, c, k, 0.1))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    assert a[2].abs().max() > 0 and a[3].abs().max() > 0


@pytest.mark.parametrize("num_freqs,ff", [(2, 0.5), (6, 1.5)])
def test_positional_encoding(ref, num_freqs, ff):
    x = torch.randn(100, 3, generator=S.gen(61)) * 3
    pe = ref.posenc.PositionalEncoding(num_freqs=num_freqs, d_in=3, freq_factor=ff, include_input=True)
    assert torch.equal(pe(x), O.positional_encoding(x, num_freqs, ff, True))


@pytest.mark.parametrize("d_hidden,d_code,d_feat,d_out", [(64, 15, 24, 16), (32, 39, 8, 9)])
def test_resnetfc_and_head(ref, d_hidden, d_code, d_feat, d_out):
    g = S.gen(71)
    w, hw, hb = S.decoder_weights(g, d_feat, d_code, d_hidden, 5, d_out, d_geo=d_out // 2, alpha=0.7)
    mlp = ref.resnetfc.ResnetFC(d_in=d_feat, d_out=d_out, n_blocks=5, d_latent=d_code, d_hidden=d_hidden)
    assert set(mlp.state_dict().keys()) == set(w.keys())
    mlp.load_state_dict(w)
    head = ref.heads3d.TSDFHeadSimple(d_out // 2)
    head.load_state_dict({"fc.weight": hw, "fc.bias": hb})
    zx = torch.randn(3, 50, d_code + d_feat, generator=g)
    with torch.no_grad():
        a = mlp(zx)
        b = O.resnetfc_forward(zx, w, 5, d_code)
        assert torch.equal(a, b)
        assert torch.equal(head(a[..., :d_out // 2]), O.tsdf_head(b[..., :d_out // 2], hw, hb))


def test_gennerf_encode_forward_end_to_end():
    """Whole path through the reference's GenNerf (CNN replaced by a pass-through; FPS front
    end bypassed by handing the point cloud to the PointNet directly is NOT possible without
    edits, so the planes come from the reference LocalPoolPointnet on the same points)."""
    GenNerf = ref_shim.ref_gennerf()
    wl = S.WORKLOADS["small"]
    g = S.gen(81)
    C = 64                                                    # num_layers 1 -> 64 channels
    cfg = ref_shim.load_model_cfg("gen_nerf", voxel_dim_train=list(wl["voxel_dim"]),
                                  voxel_dim_val=list(wl["voxel_dim"]), voxel_size=VS)
    cfg.encoder.spatial.num_layers = 1
    cfg.encoder.use_pointnet = False
    cfg.mlp.d_hidden = 64
    model = GenNerf(cfg).eval()
    w, hw, hb = S.decoder_weights(g, C, 15, 64, 5, 64, 32)
    model.mlp.load_state_dict(w)
    model.head_geo.load_state_dict({"fc.weight": hw, "fc.bias": hb})
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8).unsqueeze(0)
    feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g)
    image = torch.stack(feats, dim=1)                         # (B,T,C,H,W): pass-through "CNN"
    depth = S.depth_maps(wl["T"], wl["H"], wl["W"], g)
    xyz = S.query_points(500, wl["voxel_dim"], VS, g)
    with torch.no_grad():
        model.encode(P, image, depth, "val")
        out = model.forward(xyz)
        vol, valid, _ = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, P, feats)
        assert torch.equal(model.volume, vol) and torch.equal(model.valid, valid)
        o = O.gennerf_forward(xyz, w, hw, hb, volume=vol, valid=valid, voxel_size=VS,
                              num_freqs=cfg.code.num_freqs, freq_factor=cfg.code.freq_factor)
    for k in ("feat", "feat_geo", "feat_sem", "tsdf"):
        assert torch.equal(out[k], o[k]), k


def test_next_rows(ref):
    g = S.gen(91)
    P = S.projections(2, 24, 32, (12, 10, 6), VS, g, pull_back=0.8)
    depth = S.depth_maps(2, 24, 32, g)[0]
    a = ref.utils.get_3d_points(depth, P)
    b = O.get_3d_points(depth, P)
    assert torch.equal(a, b)
    assert torch.equal(ref.utils.get_grid_coordinates(5, 6, 7, [1.0, 2.0, 0.5], None, "cpu"),
                       O.get_grid_coordinates(5, 6, 7, [1.0, 2.0, 0.5]))
    xyz = a.reshape(2, -1, 3)
    torch.manual_seed(3)
    s_ref, c_ref = ref.utils.farthest_point_sample(xyz, 16)
    s_o, c_o = O.farthest_point_sample(xyz, 16, c_ref[:, 0])
    assert torch.equal(c_ref, c_o) and torch.equal(s_ref, s_o)


# ---- TSDF fusion (src/data/tsdf.py:320-440) -------------------------------------
@pytest.mark.parametrize("color,label,trunc_ratio,origin", [(True, True, 3, (0.1, -0.2, 0.05)), (False, False, 3, (0, 0, 0)),
                                                           (True, False, 8, (0.0, 0.0, 0.0)), (False, True, 1.5, (-0.3, 0.2, 0.0))])
def test_tsdf_fusion(color, label, trunc_ratio, origin):
    ref_shim.install()
    from src.data.tsdf import TSDFFusion as RefFusion
    g = S.gen(31)
    vd, T, H, W = (40, 36, 20), 5, 48, 64
    P = S.projections(T, H, W, vd, VS, g)
    depths = S.surface_depth_maps(T, H, W, g)
    colors = torch.rand(T, 3, H, W, generator=g)
    labels = torch.randint(0, 40, (T, H, W), generator=g)
    r = RefFusion(vd, VS, origin, trunc_ratio=trunc_ratio, device=torch.device("cpu"), color=color, label=label)
    o = O.TSDFFusion(vd, VS, origin, trunc_ratio=trunc_ratio, color=color, label=label)
    assert torch.equal(r.world, o.world)
    for t in range(T):
        r.integrate(P[t], depths[t], colors[t] if color else None, labels[t] if label else None)
        o.integrate(P[t], depths[t], colors[t] if color else None, labels[t] if label else None)
        assert torch.equal(r.tsdf_vol, o.tsdf_vol) and torch.equal(r.weight_vol, o.weight_vol)
    if color:
        assert torch.equal(r.color_vol, o.color_vol)
    if label:
        assert torch.equal(r.label_vol, o.label_vol)
    # the per-voxel formulation the CUDA kernel follows (FMA-chain projection, running update in frame order)
    e = O.tsdf_fusion_explicit(vd, VS, origin, trunc_ratio, P, depths, colors if color else None, labels if label else None)
    assert torch.equal(r.tsdf_vol, e[0]) and torch.equal(r.weight_vol, e[1])
    assert (not color) or torch.equal(r.color_vol, e[2])
    assert (not label) or torch.equal(r.label_vol, e[3])
    assert int((r.weight_vol > 0).sum()) > 100 and int(((r.weight_vol == 0) & (r.tsdf_vol == -1)).sum()) > 0
    # reset (tsdf.py:359-367)
    r.reset(), o.reset()
    assert torch.equal(r.tsdf_vol, o.tsdf_vol) and torch.equal(r.weight_vol, o.weight_vol)


# ---- training-time ray sampler (src/models/utils.py:458-540) --------------------
@pytest.mark.parametrize("N,M", [(20, 8), (5, 0), (33, 3)])
def test_sample_points_on_rays(N, M):
    ref_shim.install()
    from src.models.utils import sample_points_on_rays as ref_fn
    g = S.gen(37)
    B, Sn, H, W = 2, 64, 120, 160
    h = torch.randint(0, H, (B, Sn), generator=g)
    w = torch.randint(0, W, (B, Sn), generator=g)
    D = torch.rand(B, Sn, generator=g) * 3 + 0.4
    K = S.intrinsics(H, W).expand(B, 3, 3).contiguous()
    poses = torch.stack(list(S.camera_poses(B, (48, 48, 24), VS, g)))
    torch.manual_seed(5)
    xr, zr = ref_fn(h, w, D, K, poses, N=N, M=M, delta=0.1, min_dist=0.07, sigma=0.1)
    torch.manual_seed(5)                                       # replay the reference's draw (utils.py:496-498)
    gd = torch.stack([torch.normal(D[b].unsqueeze(-1).expand(Sn, M), 0.1 * torch.ones(Sn, M)) for b in range(B)])
    xo, zo = O.sample_points_on_rays(h, w, D, K, poses, N, M, 0.1, 0.07, gd)
    assert xo.shape == xr.shape and zo.shape == zr.shape
    # surface and gaussian depths are copied: exact; the stratified depths differ from ATen's vectorised CPU linspace
    # in the last bit of some elements (see the oracle's docstring): 1e-6 relative
    assert torch.equal(zo[..., 0], zr[..., 0]) and torch.equal(zo[..., 1 + N:], zr[..., 1 + N:])
    assert ((zo - zr).abs() <= 1e-6 * zr.abs().clamp_min(1.0)).all()
    assert ((xo - xr).abs() <= 1e-6 * xr.abs().clamp_min(1.0)).all()


def test_sample_valid_depth_pixels():
    """utils.py:340-363: the oracle's deterministic half on the reference's own randperm draws == the reference."""
    ref_shim.install()
    from src.models.utils import sample_valid_depth_pixels as ref_fn
    g = S.gen(38)
    depth = S.surface_depth_maps(3, 31, 45, g)
    depth[1, :4] = 0.0
    Sn = 150
    torch.manual_seed(9)
    b_r, h_r, w_r = ref_fn(depth, Sn)
    torch.manual_seed(9)                                       # replay the reference's draws
    ranks = torch.stack([torch.randperm(int((depth[b] != 0).sum()))[:Sn] for b in range(3)])
    h_o, w_o = O.select_valid_depth_pixels(depth, ranks)
    assert torch.equal(h_r, h_o) and torch.equal(w_r, w_o) and b_r.shape == (3, 1)

"""Parity of the sm_100a kernels (through the C ABI) with the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): voxel indices, validity masks and scatter counts bit-exact;
fp32 features / TSDF within 1e-5 relative (stated per test as rtol plus an absolute floor
of 1e-5 x the largest reference magnitude); the 16-bit (fp16 operand) tcgen05 decoder within 1e-2 absolute TSDF.
"""
import os

import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

pytestmark = pytest.mark.gpu

ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)
VS = 0.04
DEV = "cuda"


def ops():
    from gennerf_b200 import ops as _ops
    return _ops


def close(a, b, rtol=1e-5, what=""):
    """|a-b| <= rtol * max(|b|, max|b|) -- 1e-5 relative with a floor at the tensor's scale."""
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = b.abs().max().clamp_min(1e-30)
    err = (a - b).abs() / torch.maximum(b.abs(), scale)
    assert err.max().item() <= rtol, f"{what}: max rel err {err.max().item():.3e} > {rtol}"


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


# ---------------------------------------------------------------------------------------------
# lift
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,C", [("tiny", 4), ("small", 8), ("small", 3), ("small", 32), ("small", 128), ("tiny", 160)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_lift_vs_oracle(name, C, layout):
    wl = S.WORKLOADS[name]
    g = S.gen(11)
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
    feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g, B=2)
    Pb = torch.stack([P, P.flip(0)])                                           # two scenes, different cameras
    vol_o, valid_o, cnt_o = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, Pb, feats)
    fd = [f.to(DEV) for f in feats]
    if layout == "nhwc":
        fd = [f.contiguous(memory_format=torch.channels_last) for f in fd]
    vol, cnt, valid = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, Pb, fd)
    assert torch.equal(valid.cpu(), valid_o), "validity mask must be bit-exact"
    assert torch.equal(cnt.cpu(), cnt_o), "frame counts must be bit-exact"
    assert torch.equal(vol.cpu(), vol_o), "the frame-ordered sum is bit-exact (payload is a copy)"
    # reference memory layout (B,C,nx,ny,nz) contiguous
    vol2, _, _ = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, Pb, fd, volume_layout="reference")
    assert vol2.is_contiguous() and torch.equal(vol2.cpu(), vol_o)
    # masked mean (north-star variant)
    vol3, _, _ = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, Pb, fd, mean=True)
    mean_o = vol_o / cnt_o.clamp_min(1).unsqueeze(1).float()
    assert torch.equal(vol3.cpu(), mean_o)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_lift_vs_golden(golden_dir, name):
    G = load(golden_dir, f"backproject_{name}.pt")
    i, o = G["in"], G["out"]
    T = i["projection"].shape[0]
    feats = [i["features"][t:t + 1].to(DEV) for t in range(T)]
    vol, cnt, valid = ops().backproject_frames(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"].unsqueeze(0), feats)
    assert torch.equal(vol.cpu(), o["volume_sum"]) and torch.equal(valid.cpu(), o["valid_or"])
    assert torch.equal(cnt.cpu().view(-1), o["valid_per_frame"].view(T, -1).sum(0).to(torch.int32))
    # per-frame drop-in shim == reference backproject()
    from gennerf_b200.dropin import backproject
    v0, m0 = backproject(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"][0:1], feats[0])
    assert torch.equal(v0.cpu(), o["frame0_volume"]) and torch.equal(m0.cpu(), o["frame0_valid"]) and m0.dtype == torch.bool
    # accumulate in two calls (encode called repeatedly) == one call
    h = T // 2
    out = ops().backproject_frames(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"][:h].unsqueeze(0), feats[:h])
    vol2, cnt2, valid2 = ops().backproject_frames(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"][h:].unsqueeze(0),
                                                   feats[h:], out=out)
    assert torch.equal(vol2.cpu(), o["volume_sum"]) and torch.equal(valid2.cpu(), o["valid_or"])


def test_project_indices_bit_exact():
    wl = S.WORKLOADS["cfg1"]
    P = S.projections(4, wl["H"], wl["W"], wl["voxel_dim"], VS, S.gen(5))
    px_o, py_o, _, valid_o = O.project_indices(wl["voxel_dim"], VS, ORIGIN, P, wl["H"], wl["W"])
    for t in range(4):
        px, py, valid = ops().project_indices(wl["voxel_dim"], VS, ORIGIN, P[t], wl["H"], wl["W"])
        assert torch.equal(valid.cpu(), valid_o[t])
        fits = (px_o[t].abs() < 2 ** 31 - 200) & (py_o[t].abs() < 2 ** 31 - 200)
        assert torch.equal(px.cpu().long()[fits], px_o[t][fits]) and torch.equal(py.cpu().long()[fits], py_o[t][fits])
        assert fits[valid_o[t]].all()


def test_lift_more_than_64_frames():
    wl = S.WORKLOADS["tiny"]
    g = S.gen(13)
    T = 70
    P = S.projections(T, wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
    feats = S.frame_features(T, 4, wl["H"], wl["W"], g)
    vol_o, valid_o, cnt_o = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, P.unsqueeze(0), feats)
    vol, cnt, valid = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, P.unsqueeze(0), [f.to(DEV) for f in feats])
    assert torch.equal(vol.cpu(), vol_o) and torch.equal(cnt.cpu(), cnt_o) and torch.equal(valid.cpu(), valid_o)


def test_lift_full_size_properties():
    """BASELINE config 1/2 size: size-independent properties instead of a full oracle run --
    linearity in the features, count == number of per-frame valid masks, mean * count == sum."""
    wl = S.WORKLOADS["cfg2"]
    g = S.gen(1002)
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g).unsqueeze(0)
    feats = [f.to(DEV) for f in S.frame_features(wl["T"], 32, wl["H"], wl["W"], g)]
    vol, cnt, valid = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, P, feats)
    vol2, _, _ = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, P, [2.0 * f for f in feats])
    assert torch.equal(vol2, 2.0 * vol)                                            # exact: scaling by 2
    per = torch.stack([ops().project_indices(wl["voxel_dim"], VS, ORIGIN, P[0, t], wl["H"], wl["W"])[2] for t in range(wl["T"])])
    assert torch.equal(per.sum(0).to(torch.int32), cnt.view(-1))
    assert torch.equal(valid.view(-1), cnt.view(-1) > 0)
    assert (vol.permute(0, 2, 3, 4, 1)[~valid.squeeze(1)] == 0).all()
    # a slab of the grid against the oracle (x-slab of 4 voxel planes keeps the CPU time small)
    ones = [torch.ones_like(f) for f in feats]
    vol1, _, _ = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, P, ones)
    assert torch.equal(vol1[:, 0].to(torch.int32), cnt)                             # sum of ones == count


# ---------------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,layout", [(1, "ref"), (6, "ref"), (8, "cl"), (32, "cl"), (128, "cl"), (12, "cl")])
def test_trilinear_vs_oracle(C, layout):
    from gennerf_b200.dropin import trilinear_interpolation
    g = S.gen(21)
    dims = (9, 7, 5)
    vol = torch.randn(2, C, *dims, generator=g)
    xyz = S.query_points(1500, dims, VS, g, B=2)
    xyz[0, :4] = torch.tensor([[0.0, 0.0, 0.0], [9 * VS, 7 * VS, 5 * VS], [-1.0, 0.1, 9.0], [0.36, 0.28, 0.2]])
    ref = O.trilinear_interpolation(vol.permute(0, 2, 3, 4, 1), xyz, ORIGIN.squeeze(), VS)
    vd = vol.to(DEV)
    if layout == "cl":
        vd = vd.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)       # channels-last storage
    out = trilinear_interpolation(vd.permute(0, 2, 3, 4, 1), xyz.to(DEV), ORIGIN.squeeze(), VS)
    close(out, ref, 1e-5, "trilinear")


def test_trilinear_golden(golden_dir):
    from gennerf_b200.dropin import trilinear_interpolation
    G = load(golden_dir, "trilinear.pt")
    vol = G["in"]["volume_ncxyz"].to(DEV).permute(0, 2, 3, 4, 1)
    out = trilinear_interpolation(vol, G["in"]["xyz"].to(DEV), ORIGIN.squeeze(), G["in"]["voxel_size"])
    close(out, G["out"]["features"], 1e-5, "trilinear golden")


def test_plane_query_golden(golden_dir):
    G = load(golden_dir, "plane_query.pt")
    xyz = G["in"]["xyz"].to(DEV)
    planes = {k: v.to(DEV) for k, v in G["in"]["planes"].items()}
    total = 0
    for k in O.PLANES:
        out = ops().sample_features(xyz, planes={k: planes[k]}, padding=G["in"]["padding"])
        close(out.transpose(1, 2), G["out"][k], 1e-5, f"plane {k}")
        total = total + G["out"][k]
    cl = {k: v.contiguous(memory_format=torch.channels_last) for k, v in planes.items()}
    out = ops().sample_features(xyz, planes=cl, padding=G["in"]["padding"])
    close(out.transpose(1, 2), total, 1e-5, "sum of planes")


def test_map_features_vs_oracle():
    g = S.gen(23)
    dims, R, Cp, C = (12, 10, 6), 16, 8, 16
    vol = torch.randn(2, C, *dims, generator=g)
    valid = torch.rand(2, 1, *dims, generator=g) > 0.3
    vol = vol * valid
    planes = {k: torch.randn(2, Cp, R, R, generator=g) for k in O.PLANES}
    xyz = S.query_points(3000, dims, VS, g, B=2)
    ref = O.map_features(xyz, vol, valid, planes, VS, 0.1)
    out = ops().sample_features(xyz.to(DEV), volume=vol.to(DEV), planes={k: v.to(DEV) for k, v in planes.items()},
                                voxel_size=VS, padding=0.1)
    close(out, ref, 1e-5, "map_features")


# ---------------------------------------------------------------------------------------------
# triplane projection
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("domain", ["unit", "metric"])
def test_planes_golden(golden_dir, domain):
    from gennerf_b200 import dropin
    G = load(golden_dir, f"planes_{domain}.pt")
    i, o = G["in"], G["out"]
    p, c = i["p"].to(DEV), i["c"].to(DEV)
    coord, index = ops().plane_coords(p, i["padding"], i["reso"])
    for k, name in enumerate(O.PLANES):
        assert torch.equal(coord[k].cpu(), o["coord"][name]), "normalised coordinates are bit-exact"
        assert torch.equal(index[k].cpu().unsqueeze(1), o["index"][name]), "cell indices are bit-exact"
        assert torch.equal(dropin.normalize_coordinate(p, i["padding"], name).cpu(), o["coord"][name])
        assert torch.equal(dropin.coordinate2index(coord[k], i["reso"]).cpu(), o["index"][name])
    det, cnt_d = ops().scatter_mean_planes(p, c, i["reso"], i["padding"], "deterministic")
    atm, cnt_a = ops().scatter_mean_planes(p, c, i["reso"], i["padding"], "atomic")
    assert torch.equal(cnt_d, cnt_a)
    for k, name in enumerate(O.PLANES):
        _, cnt_o = O.generate_plane_features(i["p"], i["c"], name, i["reso"], i["padding"], return_count=True)
        assert torch.equal(cnt_d[k].cpu(), cnt_o), "scatter counts are bit-exact"
        assert torch.equal(det[k].cpu(), o["plane_features"][name]), "deterministic mode == CPU summation order"
        close(atm[k], o["plane_features"][name], 1e-5, f"atomic scatter {name}")
    close(ops().pool_local(p, c, i["reso"], i["padding"], "max"), o["pool_local_max"], 0.0, "pool max")
    close(ops().pool_local(p, c, i["reso"], i["padding"], "mean"), o["pool_local_mean"], 1e-5, "pool mean")


@pytest.mark.parametrize("N,Cp,R,domain", [(1, 4, 8, "unit"), (4096, 32, 256, "unit"), (70001, 32, 128, "metric"),
                                           (5000, 48, 16, "unit"), (3000, 7, 40, "metric")])
def test_scatter_vs_oracle(N, Cp, R, domain):
    g = S.gen(33)
    p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48), B=2)
    c = torch.randn(2, N, Cp, generator=g)
    det, cnt = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, "deterministic")
    atm, cnt_a = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, "atomic")
    for k, name in enumerate(O.PLANES):
        ref, cnt_o = O.generate_plane_features(p, c, name, R, 0.1, return_count=True)
        assert torch.equal(cnt[k].cpu(), cnt_o) and torch.equal(cnt_a[k].cpu(), cnt_o)
        assert torch.equal(det[k].cpu(), ref)
        close(atm[k], ref, 2e-5, f"atomic {name}")
    assert int(cnt.sum()) == 3 * 2 * N                                     # checksum: every point lands once per plane


def test_pointnet_module_vs_oracle():
    """LocalPoolPointnet.forward with the reference's structure (per-point MLP in PyTorch,
    pooling + scatter on the kernels) against the same network evaluated with oracle ops."""
    from gennerf_b200.dropin import LocalPoolPointnet
    g = S.gen(35)
    torch.manual_seed(0)
    pn = LocalPoolPointnet(c_dim=8, dim=3, hidden_dim=16, scatter_type="max", plane_resolution=32,
                           plane_type=["xz", "xy", "yz"], padding=0.1, n_blocks=3, scatter_mode="deterministic")
    for blk in pn.blocks:
        torch.nn.init.normal_(blk.fc_1.weight, std=0.1)
    p = S.plane_points(2000, g, "unit", B=2)
    with torch.no_grad():
        net = pn.fc_pos(p)
        net = pn.blocks[0](net)
        for block in pn.blocks[1:]:
            net = block(torch.cat([net, O.pool_local(p, net, 32, 0.1, scatter_type="max")], dim=2))
        c = pn.fc_c(net)
        ref = {k: O.generate_plane_features(p, c, k, 32, 0.1) for k in O.PLANES}
        out = pn.to(DEV)(p.to(DEV))
    assert list(out.keys()) == ["xz", "xy", "yz"]
    for k in O.PLANES:
        assert out[k].shape == ref[k].shape
        close(out[k], ref[k], 1e-4, f"pointnet plane {k}")                  # cuBLAS vs MKL linears in between


# ---------------------------------------------------------------------------------------------
# decoder (fp32 exact mode)
# ---------------------------------------------------------------------------------------------
def test_decoder_fp32_golden(golden_dir):
    G = load(golden_dir, "decoder.pt")
    i, o = G["in"], G["out"]
    close(ops().positional_encoding(i["pts"].to(DEV), 2, 0.5, True), o["code"], 1e-5, "posenc")
    close(ops().positional_encoding(i["pts"].to(DEV), 6, 1.5, True), o["code_nf6_ff1.5"], 1e-5, "posenc nf6")
    dw = ops().DecoderWeights(i["weights"], i["head_w"], i["head_b"], n_blocks=i["n_blocks"], d_geo=i["d_geo"],
                              use_code=True, num_freqs=2, freq_factor=0.5, include_input=True, device=DEV)
    out, tsdf = ops().decode(dw, i["pts"].to(DEV), i["feat"].to(DEV), "fp32")
    close(out, o["mlp"], 1e-5, "ResnetFC fp32")
    close(tsdf, o["tsdf"], 1e-5, "tsdf fp32")
    close(ops().tsdf_head(out[:, :i["d_geo"]], i["head_w"].to(DEV), i["head_b"].to(DEV)), o["tsdf"], 1e-5, "head")


@pytest.mark.parametrize("d_hidden,nf,d_feat,d_out,d_geo", [(512, 2, 32, 64, 32), (256, 6, 64, 65, 64), (64, 2, 24, 16, 8)])
def test_decoder_fp32_vs_oracle(d_hidden, nf, d_feat, d_out, d_geo):
    g = S.gen(41)
    d_code = 3 + 6 * nf
    w, hw, hb = S.decoder_weights(g, d_feat, d_code, d_hidden, 5, d_out, d_geo, alpha=0.8)
    xyz = S.query_points(777, (96, 96, 48), VS, g)[0]
    feat = torch.randn(777, d_feat, generator=g)
    code = O.positional_encoding(xyz, nf, 0.5, True)
    ref = O.resnetfc_forward(torch.cat((code, feat), -1), w, 5, d_code)
    ref_t = O.tsdf_head(ref[..., :d_geo], hw, hb)
    dw = ops().DecoderWeights(w, hw, hb, n_blocks=5, d_geo=d_geo, use_code=True, num_freqs=nf, freq_factor=0.5, device=DEV)
    out, tsdf = ops().decode(dw, xyz.to(DEV), feat.to(DEV), "fp32")
    close(out, ref, 2e-5, "ResnetFC fp32")
    close(tsdf, ref_t, 2e-5, "tsdf fp32")
    # stand-alone ResnetFC module (zx = [code | feat], trap T8)
    from gennerf_b200.dropin import ResnetFC
    m = ResnetFC(d_in=d_feat, d_out=d_out, n_blocks=5, d_latent=d_code, d_hidden=d_hidden)
    m.load_state_dict(w)
    y = m.to(DEV)(torch.cat((code, feat), -1).to(DEV))
    close(y, ref, 2e-5, "ResnetFC module")


# ---------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------
def _gennerf_from_golden(G, precision, fused):
    from gennerf_b200.dropin import GenNerf
    from oracle.ref_shim import to_attr
    i = G["in"]
    cfg = to_attr({
        "voxel_size": i["voxel_size"], "voxel_dim_train": list(i["voxel_dim"]), "voxel_dim_val": list(i["voxel_dim"]),
        "encoder": {"use_spatial": True, "spatial": {"num_layers": 1}, "use_pointnet": True, "use_auxiliary": False,
                    "pointnet": {"num_sparse_points": 512, "c_dim": 8, "dim": 3, "padding": i["padding"], "hidden_dim": 32,
                                 "scatter_type": "max", "plane_type": ["xz", "xy", "yz"], "plane_resolution": 16,
                                 "n_blocks": 5, "unet": False, "unet_kwargs": None, "sample_mode": "bilinear"},
                    "plane_merger": {"strategy": "average", "alpha": 0.1}},
        "mlp": {"d_out_sem": 32, "d_out_geo": 32, "n_blocks": 5, "d_hidden": 64, "combine_layer": 1000,
                "combine_type": "average", "beta": 0.0, "use_spade": False, "use_layer_norm": False, "alpha": 1.0},
        "use_code": True, "code": {"num_freqs": i["num_freqs"], "freq_factor": i["freq_factor"], "include_input": True},
    })
    model = GenNerf(cfg, precision=precision, fused=fused).eval()
    model.mlp.load_state_dict(i["weights"])
    model.head_geo.load_state_dict({"fc.weight": i["head_w"], "fc.bias": i["head_b"]})
    return model.to(DEV)


def test_gennerf_forward_fp32_golden(golden_dir):
    G = load(golden_dir, "gennerf_forward.pt")
    i, o = G["in"], G["out"]
    model = _gennerf_from_golden(G, "fp32", False)
    T = i["projection"].shape[1]
    image = i["features"].view(1, T, *i["features"].shape[1:]).to(DEV)
    cfg_pn = model.cfg.encoder.use_pointnet
    model.cfg.encoder.use_pointnet = False
    model.encode(i["projection"], image, None, "val")
    model.cfg.encoder.use_pointnet = cfg_pn
    model.c_plane = {k: v.to(DEV) for k, v in i["planes"].items()}
    with torch.no_grad():
        out = model(i["xyz"].to(DEV))
    assert set(out.keys()) == {"feat_geo", "feat_sem", "tsdf", "feat"}
    close(out["feat"], o["feat"], 1e-5, "feat")
    for k in ("feat_geo", "feat_sem", "tsdf"):
        assert out[k].shape == o[k].shape
        close(out[k], o[k], 2e-5, k)


# ---------------------------------------------------------------------------------------------
# front end of the triplane branch
# ---------------------------------------------------------------------------------------------
def test_get_3d_points_and_fps():
    g = S.gen(61)
    H, W, T = 60, 80, 3
    P = S.projections(T, H, W, (24, 24, 12), VS, g, pull_back=0.8)
    depth = S.depth_maps(T, H, W, g)[0]
    ref = O.get_3d_points(depth, P)
    out = ops().get_3d_points(depth.to(DEV), P)
    close(out, ref, 1e-5, "get_3d_points")
    # FPS is bit-identical on identical inputs (same arithmetic, first-maximum tie rule)
    xyz = ref.reshape(T, -1, 3)
    start = torch.tensor([5, 0, H * W - 1])
    s_o, c_o = O.farthest_point_sample(xyz, 64, start)
    s, c = ops().farthest_point_sample(xyz.to(DEV), 64, start.to(DEV))
    assert torch.equal(c.cpu(), c_o) and torch.equal(s.cpu(), s_o)
    # duplicates (exact ties): still the first maximum
    dup = torch.cat([xyz[:, :100], xyz[:, :100]], dim=1)
    s_o, c_o = O.farthest_point_sample(dup, 32, torch.tensor([1, 2, 3]))
    s, c = ops().farthest_point_sample(dup.to(DEV), 32, torch.tensor([1, 2, 3]).to(DEV))
    assert torch.equal(c.cpu(), c_o)


@pytest.mark.parametrize("N,npoint,B", [(100003, 48, 2), (307200, 24, 1), (16385, 40, 3), (2049, 16, 2)])
def test_fps_cluster_kernel_bit_identical(N, npoint, B):
    """The cluster FPS kernel (2..16 CTAs per cloud, slices in shared memory, candidates exchanged through DSMEM) selects
    exactly the reference's indices: ragged slices, empty last slices, ties, several clouds per launch."""
    g = S.gen(63)
    xyz = torch.rand(B, N, 3, generator=g) * 4
    xyz[:, N // 2:N // 2 + 50] = xyz[:, :50]                     # exact duplicates -> equal distances (tie rule)
    start = torch.randint(0, N, (B,), generator=g)
    s_o, c_o = O.farthest_point_sample(xyz, npoint, start)
    s, c = ops().farthest_point_sample(xyz.to(DEV), npoint, start.to(DEV))
    assert torch.equal(c.cpu(), c_o) and torch.equal(s.cpu(), s_o)
    # the single-CTA kernel (clouds too large for a cluster fall back to it) agrees as well
    from gennerf_b200 import _lib
    old = _lib.set_option("GNB_FPS_SINGLE_CTA", 1)
    try:
        s1, c1 = ops().farthest_point_sample(xyz.to(DEV), npoint, start.to(DEV))
    finally:
        _lib.set_option("GNB_FPS_SINGLE_CTA", old)
    assert torch.equal(c1.cpu(), c_o) and torch.equal(s1.cpu(), s_o)
    # ... and both multi-CTA kernels when forced: thread-block clusters (DSMEM exchange) and the co-operative grid kernel
    # (CTAs anywhere on the GPU, exchange through L2), with CTA counts that leave ragged / empty last slices
    for name, val in (("GNB_FPS_GRID", -1), ("GNB_FPS_GRID", 5), ("GNB_FPS_GRID", 18), ("GNB_FPS_GRID", 32)):
        if val > 0 and (B * val > 148 or -(-N // val) > 20 * 1024 or -(-N // val) * 12 > 225 * 1024):
            continue
        old = _lib.set_option(name, val)
        try:
            s2, c2 = ops().farthest_point_sample(xyz.to(DEV), npoint, start.to(DEV))
        finally:
            _lib.set_option(name, old)
        assert torch.equal(c2.cpu(), c_o) and torch.equal(s2.cpu(), s_o), (name, val)

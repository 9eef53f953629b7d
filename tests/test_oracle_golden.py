"""The CPU oracle against the committed golden vectors (made from the real reference by
tests/golden/make_golden.py).  Runs everywhere, including the GPU box, where it proves that
this host's ATen CPU kernels reproduce the build container's reference outputs."""
import os

import pytest
import torch

from oracle import gennerf_oracle as O

ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_backproject(golden_dir, name):
    G = load(golden_dir, f"backproject_{name}.pt")
    i, o = G["in"], G["out"]
    T = i["projection"].shape[0]
    feats = [i["features"][t:t + 1] for t in range(T)]
    vol, valid, count = O.encode_volume(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"].unsqueeze(0), feats)
    assert torch.equal(vol, o["volume_sum"]) and torch.equal(valid, o["valid_or"])
    assert torch.equal(count.view(-1), o["valid_per_frame"].view(T, -1).sum(0).to(torch.int32))
    v0, m0 = O.backproject(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"][0:1], feats[0])
    assert torch.equal(v0, o["frame0_volume"]) and torch.equal(m0, o["frame0_valid"])
    H, W = i["features"].shape[-2:]
    a = O.project_indices(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"], H, W)
    b = O.project_indices_explicit(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"], H, W)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(a[3].view(T, -1), o["valid_per_frame"].view(T, -1))


def test_trilinear(golden_dir):
    G = load(golden_dir, "trilinear.pt")
    vol = G["in"]["volume_ncxyz"].permute(0, 2, 3, 4, 1)
    out = O.trilinear_interpolation(vol, G["in"]["xyz"], ORIGIN.squeeze(), G["in"]["voxel_size"])
    assert torch.equal(out, G["out"]["features"])
    out2 = O.trilinear_interpolation_explicit(vol, G["in"]["xyz"], ORIGIN.squeeze(), G["in"]["voxel_size"])
    assert torch.allclose(out2, G["out"]["features"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("domain", ["unit", "metric"])
def test_planes(golden_dir, domain):
    G = load(golden_dir, f"planes_{domain}.pt")
    i, o = G["in"], G["out"]
    for k in O.PLANES:
        xy = O.normalize_coordinate(i["p"].clone(), i["padding"], k)
        assert torch.equal(xy, o["coord"][k])
        assert torch.equal(O.coordinate2index(xy, i["reso"]), o["index"][k])
        assert torch.equal(O.generate_plane_features(i["p"], i["c"], k, i["reso"], i["padding"]), o["plane_features"][k])
    assert torch.equal(O.pool_local(i["p"], i["c"], i["reso"], i["padding"], scatter_type="max"), o["pool_local_max"])
    assert torch.equal(O.pool_local(i["p"], i["c"], i["reso"], i["padding"], scatter_type="mean"), o["pool_local_mean"])


def test_plane_query(golden_dir):
    G = load(golden_dir, "plane_query.pt")
    for k in O.PLANES:
        out = O.sample_plane_feature(G["in"]["xyz"], G["in"]["planes"][k], k, G["in"]["padding"])
        assert torch.equal(out, G["out"][k])
        out2 = O.sample_plane_feature_explicit(G["in"]["xyz"], G["in"]["planes"][k], k, G["in"]["padding"])
        assert torch.allclose(out2, G["out"][k], rtol=1e-5, atol=1e-6)


def test_decoder(golden_dir):
    G = load(golden_dir, "decoder.pt")
    i, o = G["in"], G["out"]
    code = O.positional_encoding(i["pts"], 2, 0.5, True)
    assert torch.equal(code, o["code"])
    assert torch.equal(O.positional_encoding(i["pts"], 6, 1.5, True), o["code_nf6_ff1.5"])
    y = O.resnetfc_forward(torch.cat((code, i["feat"]), -1), i["weights"], i["n_blocks"], i["d_code"])
    # nn.Linear goes through the BLAS the host CPU selects; allow last-bit differences here
    assert torch.allclose(y, o["mlp"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(O.tsdf_head(y[..., :i["d_geo"]], i["head_w"], i["head_b"]), o["tsdf"], rtol=1e-5, atol=1e-6)


def test_gennerf_forward(golden_dir):
    G = load(golden_dir, "gennerf_forward.pt")
    i, o = G["in"], G["out"]
    T = i["projection"].shape[1]
    feats = [i["features"][t:t + 1] for t in range(T)]
    vol, valid, _ = O.encode_volume(i["voxel_dim"], i["voxel_size"], ORIGIN, i["projection"], feats)
    out = O.gennerf_forward(i["xyz"], i["weights"], i["head_w"], i["head_b"], volume=vol, valid=valid,
                            planes=i["planes"], voxel_size=i["voxel_size"], padding=i["padding"],
                            num_freqs=i["num_freqs"], freq_factor=i["freq_factor"])
    assert torch.equal(out["feat"], o["feat"])
    for k in ("feat_geo", "feat_sem", "tsdf"):
        assert torch.allclose(out[k], o[k], rtol=1e-5, atol=1e-5), k


def test_tsdf_fusion(golden_dir):
    """the oracle's TSDFFusion and its per-voxel formulation against the real reference's volumes."""
    G = load(golden_dir, "tsdf_fusion.pt")
    i, o = G["in"], G["out"]
    T = i["projection"].shape[0]
    f = O.TSDFFusion(i["voxel_dim"], i["voxel_size"], i["origin"], trunc_ratio=i["trunc_ratio"], color=True, label=True)
    for t in range(T):
        f.integrate(i["projection"][t], i["depth"][t], i["color"][t], i["label"][t].long())
        if t == 0:
            assert torch.equal(f.tsdf_vol, o["frame0"]["tsdf_vol"]) and torch.equal(f.weight_vol, o["frame0"]["weight_vol"].float())
    assert torch.equal(f.tsdf_vol, o["all"]["tsdf_vol"]) and torch.equal(f.weight_vol, o["all"]["weight_vol"].float())
    assert torch.equal(f.color_vol, o["all"]["color_vol"]) and torch.equal(f.label_vol, o["all"]["label_vol"].long())
    tsdf, color, _ = f.get_volumes()
    assert torch.equal(tsdf.reshape(-1), o["normalised"]["tsdf"]) and torch.equal(color.reshape(3, -1), o["normalised"]["color"])
    e = O.tsdf_fusion_explicit(i["voxel_dim"], i["voxel_size"], i["origin"], i["trunc_ratio"], i["projection"], i["depth"],
                               i["color"], i["label"].long())
    assert torch.equal(e[0], o["all"]["tsdf_vol"]) and torch.equal(e[1], o["all"]["weight_vol"].float())
    assert torch.equal(e[2], o["all"]["color_vol"]) and torch.equal(e[3], o["all"]["label_vol"].long())


def test_sample_points_on_rays(golden_dir):
    """the oracle's ray sampler against the real reference's output (1e-6: ATen's CPU linspace is
    vector-width dependent in the last bit, see the oracle's docstring)."""
    G = load(golden_dir, "ray_points.pt")
    i, o = G["in"], G["out"]
    xyz, z = O.sample_points_on_rays(i["h_idxs"].long(), i["w_idxs"].long(), i["depths"], i["intrinsics"], i["poses"], i["N"], i["M"],
                                     i["delta"], i["min_dist"], i["gaussian_depths"])
    assert ((z - o["z"]).abs() <= 1e-6 * o["z"].abs().clamp_min(1.0)).all()
    assert ((xyz - o["xyz_world"]).abs() <= 1e-6 * o["xyz_world"].abs().clamp_min(1.0)).all()

"""Edge cases the reference's domain has: empty and ragged inputs, frames that see nothing, points exactly on
and far outside the grid, single channels, several scenes per call."""
import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS = 0.04
ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)


def ops():
    from gennerf_b200 import ops as _ops
    return _ops


def test_lift_frames_that_see_nothing_and_single_channel():
    vd = (10, 9, 7)
    g = S.gen(71)
    P = S.projections(2, 24, 32, vd, VS, g, pull_back=0.8)
    P[1, 0, 3] += 1e7                                            # second camera: every px lands far outside the image
    behind = P[0].clone()
    behind[2] = -behind[2]                                      # every voxel behind this camera (pz <= 0)
    Pb = torch.stack([P[0], P[1], behind]).unsqueeze(0)
    feats = S.frame_features(3, 1, 24, 32, g)
    vol_o, valid_o, cnt_o = O.encode_volume(vd, VS, ORIGIN, Pb, feats)
    vol, cnt, valid = ops().backproject_frames(vd, VS, ORIGIN, Pb, [f.to(DEV) for f in feats])
    assert torch.equal(vol.cpu(), vol_o) and torch.equal(cnt.cpu(), cnt_o) and torch.equal(valid.cpu(), valid_o)
    only_blind = ops().backproject_frames(vd, VS, ORIGIN, Pb[:, 1:], [f.to(DEV) for f in feats[1:]])
    assert int(only_blind[1].sum()) == 0 and float(only_blind[0].abs().sum()) == 0.0 and not bool(only_blind[2].any())


def test_sampler_empty_single_and_extreme_points():
    g = S.gen(72)
    dims, C, Cp, R = (6, 5, 4), 8, 4, 8
    vol = torch.randn(1, C, *dims, generator=g)
    planes = {k: torch.randn(1, Cp, R, R, generator=g) for k in O.PLANES}
    valid = torch.ones(1, 1, *dims, dtype=torch.bool)
    ext = [d * VS for d in dims]
    xyz = torch.tensor([[[0.0, 0.0, 0.0], [ext[0], ext[1], ext[2]], [ext[0] * (dims[0] - 1) / dims[0], 0.0, ext[2]],
                         [-1e6, 1e6, 0.0], [1e-30, -1e-30, 5e5], [0.55, -0.55, 0.5500001], [ext[0] / 2, ext[1] / 2, ext[2] / 2]]])
    for n in (0, 1, xyz.shape[1]):
        q = xyz[:, :n].contiguous()
        out = ops().sample_features(q.to(DEV), volume=vol.to(DEV), planes={k: v.to(DEV) for k, v in planes.items()},
                                    voxel_size=VS, padding=0.1)
        assert out.shape == (1, n, Cp + C)
        if n:
            ref = O.map_features(q, vol, valid, planes, VS, 0.1)
            assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())


def test_scatter_and_pool_empty_and_one_point():
    g = S.gen(73)
    R, Cp = 8, 4
    for n in (0, 1):
        p = S.plane_points(n, g, "unit")
        c = torch.randn(1, n, Cp, generator=g)
        for mode in ("atomic", "deterministic"):
            planes, cnt = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, mode)
            assert int(cnt.sum()) == 3 * n
            for k, name in enumerate(O.PLANES):
                assert torch.equal(planes[k].cpu(), O.generate_plane_features(p, c, name, R, 0.1))
        pooled = ops().pool_local(p.to(DEV), c.to(DEV), R, 0.1, "max")
        assert torch.equal(pooled.cpu(), O.pool_local(p, c, R, 0.1, scatter_type="max")) if n else pooled.numel() == 0


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129])
def test_decoder_ragged_row_counts(n):
    g = S.gen(74)
    w, hw, hb = S.decoder_weights(g, 32, 15, 256, 5, 64, 32)
    dw = ops().DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=DEV)
    xyz = S.query_points(max(n, 1), (96, 96, 48), VS, g)[0][:n]
    feat = torch.randn(n, 32, generator=g)
    sentinel = torch.full((n + 4, 64), 7.0, device=DEV)
    for precision in ("fp32", "fp16"):
        out, tsdf = ops().decode(dw, xyz.to(DEV), feat.to(DEV), precision)
        assert out.shape == (n, 64) and tsdf.shape == (n, 1)
        if n:
            code = O.positional_encoding(xyz, 2, 0.5, True)
            ref = O.resnetfc_forward(torch.cat((code, feat), -1), w, 5, 15)
            tol = 2e-5 if precision == "fp32" else 4e-3
            assert ((out.cpu() - ref).abs().max() / ref.abs().max()).item() < tol
    assert float(sentinel.sum()) == 7.0 * sentinel.numel()


def test_fused_query_two_scenes():
    g = S.gen(75)
    wl = S.WORKLOADS["tiny"]
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
    Pb = torch.stack([P, P.flip(0)])
    feats = S.frame_features(wl["T"], 32, wl["H"], wl["W"], g, B=2)
    xyz = S.query_points(300, wl["voxel_dim"], VS, g, B=2)
    w, hw, hb = S.decoder_weights(g, 32, 15, 128, 5, 64, 32)
    vol, cnt, valid = ops().backproject_frames(wl["voxel_dim"], VS, ORIGIN, Pb, [f.to(DEV) for f in feats])
    dw = ops().DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=DEV)
    out, tsdf, feat = ops().query_fused(dw, xyz.to(DEV), volume=vol, voxel_size=VS, origin=ORIGIN)
    vol_o, valid_o, _ = O.encode_volume(wl["voxel_dim"], VS, ORIGIN, Pb, feats)
    ref = O.gennerf_forward(xyz, w, hw, hb, volume=vol_o, valid=valid_o, voxel_size=VS)
    assert ((feat.cpu() - ref["feat"]).abs().max() / ref["feat"].abs().max()).item() <= 1e-5
    assert (tsdf.cpu() - ref["tsdf"]).abs().max().item() <= 1e-2

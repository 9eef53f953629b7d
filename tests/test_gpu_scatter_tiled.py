"""Tiled scatter-mean (gnb_scatter_mean_planes_tiled) against the oracle's restatement of
LocalPoolPointnet.generate_plane_features (reference src/models/components/pointnet.py:72-89 = torch_scatter.scatter_mean):
counts bit-exact, means within the atomic mode's 2e-5 of the tensor scale (summation order differs), and equal to the
deterministic mode's counts; split tiles (more points than one work unit), ragged tiles (R not a multiple of the tile
side), several scenes, channel counts other than 32."""
import pytest
import torch

from gennerf_b200 import synthetic as S
from oracle import gennerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops():
    from gennerf_b200 import ops as _ops
    return _ops


def close(a, b, tol, what):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    err = (a - b).abs().max().item()
    scale = max(b.abs().max().item(), 1e-30)
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("N,Cp,R,domain,B", [(1, 4, 8, "unit", 1), (4096, 32, 256, "unit", 1), (70001, 32, 128, "metric", 2),
                                             (5000, 48, 16, "unit", 2), (3000, 7, 40, "metric", 1), (40000, 64, 50, "metric", 1),
                                             (20000, 130, 33, "unit", 1)])
def test_tiled_vs_oracle(N, Cp, R, domain, B):
    g = S.gen(330 + Cp)
    p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48), B=B)
    c = torch.randn(B, N, Cp, generator=g)
    til, cnt = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, "tiled")
    det, cnt_d = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, "deterministic")
    assert torch.equal(cnt, cnt_d)
    for k, name in enumerate(O.PLANES):
        ref, cnt_o = O.generate_plane_features(p, c, name, R, 0.1, return_count=True)
        assert torch.equal(cnt[k].cpu(), cnt_o), "scatter counts are bit-exact"
        close(til[k], ref, 2e-5, f"tiled {name}")
    assert int(cnt.sum()) == 3 * B * N                                     # checksum: every point lands once per plane


def test_tiled_all_points_in_one_cell_and_empty_planes():
    """every point in ONE cell: one tile split into many parts, every other tile empty (written as zeros)"""
    g = S.gen(331)
    N, Cp, R = 50000, 32, 64
    p = (torch.rand(1, N, 3, generator=g) * 1e-4 + 0.2)
    c = torch.randn(1, N, Cp, generator=g)
    planes, cnt = ops().scatter_mean_planes(p.to(DEV), c.to(DEV), R, 0.1, "tiled")
    assert int((cnt > 0).sum()) == 3 and int(cnt.max()) == N
    mean = c[0].double().mean(0).float()
    for k in range(3):
        cell = (cnt[k, 0] > 0).nonzero()[0]
        close(planes[k, 0, :, cell[0], cell[1]], mean, 1e-4, "mean of all points")
        assert float(planes[k].abs().sum()) == pytest.approx(float(planes[k, 0, :, cell[0], cell[1]].abs().sum()))


def test_tiled_config3_full_size_matches_atomic():
    """BASELINE config 3, all pixels: 8 frames x 240x320 points onto 3 x 256^2 planes x 32 channels"""
    g = S.gen(332)
    N, Cp, R = 614400, 32, 256
    for domain in ("unit", "metric"):
        p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48)).to(DEV)
        c = torch.randn(1, N, Cp, generator=g).to(DEV)
        til, cnt = ops().scatter_mean_planes(p, c, R, 0.1, "tiled")
        aut, cnt_u = ops().scatter_mean_planes(p, c, R, 0.1, "auto")
        assert torch.equal(cnt, cnt_u)
        atm, cnt_a = ops().scatter_mean_planes(p, c, R, 0.1, "atomic")
        assert torch.equal(cnt, cnt_a) and int(cnt.sum()) == 3 * N
        close(til, atm, 5e-5, f"tiled vs atomic ({domain})")

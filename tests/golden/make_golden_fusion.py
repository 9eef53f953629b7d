"""Golden vectors of the TSDF fusion path and of the training-time ray sampler, generated FROM THE REAL REFERENCE
(src/data/tsdf.py TSDFFusion, src/models/utils.py sample_points_on_rays, on the CPU).

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden_fusion.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gennerf_b200 import synthetic as S          # noqa: E402
from oracle import ref_shim                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fusion_inputs(seed=301):
    g = S.gen(seed)
    vd, vs, org, trunc_ratio = (48, 40, 24), 0.04, (0.1, -0.2, 0.05), 3
    T, H, W = 6, 60, 80
    P = S.projections(T, H, W, vd, vs, g)
    depths = S.surface_depth_maps(T, H, W, g)
    colors = torch.rand(T, 3, H, W, generator=g)
    labels = torch.randint(0, 40, (T, H, W), generator=g)
    return {"voxel_dim": vd, "voxel_size": vs, "origin": org, "trunc_ratio": trunc_ratio, "projection": P,
            "depth": depths, "color": colors, "label": labels.to(torch.int16)}


def main():
    ref_shim.install()
    from src.data.tsdf import TSDFFusion
    torch.set_grad_enabled(False)
    i = fusion_inputs()
    f = TSDFFusion(i["voxel_dim"], i["voxel_size"], i["origin"], trunc_ratio=i["trunc_ratio"], device=torch.device("cpu"),
                   color=True, label=True)
    out = {}
    T = i["projection"].shape[0]
    for t in range(T):
        f.integrate(i["projection"][t], i["depth"][t], i["color"][t], i["label"][t].long())
        if t == 0:
            out["frame0"] = {"tsdf_vol": f.tsdf_vol.clone(), "weight_vol": f.weight_vol.to(torch.int16)}
    out["all"] = {"tsdf_vol": f.tsdf_vol.clone(), "weight_vol": f.weight_vol.to(torch.int16), "color_vol": f.color_vol.clone(),
                  "label_vol": f.label_vol.to(torch.int16)}
    # get_tsdf's normalisation (tsdf.py:426-434), without building the TSDF container object
    seen = f.weight_vol > 0
    tsdf = f.tsdf_vol.clone()
    tsdf[seen] /= f.weight_vol[seen]
    color = f.color_vol.clone()
    color[:, seen] /= f.weight_vol[seen]
    out["normalised"] = {"tsdf": tsdf, "color": color}
    path = os.path.join(HERE, "tsdf_fusion.pt")
    torch.save({"in": i, "out": out}, path)
    print(f"tsdf_fusion.pt: {os.path.getsize(path) / 1024:.0f} KiB; seen voxels {int(seen.sum())} of {seen.numel()}")




def rays_main():
    """sample_points_on_rays (src/models/utils.py:458-540) on the CPU reference; the gaussian draw is replayed from the seed."""
    ref_shim.install()
    from src.models.utils import sample_points_on_rays
    g = S.gen(311)
    B, Sn, N, M, H, W = 3, 100, 20, 8, 240, 320
    h = torch.randint(0, H, (B, Sn), generator=g)
    w = torch.randint(0, W, (B, Sn), generator=g)
    D = torch.rand(B, Sn, generator=g) * 3 + 0.4
    K = S.intrinsics(H, W).expand(B, 3, 3).contiguous()
    poses = torch.stack(list(S.camera_poses(B, (96, 96, 48), 0.04, g)))
    torch.manual_seed(11)
    xyz, z = sample_points_on_rays(h, w, D, K, poses, N=N, M=M, delta=0.1, min_dist=0.07, sigma=0.1)
    torch.manual_seed(11)
    gd = torch.stack([torch.normal(D[b].unsqueeze(-1).expand(Sn, M), 0.1 * torch.ones(Sn, M)) for b in range(B)])
    assert torch.equal(z[..., 1 + N:], gd)
    path = os.path.join(HERE, "ray_points.pt")
    torch.save({"in": {"h_idxs": h.to(torch.int16), "w_idxs": w.to(torch.int16), "depths": D, "intrinsics": K, "poses": poses, "N": N, "M": M,
                       "delta": 0.1, "min_dist": 0.07, "sigma": 0.1, "gaussian_depths": gd},
                "out": {"xyz_world": xyz, "z": z}}, path)
    print(f"ray_points.pt: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
    rays_main()

"""Golden vectors of the TSDF fusion path, generated FROM THE REAL REFERENCE (src/data/tsdf.py TSDFFusion on the CPU).

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


if __name__ == "__main__":
    main()

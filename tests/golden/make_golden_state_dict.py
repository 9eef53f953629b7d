"""Golden state_dict of the REAL reference GenNerf (src/models/model.py:25-75), for the drop-in boundary test.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden_state_dict.py
Writes state_dict.pt:
  "default": the parsed configs/model/gen_nerf.yaml (unmodified: pointnet.unet True, d_hidden 512) as a plain dict and
             {key: shape} of GenNerf(cfg).state_dict() -- every checkpoint key the drop-in has to accept;
  "small":   a down-sized config (no U-Net) with the reference model's actual state_dict tensors, and the output of
             its forward() on seeded inputs, so a strict load into the drop-in can be checked numerically.
The 2D CNN (SpatialEncoder) is outside the path and downloads weights; it is replaced by a parameter-free stub
(oracle/ref_shim.py), so no `spatial.*` keys appear.
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gennerf_b200 import synthetic as S          # noqa: E402
from oracle import ref_shim                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VS = 0.04
PLANES = ("xz", "xy", "yz")


def plain(d):
    if isinstance(d, dict):
        return {k: plain(v) for k, v in d.items()}
    if isinstance(d, list):
        return [plain(v) for v in d]
    return d


def main():
    GenNerf = ref_shim.ref_gennerf()
    torch.manual_seed(1234)
    cfg = ref_shim.load_model_cfg("gen_nerf")
    model = GenNerf(cfg)
    default = {"cfg": plain(cfg), "keys": {k: tuple(v.shape) for k, v in model.state_dict().items()}}

    wl = S.WORKLOADS["e2e"]
    small = ref_shim.load_model_cfg("gen_nerf", voxel_dim_train=list(wl["voxel_dim"]), voxel_dim_val=list(wl["voxel_dim"]),
                                    voxel_size=VS)
    small.encoder.spatial.num_layers = 1
    small.encoder.pointnet.unet = False
    small.encoder.pointnet.plane_resolution = 16
    small.encoder.pointnet.c_dim = 8
    small.encoder.pointnet.hidden_dim = 16
    small.mlp.d_hidden = 64
    m = GenNerf(copy.deepcopy(small)).eval()
    g = S.gen(109)
    with torch.no_grad():
        for p in m.pointnet.parameters():             # (the PointNet's fc_1 are zero-initialised too)
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.05))
        # decoder: fc_1 of every block is zero-initialised (trap T9) -> kaiming-scale random weights, small biases, and a
        # head scaled so that |pre-tanh| stays O(1) (tanh saturation hides errors) -- gennerf_b200.synthetic
        C = 64
        w, hw, hb = S.decoder_weights(g, C + 8, 15, 64, 5, 64, 32, alpha=0.9)
        m.mlp.load_state_dict(w)
        m.head_geo.load_state_dict({"fc.weight": hw, "fc.bias": hb})
        xyz = S.query_points(300, wl["voxel_dim"], VS, g)
        m.volume = torch.randn(1, C, *wl["voxel_dim"], generator=g)
        m.valid = torch.rand(1, 1, *wl["voxel_dim"], generator=g) > 0.3
        m.c_plane = {k: torch.randn(1, 8, 16, 16, generator=g) for k in PLANES}
        out = m.forward(xyz)
    obj = {"default": default,
           "small": {"cfg": plain(small), "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
                     "in": {"xyz": xyz, "volume": m.volume, "valid": m.valid, "planes": m.c_plane},
                     "out": {k: out[k] for k in ("feat", "feat_geo", "feat_sem", "tsdf")}}}
    path = os.path.join(HERE, "state_dict.pt")
    torch.save(obj, path)
    print(f"state_dict.pt: {os.path.getsize(path) / 1024:.0f} KiB, {len(default['keys'])} default keys, "
          f"{len(obj['small']['state_dict'])} small keys")


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    main()

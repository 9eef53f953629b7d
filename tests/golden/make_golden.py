"""Generate the golden vectors in this directory FROM THE REAL REFERENCE.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Every tensor under "out" in the .pt files is the output of a reference function
(src/models/utils.py, src/models/components/*.py, src/models/model.py) called on the CPU in
fp32 through oracle/ref_shim.py; "in" holds the exact inputs.  The reference ships no
fixtures of its own, so these are the pinned known answers.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gennerf_b200 import synthetic as S          # noqa: E402
from oracle import ref_shim                      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ORIGIN = torch.tensor([0, 0, 0]).view(1, 3)
VS = 0.04
PLANES = ("xz", "xy", "yz")


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ref = ref_shim.ref_modules()
    GenNerf = ref_shim.ref_gennerf()
    torch.set_grad_enabled(False)

    # ---- back-projection + accumulation (utils.py:948, model.py:121-127) -------------
    for name, C, seed in (("tiny", 4, 101), ("small", 8, 102)):
        wl = S.WORKLOADS[name]
        g = S.gen(seed)
        P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8)
        feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g)
        vol = val = None
        valid_frames = []
        for t in range(wl["T"]):
            v, m = ref.utils.backproject(wl["voxel_dim"], VS, ORIGIN, P[t:t + 1], feats[t])
            valid_frames.append(m)
            vol = v if vol is None else vol + v
            val = m if val is None else val + m
        v0, m0 = ref.utils.backproject(wl["voxel_dim"], VS, ORIGIN, P[0:1], feats[0])
        save(f"backproject_{name}.pt", {
            "in": {"voxel_dim": wl["voxel_dim"], "voxel_size": VS, "projection": P, "features": torch.cat(feats)},
            "out": {"volume_sum": vol, "valid_or": val, "valid_per_frame": torch.cat(valid_frames),
                    "frame0_volume": v0, "frame0_valid": m0}})

    # ---- trilinear query (utils.py:999) --------------------------------------------
    g = S.gen(103)
    dims = (9, 7, 5)
    vol = torch.randn(2, 6, *dims, generator=g)
    xyz = S.query_points(601, dims, VS, g, B=2)
    xyz[0, :4] = torch.tensor([[0.0, 0.0, 0.0], [9 * VS, 7 * VS, 5 * VS], [-1.0, 0.1, 9.0], [0.36, 0.28, 0.2]])
    out = ref.utils.trilinear_interpolation(vol.permute(0, 2, 3, 4, 1), xyz, ORIGIN.squeeze(), VS)
    save("trilinear.pt", {"in": {"volume_ncxyz": vol, "xyz": xyz, "voxel_size": VS}, "out": {"features": out}})

    # ---- plane indices, scatter_mean, pool_local (utils.py:57-98, pointnet.py:72-121) ---
    for domain, seed in (("unit", 104), ("metric", 105)):
        g = S.gen(seed)
        N, Cp, R = 2500, 8, 32
        p = S.plane_points(N, g, domain, voxel_dim=(96, 96, 48), B=2)
        p[0, :7] = torch.tensor([[-0.55, 0.55, 0.0], [0.55, 0.55, 0.55], [-0.6, 0.7, -0.7], [0.0, 0.0, 0.0],
                                 [0.549999, -0.549999, 0.5500001], [1e-9, -1e-9, 3.0], [-3.0, 3.0, 0.1]])
        c = torch.randn(2, N, Cp, generator=g)
        pn = ref.pointnet.LocalPoolPointnet(c_dim=Cp, dim=3, hidden_dim=8, scatter_type="max", unet=False,
                                            plane_resolution=R, plane_type=list(PLANES), padding=0.1, n_blocks=2)
        coord = {k: ref.utils.normalize_coordinate(p.clone(), plane=k, padding=0.1) for k in PLANES}
        index = {k: ref.utils.coordinate2index(coord[k], R) for k in PLANES}
        fea = {k: pn.generate_plane_features(p, c, k) for k in PLANES}
        pooled_max = pn.pool_local(coord, index, c)
        pn.scatter = ref.pointnet.scatter_mean
        pooled_mean = pn.pool_local(coord, index, c)
        save(f"planes_{domain}.pt", {"in": {"p": p, "c": c, "reso": R, "padding": 0.1},
                                     "out": {"coord": coord, "index": index, "plane_features": fea,
                                             "pool_local_max": pooled_max, "pool_local_mean": pooled_mean}})

    # ---- plane query (model.py:153-161) ----------------------------------------------
    g = S.gen(106)
    Cp, R = 8, 16
    planes = {k: torch.randn(2, Cp, R, R, generator=g) for k in PLANES}
    xyz = S.plane_points(700, g, "unit", B=2) * 1.2
    fake = type("F", (), {})()
    fake.cfg = ref_shim.to_attr({"encoder": {"pointnet": {"padding": 0.1, "sample_mode": "bilinear"}},
                                 "loss": {"use_eikonal": False, "use_gradient": False}})
    out = {k: GenNerf.sample_plane_feature(fake, xyz, planes[k], plane=k) for k in PLANES}
    save("plane_query.pt", {"in": {"planes": planes, "xyz": xyz, "padding": 0.1}, "out": out})

    # ---- positional encoding, ResnetFC, TSDF head ----------------------------------
    g = S.gen(107)
    d_hidden, d_code, d_feat, d_out, d_geo = 64, 15, 24, 16, 8
    w, hw, hb = S.decoder_weights(g, d_feat, d_code, d_hidden, 5, d_out, d_geo, alpha=0.7)
    mlp = ref.resnetfc.ResnetFC(d_in=d_feat, d_out=d_out, n_blocks=5, d_latent=d_code, d_hidden=d_hidden)
    mlp.load_state_dict(w)
    head = ref.heads3d.TSDFHeadSimple(d_geo)
    head.load_state_dict({"fc.weight": hw, "fc.bias": hb})
    pe = ref.posenc.PositionalEncoding(num_freqs=2, d_in=3, freq_factor=0.5, include_input=True)
    pts = torch.randn(300, 3, generator=g) * 2
    code = pe(pts)
    feat = torch.randn(300, d_feat, generator=g)
    y = mlp(torch.cat((code, feat), -1))
    pe6 = ref.posenc.PositionalEncoding(num_freqs=6, d_in=3, freq_factor=1.5, include_input=True)
    save("decoder.pt", {"in": {"weights": w, "head_w": hw, "head_b": hb, "pts": pts, "feat": feat,
                               "n_blocks": 5, "d_code": d_code, "d_geo": d_geo},
                        "out": {"code": code, "code_nf6_ff1.5": pe6(pts), "mlp": y, "tsdf": head(y[..., :d_geo])}})

    # ---- whole GenNerf: encode (volume branch) + forward, and with planes ------------
    wl = S.WORKLOADS["e2e"]
    g = S.gen(108)
    C = 64
    cfg = ref_shim.load_model_cfg("gen_nerf", voxel_dim_train=list(wl["voxel_dim"]),
                                  voxel_dim_val=list(wl["voxel_dim"]), voxel_size=VS)
    cfg.encoder.spatial.num_layers = 1
    cfg.encoder.pointnet.unet = False
    cfg.encoder.pointnet.plane_resolution = 16
    cfg.encoder.pointnet.c_dim = 8
    cfg.mlp.d_hidden = 64
    model = GenNerf(cfg).eval()
    w, hw, hb = S.decoder_weights(g, C + 8, 15, 64, 5, 64, 32)
    model.mlp.load_state_dict(w)
    model.head_geo.load_state_dict({"fc.weight": hw, "fc.bias": hb})
    P = S.projections(wl["T"], wl["H"], wl["W"], wl["voxel_dim"], VS, g, pull_back=0.8).unsqueeze(0)
    feats = S.frame_features(wl["T"], C, wl["H"], wl["W"], g)
    xyz = S.query_points(400, wl["voxel_dim"], VS, g)
    planes = {k: torch.randn(1, 8, 16, 16, generator=g) for k in PLANES}
    # encode: volume branch through the reference loop; planes injected (the FPS front end
    # draws torch.randint internally and is a "next" row)
    vol = val = None
    for t in range(wl["T"]):
        v, m = ref.utils.backproject(wl["voxel_dim"], VS, ORIGIN, P[:, t], feats[t])
        vol = v if vol is None else vol + v
        val = m if val is None else val + m
    model.volume, model.valid, model.c_plane = vol, val, planes
    out = model.forward(xyz)
    save("gennerf_forward.pt", {
        "in": {"voxel_dim": wl["voxel_dim"], "voxel_size": VS, "projection": P, "features": torch.cat(feats),
               "planes": planes, "xyz": xyz, "weights": w, "head_w": hw, "head_b": hb,
               "num_freqs": cfg.code.num_freqs, "freq_factor": cfg.code.freq_factor, "padding": 0.1},
        "out": {k: out[k] for k in ("feat", "feat_geo", "feat_sem", "tsdf")}})


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    main()

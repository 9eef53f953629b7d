"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything is generated on the CPU with a torch.Generator so that the oracle, the golden
vectors and the CUDA path see identical bits; callers copy to the GPU afterwards.
Host logic only -- no kernels, no oracle.
"""
import math

import torch

#              name      T   H    W    grid            Q          R    N_points
WORKLOADS = {
    "cfg1": dict(T=8, H=240, W=320, voxel_dim=(96, 96, 48), Q=65536, R=256),
    "cfg2": dict(T=8, H=240, W=320, voxel_dim=(96, 96, 48), Q=1 << 20, R=256),
    "cfg3": dict(T=8, H=240, W=320, voxel_dim=(96, 96, 48), Q=1 << 20, R=256),
    "cfg4": dict(T=32, H=480, W=640, voxel_dim=(256, 256, 96), Q=1 << 24, R=256),
    "cfg5": dict(T=8, H=480, W=640, voxel_dim=(160, 160, 64), Q=2900 * 8, R=128),
    "tiny": dict(T=3, H=24, W=32, voxel_dim=(12, 10, 6), Q=257, R=16),
    "small": dict(T=4, H=60, W=80, voxel_dim=(24, 24, 12), Q=4099, R=64),
    "e2e": dict(T=4, H=30, W=40, voxel_dim=(16, 16, 8), Q=400, R=16),
}


def gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def intrinsics(H, W):
    """ScanNet-like pinhole: f = 0.9 W (577/640), principal point at the image centre."""
    return torch.tensor([[0.9 * W, 0.0, (W - 1) / 2.0],
                         [0.0, 0.9 * W, (H - 1) / 2.0],
                         [0.0, 0.0, 1.0]], dtype=torch.float32)


def camera_poses(T, voxel_dim, voxel_size, g, pull_back=0.0):
    """cam->world poses (T,4,4): Rz(U(0,2pi)) Rx(U(0,2pi)), centre = volume centre + N(0,0.3^2) m,
    moved `pull_back` metres against the viewing axis (small test grids: so frames see them)."""
    az = torch.rand(T, generator=g) * 2 * math.pi
    ax = torch.rand(T, generator=g) * 2 * math.pi
    centre = torch.tensor([d * voxel_size / 2.0 for d in voxel_dim], dtype=torch.float32)
    trans = centre + 0.3 * torch.randn(T, 3, generator=g)
    poses = torch.zeros(T, 4, 4)
    for t in range(T):
        cz, sz = math.cos(az[t]), math.sin(az[t])
        cx, sx = math.cos(ax[t]), math.sin(ax[t])
        Rz = torch.tensor([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
        Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
        poses[t, :3, :3] = Rz @ Rx
        poses[t, :3, 3] = trans[t] - pull_back * poses[t, :3, 2]
        poses[t, 3, 3] = 1.0
    return poses


def projections(T, H, W, voxel_dim, voxel_size, g, pull_back=0.0):
    """world->pixel (T,3,4) = K @ inverse(pose)[:3]  (reference src/data/transforms.py:59)."""
    K = intrinsics(H, W)
    poses = camera_poses(T, voxel_dim, voxel_size, g, pull_back)
    return torch.stack([K @ torch.inverse(poses[t])[:3, :] for t in range(T)]).contiguous()


def frame_features(T, C, H, W, g, B=1):
    """T tensors (B,C,H,W) ~ N(0,1): what the 2D CNN hands to the path."""
    return [torch.randn(B, C, H, W, generator=g) for _ in range(T)]


def depth_maps(T, H, W, g, B=1):
    return 0.5 + 2.5 * torch.rand(B, T, H, W, generator=g)


def query_points(Q, voxel_dim, voxel_size, g, B=1):
    """uniform in [-0.05, 1.05] * n * vs per axis (about a quarter of the points have at
    least one coordinate outside the grid, which exercises the border clamp)."""
    ext = torch.tensor([d * voxel_size for d in voxel_dim], dtype=torch.float32)
    u = torch.rand(B, Q, 3, generator=g) * 1.1 - 0.05
    return (u * ext).contiguous()


def plane_points(N, g, domain="unit", voxel_dim=None, voxel_size=0.04, B=1):
    """'unit': U(-0.55,0.55)^3, the domain normalize_coordinate was designed for;
    'metric': U(0, n*vs), what GenNerf.encode really feeds it (SURVEY trap T6)."""
    if domain == "unit":
        return (torch.rand(B, N, 3, generator=g) * 1.1 - 0.55).contiguous()
    ext = torch.tensor([d * voxel_size for d in voxel_dim], dtype=torch.float32)
    return (torch.rand(B, N, 3, generator=g) * ext).contiguous()


def decoder_weights(g, d_feat, d_code, d_hidden=512, n_blocks=5, d_out=64, d_geo=32, alpha=1.0):
    """state_dict with the reference's ResnetFC / TSDFHeadSimple keys.  kaiming-normal
    (fan_in) like the reference, but fc_1 is NOT zero (SURVEY trap T9) and biases are
    small random values so that every term of the network is exercised.  The head weight is
    scaled so |pre-tanh| stays around 1 (tanh saturation hides errors)."""
    def kaiming(out_f, in_f):
        return torch.randn(out_f, in_f, generator=g) * math.sqrt(2.0 / in_f)

    def bias(n):
        return 0.05 * torch.randn(n, generator=g)

    w = {"lin_in.weight": kaiming(d_hidden, d_feat), "lin_in.bias": bias(d_hidden),
         "lin_out.weight": kaiming(d_out, d_hidden), "lin_out.bias": bias(d_out),
         "alpha": torch.tensor(float(alpha))}
    for i in range(n_blocks):
        w[f"lin_z.{i}.weight"] = kaiming(d_hidden, d_code)
        w[f"lin_z.{i}.bias"] = bias(d_hidden)
        w[f"blocks.{i}.fc_0.weight"] = kaiming(d_hidden, d_hidden)
        w[f"blocks.{i}.fc_0.bias"] = bias(d_hidden)
        # a residual branch of moderate gain keeps the stream O(1) over 5 blocks
        w[f"blocks.{i}.fc_1.weight"] = 0.5 * kaiming(d_hidden, d_hidden)
        w[f"blocks.{i}.fc_1.bias"] = bias(d_hidden)
    head_w = torch.randn(1, d_geo, generator=g) * (0.5 / math.sqrt(d_geo))
    head_b = 0.05 * torch.randn(1, generator=g)
    return w, head_w, head_b


def surface_depth_maps(T, H, W, g, mean=0.9, holes=True):
    """(T,H,W) smooth depth maps around `mean` metres (a wavy surface plus 2 cm of noise) with a regular pattern of
    zero-depth pixels ("no measurement"), for the TSDF fusion path (reference src/data/tsdf.py:369-418)."""
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    d = torch.stack([mean + 0.4 * torch.sin(xx * (8.0 / W) + t) * torch.cos(yy * (6.0 / H))
                     + 0.02 * torch.randn(H, W, generator=g) for t in range(T)])
    if holes:
        d[:, ::7, ::5] = 0.0
    return d.contiguous()

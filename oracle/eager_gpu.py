"""The reference's own code path for the lift-and-query hot path, restated device-generically so that it can run on the
B200 through ATen's CUDA kernels (bmm, index_put_, grid_sampler_3d/2d, cuBLAS linear) -- the "generic library path on the
same B200" of SURVEY.md section 2.2.

TEST / BENCH INFRASTRUCTURE ONLY, like the rest of oracle/: bench.py's informational `gpu_eager_baseline` leg runs it;
the product never imports it.  It follows the same reference lines as gennerf_oracle.py (file:line relative to the
reference root) but allocates on the inputs' device, and is NOT a parity oracle: on the GPU ATen's division by a host
scalar and TF32 matmuls differ from the CPU run (SURVEY traps T6 / T10), which is why the CPU run is the oracle.
"""
import torch
import torch.nn.functional as F

PLANES = ("xz", "xy", "yz")
_AXES = {"xz": [0, 2], "xy": [0, 1], "yz": [1, 2]}


def coordinates(voxel_dim, device):
    """src/data/tsdf.py:25-40."""
    nx, ny, nz = voxel_dim
    x = torch.arange(nx, dtype=torch.long, device=device)
    y = torch.arange(ny, dtype=torch.long, device=device)
    z = torch.arange(nz, dtype=torch.long, device=device)
    x, y, z = torch.meshgrid(x, y, z, indexing="ij")
    return torch.stack((x.flatten(), y.flatten(), z.flatten()))


def backproject(voxel_dim, voxel_size, origin, projection, features):
    """src/models/utils.py:948-996, line by line, on features.device."""
    batch = features.size(0)
    channels = features.size(1)
    device = features.device
    nx, ny, nz = voxel_dim
    coords = coordinates(voxel_dim, device).unsqueeze(0).expand(batch, -1, -1)
    world = coords.type_as(projection) * voxel_size + origin.to(device).unsqueeze(2)
    world = torch.cat((world, torch.ones_like(world[:, :1])), dim=1)
    camera = torch.bmm(projection, world)
    px = (camera[:, 0, :] / camera[:, 2, :]).round().type(torch.long)
    py = (camera[:, 1, :] / camera[:, 2, :]).round().type(torch.long)
    pz = camera[:, 2, :]
    height, width = features.size()[2:]
    valid = (px >= 0) & (py >= 0) & (px < width) & (py < height) & (pz > 0)
    volume = torch.zeros(batch, channels, nx * ny * nz, dtype=features.dtype, device=device)
    for b in range(batch):
        volume[b, :, valid[b]] = features[b, :, py[b, valid[b]], px[b, valid[b]]]
    return volume.view(batch, channels, nx, ny, nz), valid.view(batch, 1, nx, ny, nz)


def encode_volume(voxel_dim, voxel_size, origin, projections, features):
    """GenNerf.encode's volume branch (src/models/model.py:100-127)."""
    volume = valid = None
    for t, feat in enumerate(features):
        v, m = backproject(voxel_dim, voxel_size, origin, projections[:, t], feat)
        if volume is None:
            volume, valid = v, m
        else:
            volume = volume + v
            valid = valid + m
    return volume, valid


def normalize_coordinate(p, padding=0.1, plane="xz"):
    """src/models/utils.py:75-98."""
    xy = p[:, :, _AXES[plane]]
    xy_new = xy / (1 + padding + 10e-6)
    xy_new = xy_new + 0.5
    xy_new = torch.where(xy_new >= 1, torch.full_like(xy_new, 1 - 10e-6), xy_new)
    xy_new = torch.where(xy_new < 0, torch.zeros_like(xy_new), xy_new)
    return xy_new


def generate_plane_features(p, c, plane, reso, padding=0.1):
    """pointnet.py:72-89 with torch_scatter.scatter_mean's published semantics (scatter_add_ + count + divide)."""
    xy = normalize_coordinate(p.clone(), padding, plane)
    x = (xy * reso).long()
    index = (x[:, :, 0] + reso * x[:, :, 1])[:, None, :]
    B, N, Cp = c.shape
    src = c.permute(0, 2, 1)
    out = torch.zeros(B, Cp, reso * reso, device=c.device).scatter_add_(2, index.expand(B, Cp, N), src)
    cnt = torch.zeros(B, 1, reso * reso, device=c.device).scatter_add_(2, index, torch.ones(B, 1, N, device=c.device))
    return (out / cnt.clamp_min(1)).reshape(B, Cp, reso, reso)


def map_features(xyz, volume, valid, planes, voxel_size, padding):
    """GenNerf.map_features (model.py:163-204): per call volume/valid normalisation, grid_sample 3-D and 2-D."""
    feats = []
    if planes is not None:
        fp = 0
        for name in PLANES:
            xy = normalize_coordinate(xyz.clone(), padding, name)
            vgrid = 2.0 * xy[:, :, None] - 1.0
            fp = fp + F.grid_sample(planes[name], vgrid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1)
        feats.append(fp.transpose(1, 2))
    if volume is not None:
        vol = volume / valid
        vol = vol.transpose(0, 1)
        vol[:, valid.squeeze(1) == 0] = 0
        vol = vol.transpose(0, 1).permute(0, 2, 3, 4, 1)                 # (B,nx,ny,nz,C), model.py:201
        B, nx, ny, nz, C = vol.shape
        size = torch.tensor([nx, ny, nz], device=xyz.device) * voxel_size
        g = 2.0 * (xyz / size) - 1.0                                       # utils.py:1017-1021 (origin 0)
        out = F.grid_sample(vol.permute(0, 4, 3, 2, 1), g[:, :, None, None, :], mode="bilinear", padding_mode="border",
                            align_corners=True)
        feats.append(out[:, :, :, 0, 0].permute(0, 2, 1))
    return torch.cat(feats, dim=-1)


def positional_encoding(x, num_freqs, freq_factor):
    """positional_encoding.py:28-40 (include_input=True)."""
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs, device=x.device)
    f = torch.repeat_interleave(freqs, 2).view(1, -1, 1)
    ph = torch.zeros(2 * num_freqs, device=x.device)
    ph[1::2] = torch.pi * 0.5
    emb = torch.sin(torch.addcmul(ph.view(1, -1, 1), x.unsqueeze(1).repeat(1, num_freqs * 2, 1), f))
    return torch.cat((x, emb.view(x.shape[0], -1)), dim=-1)


def resnetfc(zx, w, n_blocks, d_latent):
    """resnetfc.py:134-189 with the default options."""
    z, x = zx[..., :d_latent], zx[..., d_latent:]
    x = F.linear(x, w["lin_in.weight"], w["lin_in.bias"])
    for i in range(n_blocks):
        x = x + w["alpha"] * F.linear(z, w[f"lin_z.{i}.weight"], w[f"lin_z.{i}.bias"])
        net = F.linear(F.relu(x), w[f"blocks.{i}.fc_0.weight"], w[f"blocks.{i}.fc_0.bias"])
        x = x + F.linear(F.relu(net), w[f"blocks.{i}.fc_1.weight"], w[f"blocks.{i}.fc_1.bias"])
    return F.linear(F.relu(x), w["lin_out.weight"], w["lin_out.bias"])


def forward(xyz, w, head_w, head_b, volume, valid, planes, voxel_size, padding, num_freqs, freq_factor, n_blocks, d_geo):
    """GenNerf.forward (model.py:207-248) -> tsdf (B,Q,1)."""
    B, Q, _ = xyz.shape
    feat = map_features(xyz, volume, valid, planes, voxel_size, padding)
    code = positional_encoding(xyz.reshape(-1, 3), num_freqs, freq_factor).reshape(B, Q, -1)
    out = resnetfc(torch.cat((code, feat), dim=-1), w, n_blocks, code.shape[-1])
    return torch.tanh(F.linear(out[..., :d_geo], head_w, head_b))


def predict_chunks(xyz, chunk, *args, **kw):
    """The reference's dense-extraction loop (model.py:769-777): 10 000-point chunks, .cpu() per chunk."""
    outs = []
    for c in torch.split(xyz, chunk, dim=1):
        outs.append(forward(c, *args, **kw).detach().cpu())
    return torch.cat(outs, dim=1)

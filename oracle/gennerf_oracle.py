"""CPU oracle for the gen-nerf lift-and-query hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU in fp32, the algorithm of the reference path named by
BASELINE.json's north_star.  It is the checker the CUDA kernels
are compared with; it is never the product.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(gennerf_b200/) must never import anything from oracle/.

Pinning status: the reference ships no golden vectors or tests, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container through oracle/ref_shim.py:
  * tests/test_oracle_pinning.py   oracle == real reference, bit-for-bit, live (skipped
                                   where /root/reference is absent)
  * tests/golden/*.pt              vectors produced by tests/golden/make_golden.py from
                                   the real reference; the oracle and the CUDA path are
                                   both checked against them everywhere.

The reference is PyTorch; wherever the reference's arithmetic is a call into ATen
(torch.bmm, F.grid_sample, nn.Linear, Tensor.scatter_add_) the oracle calls the same
ATen CPU kernel, so that it is the reference's own arithmetic and not a look-alike.
Explicit restatements of those library kernels (the *_explicit functions) document the
arithmetic the CUDA kernels implement and are tested against the ATen calls.

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PLANES = ("xz", "xy", "yz")
_PLANE_AXES = {"xz": (0, 2), "xy": (0, 1), "yz": (1, 2)}


# ----------------------------------------------------------------------------------------
# a1  voxel index grid                                           src/data/tsdf.py:25-40
# ----------------------------------------------------------------------------------------
def coordinates(voxel_dim):
    """int64 (3, V) voxel indices, linear id v = (x*ny + y)*nz + z  (tsdf.py:35-40)."""
    nx, ny, nz = (int(d) for d in voxel_dim)
    v = torch.arange(nx * ny * nz, dtype=torch.long)
    z = v % nz
    y = (v // nz) % ny
    x = v // (nz * ny)
    return torch.stack((x, y, z))


# ----------------------------------------------------------------------------------------
# a2  per-frame back-projection                               src/models/utils.py:948-996
# ----------------------------------------------------------------------------------------
def project_indices(voxel_dim, voxel_size, origin, projection, height, width):
    """px, py (int64 (B,V)), pz (fp32 (B,V)), valid (bool (B,V)).

    utils.py:973-985.  world = fl(i)*voxel_size + origin (two rounded ops), homogeneous 1,
    camera = bmm(projection, world), px = round-half-even(cx/cz) cast to int64.
    """
    B = projection.shape[0]
    coords = coordinates(voxel_dim).unsqueeze(0).expand(B, -1, -1)
    world = coords.type_as(projection) * voxel_size + origin.unsqueeze(2)
    world = torch.cat((world, torch.ones_like(world[:, :1])), dim=1)
    camera = torch.bmm(projection, world)
    px = (camera[:, 0, :] / camera[:, 2, :]).round().type(torch.long)
    py = (camera[:, 1, :] / camera[:, 2, :]).round().type(torch.long)
    pz = camera[:, 2, :]
    valid = (px >= 0) & (py >= 0) & (px < width) & (py < height) & (pz > 0)
    return px, py, pz, valid


def _fma32(a, b, c):
    """Correctly rounded fp32 fma(a,b,c) from float64 arithmetic: the product of two fp32
    values is exact in fp64; the fp64 sum is rounded once more to fp32 (double rounding is
    possible only on exact fp32 ties of the fp64-rounded sum, handled by the error term)."""
    a64, b64, c64 = a.double(), b.double(), c.double()
    p = a64 * b64                      # exact
    s = p + c64                        # rounded to fp64
    # two-sum error of the fp64 addition, used to break fp32 ties correctly
    bb = s - p
    err = (p - (s - bb)) + (c64 - bb)
    r = s.float()
    # if s sits exactly on an fp32 rounding tie, nudge by the sign of the fp64 error
    r64 = r.double()
    up = torch.nextafter(r, torch.full_like(r, float("inf"))).double()
    dn = torch.nextafter(r, torch.full_like(r, float("-inf"))).double()
    tie_up = (s - r64) == (up - s)
    tie_dn = (r64 - s) == (s - dn)
    r = torch.where(tie_up & (err > 0), up.float(), r)
    r = torch.where(tie_dn & (err < 0), dn.float(), r)
    return r


def project_indices_explicit(voxel_dim, voxel_size, origin, projection, height, width):
    """Same result as project_indices, with the 3x4 . 4xV product written as the fused
    multiply-add chain the CUDA kernel uses (SURVEY.md trap T3):
        cam_r = fma(P[r,3], 1, fma(P[r,2], wz, fma(P[r,1], wy, P[r,0]*wx)))
    """
    B = projection.shape[0]
    coords = coordinates(voxel_dim).float()
    vs = torch.tensor(float(voxel_size), dtype=torch.float32)
    org = origin.reshape(3).float()
    w = [coords[i] * vs + org[i] for i in range(3)]            # fl(fl(i)*vs) + origin
    one = torch.ones_like(w[0])
    px_l, py_l, pz_l, valid_l = [], [], [], []
    for b in range(B):
        P = projection[b]
        cam = []
        for r in range(3):
            acc = P[r, 0] * w[0]
            acc = _fma32(P[r, 1].expand_as(acc), w[1], acc)
            acc = _fma32(P[r, 2].expand_as(acc), w[2], acc)
            acc = _fma32(P[r, 3].expand_as(acc), one, acc)
            cam.append(acc)
        fx = (cam[0] / cam[2]).round()
        fy = (cam[1] / cam[2]).round()
        px = fx.type(torch.long)
        py = fy.type(torch.long)
        valid = (px >= 0) & (py >= 0) & (px < width) & (py < height) & (cam[2] > 0)
        px_l.append(px), py_l.append(py), pz_l.append(cam[2]), valid_l.append(valid)
    return torch.stack(px_l), torch.stack(py_l), torch.stack(pz_l), torch.stack(valid_l)


def backproject(voxel_dim, voxel_size, origin, projection, features):
    """volume (B,C,nx,ny,nz) fp32, valid (B,1,nx,ny,nz) bool     (utils.py:948-996).

    volume[b,:,v] = features[b,:,py,px] where valid else 0 (nearest pixel, trap T1).
    """
    B, C, H, W = features.shape
    nx, ny, nz = voxel_dim
    px, py, _, valid = project_indices(voxel_dim, voxel_size, origin, projection, H, W)
    volume = torch.zeros(B, C, nx * ny * nz, dtype=features.dtype)
    for b in range(B):
        sel = valid[b]
        volume[b][:, sel] = features[b][:, py[b][sel], px[b][sel]]
    return volume.view(B, C, nx, ny, nz), valid.view(B, 1, nx, ny, nz)


# ----------------------------------------------------------------------------------------
# a3/a4  accumulation over frames and normalisation     src/models/model.py:100-127,195-199
# ----------------------------------------------------------------------------------------
def encode_volume(voxel_dim, voxel_size, origin, projections, features):
    """Accumulate T frames (model.py:121-127; voxel_net.py:120-126).

    projections (B,T,3,4); features: sequence of T tensors (B,C,H,W).
    Returns volume (B,C,nx,ny,nz) = SUM over frames in frame order, valid bool = OR over
    frames (trap T2: bool + bool is logical or), and count int32 (B,nx,ny,nz) -- the number
    of frames that see a voxel, which the reference does not keep but the CUDA kernel does.
    """
    volume = valid = count = None
    for t, feat in enumerate(features):
        vol_t, valid_t = backproject(voxel_dim, voxel_size, origin, projections[:, t], feat)
        if volume is None:
            volume, valid = vol_t, valid_t
            count = valid_t.squeeze(1).to(torch.int32)
        else:
            volume = volume + vol_t
            valid = valid + valid_t
            count = count + valid_t.squeeze(1).to(torch.int32)
    return volume, valid, count


def normalize_volume(volume, valid):
    """model.py:195-199: volume/valid with the NaNs (valid == 0) replaced by 0."""
    out = volume / valid
    out = out.transpose(0, 1)
    out[:, valid.squeeze(1) == 0] = 0
    return out.transpose(0, 1)


# ----------------------------------------------------------------------------------------
# a5  trilinear query                                       src/models/utils.py:999-1042
# ----------------------------------------------------------------------------------------
def _normalize_query(xyz, dims, origin, voxel_size):
    """utils.py:1017-1021 -- note the normalisation by n*voxel_size (trap T4)."""
    xyz = xyz - origin
    xyz = xyz / (torch.tensor(list(dims)) * voxel_size)
    xyz = 2 * xyz - 1
    return xyz.float()


def trilinear_interpolation(voxel_volume, xyz, origin, voxel_size, mode="bilinear"):
    """voxel_volume (B,nx,ny,nz,C) (any strides), xyz (B,N,3) -> (B,N,C).
    Calls the same ATen grid_sampler_3d CPU kernel the reference calls (utils.py:1035)."""
    B, nx, ny, nz, C = voxel_volume.shape
    N = xyz.shape[1]
    g = _normalize_query(xyz, (nx, ny, nz), origin, voxel_size)
    vol = voxel_volume.permute(0, 4, 3, 2, 1)
    out = F.grid_sample(vol, g.view(B, N, 1, 1, 3), mode=mode, align_corners=True, padding_mode="border")
    return out.view(B, C, N).permute(0, 2, 1)


def _unnormalize_clip(g, size):
    """ATen GridSampler.h: align_corners=True unnormalise, then border clip."""
    x = ((g + 1) / 2) * (size - 1)
    return torch.clamp(x, min=0.0, max=float(size - 1))


def trilinear_interpolation_explicit(voxel_volume, xyz, origin, voxel_size):
    """The arithmetic of ATen's grid_sampler_3d (bilinear, border, align_corners=True)
    written out: the 8-corner weighted sum the CUDA sampler implements.  Out-of-range
    corners (index == n, weight 0) are skipped."""
    B, nx, ny, nz, C = voxel_volume.shape
    g = _normalize_query(xyz, (nx, ny, nz), origin, voxel_size)
    ix = _unnormalize_clip(g[..., 0], nx)
    iy = _unnormalize_clip(g[..., 1], ny)
    iz = _unnormalize_clip(g[..., 2], nz)
    x0, y0, z0 = ix.floor(), iy.floor(), iz.floor()
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    out = torch.zeros(B, xyz.shape[1], C)
    bidx = torch.arange(B).view(B, 1).expand(B, xyz.shape[1])
    for (xc, wx) in ((x0, x1 - ix), (x1, ix - x0)):
        for (yc, wy) in ((y0, y1 - iy), (y1, iy - y0)):
            for (zc, wz) in ((z0, z1 - iz), (z1, iz - z0)):
                inb = (xc <= nx - 1) & (yc <= ny - 1) & (zc <= nz - 1)
                xi = xc.clamp(max=nx - 1).long()
                yi = yc.clamp(max=ny - 1).long()
                zi = zc.clamp(max=nz - 1).long()
                val = voxel_volume[bidx, xi, yi, zi]                    # (B,N,C)
                out = out + torch.where(inb, wx * wy * wz, torch.zeros(())).unsqueeze(-1) * val
    return out


# ----------------------------------------------------------------------------------------
# a6  plane coordinates and cell indices                      src/models/utils.py:57-98
# ----------------------------------------------------------------------------------------
def normalize_coordinate(p, padding=0.1, plane="xz"):
    """(B,N,3) -> (B,N,2) in [0, 1-1e-5]   (utils.py:75-98, trap T6: 10e-6 == 1e-5)."""
    a0, a1 = _PLANE_AXES[plane]
    xy = p[:, :, [a0, a1]]
    xy_new = xy / (1 + padding + 10e-6)
    xy_new = xy_new + 0.5
    xy_new = torch.where(xy_new >= 1, torch.tensor(1 - 10e-6, dtype=xy_new.dtype), xy_new)
    xy_new = torch.where(xy_new < 0, torch.zeros((), dtype=xy_new.dtype), xy_new)
    return xy_new


def coordinate2index(x, reso):
    """(B,N,2) in [0,1) -> int64 (B,1,N): x0 + reso*x1 (utils.py:57-72, '2d')."""
    xi = (x * reso).long()
    return (xi[:, :, 0] + reso * xi[:, :, 1])[:, None, :]


# ----------------------------------------------------------------------------------------
# torch_scatter semantics (third-party, absent from /root/reference; README.md:42, no pin)
# restated from upstream torch_scatter/scatter.py: scatter_sum = scatter_add_,
# scatter_mean = sum / clamp(count,1), scatter_max -> untouched cells hold 0.
# ----------------------------------------------------------------------------------------
def scatter_mean(src, index, dim_size):
    """src (B,C,N), index (B,1,N) -> mean (B,C,dim_size), count int32 (B,dim_size)."""
    B, C, N = src.shape
    idx = index.expand(B, C, N)
    out = torch.zeros(B, C, dim_size, dtype=src.dtype).scatter_add_(2, idx, src)
    ones = torch.ones(B, 1, N, dtype=src.dtype)
    cnt = torch.zeros(B, 1, dim_size, dtype=src.dtype).scatter_add_(2, index, ones)
    count = cnt.squeeze(1).to(torch.int32)
    cnt = torch.where(cnt < 1, torch.ones(()), cnt)
    out.true_divide_(cnt)
    return out, count


def scatter_max(src, index, dim_size):
    """src (B,C,N), index (B,1,N) -> max (B,C,dim_size), 0 where no point falls."""
    B, C, N = src.shape
    idx = index.expand(B, C, N)
    out = torch.full((B, C, dim_size), float("-inf"), dtype=src.dtype)
    out.scatter_reduce_(2, idx, src, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros(()), out)


# ----------------------------------------------------------------------------------------
# a7/a8  triplane scatter and local pooling    src/models/components/pointnet.py:72-121
# ----------------------------------------------------------------------------------------
def generate_plane_features(p, c, plane, reso, padding=0.1, return_count=False):
    """p (B,N,3), c (B,N,C_p) -> (B,C_p,reso,reso): scatter_mean of point features onto one
    plane (pointnet.py:72-89, without the optional U-Net; trap T7: always mean)."""
    xy = normalize_coordinate(p.clone(), plane=plane, padding=padding)
    index = coordinate2index(xy, reso)
    fea, count = scatter_mean(c.permute(0, 2, 1).float(), index, reso * reso)
    fea = fea.reshape(p.size(0), c.size(2), reso, reso)
    if return_count:
        return fea, count.reshape(p.size(0), reso, reso)
    return fea


def pool_local(p, c, reso, padding=0.1, planes=PLANES, scatter_type="max"):
    """c (B,N,hidden) -> (B,N,hidden): for every plane scatter (max|mean) into cells then
    gather back to the points; summed over planes (pointnet.py:105-121)."""
    B, N, Hd = c.shape
    out = 0
    for plane in planes:
        xy = normalize_coordinate(p.clone(), plane=plane, padding=padding)
        index = coordinate2index(xy, reso)
        if scatter_type == "max":
            fea = scatter_max(c.permute(0, 2, 1), index, reso * reso)
        else:
            fea, _ = scatter_mean(c.permute(0, 2, 1), index, reso * reso)
        out = out + fea.gather(2, index.expand(-1, Hd, -1))
    return out.permute(0, 2, 1)


# ----------------------------------------------------------------------------------------
# a9  plane query                                            src/models/model.py:153-161
# ----------------------------------------------------------------------------------------
def sample_plane_feature(p, c, plane, padding=0.1, mode="bilinear"):
    """p (B,Q,3), c (B,C_p,R,R) -> (B,C_p,Q): bilinear, border, align_corners=True, i.e.
    pixel coordinate = u*(R-1) (model.py:153-161); same ATen kernel as the reference."""
    xy = normalize_coordinate(p.clone(), plane=plane, padding=padding)
    vgrid = 2.0 * xy[:, :, None].float() - 1.0
    return F.grid_sample(c, vgrid, padding_mode="border", align_corners=True, mode=mode).squeeze(-1)


def sample_plane_feature_explicit(p, c, plane, padding=0.1):
    """The 4-corner arithmetic of ATen's grid_sampler_2d written out.  Grid x (= first plane
    coordinate) indexes the last (W) axis of c, grid y the H axis."""
    B, Cp, R, _ = c.shape
    xy = normalize_coordinate(p.clone(), plane=plane, padding=padding)
    g = 2.0 * xy.float() - 1.0
    ix = _unnormalize_clip(g[..., 0], R)
    iy = _unnormalize_clip(g[..., 1], R)
    x0, y0 = ix.floor(), iy.floor()
    x1, y1 = x0 + 1, y0 + 1
    out = torch.zeros(B, Cp, p.shape[1])
    bidx = torch.arange(B).view(B, 1).expand(B, p.shape[1])
    for (xc, wx) in ((x0, x1 - ix), (x1, ix - x0)):
        for (yc, wy) in ((y0, y1 - iy), (y1, iy - y0)):
            inb = (xc <= R - 1) & (yc <= R - 1)
            xi = xc.clamp(max=R - 1).long()
            yi = yc.clamp(max=R - 1).long()
            val = c[bidx, :, yi, xi]                                       # (B,Q,Cp)
            out = out + (torch.where(inb, wx * wy, torch.zeros(())).unsqueeze(-1) * val).permute(0, 2, 1)
    return out


def grid_sample_2d(image, optical):
    """The reference's double-differentiable stand-in for F.grid_sample 2-D (bilinear, border, align_corners=True):
    src/models/utils.py:1117-1174, chosen by model.py:157-158 when loss.use_eikonal / loss.use_gradient is set.
    image (N,C,IH,IW), optical (N,H,W,2) -> (N,C,H,W).  Unlike ATen's kernel it does not clip the coordinates, only the
    corner indices; inside [-1,1] (where normalize_coordinate keeps vgrid) both agree."""
    N, C, IH, IW = image.shape
    _, H, W, _ = optical.shape
    ix = ((optical[..., 0] + 1) / 2) * (IW - 1)
    iy = ((optical[..., 1] + 1) / 2) * (IH - 1)
    with torch.no_grad():
        ix_nw, iy_nw = torch.floor(ix), torch.floor(iy)
        ix_ne, iy_ne = ix_nw + 1, iy_nw
        ix_sw, iy_sw = ix_nw, iy_nw + 1
        ix_se, iy_se = ix_nw + 1, iy_nw + 1
    nw = (ix_se - ix) * (iy_se - iy)
    ne = (ix - ix_sw) * (iy_sw - iy)
    sw = (ix_ne - ix) * (iy - iy_ne)
    se = (ix - ix_nw) * (iy - iy_nw)
    flat = image.view(N, C, IH * IW)

    def corner(cx, cy):
        with torch.no_grad():
            idx = (cy.clamp(0, IH - 1) * IW + cx.clamp(0, IW - 1)).long().view(N, 1, H * W).repeat(1, C, 1)
        return torch.gather(flat, 2, idx).view(N, C, H, W)

    return (corner(ix_nw, iy_nw) * nw.view(N, 1, H, W) + corner(ix_ne, iy_ne) * ne.view(N, 1, H, W) +
            corner(ix_sw, iy_sw) * sw.view(N, 1, H, W) + corner(ix_se, iy_se) * se.view(N, 1, H, W))


def sample_plane_feature_eikonal(p, c, plane, padding=0.1):
    """model.py:153-161 with loss.use_eikonal / use_gradient: normalize_coordinate + grid_sample_2d -> (B,C_p,Q)."""
    xy = normalize_coordinate(p.clone(), plane=plane, padding=padding)
    vgrid = 2.0 * xy[:, :, None].float() - 1.0
    return grid_sample_2d(c, vgrid).squeeze(-1)


def map_features_twice_differentiable(xyz, volume=None, planes=None, voxel_size=0.04, padding=0.1, origin=None):
    """map_features (model.py:163-204) from operations autograd can differentiate twice: the planes through the reference's
    own grid_sample_2d, the volume through the written-out 8-corner sum (ATen has no double backward for grid_sampler_3d,
    so the reference itself cannot run its eikonal loss with a volume that requires grad).  Checker of
    gnb_sample_features_bwd2."""
    feats = []
    if planes is not None:
        fp = 0
        for name in PLANES:
            if name in planes:
                fp = fp + sample_plane_feature_eikonal(xyz, planes[name], name, padding)
        feats.append(fp.transpose(1, 2))
    if volume is not None:
        org = torch.zeros(3, dtype=torch.long) if origin is None else origin
        feats.append(trilinear_interpolation_explicit(volume.permute(0, 2, 3, 4, 1), xyz, org, voxel_size))
    return torch.cat(feats, dim=-1)


# ----------------------------------------------------------------------------------------
# a10  feature lookup                                        src/models/model.py:163-204
# ----------------------------------------------------------------------------------------
def map_features(xyz, volume=None, valid=None, planes=None, voxel_size=0.04, padding=0.1,
                 origin=None):
    """(B,Q,3) -> (B,Q,C_p + C): plane features FIRST, then volume features
    (model.py:176-203).  planes: dict plane-name -> (B,C_p,R,R) in the reference's key order."""
    B, Q, _ = xyz.shape
    feats = []
    if planes is not None:
        fp = 0
        for name in PLANES:                                   # model.py:185-190 order
            if name in planes:
                fp = fp + sample_plane_feature(xyz, planes[name], name, padding)
        feats.append(fp.transpose(1, 2))
    if volume is not None:
        vol = normalize_volume(volume, valid).permute(0, 2, 3, 4, 1)
        org = torch.zeros(3, dtype=torch.long) if origin is None else origin
        feats.append(trilinear_interpolation(vol, xyz, org, voxel_size))
    return torch.cat(feats, dim=-1)


# ----------------------------------------------------------------------------------------
# a11  positional encoding      src/models/components/positional_encoding.py:10-40
# ----------------------------------------------------------------------------------------
def positional_encoding(x, num_freqs, freq_factor=np.pi, include_input=True):
    """(M,3) -> (M, 3 + 6*num_freqs): [x, sin(f0 x), sin(f0 x + pi/2), sin(f1 x), ...]
    with f_k = freq_factor * 2^k and the phase added through addcmul (trap T12)."""
    d_in = x.shape[-1]
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
    _freqs = torch.repeat_interleave(freqs, 2).view(1, -1, 1)
    _phases = torch.zeros(2 * num_freqs)
    _phases[1::2] = np.pi * 0.5
    _phases = _phases.view(1, -1, 1)
    embed = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    embed = torch.sin(torch.addcmul(_phases, embed, _freqs))
    embed = embed.view(x.shape[0], -1)
    if include_input:
        embed = torch.cat((x, embed), dim=-1)
    return embed


# ----------------------------------------------------------------------------------------
# a12/a13  ResNet-MLP decoder and TSDF head
#          src/models/components/resnetfc.py:54-63,134-189 ; heads3d.py:36-50
# ----------------------------------------------------------------------------------------
def resnetfc_forward(zx, w, n_blocks, d_latent, beta=0.0):
    """zx (..., d_latent + d_in): the first d_latent columns are the code injected into
    every block through lin_z (trap T8), the rest goes through lin_in.  `w` is a state_dict
    with the reference's keys (lin_in.*, lin_z.{i}.*, blocks.{i}.fc_{0,1}.*, lin_out.*, alpha).
    Default options only: no spade, no layer norm, combine_layer > n_blocks."""
    act = F.relu if beta <= 0 else (lambda t: F.softplus(t, beta=beta))
    z, x = zx[..., :d_latent], zx[..., d_latent:]
    x = F.linear(x, w["lin_in.weight"], w["lin_in.bias"])
    for i in range(n_blocks):
        if d_latent > 0:
            tz = F.linear(z, w[f"lin_z.{i}.weight"], w[f"lin_z.{i}.bias"])
            x = x + w["alpha"] * tz
        net = F.linear(act(x), w[f"blocks.{i}.fc_0.weight"], w[f"blocks.{i}.fc_0.bias"])
        dx = F.linear(act(net), w[f"blocks.{i}.fc_1.weight"], w[f"blocks.{i}.fc_1.bias"])
        x = x + dx
    return F.linear(act(x), w["lin_out.weight"], w["lin_out.bias"])


def tsdf_head(feat_geo, weight, bias):
    """tanh(Linear(d_geo -> 1))   (heads3d.py:44-45)."""
    return torch.tanh(F.linear(feat_geo, weight, bias))


# ----------------------------------------------------------------------------------------
# a14  the whole query                                       src/models/model.py:207-248
# ----------------------------------------------------------------------------------------
def gennerf_forward(xyz, mlp_w, head_w, head_b, *, volume=None, valid=None, planes=None,
                    voxel_size=0.04, padding=0.1, num_freqs=2, freq_factor=0.5,
                    include_input=True, use_code=True, n_blocks=5, d_out_geo=32, d_out_sem=32,
                    beta=0.0, twice_differentiable=False):
    """twice_differentiable: the feature lookup the eikonal / gradient losses need (model.py:157-158: grid_sample_2d for
    the planes; written-out trilinear sum for an already normalised volume, `valid` ignored)."""
    B, Q, _ = xyz.shape
    if twice_differentiable:
        feat = map_features_twice_differentiable(xyz, volume, planes, voxel_size, padding)
    else:
        feat = map_features(xyz, volume, valid, planes, voxel_size, padding)
    code = xyz
    if use_code:
        code = positional_encoding(xyz.reshape(-1, 3), num_freqs, freq_factor, include_input).reshape(B, Q, -1)
    out = resnetfc_forward(torch.cat((code, feat), dim=-1), mlp_w, n_blocks, code.shape[-1], beta)
    feat_geo = out[..., :d_out_geo]
    feat_sem = out[..., d_out_geo:d_out_geo + d_out_sem]
    return {"feat_geo": feat_geo, "feat_sem": feat_sem, "tsdf": tsdf_head(feat_geo, head_w, head_b),
            "feat": feat}


# ----------------------------------------------------------------------------------------
# next rows                               src/models/utils.py:120-202,926-935
# ----------------------------------------------------------------------------------------
def get_grid_coordinates(nx, ny, nz, volume_size):
    """(nx,ny,nz,3) query grid, linspace(0, size, n) inclusive  (utils.py:926-935)."""
    x = torch.linspace(0, volume_size[0], nx)
    y = torch.linspace(0, volume_size[1], ny)
    z = torch.linspace(0, volume_size[2], nz)
    gx, gy, gz = torch.meshgrid(x, y, z, indexing="ij")
    return torch.stack([gx, gy, gz], dim=-1)


def get_3d_points(depth_map, projection):
    """depth (B,H,W), projection (B,3,4) world->pixel -> world points (B,H,W,3)
    (utils.py:120-175): [u*d, v*d, d, 1] . inverse([P;0 0 0 1])^T, dehomogenised."""
    B, H, W = depth_map.shape
    u = torch.arange(0, W).view(1, -1).expand(H, -1).float()
    v = torch.arange(0, H).view(-1, 1).expand(-1, W).float()
    uv1 = torch.stack((u, v, torch.ones_like(u)), dim=-1).view(1, -1, 3).expand(B, -1, -1)
    pts2d = uv1 * depth_map.view(B, -1).unsqueeze(-1)
    bottom = torch.tensor([0, 0, 0, 1], dtype=projection.dtype).view(1, 1, 4).repeat(B, 1, 1)
    inv = torch.inverse(torch.cat((projection, bottom), dim=1))
    hom = torch.cat((pts2d, torch.ones_like(pts2d[..., :1])), dim=-1)
    p3 = torch.matmul(hom, inv.transpose(-1, -2))
    return (p3[..., :3] / p3[..., 3:4]).reshape(B, H, W, 3)


def farthest_point_sample(xyz, npoint, start):
    """xyz (B,N,3), start (B,) int64 first index (the reference draws it with
    torch.randint, utils.py:191; parity needs it as an input) -> (B,npoint,3), (B,npoint)."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = start.clone()
    bi = torch.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[bi, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        distance = torch.where(dist < distance, dist, distance)
        farthest = torch.max(distance, -1)[1]
    return xyz[bi[:, None], centroids], centroids


# ----------------------------------------------------------------------------------------
# f-3  TSDF fusion (GT generation / evaluation re-fusion)         src/data/tsdf.py:320-440
# ----------------------------------------------------------------------------------------
class TSDFFusion:
    """CPU restatement of the reference's TSDFFusion (tsdf.py:320-440): running TSDF / weight
    (/ colour / label) volumes over the flat voxel index v = (x*ny + y)*nz + z.

    integrate() follows tsdf.py:369-418 op by op with the same ATen CPU kernels (2-D `@`,
    true fp32 division by the Python-float truncation margin, masked assignments)."""

    def __init__(self, voxel_dim, voxel_size, origin, trunc_ratio=3, color=True, label=False):
        nx, ny, nz = (int(d) for d in voxel_dim)
        self.voxel_dim = (nx, ny, nz)
        self.voxel_size = voxel_size
        self.origin = torch.tensor(origin, dtype=torch.float).view(1, 3)
        self.trunc_margin = voxel_size * trunc_ratio                     # Python float (tsdf.py:341)
        world = coordinates(voxel_dim).type(torch.float) * voxel_size + self.origin.T   # tsdf.py:344
        self.world = torch.cat((world, torch.ones_like(world[:1])), dim=0)
        V = nx * ny * nz
        self.tsdf_vol = torch.ones(V)
        self.weight_vol = torch.zeros(V)
        self.color_vol = torch.zeros(3, V) if color else None
        self.label_vol = -torch.ones(V, dtype=torch.long) if label else None

    def reset(self):
        self.tsdf_vol.fill_(1)
        self.weight_vol.fill_(0)
        if self.color_vol is not None:
            self.color_vol.fill_(0)
        if self.label_vol is not None:
            self.label_vol.fill_(-1)

    def integrate(self, projection, depth, color=None, label=None):
        camera = projection @ self.world                                  # tsdf.py:380
        px = (camera[0, :] / camera[2, :]).round().type(torch.long)
        py = (camera[1, :] / camera[2, :]).round().type(torch.long)
        pz = camera[2, :]
        height, width = depth.size()
        valid = (px >= 0) & (py >= 0) & (px < width) & (py < height) & (pz > 0)
        valid_ = valid.clone()
        valid[valid_] *= depth[py[valid_], px[valid_]] > 0                # tsdf.py:391
        dist = pz[valid] - depth[py[valid], px[valid]]
        dist = torch.clamp(dist / self.trunc_margin, min=-1)              # tsdf.py:395
        valid1 = dist < 1
        valid_ = valid.clone()
        valid[valid_] *= valid1
        dist = dist[valid1]
        mask1 = self.weight_vol == 0
        self.tsdf_vol[valid & mask1] = dist[mask1[valid]]                 # first observation: copy (tsdf.py:405)
        mask2 = valid.clone()
        valid2 = dist > -1
        mask2[valid] *= valid2                                            # near surface
        mask3 = ~mask1 & mask2
        self.tsdf_vol[mask3] += dist[mask3[valid]]
        self.weight_vol[mask2] += 1
        if self.color_vol is not None:
            self.color_vol[:, mask2] += color[:, py[mask2], px[mask2]]
        if self.label_vol is not None:
            self.label_vol[mask2] = label[py[mask2], px[mask2]]           # newest label wins

    def get_volumes(self):
        """The arithmetic of get_tsdf (tsdf.py:420-440) without the TSDF container object:
        tsdf (nx,ny,nz), colour (3,nx,ny,nz) or None, label (nx,ny,nz) or None."""
        nx, ny, nz = self.voxel_dim
        seen = self.weight_vol > 0
        tsdf = self.tsdf_vol.clone()
        tsdf[seen] /= self.weight_vol[seen]
        color = None
        if self.color_vol is not None:
            color = self.color_vol.clone()
            color[:, seen] /= self.weight_vol[seen]
            color = color.view(3, nx, ny, nz)
        label = self.label_vol.view(nx, ny, nz).clone() if self.label_vol is not None else None
        return tsdf.view(nx, ny, nz), color, label


def tsdf_fusion_explicit(voxel_dim, voxel_size, origin, trunc_ratio, projections, depths, colors=None, labels=None):
    """The same fusion written per voxel the way the CUDA kernel computes it: FMA-chain projection (as
    project_indices_explicit), round-half-even pixel, then the running update in frame order.  Returns the raw
    (tsdf_vol, weight_vol, color_vol, label_vol); must equal TSDFFusion.integrate called frame by frame."""
    V = int(voxel_dim[0]) * int(voxel_dim[1]) * int(voxel_dim[2])
    coords = coordinates(voxel_dim).float()
    vs = torch.tensor(float(voxel_size), dtype=torch.float32)
    org = torch.as_tensor(origin, dtype=torch.float32).reshape(3)
    w = [coords[i] * vs + org[i] for i in range(3)]
    one = torch.ones_like(w[0])
    trunc = torch.tensor(voxel_size * trunc_ratio, dtype=torch.float32)    # the Python float, rounded to fp32 once
    tsdf = torch.ones(V)
    weight = torch.zeros(V)
    color_vol = torch.zeros(3, V) if colors is not None else None
    label_vol = -torch.ones(V, dtype=torch.long) if labels is not None else None
    for f in range(projections.shape[0]):
        P = projections[f]
        H, W = depths[f].shape
        cam = []
        for r in range(3):
            acc = P[r, 0] * w[0]
            acc = _fma32(P[r, 1].expand_as(acc), w[1], acc)
            acc = _fma32(P[r, 2].expand_as(acc), w[2], acc)
            acc = _fma32(P[r, 3].expand_as(acc), one, acc)
            cam.append(acc)
        fx = (cam[0] / cam[2]).round()
        fy = (cam[1] / cam[2]).round()
        inb = (fx >= 0) & (fy >= 0) & (fx < W) & (fy < H) & (cam[2] > 0)
        ix = torch.where(inb, fx, torch.zeros_like(fx)).long()
        iy = torch.where(inb, fy, torch.zeros_like(fy)).long()
        dpt = depths[f][iy, ix]
        ok = inb & (dpt > 0)
        dist = torch.clamp((cam[2] - dpt) / trunc, min=-1)
        ok = ok & (dist < 1)
        first = ok & (weight == 0)
        near = ok & (dist > -1)
        tsdf = torch.where(first, dist, tsdf)
        tsdf = torch.where(near & ~first, tsdf + dist, tsdf)
        weight = torch.where(near, weight + 1, weight)
        if color_vol is not None:
            color_vol = torch.where(near.unsqueeze(0), color_vol + colors[f][:, iy, ix], color_vol)
        if label_vol is not None:
            label_vol = torch.where(near, labels[f][iy, ix], label_vol)
    return tsdf, weight, color_vol, label_vol


# ----------------------------------------------------------------------------------------
# f-4  training-time ray sampler                              src/models/utils.py:458-540
# ----------------------------------------------------------------------------------------
def sample_points_on_rays(h_idxs, w_idxs, depths, intrinsics, poses, N, M, delta, min_dist, gaussian_depths):
    """CPU restatement of sample_points_on_rays (utils.py:458-540, iSDF ray sampling) with the random draw handed in:
    `gaussian_depths` (B,S,M) is what the reference draws per camera with torch.normal(D, sigma) (utils.py:496-498).

    Returns xyz_world (B,S,1+N+M,3) and z (B,S,1+N+M) = [surface depth | N stratified depths in
    [min_dist, D+delta] | M gaussian depths].  The stratified depths use torch.linspace's element formula
    (start + step*i below the midpoint, end - step*(N-1-i) above; step = (end-start)/(N-1) in fp32): the formula
    of ATen's CUDA kernel, which the reference runs in training.  ATen's vectorised CPU kernel evaluates the same
    expression with a different association inside each SIMD chunk and differs from it in the last bit of some
    elements (by an amount that depends on the host's vector width), so this function is pinned against the real
    reference to 1e-6 relative, not bit for bit."""
    B, S = depths.shape
    D = depths.float()
    start = torch.tensor(float(min_dist), dtype=torch.float32)
    end = D + delta                                                      # fp32 tensor + python float (utils.py:493)
    step = (end - start) / (N - 1)
    i = torch.arange(N, dtype=torch.float32)
    lo = start + step.unsqueeze(-1) * i
    hi = end.unsqueeze(-1) - step.unsqueeze(-1) * (N - 1 - i)
    strat = torch.where(torch.arange(N) < N // 2, lo, hi)                # (B,S,N)
    z = torch.cat((D.unsqueeze(-1), strat, gaussian_depths.float()), dim=-1)          # (B,S,1+N+M)
    w_norm = (w_idxs - intrinsics[:, 0, 2].unsqueeze(-1)) / intrinsics[:, 0, 0].unsqueeze(-1)
    h_norm = (h_idxs - intrinsics[:, 1, 2].unsqueeze(-1)) / intrinsics[:, 1, 1].unsqueeze(-1)
    x = w_norm.unsqueeze(-1) * z
    y = h_norm.unsqueeze(-1) * z
    P = z.shape[2] * S
    cam = torch.stack((x.reshape(B, P), y.reshape(B, P), z.reshape(B, P)), dim=-1)
    hom = torch.cat((cam, torch.ones(B, P, 1)), dim=-1)
    world = torch.bmm(poses, hom.permute(0, 2, 1)).permute(0, 2, 1)
    xyz = world[:, :, :3] / world[:, :, 3:]
    return xyz.reshape(B, S, -1, 3), z


def select_valid_depth_pixels(depth, ranks):
    """The deterministic part of sample_valid_depth_pixels (reference src/models/utils.py:340-363):
    idxs[b] = argwhere(depth[b] != 0)[ranks[b]] -> h_idxs (B,S), w_idxs (B,S).  The reference draws
    ranks[b] = randperm(n_valid_b)[:S]; the draw itself is RNG-defined and stays torch's."""
    hs, ws = [], []
    for b in range(depth.shape[0]):
        valid_indices = torch.argwhere(depth[b] != 0)
        sel = valid_indices[ranks[b]]
        hs.append(sel[:, 0])
        ws.append(sel[:, 1])
    return torch.stack(hs), torch.stack(ws)

"""`torch.library` custom-op layer over the C ABI (north_star: "a thin C-ABI torch custom-op layer").

Every hot op of the path is registered as `torch.ops.gennerf_b200.<name>` with
  * the CUDA implementation = the ctypes call into libgennerf_b200.so (gennerf_b200/ops.py),
  * a fake (meta) implementation = output shapes / strides only, so FakeTensor tracing, `torch.compile`
    (the `compile:` switch of configs/model/gen_nerf.yaml:121-122) and `torch.export` see through the ops, and
  * an autograd formula wired to the `gnb_*_bwd` kernels.
The backward ops are custom ops themselves.  The sampler's backward HAS an autograd formula (`sample_features_bwd2`, the
double backward the reference's eikonal / gradient losses need: create_graph=True, src/models/utils.py:636-649); the other
backward ops have none, so differentiating them raises PyTorch's "no autograd formula registered" error instead of
silently dropping a second-order term.  CUDA tensors only; there is no CPU implementation.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

NS = "gennerf_b200"
_op = torch.library.custom_op


def _cl3d_empty(like, B, C, nx, ny, nz):
    return like.new_empty((B, nx, ny, nz, C), dtype=torch.float32).permute(0, 4, 1, 2, 3)


# ------------------------------------------------------------------------------------------------
# lift (utils.py:948-996 + model.py:121-127)
# ------------------------------------------------------------------------------------------------
@_op(f"{NS}::backproject_frames", mutates_args=())
def backproject_frames(features: List[Tensor], projections: Tensor, voxel_dim: List[int], voxel_size: float,
                       origin: List[float], mean: bool) -> Tuple[Tensor, Tensor, Tensor]:
    return ops.backproject_frames(voxel_dim, voxel_size, origin, projections, features, mean=mean)


@backproject_frames.register_fake
def _(features, projections, voxel_dim, voxel_size, origin, mean):
    f = features[0]
    B, C = f.shape[0], f.shape[1]
    nx, ny, nz = voxel_dim
    return (_cl3d_empty(f, B, C, nx, ny, nz), f.new_empty((B, nx, ny, nz), dtype=torch.int32),
            f.new_empty((B, 1, nx, ny, nz), dtype=torch.bool))


@_op(f"{NS}::backproject_frames_bwd", mutates_args=())
def backproject_frames_bwd(grad_volume: Tensor, count: Tensor, projections: Tensor, voxel_dim: List[int], voxel_size: float,
                           origin: List[float], mean: bool, feat_shape: List[int], n_frames: int, nhwc: bool) -> List[Tensor]:
    return ops.backproject_frames_bwd(voxel_dim, voxel_size, origin, projections, grad_volume, tuple(feat_shape), n_frames,
                                      nhwc=nhwc, mean=mean, count=count)


@backproject_frames_bwd.register_fake
def _(grad_volume, count, projections, voxel_dim, voxel_size, origin, mean, feat_shape, n_frames, nhwc):
    B, C, H, W = feat_shape
    if nhwc:
        return [grad_volume.new_empty((B, H, W, C)).permute(0, 3, 1, 2) for _ in range(n_frames)]
    return [grad_volume.new_empty((B, C, H, W)) for _ in range(n_frames)]


def _lift_setup(ctx, inputs, output):
    features, projections, voxel_dim, voxel_size, origin, mean = inputs
    ctx.meta = (list(voxel_dim), float(voxel_size), list(origin), bool(mean), list(features[0].shape), len(features),
                all(f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in features))
    ctx.save_for_backward(output[1], projections)
    ctx.set_materialize_grads(False)


def _lift_bwd(ctx, gvol, _gcount, _gvalid):
    voxel_dim, voxel_size, origin, mean, shape, T, nhwc = ctx.meta
    count, projections = ctx.saved_tensors
    if gvol is None:
        return None, None, None, None, None, None
    grads = backproject_frames_bwd(gvol, count, projections, voxel_dim, voxel_size, origin, mean, shape, T, nhwc)
    return grads, None, None, None, None, None


backproject_frames.register_autograd(_lift_bwd, setup_context=_lift_setup)


# ------------------------------------------------------------------------------------------------
# point-query sampler (model.py:163-204)
# ------------------------------------------------------------------------------------------------
def _planes(p_xz, p_xy, p_yz):
    d = {k: v for k, v in zip(ops.PLANES, (p_xz, p_xy, p_yz)) if v is not None}
    return d or None


@_op(f"{NS}::sample_features", mutates_args=())
def sample_features(xyz: Tensor, volume: Optional[Tensor], p_xz: Optional[Tensor], p_xy: Optional[Tensor], p_yz: Optional[Tensor],
                    voxel_size: float, origin: List[float], padding: float) -> Tensor:
    return ops.sample_features(xyz, volume=volume, planes=_planes(p_xz, p_xy, p_yz), voxel_size=voxel_size, origin=origin,
                               padding=padding)


@sample_features.register_fake
def _(xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding):
    pl = [p for p in (p_xz, p_xy, p_yz) if p is not None]
    c = (volume.shape[1] if volume is not None else 0) + (pl[0].shape[1] if pl else 0)
    return xyz.new_empty((xyz.shape[0], xyz.shape[1], c), dtype=torch.float32)


@_op(f"{NS}::sample_features_bwd", mutates_args=())
def sample_features_bwd(grad_out: Tensor, xyz: Tensor, volume: Optional[Tensor], p_xz: Optional[Tensor], p_xy: Optional[Tensor],
                        p_yz: Optional[Tensor], voxel_size: float, origin: List[float], padding: float, need_xyz: bool,
                        need_volume: bool, need_planes: bool) -> List[Tensor]:
    """-> [grad_xyz, grad_volume, grad_xz, grad_xy, grad_yz]; entries that were not asked for are empty (0-element) tensors."""
    gxyz, gvol, gpl = ops.sample_features_bwd(grad_out, xyz, volume, _planes(p_xz, p_xy, p_yz), voxel_size=voxel_size,
                                              origin=origin, padding=padding, need_volume=need_volume and volume is not None,
                                              need_planes=need_planes, need_xyz=need_xyz)
    gs = [gxyz, gvol] + [(gpl.get(k) if gpl else None) for k in ops.PLANES]
    return [g if g is not None else grad_out.new_empty(0) for g in gs]


@sample_features_bwd.register_fake
def _(grad_out, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding, need_xyz, need_volume, need_planes):
    e = lambda: grad_out.new_empty(0)      # noqa: E731
    out = [torch.empty_like(xyz, dtype=torch.float32) if need_xyz else e(),
           torch.empty_like(volume) if (need_volume and volume is not None) else e()]
    for p in (p_xz, p_xy, p_yz):
        out.append(torch.empty_like(p) if (need_planes and p is not None) else e())
    return out


@_op(f"{NS}::sample_features_bwd2", mutates_args=())
def sample_features_bwd2(grad_out: Tensor, gg_xyz: Tensor, xyz: Tensor, volume: Optional[Tensor], p_xz: Optional[Tensor],
                         p_xy: Optional[Tensor], p_yz: Optional[Tensor], voxel_size: float, origin: List[float], padding: float,
                         need_grad_out: bool, need_xyz: bool, need_volume: bool, need_planes: bool) -> List[Tensor]:
    """Double backward (gnb_sample_features_bwd2) -> [g_grad_out, g_xyz, g_volume, g_xz, g_xy, g_yz]; entries that were not
    asked for are empty.  No autograd formula of its own: a third derivative raises."""
    ggo, gvol, gpl, gx = ops.sample_features_bwd2(grad_out, gg_xyz, xyz, volume, _planes(p_xz, p_xy, p_yz), voxel_size=voxel_size,
                                                  origin=origin, padding=padding, need_grad_out=need_grad_out,
                                                  need_volume=need_volume and volume is not None, need_planes=need_planes,
                                                  need_xyz=need_xyz)
    gs = [ggo, gx, gvol] + [(gpl.get(k) if gpl else None) for k in ops.PLANES]
    return [g if g is not None else grad_out.new_empty(0) for g in gs]


@sample_features_bwd2.register_fake
def _(grad_out, gg_xyz, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding, need_grad_out, need_xyz, need_volume, need_planes):
    e = lambda: grad_out.new_empty(0)      # noqa: E731
    out = [torch.empty_like(grad_out, dtype=torch.float32, memory_format=torch.contiguous_format) if need_grad_out else e(),
           torch.empty_like(xyz, dtype=torch.float32) if need_xyz else e(),
           torch.empty_like(volume) if (need_volume and volume is not None) else e()]
    for p in (p_xz, p_xy, p_yz):
        out.append(torch.empty_like(p) if (need_planes and p is not None) else e())
    return out


def _sample_bwd_setup(ctx, inputs, output):
    grad_out, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding = inputs[:9]
    ctx.save_for_backward(grad_out, xyz, volume, p_xz, p_xy, p_yz)
    ctx.meta = (float(voxel_size), list(origin), float(padding))
    ctx.set_materialize_grads(False)


def _sample_bwd_bwd(ctx, grads):
    """Backward of the sampler's backward (create_graph=True: eikonal / gradient losses, reference utils.py:636-649,
    model.py:385-400).  The coordinate gradient's derivative is one kernel; the (rare) derivatives of grad_volume /
    grad_planes are the sampler and its backward applied to the incoming gradients, because
    grad_volume = sum_q w(xyz_q) grad_out_q is the adjoint of the sampler itself."""
    gg_xyz, gg_vol, gg_xz, gg_xy, gg_yz = grads
    grad_out, xyz, volume, p_xz, p_xy, p_yz = ctx.saved_tensors
    voxel_size, origin, padding = ctx.meta
    need = ctx.needs_input_grad                    # grad_out, xyz, volume, p_xz, p_xy, p_yz, ...
    planes = (p_xz, p_xy, p_yz)
    out = [None] * 6
    if gg_xyz is not None and gg_xyz.numel() > 0 and any(need[:6]):
        r = sample_features_bwd2(grad_out, gg_xyz, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding,
                                 need[0], need[1], need[2], any(need[3:6]))
        for i in range(6):
            if need[i] and r[i].numel() > 0:
                out[i] = r[i]
    ggp = (gg_xz, gg_xy, gg_yz)
    if (gg_vol is not None and gg_vol.numel() > 0) or any(g is not None and g.numel() > 0 for g in ggp):
        zv = None if volume is None else (gg_vol if (gg_vol is not None and gg_vol.numel() > 0) else torch.zeros_like(volume))
        zp = [None if p is None else (g if (g is not None and g.numel() > 0) else torch.zeros_like(p)) for g, p in zip(ggp, planes)]
        if need[0]:
            t = sample_features(xyz, zv, zp[0], zp[1], zp[2], voxel_size, origin, padding)
            out[0] = t if out[0] is None else out[0] + t
        if need[1]:
            t = sample_features_bwd(grad_out, xyz, zv, zp[0], zp[1], zp[2], voxel_size, origin, padding, True, False, False)[0]
            out[1] = t if out[1] is None else out[1] + t
    return tuple(out) + (None,) * 6


sample_features_bwd.register_autograd(_sample_bwd_bwd, setup_context=_sample_bwd_setup)


def _sample_setup(ctx, inputs, output):
    xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding = inputs
    ctx.save_for_backward(xyz, volume, p_xz, p_xy, p_yz)
    ctx.meta = (float(voxel_size), list(origin), float(padding))


def _sample_bwd(ctx, gout):
    xyz, volume, p_xz, p_xy, p_yz = ctx.saved_tensors
    voxel_size, origin, padding = ctx.meta
    need = ctx.needs_input_grad
    g = sample_features_bwd(gout, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding, need[0], need[1], any(need[2:5]))
    pick = lambda t, want: t if (want and t.numel() > 0) else None     # noqa: E731
    return (pick(g[0], need[0]), pick(g[1], need[1]), pick(g[2], need[2]), pick(g[3], need[3]), pick(g[4], need[4]),
            None, None, None)


sample_features.register_autograd(_sample_bwd, setup_context=_sample_setup)


# ------------------------------------------------------------------------------------------------
# triplane scatter-mean and local pooling (pointnet.py:72-121)
# ------------------------------------------------------------------------------------------------
@_op(f"{NS}::scatter_mean_planes", mutates_args=())
def scatter_mean_planes(p: Tensor, c: Tensor, reso: int, padding: float, mode: str) -> Tuple[Tensor, Tensor]:
    return ops.scatter_mean_planes(p, c, reso, padding, mode)


@scatter_mean_planes.register_fake
def _(p, c, reso, padding, mode):
    B, Cp = c.shape[0], c.shape[2]
    return (c.new_empty((3, B, reso, reso, Cp), dtype=torch.float32).permute(0, 1, 4, 2, 3),
            c.new_empty((3, B, reso, reso), dtype=torch.int32))


@_op(f"{NS}::scatter_mean_planes_bwd", mutates_args=())
def scatter_mean_planes_bwd(p: Tensor, grad_planes: Tensor, count: Tensor, padding: float) -> Tensor:
    return ops.scatter_mean_planes_bwd(p, grad_planes, count, padding)


@scatter_mean_planes_bwd.register_fake
def _(p, grad_planes, count, padding):
    return p.new_empty((p.shape[0], p.shape[1], grad_planes.shape[2]), dtype=torch.float32)


def _scatter_setup(ctx, inputs, output):
    p, c, reso, padding, mode = inputs
    ctx.save_for_backward(p, output[1])
    ctx.padding = float(padding)
    ctx.set_materialize_grads(False)


def _scatter_bwd(ctx, gplanes, _gcount):
    p, count = ctx.saved_tensors
    if gplanes is None:
        return None, None, None, None, None
    return None, scatter_mean_planes_bwd(p, gplanes, count, ctx.padding), None, None, None


scatter_mean_planes.register_autograd(_scatter_bwd, setup_context=_scatter_setup)


@_op(f"{NS}::pool_local", mutates_args=())
def pool_local(p: Tensor, c: Tensor, reso: int, padding: float, scatter_type: str) -> Tuple[Tensor, Tensor]:
    """-> (pooled (B,N,Hd), scratch): `scratch` is the arg-max / count state the backward kernel reads."""
    return ops.pool_local_fwd_keep(p, c, reso, padding, scatter_type)


@pool_local.register_fake
def _(p, c, reso, padding, scatter_type):
    from ._lib import lib
    B, N, Hd = c.shape
    n = lib().gnb_pool_scratch_bytes(int(B), int(N), int(Hd), int(reso))       # host-side size function, no GPU work
    return torch.empty_like(c, dtype=torch.float32), c.new_empty((n,), dtype=torch.uint8)


@_op(f"{NS}::pool_local_bwd", mutates_args=())
def pool_local_bwd(p: Tensor, c: Tensor, grad_out: Tensor, scratch: Tensor, reso: int, padding: float, scatter_type: str) -> Tensor:
    return ops.pool_local_bwd(p, c, grad_out, scratch, reso, padding, scatter_type)


@pool_local_bwd.register_fake
def _(p, c, grad_out, scratch, reso, padding, scatter_type):
    return torch.empty_like(c, dtype=torch.float32)


def _pool_setup(ctx, inputs, output):
    p, c, reso, padding, scatter_type = inputs
    ctx.save_for_backward(p, c, output[1])
    ctx.meta = (int(reso), float(padding), str(scatter_type))
    ctx.set_materialize_grads(False)


def _pool_bwd(ctx, gout, _gscratch):
    p, c, scratch = ctx.saved_tensors
    if gout is None:
        return None, None, None, None, None
    reso, padding, scatter_type = ctx.meta
    return None, pool_local_bwd(p, c, gout, scratch, reso, padding, scatter_type), None, None, None


pool_local.register_autograd(_pool_bwd, setup_context=_pool_setup)


# ------------------------------------------------------------------------------------------------
# decoder (positional_encoding.py:28-40, resnetfc.py:134-189, heads3d.py:36-50) and the fused query (model.py:207-248)
# ------------------------------------------------------------------------------------------------
def mlp_keys(n_blocks):
    """Fixed order of a ResnetFC state_dict (reference resnetfc.py:66-132) in the Tensor[] argument of the decoder ops."""
    keys = ["lin_in.weight", "lin_in.bias", "lin_out.weight", "lin_out.bias", "alpha"]
    for i in range(n_blocks):
        keys += [f"lin_z.{i}.weight", f"lin_z.{i}.bias", f"blocks.{i}.fc_0.weight", f"blocks.{i}.fc_0.bias",
                 f"blocks.{i}.fc_1.weight", f"blocks.{i}.fc_1.bias"]
    return keys


def mlp_param_list(sd, n_blocks):
    return [sd[k] for k in mlp_keys(n_blocks)]


def _weights(params, head_w, head_b, n_blocks, d_geo, use_code, num_freqs, freq_factor, include_input, packed, precision):
    sd = dict(zip(mlp_keys(n_blocks), params))
    dw = ops.DecoderWeights(sd, head_w, head_b, n_blocks=n_blocks, d_geo=d_geo, use_code=bool(use_code), num_freqs=num_freqs,
                            freq_factor=freq_factor, include_input=include_input, device=params[0].device)
    if not use_code:
        dw.w.use_code, dw.w.d_code = 0, 3
    if packed is not None and precision != "fp32":
        dw.packed, dw.packed_dtype = packed, precision
        dw.w.tc_dtype = {"fp16": 0, "bf16": 1}[precision]
    return dw


@_op(f"{NS}::decode", mutates_args=())
def decode(xyz: Tensor, feat: Tensor, params: List[Tensor], head_w: Tensor, head_b: Tensor, packed: Optional[Tensor],
           n_blocks: int, d_geo: int, use_code: bool, num_freqs: int, freq_factor: float, include_input: bool,
           precision: str) -> Tuple[Tensor, Tensor]:
    dw = _weights(params, head_w, head_b, n_blocks, d_geo, use_code, num_freqs, freq_factor, include_input, packed, precision)
    return ops.decode(dw, xyz, feat, precision)


@decode.register_fake
def _(xyz, feat, params, head_w, head_b, packed, n_blocks, d_geo, use_code, num_freqs, freq_factor, include_input, precision):
    lead = xyz.shape[:-1]
    d_out = params[2].shape[0]
    return xyz.new_empty((*lead, d_out), dtype=torch.float32), xyz.new_empty((*lead, 1), dtype=torch.float32)


@_op(f"{NS}::query_fused", mutates_args=())
def query_fused(xyz: Tensor, volume: Optional[Tensor], p_xz: Optional[Tensor], p_xy: Optional[Tensor], p_yz: Optional[Tensor],
                params: List[Tensor], head_w: Tensor, head_b: Tensor, packed: Optional[Tensor], voxel_size: float,
                origin: List[float], padding: float, n_blocks: int, d_geo: int, use_code: bool, num_freqs: int,
                freq_factor: float, include_input: bool, precision: str) -> Tuple[Tensor, Tensor, Tensor]:
    dw = _weights(params, head_w, head_b, n_blocks, d_geo, use_code, num_freqs, freq_factor, include_input, packed, precision)
    return ops.query_fused(dw, xyz, volume=volume, planes=_planes(p_xz, p_xy, p_yz), voxel_size=voxel_size, origin=origin,
                           padding=padding, precision=precision)


@query_fused.register_fake
def _(xyz, volume, p_xz, p_xy, p_yz, params, head_w, head_b, packed, voxel_size, origin, padding, n_blocks, d_geo, use_code,
      num_freqs, freq_factor, include_input, precision):
    B, Q = xyz.shape[0], xyz.shape[1]
    pl = [p for p in (p_xz, p_xy, p_yz) if p is not None]
    c = (volume.shape[1] if volume is not None else 0) + (pl[0].shape[1] if pl else 0)
    f32 = dict(dtype=torch.float32)
    return xyz.new_empty((B, Q, params[2].shape[0]), **f32), xyz.new_empty((B, Q, 1), **f32), xyz.new_empty((B, Q, c), **f32)


OPS = ("backproject_frames", "backproject_frames_bwd", "sample_features", "sample_features_bwd", "scatter_mean_planes",
       "scatter_mean_planes_bwd", "pool_local", "pool_local_bwd", "decode", "query_fused")

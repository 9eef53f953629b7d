"""torch.autograd wiring of the backward kernels.

Each Function's forward is the same C-ABI forward the inference path uses; backward calls the
matching `gnb_*_bwd` entry point.  Scatter-add gradients are atomic (order-nondeterministic), like
the reference's CUDA index_put_ / grid_sampler backward.
"""
import torch

from . import ops

PLANES = ops.PLANES


class LiftFn(torch.autograd.Function):
    """backproject + accumulate over frames (utils.py:948-996, model.py:121-127)."""

    @staticmethod
    def forward(ctx, voxel_dim, voxel_size, origin, projections, mean, *features):
        volume, count, valid = ops.backproject_frames(voxel_dim, voxel_size, origin, projections, features, mean=mean)
        ctx.meta = (tuple(voxel_dim), voxel_size, origin, torch.as_tensor(projections).detach().cpu(), bool(mean),
                    tuple(features[0].shape), len(features),
                    all(f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in features))
        ctx.save_for_backward(count)
        ctx.mark_non_differentiable(count, valid)
        return volume, count, valid

    @staticmethod
    def backward(ctx, gvol, _gc, _gv):
        voxel_dim, voxel_size, origin, P, mean, shape, T, nhwc = ctx.meta
        (count,) = ctx.saved_tensors
        grads = ops.backproject_frames_bwd(voxel_dim, voxel_size, origin, P, gvol, shape, T, nhwc=nhwc, mean=mean, count=count)
        return (None, None, None, None, None) + tuple(grads)


def backproject_frames(voxel_dim, voxel_size, origin, projections, features, mean=False):
    return LiftFn.apply(voxel_dim, voxel_size, origin, projections, mean, *features)


class SampleFn(torch.autograd.Function):
    """map_features (model.py:163-204): trilinear volume + 3 bilinear planes."""

    @staticmethod
    def forward(ctx, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding):
        planes = {k: v for k, v in zip(PLANES, (p_xz, p_xy, p_yz)) if v is not None}
        out = ops.sample_features(xyz, volume=volume, planes=planes or None, voxel_size=voxel_size, origin=origin, padding=padding)
        ctx.save_for_backward(xyz, volume, p_xz, p_xy, p_yz)
        ctx.meta = (voxel_size, origin, padding)
        return out

    @staticmethod
    def backward(ctx, gout):
        xyz, volume, p_xz, p_xy, p_yz = ctx.saved_tensors
        voxel_size, origin, padding = ctx.meta
        planes = {k: v for k, v in zip(PLANES, (p_xz, p_xy, p_yz)) if v is not None}
        need = ctx.needs_input_grad
        gxyz, gvol, gpl = ops.sample_features_bwd(gout, xyz, volume, planes or None, voxel_size=voxel_size, origin=origin,
                                                  padding=padding, need_volume=need[1], need_planes=any(need[2:5]),
                                                  need_xyz=need[0])
        gp = [gpl.get(k) if (gpl and need[2 + i]) else None for i, k in enumerate(PLANES)]
        return gxyz, gvol, gp[0], gp[1], gp[2], None, None, None


def sample_features(xyz, volume=None, planes=None, voxel_size=0.04, origin=None, padding=0.1):
    planes = planes or {}
    return SampleFn.apply(xyz, volume, planes.get("xz"), planes.get("xy"), planes.get("yz"), voxel_size, origin, padding)


class ScatterMeanFn(torch.autograd.Function):
    """generate_plane_features x3 (pointnet.py:72-89).  No gradient reaches the point positions: the
    reference indexes with integer cell ids."""

    @staticmethod
    def forward(ctx, p, c, reso, padding, mode):
        planes, count = ops.scatter_mean_planes(p, c, reso, padding, mode)
        ctx.save_for_backward(p, count)
        ctx.meta = (padding,)
        ctx.mark_non_differentiable(count)
        return planes, count

    @staticmethod
    def backward(ctx, gplanes, _gcount):
        p, count = ctx.saved_tensors
        return None, ops.scatter_mean_planes_bwd(p, gplanes, count, ctx.meta[0]), None, None, None


def scatter_mean_planes(p, c, reso, padding=0.1, mode="atomic"):
    return ScatterMeanFn.apply(p, c, reso, padding, mode)


class PoolLocalFn(torch.autograd.Function):
    """pool_local (pointnet.py:105-121)."""

    @staticmethod
    def forward(ctx, p, c, reso, padding, scatter_type):
        out, scratch = ops.pool_local_fwd_keep(p, c, reso, padding, scatter_type)
        ctx.save_for_backward(p, c, scratch)
        ctx.meta = (reso, padding, scatter_type)
        return out

    @staticmethod
    def backward(ctx, gout):
        p, c, scratch = ctx.saved_tensors
        reso, padding, scatter_type = ctx.meta
        return None, ops.pool_local_bwd(p, c, gout, scratch, reso, padding, scatter_type), None, None, None


def pool_local(p, c, reso, padding=0.1, scatter_type="max"):
    return PoolLocalFn.apply(p, c, reso, padding, scatter_type)

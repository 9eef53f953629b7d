"""Differentiable entry points of the path.

The forward is the same C-ABI kernel the inference path uses, the backward calls the matching `gnb_*_bwd` kernel.
Scatter-add gradients are atomic (order-nondeterministic), like the reference's CUDA index_put_ / grid_sampler backward.

Two spellings of the same formulas:
  * the `torch.ops.gennerf_b200.*` custom ops (gennerf_b200/torch_ops.py: fake kernels + registered autograd) whenever a
    tracer or a dispatch mode is active (torch.compile, FakeTensor, torch.export, make_fx), and
  * plain `torch.autograd.Function`s around the same `ops.*` calls in eager mode: a torch.library Python op costs 45-90 us
    of host time per forward call and about as much again per backward call (tools/host_overhead_probe.py), and a training
    step at the reference's sizes is bound by the host's launch rate (14 such calls per step).
The sampler's backward is differentiable again (gnb_sample_features_bwd2: the eikonal / gradient losses' create_graph=True,
reference utils.py:636-649): under grad mode the eager Function hands its backward to the custom op that carries that
formula.  The other backwards are once-differentiable (linear in the incoming gradient, no dependence on differentiable
saved tensors): a double backward through them raises instead of dropping terms.
"""
import torch
from torch.autograd.function import once_differentiable
from torch.utils._python_dispatch import _get_current_dispatch_mode

from . import ops
from . import torch_ops as T

PLANES = ops.PLANES
EAGER_FUNCTIONS = True        # False: always go through the custom ops (what the tests compare the eager path with)


def _eager():
    return EAGER_FUNCTIONS and not torch.compiler.is_compiling() and _get_current_dispatch_mode() is None


class _Lift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P, voxel_dim, voxel_size, origin, mean, *features):
        vol, cnt, valid = ops.backproject_frames(voxel_dim, voxel_size, origin, P, list(features), mean=mean)
        ctx.meta = (voxel_dim, voxel_size, origin, mean, tuple(features[0].shape), len(features),
                    all(f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in features))
        ctx.save_for_backward(cnt, P)
        ctx.mark_non_differentiable(cnt, valid)
        ctx.set_materialize_grads(False)
        return vol, cnt, valid

    @staticmethod
    @once_differentiable
    def backward(ctx, gvol, _gcount, _gvalid):
        voxel_dim, voxel_size, origin, mean, shape, n_frames, nhwc = ctx.meta
        if gvol is None:
            return (None,) * (5 + n_frames)
        cnt, P = ctx.saved_tensors
        grads = ops.backproject_frames_bwd(voxel_dim, voxel_size, origin, P, gvol, shape, n_frames, nhwc=nhwc, mean=mean, count=cnt)
        return (None, None, None, None, None, *grads)


class _Sample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding):
        planes = {k: v for k, v in zip(PLANES, (p_xz, p_xy, p_yz)) if v is not None} or None
        out = ops.sample_features(xyz, volume=volume, planes=planes, voxel_size=voxel_size, origin=origin, padding=padding)
        ctx.save_for_backward(xyz, volume, p_xz, p_xy, p_yz)
        ctx.meta = (voxel_size, origin, padding)
        return out

    @staticmethod
    def backward(ctx, gout):
        xyz, volume, p_xz, p_xy, p_yz = ctx.saved_tensors
        voxel_size, origin, padding = ctx.meta
        need = ctx.needs_input_grad
        if torch.is_grad_enabled():
            # create_graph=True: the custom op carries the double-backward formula (torch_ops._sample_bwd_bwd)
            g = T.sample_features_bwd(gout, xyz, volume, p_xz, p_xy, p_yz, voxel_size, origin, padding, need[0], need[1], any(need[2:5]))
            pick = lambda t, want: t if (want and t.numel() > 0) else None     # noqa: E731
            return (pick(g[0], need[0]), pick(g[1], need[1]), pick(g[2], need[2]), pick(g[3], need[3]), pick(g[4], need[4]),
                    None, None, None)
        planes = {k: v for k, v in zip(PLANES, (p_xz, p_xy, p_yz)) if v is not None} or None
        gxyz, gvol, gpl = ops.sample_features_bwd(gout, xyz, volume, planes, voxel_size=voxel_size, origin=origin, padding=padding,
                                                  need_volume=need[1] and volume is not None, need_planes=any(need[2:5]), need_xyz=need[0])
        gp = [(gpl.get(k) if (gpl and need[2 + i]) else None) for i, k in enumerate(PLANES)]
        return (gxyz if need[0] else None, gvol if need[1] else None, gp[0], gp[1], gp[2], None, None, None)


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, c, reso, padding, mode):
        planes, count = ops.scatter_mean_planes(p, c, reso, padding, mode)
        ctx.save_for_backward(p, count)
        ctx.padding = padding
        ctx.mark_non_differentiable(count)
        ctx.set_materialize_grads(False)
        return planes, count

    @staticmethod
    @once_differentiable
    def backward(ctx, gplanes, _gcount):
        if gplanes is None:
            return None, None, None, None, None
        p, count = ctx.saved_tensors
        return None, ops.scatter_mean_planes_bwd(p, gplanes, count, ctx.padding), None, None, None


class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, c, reso, padding, scatter_type):
        pooled, scratch = ops.pool_local_fwd_keep(p, c, reso, padding, scatter_type)
        ctx.save_for_backward(p, c, scratch)
        ctx.meta = (reso, padding, scatter_type)
        ctx.set_materialize_grads(False)
        return pooled

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        if gout is None:
            return None, None, None, None, None
        p, c, scratch = ctx.saved_tensors
        reso, padding, scatter_type = ctx.meta
        return None, ops.pool_local_bwd(p, c, gout, scratch, reso, padding, scatter_type), None, None, None


def backproject_frames(voxel_dim, voxel_size, origin, projections, features, mean=False):
    """backproject + accumulate over frames (utils.py:948-996, model.py:121-127) -> volume, count, valid."""
    P = torch.as_tensor(projections).detach()
    vd, vs, o3 = [int(d) for d in voxel_dim], float(voxel_size), ops._origin3(origin)
    if _eager():
        return _Lift.apply(P, vd, vs, o3, bool(mean), *features)
    return T.backproject_frames(list(features), P, vd, vs, o3, bool(mean))


def sample_features(xyz, volume=None, planes=None, voxel_size=0.04, origin=None, padding=0.1):
    """map_features (model.py:163-204): trilinear volume + 3 bilinear planes."""
    planes = planes or {}
    fn = _Sample.apply if _eager() else T.sample_features
    return fn(xyz, volume, planes.get("xz"), planes.get("xy"), planes.get("yz"), float(voxel_size), ops._origin3(origin), float(padding))


def scatter_mean_planes(p, c, reso, padding=0.1, mode="atomic"):
    """generate_plane_features x3 (pointnet.py:72-89).  No gradient reaches the point positions: the reference indexes
    with integer cell ids."""
    fn = _Scatter.apply if _eager() else T.scatter_mean_planes
    return fn(p, c, int(reso), float(padding), str(mode))


def pool_local(p, c, reso, padding=0.1, scatter_type="max"):
    """pool_local (pointnet.py:105-121)."""
    if _eager():
        return _Pool.apply(p, c, int(reso), float(padding), str(scatter_type))
    return T.pool_local(p, c, int(reso), float(padding), str(scatter_type))[0]

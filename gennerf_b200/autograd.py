"""Differentiable entry points of the path.

Thin wrappers over the `torch.ops.gennerf_b200.*` custom ops (gennerf_b200/torch_ops.py): the forward is the same C-ABI
kernel the inference path uses, the autograd formula calls the matching `gnb_*_bwd` kernel.  Scatter-add gradients are
atomic (order-nondeterministic), like the reference's CUDA index_put_ / grid_sampler backward.  The sampler's
backward is differentiable again (gnb_sample_features_bwd2: the eikonal / gradient losses' create_graph=True, reference
utils.py:636-649); the other formulas are once-differentiable: a double backward raises instead of dropping terms.
"""
import torch

from . import ops
from . import torch_ops as T

PLANES = ops.PLANES


def backproject_frames(voxel_dim, voxel_size, origin, projections, features, mean=False):
    """backproject + accumulate over frames (utils.py:948-996, model.py:121-127) -> volume, count, valid."""
    P = torch.as_tensor(projections).detach()
    return T.backproject_frames(list(features), P, [int(d) for d in voxel_dim], float(voxel_size), ops._origin3(origin), bool(mean))


def sample_features(xyz, volume=None, planes=None, voxel_size=0.04, origin=None, padding=0.1):
    """map_features (model.py:163-204): trilinear volume + 3 bilinear planes."""
    planes = planes or {}
    return T.sample_features(xyz, volume, planes.get("xz"), planes.get("xy"), planes.get("yz"), float(voxel_size),
                             ops._origin3(origin), float(padding))


def scatter_mean_planes(p, c, reso, padding=0.1, mode="atomic"):
    """generate_plane_features x3 (pointnet.py:72-89).  No gradient reaches the point positions: the reference indexes
    with integer cell ids."""
    return T.scatter_mean_planes(p, c, int(reso), float(padding), str(mode))


def pool_local(p, c, reso, padding=0.1, scatter_type="max"):
    """pool_local (pointnet.py:105-121)."""
    return T.pool_local(p, c, int(reso), float(padding), str(scatter_type))[0]

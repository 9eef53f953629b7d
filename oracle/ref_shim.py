"""Import shim for the real reference tree (TEST INFRASTRUCTURE, container-only).

The reference (mrchris7/gen-nerf) is pure Python; it lives read-only at
/root/reference in the build container and does NOT exist on the GPU box.  This
module makes `src.models.*` importable on CPU by stubbing the third-party
packages that are absent here (lightning, hydra, trimesh, open3d, ...), and by
providing the upstream semantics of `torch_scatter` (rusty1s/pytorch_scatter,
unpinned in the reference: README.md:42) as a pure-torch stand-in.

It is used by
  * tests/golden/make_golden.py   -- generates the committed golden vectors
  * tests/test_oracle_pinning.py  -- asserts oracle == real reference (skipped
                                     when /root/reference is absent)
Nothing in the product path may import this file.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import torch

REFERENCE_ROOT = os.environ.get("GENNERF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


# --- torch_scatter stand-in (upstream torch_scatter/scatter.py semantics) -------------
def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    """Upstream returns (out, argmax); cells no point falls into hold 0 and
    argmax == src.size(dim)."""
    index_b = _broadcast(index, src, dim)
    size = list(src.size())
    if out is not None:
        size[dim] = out.size(dim)
    elif dim_size is not None:
        size[dim] = dim_size
    else:
        size[dim] = int(index.max()) + 1
    res = torch.zeros(size, dtype=src.dtype, device=src.device)
    res.scatter_reduce_(dim, index_b, src, reduce="amax", include_self=False)
    # argmax: first position attaining the max (upstream CPU kernel keeps the first)
    n = src.size(dim)
    pos = torch.arange(n, device=src.device)
    shape = [1] * src.dim()
    shape[dim] = n
    pos = pos.view(shape).expand(src.size())
    hit = src == res.gather(dim, index_b)
    cand = torch.where(hit, pos, torch.full_like(pos, n))
    arg = torch.full(size, n, dtype=torch.long, device=src.device)
    arg.scatter_reduce_(dim, index_b, cand, reduce="amin", include_self=True)
    return res, arg


_STUBS = [
    "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "skimage", "skimage.measure", "trimesh",
    "open3d", "pyrender", "lightning", "lightning.pytorch", "lightning.pytorch.loggers",
    "lightning.pytorch.loggers.wandb", "lightning.pytorch.utilities",
    "lightning.pytorch.utilities.rank_zero", "lightning.pytorch.trainer",
    "lightning.pytorch.callbacks", "lightning_utilities", "lightning_utilities.core",
    "lightning_utilities.core.rank_zero", "hydra", "hydra.utils", "hydra.core",
    "hydra.core.hydra_config", "rich.prompt", "omegaconf", "pytorch_lightning", "rootutils",
    "torch_scatter", "torch_cluster", "wandb", "cv2",
]


class _LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass


class _Trivial:
    def __init__(self, *a, **k):
        pass


_installed = False


def install():
    """Idempotently make `import src.models...` work against REFERENCE_ROOT."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = MagicMock(name=name)
    L = sys.modules["lightning"]
    if isinstance(L, MagicMock):
        L.LightningModule = _LightningModule
        L.LightningDataModule = object
        L.Callback = object
        lp = sys.modules["lightning.pytorch"]
        lp.LightningModule = _LightningModule
        lp.Callback = object
        sys.modules["lightning.pytorch.loggers"].Logger = _Trivial
        sys.modules["lightning.pytorch.loggers"].WandbLogger = _Trivial
        sys.modules["lightning.pytorch.loggers.wandb"].WandbLogger = _Trivial
        sys.modules["lightning.pytorch.callbacks"].Callback = object
    ts = sys.modules["torch_scatter"]
    if isinstance(ts, MagicMock):
        ts.scatter_mean = scatter_mean
        ts.scatter_max = scatter_max
        ts.scatter_sum = scatter_sum
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class AttrDict(dict):
    """Attribute-access config object (OmegaConf is absent here)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, list):
        return [to_attr(v) for v in d]
    return d


def load_model_cfg(name="gen_nerf", **data):
    """Load configs/model/<name>.yaml from the reference, resolving the ${data.*}
    interpolations from `data` (OmegaConf stand-in)."""
    import yaml

    with open(os.path.join(REFERENCE_ROOT, "configs", "model", f"{name}.yaml")) as f:
        raw = yaml.safe_load(f)
    defaults = dict(voxel_size=0.04, voxel_dim_train=[160, 160, 64], voxel_dim_val=[256, 256, 96],
                    voxel_dim_test=[416, 416, 128])
    defaults.update(data)

    def resolve(v):
        if isinstance(v, dict):
            return {k: resolve(x) for k, x in v.items()}
        if isinstance(v, str) and v.startswith("${data."):
            return defaults[v[len("${data."):-1]]
        if isinstance(v, str) and v.startswith("${"):
            return None
        return v

    return to_attr(resolve(raw))


def ref_modules():
    """Returns the reference modules of the hot path as a namespace."""
    install()
    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("src.models.utils")
    ns.tsdf = importlib.import_module("src.data.tsdf")
    ns.pointnet = importlib.import_module("src.models.components.pointnet")
    ns.resnetfc = importlib.import_module("src.models.components.resnetfc")
    ns.posenc = importlib.import_module("src.models.components.positional_encoding")
    ns.heads3d = importlib.import_module("src.models.components.heads3d")
    return ns


def ref_gennerf():
    """Returns the reference GenNerf class with SpatialEncoder replaced by a stub that
    passes features through (the 2D CNN is outside the path and downloads weights)."""
    install()
    model = importlib.import_module("src.models.model")

    class _PassThroughSpatial(torch.nn.Module):
        """image IS the synthetic feature map (B,C,H,W)."""

        @classmethod
        def from_conf(cls, cfg):
            return cls()

        def forward(self, image):
            return image

    model.SpatialEncoder = _PassThroughSpatial
    return model.GenNerf

"""Out-of-bounds check of our own (compute-sanitizer is closed on this GPU pool, profiles/r2_sanitizer_unavailable.txt):
every CUDA buffer gennerf_b200.ops allocates for a kernel (outputs, scratch, operand images) is placed between two 4 KB guard
bands filled with a canary byte; after the op the bands must be untouched.  Shapes are small and deliberately ragged
(sizes that are not multiples of a tile, a warp or a vector)."""
import numpy as np
import pytest
import torch

from gennerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS = 0.04
ORIGIN = torch.zeros(1, 3)
GUARD, CANARY = 4096, 0xA5


class GuardedTorch:
    """Stands in for the `torch` module inside gennerf_b200.ops: empty / zeros on a CUDA device come with guard bands."""

    def __init__(self):
        self.bands = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, size, dtype, device, zero):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        dtype = dtype or torch.float32
        if device is None or torch.device(device).type != "cuda":
            return (torch.zeros if zero else torch.empty)(size, dtype=dtype, device=device)
        item = torch.empty((), dtype=dtype).element_size()
        n = int(np.prod(size)) * item if len(size) else item
        pad = (-n) % 256
        raw = torch.full((GUARD + n + pad + GUARD,), CANARY, dtype=torch.uint8, device=device)
        self.bands.append((raw, n))
        t = raw[GUARD:GUARD + n].view(dtype).view(size)
        if zero:
            t.zero_()
        return t

    def empty(self, *size, dtype=None, device=None, **kw):
        return self._alloc(size, dtype, device, False)

    def zeros(self, *size, dtype=None, device=None, **kw):
        return self._alloc(size, dtype, device, True)

    def check(self):
        torch.cuda.synchronize()
        for raw, n in self.bands:
            lo, hi = raw[:GUARD], raw[GUARD + n:]
            assert bool((lo == CANARY).all()), f"write BEFORE a {n}-byte buffer"
            assert bool((hi == CANARY).all()), f"write PAST a {n}-byte buffer"
        k = len(self.bands)
        self.bands = []
        return k


@pytest.fixture
def guarded(monkeypatch):
    from gennerf_b200 import ops
    g = GuardedTorch()
    monkeypatch.setattr(ops, "torch", g)
    return g


def test_lift_scatter_pool_guard_bands(guarded):
    from gennerf_b200 import ops
    g = S.gen(61)
    vd = (13, 11, 7)
    T, C, H, W = 3, 12, 23, 31
    P = S.projections(T, H, W, vd, VS, g, pull_back=0.8).unsqueeze(0)
    feats = [f.to(DEV) for f in S.frame_features(T, C, H, W, g)]
    vol, cnt, valid = ops.backproject_frames(vd, VS, ORIGIN, P, feats)
    assert guarded.check() >= 3
    ops.backproject_frames_bwd(vd, VS, ORIGIN, P, torch.randn(1, C, *vd, generator=g).to(DEV), (1, C, H, W), T)
    assert guarded.check() >= 1
    N, Cp, R = 1003, 12, 19
    p = S.plane_points(N, g, "unit").to(DEV)
    c = torch.randn(1, N, Cp, generator=g).to(DEV)
    for mode in ("atomic", "deterministic", "sum"):
        ops.scatter_mean_planes(p, c, R, 0.1, mode)
        assert guarded.check() >= 2
    for st in ("max", "mean"):
        ops.pool_local(p, c, R, 0.1, st)
        assert guarded.check() >= 1
    ops.plane_coords(p, 0.1, R)
    assert guarded.check() >= 2


@pytest.mark.parametrize("Q", [1, 127, 129, 4099, 70001])
def test_query_paths_guard_bands(guarded, Q):
    from gennerf_b200 import ops
    g = S.gen(62)
    vd = (24, 24, 12)
    Cv, Cp, R = 32, 32, 17
    xyz = S.query_points(Q, vd, VS, g).to(DEV)
    vol = torch.randn(1, *vd, Cv, generator=g).to(DEV).permute(0, 4, 1, 2, 3)
    planes = {k: torch.randn(1, Cp, R, R, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) for k in ops.PLANES}
    w, hw, hb = S.decoder_weights(g, Cv + Cp, 15, 512, 5, 64, 32)
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=DEV)
    kw = dict(volume=vol, planes=planes, voxel_size=VS, origin=ORIGIN, padding=0.1)
    feat = ops.sample_features(xyz, **kw)
    assert guarded.check() >= 1
    if Q >= 4099:
        ops.sample_features(xyz, binned=True, **kw)
        assert guarded.check() >= 2
    a = ops.query_fused(dw, xyz, want_feat=True, mode="fused", presort=False, **kw)
    assert guarded.check() >= 3
    b = ops.query_image(dw, xyz, want_feat=True, chunk=1000 if Q > 2000 else None, **kw)
    assert guarded.check() >= 4
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    if Q >= 4099:
        ops.query_fused(dw, xyz, want_feat=False, mode="fused", presort=True, **kw)
        assert guarded.check() >= 3
    ops.decode(dw, xyz, feat, "fp16")
    ops.decode(dw, xyz, feat, "fp32")
    assert guarded.check() >= 4
    ops.sample_features_bwd(torch.randn(1, Q, Cv + Cp, generator=g).to(DEV), xyz, **kw)
    assert guarded.check() >= 1


def test_dense_grid_query_guard_bands(guarded):
    from gennerf_b200 import ops
    g = S.gen(63)
    vd = (10, 9, 7)
    vol = torch.randn(1, *vd, 32, generator=g).to(DEV).permute(0, 4, 1, 2, 3)
    w, hw, hb = S.decoder_weights(g, 32, 15, 512, 5, 64, 32)
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=5, d_geo=32, device=DEV)
    axes = [torch.linspace(0, (n - 1) * VS, 2 * n - 1, device=DEV) for n in vd]
    grid = tuple(2 * n - 1 for n in vd)
    ops.query_grid_fused(dw, grid, axes, volume=vol, voxel_size=VS, origin=ORIGIN, want_out=True)
    assert guarded.check() >= 2


def test_training_decoder_backward_guard_bands(guarded):
    """gnb_decode_tc_save + gnb_decode_train_bwd (and the stand-alone link / head kernels) at ragged row counts: activations,
    gradient buffers and the workspace sit between canaries."""
    from gennerf_b200 import ops
    g = S.gen(67)
    n = 333
    w, hw, hb = S.decoder_weights(g, 24, 15, 128, 3, 40, 12)
    w = {k: v.to(DEV) for k, v in w.items()}
    dw = ops.DecoderWeights(w, hw, hb, n_blocks=3, d_geo=12, use_code=2, num_freqs=0, freq_factor=0.0, include_input=False,
                            d_code=15, device=DEV, alpha_on_device=True)
    code = torch.randn(n, 15, generator=g).to(DEV)
    feat = torch.randn(n, 24, generator=g).to(DEV)
    out, tsdf, acts = ops.decode_save(dw, code, feat, "fp16")
    assert guarded.check() >= 3
    g_out = torch.randn(n, 40, generator=g).to(DEV)
    g_tsdf = torch.randn(n, 1, generator=g).to(DEV)
    grads, d_hw, d_hb, g_code, g_feat = ops.decode_train_bwd(dw, code, feat, out, tsdf, acts, g_out, g_tsdf)
    assert guarded.check() >= 5
    assert all(torch.isfinite(v).all() for v in grads.values()) and torch.isfinite(g_code).all() and torch.isfinite(g_feat).all()
    pre = torch.randn(n, 128, generator=g).to(DEV)
    col = ops.torch.zeros(128, device=DEV)
    o, a32 = ops.mlp_grad_link(pre, acts[1], colsum=col, want_act32=True)
    G, *_ = ops.mlp_grad_head(g_out, g_tsdf, out, tsdf, hw.to(DEV), 12)
    assert guarded.check() >= 4

"""Host-side checks that need no GPU: the C-ABI library builds for sm_100a, loads, exports every
symbol include/gennerf_b200.h declares, the ctypes structs match the compiled layout, and the
product refuses to run without CUDA (no CPU fallback)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    from gennerf_b200 import build
    build.build()
    from gennerf_b200 import _lib
    return _lib


def test_header_symbols_exported(library):
    header = open(os.path.join(ROOT, "include", "gennerf_b200.h")).read()
    declared = set(re.findall(r"\b(gnb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    L = library.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(library.SIGNATURES), "ctypes binding and header disagree"
    assert L.gnb_version() == 100


def test_struct_layout_matches(library):
    import ctypes as C
    L = library.lib()
    for which, st in enumerate((library.GnbLiftParams, library.GnbSampleParams, library.GnbDecoderWeights, library.GnbFusionParams,
                                library.GnbDecoderGrads)):
        assert L.gnb_struct_size(which) == C.sizeof(st)


def test_sass_is_sm100a(library):
    out = subprocess.run(["cuobjdump", "-lelf", library.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors_do_not_touch_the_gpu(library):
    import ctypes as C
    L = library.lib()
    p = library.GnbLiftParams()
    assert L.gnb_backproject_frames(C.byref(p), None) == -1
    assert b"voxel grid" in L.gnb_last_error()
    assert L.gnb_plane_coords(None, 10, 0.1, 8, None, None, None) == -1


def test_binned_sampler_plan_is_host_only(library):
    """gnb_sample_binned_scratch_bytes decides on the host whether the brick-binned path applies (channels-last fp32
    volume, C % 4 == 0, z-rows contiguous) and how much scratch it needs; 0 sends the caller to gnb_sample_features."""
    import ctypes as C
    L = library.lib()

    def params(nx, ny, nz, ch, Q, channels_last=True, B=1):
        s = library.GnbSampleParams()
        s.batch, s.n_query, s.xyz = B, Q, 0x10000
        s.volume = 0x20000
        s.nx, s.ny, s.nz, s.C = nx, ny, nz, ch
        if channels_last:
            s.vol_stride_c, s.vol_stride_z, s.vol_stride_y, s.vol_stride_x = 1, ch, nz * ch, ny * nz * ch
        else:
            s.vol_stride_z, s.vol_stride_y, s.vol_stride_x, s.vol_stride_c = 1, nz, ny * nz, nx * ny * nz
        s.vol_stride_b = nx * ny * nz * ch
        s.voxel_size = 0.04
        return s

    Q = 1 << 20
    n = L.gnb_sample_binned_scratch_bytes(C.byref(params(96, 96, 48, 32, Q)))
    assert n >= Q * 20                                   # 16-byte sorted records + 4-byte bin ids per query, plus the tables
    assert n < Q * 20 + (64 << 20)
    assert L.gnb_sample_binned_scratch_bytes(C.byref(params(96, 96, 48, 32, Q, channels_last=False))) == 0
    assert L.gnb_sample_binned_scratch_bytes(C.byref(params(96, 96, 48, 30, Q))) == 0        # C % 4 != 0
    assert L.gnb_sample_binned_scratch_bytes(C.byref(params(96, 96, 48, 1, Q))) == 0
    more = L.gnb_sample_binned_scratch_bytes(C.byref(params(96, 96, 48, 32, 2 * Q)))
    assert more > n
    assert L.gnb_sample_binned_scratch_bytes(C.byref(params(256, 256, 96, 32, Q, B=2))) > 0
    s = params(96, 96, 48, 32, Q)
    assert L.gnb_sample_features_binned(C.byref(s), None, 0, None) != 0 and b"binned" in L.gnb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from gennerf_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.sample_features(torch.zeros(1, 4, 3), volume=torch.zeros(1, 2, 3, 3, 3))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.backproject_frames((4, 4, 4), 0.04, None, torch.zeros(1, 1, 3, 4), [torch.zeros(1, 2, 4, 4)])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gennerf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_extent_arithmetic_matches_torch():
    """The sampler computes fl32(n) * fl32(voxel_size); the reference computes
    torch.tensor([nx,ny,nz]) * voxel_size (utils.py:1019)."""
    import numpy as np
    for vs in (0.04, 0.02, 0.05, 0.16):
        for n in (48, 50, 60, 96, 128, 160, 180, 190, 256, 416):
            assert float((torch.tensor([n]) * vs)[0]) == float(np.float32(n) * np.float32(vs))
    for pad in (0.1, 0.0, 0.05):
        den = np.float32(1 + pad + 10e-6)
        x = torch.randn(1000)
        assert torch.equal(x / (1 + pad + 10e-6), torch.from_numpy(x.numpy() / den))


def test_training_backward_plan_and_argument_checks_are_host_only(library):
    """gnb_decode_train_bwd_workspace_bytes is a host-side plan (G, the stream gradients of every block, one layer's
    pre-activation / masked gradient / fp32 activation, the padded lin_z operands); the argument checks of the training entry
    points return GNB_E_INVALID before any CUDA call, and cuBLAS is not a link-time dependency of the library."""
    import ctypes as C
    L = library.lib()
    w = library.GnbDecoderWeights()
    w.d_feat, w.d_code, w.d_hidden, w.n_blocks, w.d_out, w.d_geo, w.use_code = 64, 15, 512, 5, 64, 32, 2
    n = 23200
    b = L.gnb_decode_train_bwd_workspace_bytes(C.byref(w), n)
    floats = n * 64 + n * 6 * 512 + 3 * n * 512 + 2 * n * 16 + 2 * 5 * 512 * 16
    assert floats * 4 <= b <= floats * 4 + 16 * 64 * 4                       # each of the 9 regions rounded up to 64 floats
    assert L.gnb_decode_train_bwd_workspace_bytes(C.byref(w), 2 * n) > b
    assert L.gnb_decode_train_bwd_workspace_bytes(None, n) == -1
    g = library.GnbDecoderGrads()
    assert L.gnb_decode_train_bwd(C.byref(w), None, None, None, None, None, None, None, 10, C.byref(g), None, 0, 1, None) == -1
    assert b"null pointer" in L.gnb_last_error()
    w.use_code = 1
    assert L.gnb_decode_train_bwd(C.byref(w), None, None, None, None, None, None, None, 10, C.byref(g), None, 0, 1, None) == -1
    assert b"use_code" in L.gnb_last_error()
    assert L.gnb_mlp_grad_link(None, 512, None, 512, 0, None, 0, None, 512, None, 0, None, 10, 510, None) == -1     # d % 4
    assert L.gnb_mlp_grad_link(None, 512, None, 512, 0, None, 0, None, 512, None, 0, None, 10, 512, None) == -1     # null pointers
    assert L.gnb_mlp_grad_head(None, None, None, None, None, 10, 64, 32, None, None, None, None, None) == -1
    out = subprocess.run(["ldd", library.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in out and "libcuda.so" not in out

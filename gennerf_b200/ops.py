"""Torch-facing operator layer over the C ABI (include/gennerf_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every computation below is a
hand-written sm_100a kernel reached through ctypes with raw device pointers and the
current CUDA stream.  CUDA tensors only -- no CPU path exists.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GNB_MAX_FRAMES, GnbDecoderWeights, GnbFusionParams, GnbLiftParams, GnbSampleParams, check, lib

PLANES = ("xz", "xy", "yz")


def _nvtx(fn):
    """NVTX range `gnb.<op>` around the op (Nsight timelines).  A push / pop pair costs ~1 us with no
    profiler attached."""
    import functools
    name = "gnb." + fn.__name__

    @functools.wraps(fn)
    def wrapped(*a, **k):
        torch.cuda.nvtx.range_push(name)
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.nvtx.range_pop()
    return wrapped


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gennerf_b200 ops take CUDA tensors only (there is no CPU fallback)")


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def _origin3(origin):
    if origin is None:
        return [0.0, 0.0, 0.0]
    if torch.is_tensor(origin):
        return [float(v) for v in origin.reshape(-1).tolist()]
    return [float(v) for v in origin]


# ------------------------------------------------------------------------------------------
# lift
# ------------------------------------------------------------------------------------------
@_nvtx
def backproject_frames(voxel_dim, voxel_size, origin, projections, features, *, mean=False,
                       volume_layout="channels_last", out=None, x_range=None, accumulate=None):
    """Fused back-projection of T frames (reference utils.py:948 + model.py:121-127,195-199).

    projections: (B,T,3,4) (CPU or CUDA; a CUDA tensor costs one small D2H copy);
    features: sequence of T CUDA tensors (B,C,H,W), NCHW-contiguous (reference layout,
    transposed on the fly) or channels_last (zero-copy).
    Returns volume (B,C,nx,ny,nz) [channels_last_3d strides unless volume_layout='reference'],
    count (B,nx,ny,nz) int32, valid (B,1,nx,ny,nz) bool.  volume == the reference's
    accumulated self.volume (a SUM, bit-exact); mean=True divides by count instead.
    `out=(volume, count, valid)` accumulates further frames into earlier results
    (accumulate=False: just write into those buffers).  `x_range=(x0, x1)` computes and writes only
    that slab of the grid (multi-GPU sharding); the rest of the buffers is left untouched.
    """
    nx, ny, nz = (int(d) for d in voxel_dim)
    feats = [_f32(f) for f in features]
    _need_cuda(*feats)
    T = len(feats)
    B, Cc, H, W = feats[0].shape
    P = torch.as_tensor(projections).detach().to("cpu", torch.float32).reshape(-1, T, 3, 4).contiguous()
    if P.shape[0] != B:
        raise RuntimeError(f"projections batch {P.shape[0]} != features batch {B}")
    dev = feats[0].device
    V = nx * ny * nz
    if out is None:
        if volume_layout == "channels_last":
            vol_store = torch.empty((B, nx, ny, nz, Cc), device=dev, dtype=torch.float32)
            volume = vol_store.permute(0, 4, 1, 2, 3)
        elif volume_layout == "reference":
            volume = torch.empty((B, Cc, nx, ny, nz), device=dev, dtype=torch.float32)
        else:
            raise ValueError(volume_layout)
        count = torch.empty((B, nx, ny, nz), device=dev, dtype=torch.int32)
        valid = torch.empty((B, 1, nx, ny, nz), device=dev, dtype=torch.bool)
        accumulate = False
    else:
        volume, count, valid = out
        accumulate = True if accumulate is None else bool(accumulate)
    sb, sc, sx, sy, sz = volume.stride()
    if not (sz * nz == sy and sy * ny == sx):
        raise RuntimeError("volume must be dense over (nx,ny,nz)")
    # channels_last (NHWC) feature maps are consumed in place; anything else is made NCHW-contiguous
    # (the reference's layout) and transposed by the library
    nhwc = all(f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in feats)
    if not nhwc:
        feats = [f.contiguous() for f in feats]
    with torch.cuda.device(dev):
        for t0 in range(0, T, _lib.GNB_MAX_FRAMES):
            chunk = feats[t0:t0 + _lib.GNB_MAX_FRAMES]
            n = len(chunk)
            p = GnbLiftParams()
            p.nx, p.ny, p.nz = nx, ny, nz
            p.voxel_size = float(voxel_size)
            p.origin[:] = _origin3(origin)
            p.batch, p.n_frames, p.C, p.H, p.W = B, n, Cc, H, W
            p.feat_layout = _lib.LAYOUT_NHWC if nhwc else _lib.LAYOUT_NCHW
            for i, f in enumerate(chunk):
                p.features[i] = f.data_ptr()
            Pc = P[:, t0:t0 + n].contiguous()
            p.h_projection = Pc.data_ptr()
            scratch = None
            if not nhwc:
                scratch = torch.empty(n * B * H * W * Cc, device=dev, dtype=torch.float32)
                p.scratch = scratch.data_ptr()
            p.volume = volume.data_ptr()
            p.vol_stride_b, p.vol_stride_v, p.vol_stride_c = sb, sz, sc
            p.count, p.valid = count.data_ptr(), valid.data_ptr()
            p.accumulate = int(accumulate or t0 > 0)
            p.mean = int(bool(mean) and t0 + n >= T)
            if x_range is not None:
                p.x_begin, p.x_end = int(x_range[0]), int(x_range[1])
                if p.x_end <= p.x_begin:
                    continue
            check(lib().gnb_backproject_frames(C.byref(p), _stream()), "gnb_backproject_frames")
    return volume, count, valid


def project_indices(voxel_dim, voxel_size, origin, projection, H, W, device="cuda"):
    """px, py (int32 (V,)), valid (bool (V,)) for ONE 3x4 projection: the integer part of the
    reference's backproject (utils.py:979-985), for bit-exact parity tests."""
    nx, ny, nz = (int(d) for d in voxel_dim)
    V = nx * ny * nz
    px = torch.empty(V, device=device, dtype=torch.int32)
    py = torch.empty(V, device=device, dtype=torch.int32)
    valid = torch.empty(V, device=device, dtype=torch.bool)
    o = (C.c_float * 3)(*_origin3(origin))
    P = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(projection).reshape(-1).tolist()])
    with torch.cuda.device(px.device):
        check(lib().gnb_project_indices(nx, ny, nz, float(voxel_size), o, P, int(H), int(W), px.data_ptr(),
                                        py.data_ptr(), valid.data_ptr(), _stream()), "gnb_project_indices")
    return px, py, valid


@_nvtx
def nchw_to_nhwc(frames, out=None):
    """T tensors (B,C,H,W) contiguous -> one (T,B,H,W,C) tensor (layout helper); `out` = a contiguous (T,B,H,W,C) buffer
    (e.g. a rank's slice of parallel.FrameBuffer.flat) to write into."""
    frames = [_f32(f).contiguous() for f in frames]
    _need_cuda(*frames)
    B, Cc, H, W = frames[0].shape
    if out is None:
        dst = torch.empty((len(frames), B, H, W, Cc), device=frames[0].device, dtype=torch.float32)
    else:
        dst = out
        if tuple(dst.shape) != (len(frames), B, H, W, Cc) or not dst.is_contiguous() or dst.dtype != torch.float32:
            raise RuntimeError("nchw_to_nhwc: out must be a contiguous fp32 (T,B,H,W,C) tensor")
    ptrs = (C.c_void_p * len(frames))(*[f.data_ptr() for f in frames])
    with torch.cuda.device(dst.device):
        check(lib().gnb_nchw_to_nhwc(ptrs, len(frames), dst.data_ptr(), B, Cc, H, W, _stream()), "gnb_nchw_to_nhwc")
    return dst


# ------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------
def _fill_sample_params(xyz, volume, planes, voxel_size, origin, padding):
    """volume: (B,C,nx,ny,nz) logical, any strides.  planes: dict name -> (B,C_p,R,R) logical,
    any (common) strides, or None.  Returns (params, keepalive, B, Q, C_p, C)."""
    _need_cuda(xyz, volume)
    xyz = _f32(xyz).contiguous()
    B, Q, _ = xyz.shape
    s = GnbSampleParams()
    keep = [xyz]
    s.batch, s.n_query, s.xyz = B, Q, xyz.data_ptr()
    Cv = Cp = 0
    if volume is not None:
        volume = _f32(volume)
        if volume.shape[0] != B:
            raise RuntimeError("volume batch != xyz batch")
        _, Cv, nx, ny, nz = volume.shape
        s.volume = volume.data_ptr()
        s.nx, s.ny, s.nz, s.C = nx, ny, nz, Cv
        s.vol_stride_b, s.vol_stride_c, s.vol_stride_x, s.vol_stride_y, s.vol_stride_z = volume.stride()
        s.voxel_size = float(voxel_size)
        s.origin[:] = _origin3(origin)
        keep.append(volume)
    if planes:
        ref = None
        for k, name in enumerate(PLANES):
            pl = planes.get(name)
            if pl is None:
                continue
            pl = _f32(pl)
            _need_cuda(pl)
            if ref is None:
                ref = pl
            elif pl.stride() != ref.stride() or pl.shape != ref.shape:
                pl = pl.contiguous(memory_format=torch.channels_last) if ref.is_contiguous(
                    memory_format=torch.channels_last) else pl.contiguous()
                if pl.stride() != ref.stride():
                    raise RuntimeError("the three planes must share shape and strides")
            s.plane[k] = pl.data_ptr()
            keep.append(pl)
        if ref is not None:
            if ref.shape[0] != B or ref.shape[2] != ref.shape[3]:
                raise RuntimeError("planes must be (B,C_p,R,R)")
            Cp, R = ref.shape[1], ref.shape[2]
            s.R, s.Cp = R, Cp
            s.pl_stride_b, s.pl_stride_c, s.pl_stride_h, s.pl_stride_w = ref.stride()
            s.padding = float(padding)
    return s, keep, B, Q, Cp, Cv


@_nvtx
def sample_features(xyz, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1, binned="auto"):
    """GenNerf.map_features (reference model.py:163-204): (B,Q,3) -> (B,Q,C_p + C), plane
    features first.  `volume` is the accumulated (already normalised == summed) volume.

    binned: "auto" sorts the queries into voxel bricks and gathers from shared-memory tiles
    (gnb_sample_features_binned, same bits) when there is at least one query per three voxels
    and the volume is channels-last; True forces it (raises when it does not apply); False never."""
    s, keep, B, Q, Cp, Cv = _fill_sample_params(xyz, volume, planes, voxel_size, origin, padding)
    out = torch.empty((B, Q, Cp + Cv), device=xyz.device, dtype=torch.float32)
    s.out, s.out_stride = out.data_ptr(), Cp + Cv
    with torch.cuda.device(out.device):
        use = False
        if binned and volume is not None and Q > 0:
            # (above 256 channels a brick's tile arrives as row copies instead of one tensor-map copy and the staged kernel, which
            #  already runs at the bandwidth of its random 8-line gathers, is faster: C = 512, 1 Mi queries 2.5 vs 3.5 ms)
            dense = binned is True or (B * Q >= (1 << 16) and 3 * Q >= volume.shape[2] * volume.shape[3] * volume.shape[4]
                                       and volume.shape[1] <= 256)
            nbytes = lib().gnb_sample_binned_scratch_bytes(C.byref(s)) if dense else 0
            if binned is True and nbytes == 0:
                raise RuntimeError("gennerf_b200: the binned sampler needs a channels-last fp32 volume with C % 4 == 0")
            use = nbytes > 0
        if use:
            scratch = torch.empty(nbytes, device=out.device, dtype=torch.uint8)
            check(lib().gnb_sample_features_binned(C.byref(s), scratch.data_ptr(), nbytes, _stream()), "gnb_sample_features_binned")
        else:
            check(lib().gnb_sample_features(C.byref(s), _stream()), "gnb_sample_features")
    return out


# ------------------------------------------------------------------------------------------
# triplane projection
# ------------------------------------------------------------------------------------------
@_nvtx
def plane_coords(p, padding, reso):
    """normalize_coordinate + coordinate2index for the three planes (reference utils.py:57-98).
    p (B,N,3) -> coord (3,B,N,2) fp32, index (3,B,N) int64; plane order xz, xy, yz."""
    _need_cuda(p)
    p = _f32(p).contiguous()
    B, N, _ = p.shape
    coord = torch.empty((3, B, N, 2), device=p.device, dtype=torch.float32)
    index = torch.empty((3, B, N), device=p.device, dtype=torch.int64)
    with torch.cuda.device(p.device):
        check(lib().gnb_plane_coords(p.data_ptr(), B * N, float(padding), int(reso), coord.data_ptr(),
                                     index.data_ptr(), _stream()), "gnb_plane_coords")
    return coord, index


@_nvtx
def scatter_mean_planes(p, c, reso, padding=0.1, mode="atomic"):
    """LocalPoolPointnet.generate_plane_features for xz, xy, yz at once (reference
    pointnet.py:72-89, without the U-Net).  p (B,N,3), c (B,N,C_p) ->
    planes (3,B,C_p,R,R) logical (channels-last storage), count (3,B,R,R) int32.
    mode 'atomic' (16-byte vector reductions; fast), 'auto' (= atomic), 'deterministic' (bit-identical to the CPU
    scatter order), 'sum' (atomic, sums left undivided for an all-reduce)."""
    _need_cuda(p, c)
    p, c = _f32(p).contiguous(), _f32(c).contiguous()
    B, N, _ = p.shape
    Cp = c.shape[2]
    R = int(reso)
    store = torch.empty((3, B, R, R, Cp), device=p.device, dtype=torch.float32)
    count = torch.empty((3, B, R, R), device=p.device, dtype=torch.int32)
    if mode == "auto":
        mode = "atomic"
    m = {"atomic": _lib.SCATTER_ATOMIC, "deterministic": _lib.SCATTER_DETERMINISTIC, "sum": _lib.SCATTER_ATOMIC_SUM}[mode]
    nbytes = lib().gnb_scatter_scratch_bytes(B, N, R, m)
    scratch = torch.empty(max(nbytes, 1), device=p.device, dtype=torch.uint8)
    with torch.cuda.device(p.device):
        check(lib().gnb_scatter_mean_planes(p.data_ptr(), c.data_ptr(), B, N, Cp, R, float(padding), m,
                                            store.data_ptr(), count.data_ptr(), scratch.data_ptr(), nbytes,
                                            _stream()), "gnb_scatter_mean_planes")
    return store.permute(0, 1, 4, 2, 3), count


@_nvtx
def scatter_finalize(planes, count):
    """sums (3,B,C_p,R,R) [channels-last storage] / max(count,1) in place (after an all-reduce)."""
    store = planes.permute(0, 1, 3, 4, 2)
    if not store.is_contiguous():
        raise RuntimeError("scatter_finalize expects the channels-last storage scatter_mean_planes returns")
    with torch.cuda.device(planes.device):
        check(lib().gnb_scatter_finalize(store.data_ptr(), count.data_ptr(), count.numel(), planes.shape[2], _stream()),
              "gnb_scatter_finalize")
    return planes


@_nvtx
def pool_local(p, c, reso, padding=0.1, scatter_type="max"):
    """LocalPoolPointnet.pool_local (reference pointnet.py:105-121): c (B,N,Hd) -> (B,N,Hd)."""
    _need_cuda(p, c)
    p, c = _f32(p).contiguous(), _f32(c).contiguous()
    B, N, _ = p.shape
    Hd = c.shape[2]
    t = {"max": _lib.POOL_MAX, "mean": _lib.POOL_MEAN}[scatter_type]
    out = torch.empty_like(c)
    nbytes = lib().gnb_pool_scratch_bytes(B, N, Hd, int(reso))
    scratch = torch.empty(nbytes, device=p.device, dtype=torch.uint8)
    with torch.cuda.device(p.device):
        check(lib().gnb_pool_local(p.data_ptr(), c.data_ptr(), B, N, Hd, int(reso), float(padding), t,
                                   out.data_ptr(), scratch.data_ptr(), nbytes, _stream()), "gnb_pool_local")
    return out


# ------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------
class DecoderWeights:
    """Device-side view of the reference's ResnetFC + TSDFHeadSimple parameters.

    Built from state_dicts with the reference's keys (lin_in.*, lin_z.{i}.*,
    blocks.{i}.fc_{0,1}.*, lin_out.*, alpha; fc.weight / fc.bias for the head).  The tensors
    stay owned by the nn.Modules (checkpoint compatibility); this object only holds pointers
    and, for the bf16 tcgen05 path, a packed copy refreshed by `pack()`.
    """

    def __init__(self, mlp_sd, head_w, head_b, *, n_blocks, d_geo, use_code=True, num_freqs=2,
                 freq_factor=0.5, include_input=True, d_code=None, device="cuda", alpha_on_device=False):
        """alpha_on_device: the kernels read ResnetFC.alpha from the parameter's device memory (GnbDecoderWeights.alpha_dev)
        instead of a host copy -- no `.item()`, i.e. no stream sync when the weights object is built inside a training step."""
        def dev(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        self.t = {k: dev(v) for k, v in mlp_sd.items()}
        self.head_w, self.head_b = dev(head_w).reshape(-1), dev(head_b).reshape(-1)
        w = GnbDecoderWeights()
        w.d_hidden, w.d_feat = self.t["lin_in.weight"].shape
        w.d_out = self.t["lin_out.weight"].shape[0]
        w.n_blocks, w.d_geo = int(n_blocks), int(d_geo)
        if alpha_on_device and "alpha" in self.t:
            w.alpha, w.alpha_dev = 1.0, self.t["alpha"].data_ptr()
        else:
            w.alpha = float(self.t["alpha"].item()) if "alpha" in self.t else 1.0
        w.use_code, w.num_freqs, w.freq_factor, w.include_input = int(use_code), int(num_freqs), float(freq_factor), int(include_input)
        if int(use_code) == 2:         # codes are given (stand-alone ResnetFC.forward)
            w.d_code = int(d_code)
        else:
            w.d_code = ((3 if include_input else 0) + 6 * int(num_freqs)) if use_code else 3
        w.lin_in_w, w.lin_in_b = self.t["lin_in.weight"].data_ptr(), self.t["lin_in.bias"].data_ptr()
        for i in range(w.n_blocks):
            w.lin_z_w[i], w.lin_z_b[i] = self.t[f"lin_z.{i}.weight"].data_ptr(), self.t[f"lin_z.{i}.bias"].data_ptr()
            w.fc0_w[i], w.fc0_b[i] = self.t[f"blocks.{i}.fc_0.weight"].data_ptr(), self.t[f"blocks.{i}.fc_0.bias"].data_ptr()
            w.fc1_w[i], w.fc1_b[i] = self.t[f"blocks.{i}.fc_1.weight"].data_ptr(), self.t[f"blocks.{i}.fc_1.bias"].data_ptr()
        w.lin_out_w, w.lin_out_b = self.t["lin_out.weight"].data_ptr(), self.t["lin_out.bias"].data_ptr()
        w.head_w, w.head_b = self.head_w.data_ptr(), self.head_b.data_ptr()
        # device status word: the tensor-core kernels set bit 0 when an fp16 operand saturated (see overflowed())
        self.status = torch.zeros(1, device=device, dtype=torch.int32)
        w.status = self.status.data_ptr()
        self.w = w
        self.device = torch.device(device)
        self.packed = None

    def pack(self, dtype="fp16"):
        """(Re)build the 16-bit tensor-core operand image of the weights ('fp16' or 'bf16')."""
        self.w.tc_dtype = {"fp16": _lib.TC_FP16, "bf16": _lib.TC_BF16}[dtype]
        n = lib().gnb_decoder_packed_bytes(C.byref(self.w))
        if n <= 0:
            raise RuntimeError("gennerf_b200: this decoder shape has no tcgen05 path: "
                               + lib().gnb_last_error().decode())
        if self.packed is None or self.packed.numel() != n:
            self.packed = torch.empty(n, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            check(lib().gnb_decoder_pack_tc(C.byref(self.w), self.packed.data_ptr(), _stream()), "gnb_decoder_pack_tc")
        self.packed_dtype = dtype
        return self.packed

    def overflowed(self, reset=True):
        """True when a tensor-core launch since the last reset saturated an fp16 operand at +-65504 (input feature,
        positional code or hidden activation): that result is outside the 1e-2 TSDF contract -- decode with
        precision='fp32'.  Costs one 4-byte device-to-host read (a stream sync)."""
        hit = bool(self.status.item())
        if hit and reset:
            self.status.zero_()
        return hit

    @property
    def wide(self):
        """More than 8 k-chunks of lin_in (d_feat > 512, e.g. the reference's default latent 512 + 32): the tcgen05 decoder
        takes such features through the operand image only (gnb_features_to_image / the sampler's image output)."""
        return self.w.d_feat > 512

    def tc_image(self, dtype):
        if self.packed is None or getattr(self, "packed_dtype", None) != dtype:
            self.pack(dtype)
        return self.packed


@_nvtx
def decode(weights, xyz, feat, precision="fp32"):
    """PositionalEncoding -> ResnetFC -> TSDFHeadSimple (reference model.py:226-246).
    precision: 'fp32' (CUDA cores, exact mode, 1e-5) | 'fp16' (tcgen05, fp32 accumulate, |dTSDF| <= 1e-2).
    ('bf16' selects the C ABI's experimental GNB_TC_BF16 operand type: ~3e-2 TSDF on this network, outside the contract.)
    xyz (..., 3), feat (..., d_feat) -> out (..., d_out), tsdf (..., 1)."""
    _need_cuda(xyz, feat)
    lead = xyz.shape[:-1]
    xyz2 = _f32(xyz).reshape(-1, 3).contiguous()
    feat2 = _f32(feat).reshape(-1, feat.shape[-1]).contiguous()
    n = xyz2.shape[0]
    out = torch.empty((n, weights.w.d_out), device=xyz.device, dtype=torch.float32)
    tsdf = torch.empty((n, 1), device=xyz.device, dtype=torch.float32)
    with torch.cuda.device(xyz.device):
        if precision == "fp32":
            check(lib().gnb_decode_fp32(C.byref(weights.w), xyz2.data_ptr(), feat2.data_ptr(), n, out.data_ptr(),
                                        tsdf.data_ptr(), _stream()), "gnb_decode_fp32")
        elif precision in ("fp16", "bf16") and weights.wide:
            # wide latent: fp32 rows -> 16-bit operand image (chunks of <= 512 MB) -> decoder streaming the image's k-chunks
            packed = weights.tc_image(precision)
            kf = lib().gnb_decoder_image_kchunks(C.byref(weights.w))
            if kf <= 0:
                raise RuntimeError("gennerf_b200: no tcgen05 path for these decoder dimensions / options: " + lib().gnb_last_error().decode())
            step = _image_rows(kf)
            image = torch.empty(((min(step, n) + 127) // 128) * kf * 16384, device=xyz.device, dtype=torch.uint8)
            dt = _lib.TC_BF16 if precision == "bf16" else _lib.TC_FP16
            for r0 in range(0, n, step):
                r1 = min(r0 + step, n)
                check(lib().gnb_features_to_image(feat2[r0:r1].data_ptr(), r1 - r0, feat2.shape[1], feat2.stride(0), kf, dt,
                                                  image.data_ptr(), weights.status.data_ptr(), _stream()), "gnb_features_to_image")
                check(lib().gnb_decode_image_tc(C.byref(weights.w), packed.data_ptr(), xyz2[r0:r1].data_ptr(), image.data_ptr(), r1 - r0,
                                                  out[r0:r1].data_ptr(), tsdf[r0:r1].data_ptr(), _stream()), "gnb_decode_image_tc")
        elif precision in ("fp16", "bf16"):
            packed = weights.tc_image(precision)
            check(lib().gnb_decode_tc(C.byref(weights.w), packed.data_ptr(), xyz2.data_ptr(), feat2.data_ptr(), n,
                                        out.data_ptr(), tsdf.data_ptr(), _stream()), "gnb_decode_tc")
        else:
            raise ValueError(precision)
    return out.reshape(*lead, weights.w.d_out), tsdf.reshape(*lead, 1)


@_nvtx
def decode_save(weights, xyz, feat, precision="fp16"):
    """ops.decode on the tcgen05 kernel that also returns the 16-bit activations every layer consumed -- the ReLU masks and
    left operands the backward pass of a training step needs (gnb_decode_tc_save).  Flat inputs: xyz (n,3) [or (n,d_code)
    codes when the weights were built with use_code=2], feat (n,d_feat).  Returns out (n,d_out), tsdf (n,1),
    activations (2*n_blocks+1, n, d_hidden) fp16 / bf16."""
    _need_cuda(xyz, feat)
    if weights.wide:
        raise RuntimeError("gennerf_b200: the activation-saving decoder (train_precision='fp16') takes latent codes up to 512 wide; "
                           "train this model with train_precision='fp32'")
    xyz2, feat2 = _f32(xyz).contiguous(), _f32(feat).contiguous()
    n = xyz2.shape[0]
    out = torch.empty((n, weights.w.d_out), device=xyz.device, dtype=torch.float32)
    tsdf = torch.empty((n, 1), device=xyz.device, dtype=torch.float32)
    acts = torch.empty((2 * weights.w.n_blocks + 1, n, weights.w.d_hidden), device=xyz.device,
                       dtype=torch.float16 if precision == "fp16" else torch.bfloat16)
    packed = weights.tc_image(precision)
    with torch.cuda.device(xyz.device):
        check(lib().gnb_decode_tc_save(C.byref(weights.w), packed.data_ptr(), xyz2.data_ptr(), feat2.data_ptr(), n,
                                       out.data_ptr(), tsdf.data_ptr(), acts.data_ptr(), _stream()), "gnb_decode_tc_save")
    return out, tsdf, acts


def _row_major(t, what):
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        raise ValueError(f"gennerf_b200: {what} must be a 2-D row-major tensor (or a column slice of one) with a row stride that is a "
                         "multiple of 4 elements and a 16-byte aligned start")
    return t


def mlp_grad_link(pre, act, res=None, *, out=None, colsum=None, want_act32=False):
    """One Linear + ReLU link of the ResNet-MLP's backward pass (gnb_mlp_grad_link; reference resnetfc.py:54-63 under autograd:
    threshold_backward, the skip connection's add and the bias gradient's column sum in ONE pass):
    out = where(act > 0, pre, 0) [+ res];  colsum (d,) += out.sum(0) (atomic order; must arrive zeroed).
    pre / res / out (n,d) fp32, act (n,d) fp16 / bf16 as saved by decode_save; all may be column slices of wider row-major
    tensors.  Returns out, or (out, act.float()) with want_act32 (the fp32 left operand of the weight-gradient GEMM)."""
    _need_cuda(pre, act, res, out, colsum)
    n, d = pre.shape
    if act.dtype not in (torch.float16, torch.bfloat16) or tuple(act.shape) != (n, d) or act.stride(1) != 1:
        raise ValueError("gennerf_b200: mlp_grad_link takes (n,d) fp16 / bf16 activations")
    if pre.dtype != torch.float32 or (res is not None and (res.dtype != torch.float32 or tuple(res.shape) != (n, d))):
        raise ValueError("gennerf_b200: mlp_grad_link takes fp32 gradients of one shape")
    _row_major(pre, "pre")
    if res is not None:
        _row_major(res, "res")
    if out is None:
        out = torch.empty((n, d), device=pre.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or tuple(out.shape) != (n, d):
        raise ValueError("gennerf_b200: mlp_grad_link: out must be (n,d) fp32")
    _row_major(out, "out")
    if colsum is not None and (colsum.dtype != torch.float32 or colsum.numel() != d or not colsum.is_contiguous()):
        raise ValueError("gennerf_b200: mlp_grad_link: colsum must be a contiguous fp32 vector of d elements")
    a32 = torch.empty((n, d), device=pre.device, dtype=torch.float32) if want_act32 else None
    with torch.cuda.device(pre.device):
        check(lib().gnb_mlp_grad_link(pre.data_ptr(), pre.stride(0), act.data_ptr(), act.stride(0),
                                      _lib.TC_FP16 if act.dtype == torch.float16 else _lib.TC_BF16,
                                      res.data_ptr() if res is not None else None, res.stride(0) if res is not None else 0,
                                      out.data_ptr(), out.stride(0), a32.data_ptr() if want_act32 else None, d,
                                      colsum.data_ptr() if colsum is not None else None, n, d, _stream()), "gnb_mlp_grad_link")
    return (out, a32) if want_act32 else out


def mlp_grad_head(g_out, g_tsdf, out, tsdf, head_w, d_geo):
    """Gradient entering the decoder from its outputs (gnb_mlp_grad_head; reference heads3d.py:36-50 + model.py:226-246 under
    autograd).  g_out (n,d_out) or None, g_tsdf (n,1) or None.  Returns G (n,d_out), d_head_w (d_geo,), d_head_b (1,),
    d_lin_out_b (d_out,)."""
    _need_cuda(g_out, g_tsdf, out, tsdf, head_w)
    n, d_out = out.shape
    dev = out.device
    g_out = None if g_out is None else _f32(g_out).contiguous()
    g_tsdf = None if g_tsdf is None else _f32(g_tsdf).contiguous()
    out, tsdf, hw = _f32(out).contiguous(), _f32(tsdf).contiguous(), _f32(head_w).reshape(-1).contiguous()
    G = torch.empty((n, d_out), device=dev, dtype=torch.float32)
    sums = torch.zeros(d_geo + 1 + d_out, device=dev, dtype=torch.float32)
    d_hw, d_hb, d_lb = sums[:d_geo], sums[d_geo:d_geo + 1], sums[d_geo + 1:]
    with torch.cuda.device(dev):
        check(lib().gnb_mlp_grad_head(g_out.data_ptr() if g_out is not None else None,
                                      g_tsdf.data_ptr() if g_tsdf is not None else None, out.data_ptr(), tsdf.data_ptr(),
                                      hw.data_ptr(), n, d_out, int(d_geo), G.data_ptr(), d_hw.data_ptr(), d_hb.data_ptr(),
                                      d_lb.data_ptr(), _stream()), "gnb_mlp_grad_head")
    return G, d_hw, d_hb, d_lb


@_nvtx
def decode_train_bwd(weights, code, feat, out, tsdf, acts, g_out, g_tsdf, *, need_code=True, need_feat=True, tf32=None):
    """The decoder's whole backward pass in ONE library call (gnb_decode_train_bwd; loss.backward() through ResnetFC +
    TSDFHeadSimple in the reference, resnetfc.py:134-189 + heads3d.py:36-50): own kernels for the output head, the ReLU masks /
    skip connections / bias sums of every link and the lin_z tail, cuBLAS SGEMMs (TF32 when torch's float32 matmul precision is
    not "highest", the reference's training setting) for the dgrad / wgrad products.  `weights`: the DecoderWeights of the
    forward (use_code=2); code (n,d_code), feat (n,d_feat), out / tsdf / acts as decode_save returned them; g_out (n,d_out) /
    g_tsdf (n,1) or None.  Returns (grads: dict keyed like the ResnetFC state_dict, d_head_w (d_geo,) | None, d_head_b (1,) |
    None, g_code | None, g_feat | None)."""
    _need_cuda(code, feat, out, tsdf, acts, g_out, g_tsdf)
    w = weights.w
    n = code.shape[0]
    H, nb, dc, df, dout, dgeo = w.d_hidden, w.n_blocks, w.d_code, w.d_feat, w.d_out, w.d_geo
    dev = code.device
    code, feat = _f32(code).contiguous(), _f32(feat).contiguous()
    out, tsdf = _f32(out).contiguous(), _f32(tsdf).contiguous()
    g_out = None if g_out is None else _f32(g_out).contiguous()
    g_tsdf = None if g_tsdf is None else _f32(g_tsdf).contiguous()
    if tuple(acts.shape) != (2 * nb + 1, n, H) or not acts.is_contiguous():
        raise ValueError("gennerf_b200: decode_train_bwd takes the activations decode_save returned")
    if tf32 is None:
        tf32 = torch.get_float32_matmul_precision() != "highest" or torch.backends.cuda.matmul.allow_tf32
    r4 = lambda v: (v + 3) & ~3                                       # noqa: E731  (16-byte aligned slices)
    # accumulated (zeroed) gradients in one buffer, overwritten ones in another
    zsizes = [("lin_in.bias", H)] + [(f"blocks.{i}.fc_0.bias", H) for i in range(nb)] + [(f"blocks.{i}.fc_1.bias", H) for i in range(nb)] \
        + [("lin_out.bias", dout), ("head_w", max(dgeo, 1)), ("head_b", 1), ("alpha", 1)]
    esizes = [("lin_in.weight", H * df), ("lin_out.weight", dout * H)] + [(f"lin_z.{i}.weight", H * dc) for i in range(nb)] \
        + [(f"lin_z.{i}.bias", H) for i in range(nb)] + [(f"blocks.{i}.fc_0.weight", H * H) for i in range(nb)] \
        + [(f"blocks.{i}.fc_1.weight", H * H) for i in range(nb)]
    zbuf = torch.zeros(sum(r4(s) for _, s in zsizes), device=dev, dtype=torch.float32)
    ebuf = torch.empty(sum(r4(s) for _, s in esizes), device=dev, dtype=torch.float32)
    t = {}
    for buf, sizes in ((zbuf, zsizes), (ebuf, esizes)):
        o = 0
        for k, s in sizes:
            t[k] = buf[o:o + s]
            o += r4(s)
    g = _lib.GnbDecoderGrads()
    ptr = lambda k: t[k].data_ptr()                                   # noqa: E731
    g.lin_in_w, g.lin_in_b, g.lin_out_w, g.lin_out_b = ptr("lin_in.weight"), ptr("lin_in.bias"), ptr("lin_out.weight"), ptr("lin_out.bias")
    for i in range(nb):
        g.lin_z_w[i], g.lin_z_b[i] = ptr(f"lin_z.{i}.weight"), ptr(f"lin_z.{i}.bias")
        g.fc0_w[i], g.fc0_b[i] = ptr(f"blocks.{i}.fc_0.weight"), ptr(f"blocks.{i}.fc_0.bias")
        g.fc1_w[i], g.fc1_b[i] = ptr(f"blocks.{i}.fc_1.weight"), ptr(f"blocks.{i}.fc_1.bias")
    g.head_w, g.head_b, g.alpha = ptr("head_w"), ptr("head_b"), ptr("alpha")
    g_code = torch.empty((n, dc), device=dev, dtype=torch.float32) if need_code else None
    g_feat = torch.empty((n, df), device=dev, dtype=torch.float32) if need_feat else None
    g.g_code = g_code.data_ptr() if need_code else None
    g.g_feat = g_feat.data_ptr() if need_feat else None
    nbytes = lib().gnb_decode_train_bwd_workspace_bytes(C.byref(w), n)
    ws = torch.empty(max(nbytes, 256), device=dev, dtype=torch.uint8)
    saved_dtype = w.tc_dtype
    w.tc_dtype = _lib.TC_FP16 if acts.dtype == torch.float16 else _lib.TC_BF16
    try:
        with torch.cuda.device(dev):
            check(lib().gnb_decode_train_bwd(C.byref(w), code.data_ptr(), feat.data_ptr(), out.data_ptr(), tsdf.data_ptr(), acts.data_ptr(),
                                             g_out.data_ptr() if g_out is not None else None,
                                             g_tsdf.data_ptr() if g_tsdf is not None else None, n, C.byref(g), ws.data_ptr(), nbytes,
                                             int(bool(tf32)), _stream()), "gnb_decode_train_bwd")
    finally:
        w.tc_dtype = saved_dtype
    grads = {"lin_in.weight": t["lin_in.weight"].view(H, df), "lin_in.bias": t["lin_in.bias"],
             "lin_out.weight": t["lin_out.weight"].view(dout, H), "lin_out.bias": t["lin_out.bias"], "alpha": t["alpha"]}
    for i in range(nb):
        grads[f"lin_z.{i}.weight"], grads[f"lin_z.{i}.bias"] = t[f"lin_z.{i}.weight"].view(H, dc), t[f"lin_z.{i}.bias"]
        grads[f"blocks.{i}.fc_0.weight"], grads[f"blocks.{i}.fc_0.bias"] = t[f"blocks.{i}.fc_0.weight"].view(H, H), t[f"blocks.{i}.fc_0.bias"]
        grads[f"blocks.{i}.fc_1.weight"], grads[f"blocks.{i}.fc_1.bias"] = t[f"blocks.{i}.fc_1.weight"].view(H, H), t[f"blocks.{i}.fc_1.bias"]
    has_t = g_tsdf is not None
    return grads, (t["head_w"][:dgeo] if has_t else None), (t["head_b"] if has_t else None), g_code, g_feat


IMAGE_CHUNK = 1 << 22          # queries per sampler + decoder launch pair of query_image (512 MB of operand image per 64 features)


def _image_rows(kf):
    """Rows per operand-image chunk: IMAGE_CHUNK (a 512 MB image) for up to 64 features; for wider latents as many rows as a
    2 GB image holds -- the brick-binned sampler stages a tile per brick, so the more queries of a brick are in one chunk the
    fewer bytes it moves per query, and fewer launches either way."""
    kf = max(1, int(kf))
    if kf <= 1:
        return IMAGE_CHUNK
    return max(128, min(IMAGE_CHUNK, ((2 << 30) // (kf * 128 * 128)) * 128))


@_nvtx
def query_image(weights, xyz, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1,
                want_feat=False, precision="fp16", chunk=None, out=None, tsdf=None, want_out=True):
    """GenNerf.forward as TWO kernels per chunk of queries: the sampler (brick-binned where the queries are dense) writes the
    features straight as the decoder's 16-bit lin_in operand image, and the tcgen05 decoder brings each tile's image into
    shared memory with one bulk copy.  Same bits as query_fused (same sampling arithmetic, same rounding to 16 bits); faster
    for many queries, because the decoder no longer gathers and converts features between its tiles.
    Returns out (B,Q,d_out), tsdf (B,Q,1), feat (B,Q,C_lat) or None."""
    _need_cuda(xyz)
    xyz = _f32(xyz).contiguous()
    B, Q, _ = xyz.shape
    dev = xyz.device
    d_feat = weights.w.d_feat
    packed = weights.tc_image(precision)
    kf = lib().gnb_decoder_image_kchunks(C.byref(weights.w))
    if kf <= 0 and B * Q > 0:          # no early-staging variant for these dimensions / options: the single fused kernel
        o, t, f = query_fused(weights, xyz, volume, planes, voxel_size=voxel_size, origin=origin, padding=padding,
                              want_feat=want_feat, precision=precision, mode="fused")
        if out is not None:
            out.copy_(o)
        if tsdf is not None:
            tsdf.copy_(t)
        return (o if out is None else out), (t if tsdf is None else tsdf), f
    # (out / tsdf: optional contiguous fp32 destinations of shape (B,Q,d_out) / (B,Q,1), e.g. slices of a larger result)
    if out is None and want_out:
        out = torch.empty((B, Q, weights.w.d_out), device=dev, dtype=torch.float32)
    tsdf = torch.empty((B, Q, 1), device=dev, dtype=torch.float32) if tsdf is None else tsdf
    feat = torch.empty((B, Q, d_feat), device=dev, dtype=torch.float32) if want_feat else None
    if B * Q == 0:
        return out, tsdf, feat
    step = int(chunk or _image_rows(kf))
    rows = min(step, Q)
    image = torch.zeros(((rows + 127) // 128) * kf * 16384, device=dev, dtype=torch.uint8)   # zeros: operand columns past d_feat
    with torch.cuda.device(dev):
        for b in range(B):
            vol_b = volume[b:b + 1] if volume is not None else None
            pl_b = {k: (v[b:b + 1] if v is not None else None) for k, v in planes.items()} if planes else None
            for q0 in range(0, Q, step):
                q1 = min(q0 + step, Q)
                xq = xyz[b:b + 1, q0:q1]
                s, keep, _, n, Cp, Cv = _fill_sample_params(xq, vol_b, pl_b, voxel_size, origin, padding)
                if Cp + Cv != d_feat:
                    raise RuntimeError(f"query_image: d_feat {d_feat} != C_p + C = {Cp + Cv}")
                if feat is not None:
                    s.out, s.out_stride = feat[b, q0:q1].data_ptr(), d_feat
                s.image, s.image_kchunks = image.data_ptr(), kf
                s.image_dtype = _lib.TC_BF16 if precision == "bf16" else _lib.TC_FP16
                s.image_status = weights.status.data_ptr()
                nbytes = 0
                if vol_b is not None and n >= (1 << 16) and 3 * n >= vol_b.shape[2] * vol_b.shape[3] * vol_b.shape[4] and Cv <= 256:
                    nbytes = lib().gnb_sample_binned_scratch_bytes(C.byref(s))
                if nbytes > 0:
                    scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                    check(lib().gnb_sample_features_binned(C.byref(s), scratch.data_ptr(), nbytes, _stream()), "gnb_sample_features_binned")
                else:
                    check(lib().gnb_sample_features(C.byref(s), _stream()), "gnb_sample_features")
                check(lib().gnb_decode_image_tc(C.byref(weights.w), packed.data_ptr(), xq.data_ptr(), image.data_ptr(), n,
                                                  out[b, q0:q1].data_ptr() if out is not None else None,
                                                  tsdf[b, q0:q1].data_ptr(), _stream()), "gnb_decode_image_tc")
    return out, tsdf, feat


@_nvtx
def query_fused(weights, xyz, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1,
                want_feat=True, precision="fp16", presort="auto", mode="auto"):
    """GenNerf.forward in one kernel (sampler fused into the tcgen05 decoder).
    Returns out (B,Q,d_out), tsdf (B,Q,1), feat (B,Q,C_lat) or None.

    mode: "auto" hands 65 536 queries or more to query_image (sampler kernel writing the decoder's operand image, then the
    decoder: same bits, less time per query); "fused" always runs the single fused kernel; "image" always query_image.
    presort (fused kernel): "auto" first counting-sorts the queries by voxel brick (same bits, outputs in the caller's order)
    when there is at least one query per three voxels -- the sampling prologue then reads the volume brick by brick; True
    forces it (raises when there is no channels-last volume to sort by), False never."""
    if mode == "image" or (weights.wide and mode != "fused") or (mode == "auto" and presort == "auto" and xyz.shape[0] * xyz.shape[1] >= (1 << 16)
                                           and not _lib.get_option("GNB_QUERY_FUSED")):
        return query_image(weights, xyz, volume, planes, voxel_size=voxel_size, origin=origin, padding=padding,
                           want_feat=want_feat, precision=precision)
    s, keep, B, Q, Cp, Cv = _fill_sample_params(xyz, volume, planes, voxel_size, origin, padding)
    feat = None
    if want_feat:
        feat = torch.empty((B, Q, Cp + Cv), device=xyz.device, dtype=torch.float32)
        s.out, s.out_stride = feat.data_ptr(), Cp + Cv
    out = torch.empty((B, Q, weights.w.d_out), device=xyz.device, dtype=torch.float32)
    tsdf = torch.empty((B, Q, 1), device=xyz.device, dtype=torch.float32)
    packed = weights.tc_image(precision)
    with torch.cuda.device(xyz.device):
        nbytes = 0
        if presort and volume is not None and Q > 0:
            dense = presort is True or (B * Q >= (1 << 16) and 3 * Q >= volume.shape[2] * volume.shape[3] * volume.shape[4])
            nbytes = lib().gnb_query_fused_sorted_scratch_bytes(C.byref(s)) if dense else 0
            if presort is True and nbytes == 0:
                raise RuntimeError("gennerf_b200: presort needs a channels-last fp32 volume with C % 4 == 0")
        if nbytes > 0:
            scratch = torch.empty(nbytes, device=xyz.device, dtype=torch.uint8)
            check(lib().gnb_query_fused_sorted_tc(C.byref(s), C.byref(weights.w), packed.data_ptr(), out.data_ptr(),
                                                    tsdf.data_ptr(), scratch.data_ptr(), nbytes, _stream()), "gnb_query_fused_sorted_tc")
        else:
            check(lib().gnb_query_fused_tc(C.byref(s), C.byref(weights.w), packed.data_ptr(), out.data_ptr(),
                                             tsdf.data_ptr(), _stream()), "gnb_query_fused_tc")
    return out, tsdf, feat


@_nvtx
def query_grid_fused(weights, grid_dim, axes, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1,
                     want_out=False, precision="fp16", mode="auto"):
    """GenNerf.predict_tsdf's query (reference model.py:752-790) in one kernel without a materialised query grid:
    `axes` = (ax (nx,), ay (ny,), az (nz,)) CUDA fp32 coordinate axes (torch.linspace of get_grid_coordinates,
    utils.py:926-935).  Returns tsdf (B,nx,ny,nz) and, with want_out, out (B,nx*ny*nz,d_out)."""
    nx, ny, nz = (int(d) for d in grid_dim)
    ref = volume if volume is not None else next(p for p in planes.values() if p is not None)
    B, dev = ref.shape[0], ref.device
    ax = torch.cat([_f32(a).reshape(-1) for a in axes]).to(dev).contiguous()
    if ax.numel() != nx + ny + nz:
        raise RuntimeError("query_grid_fused: axes must hold nx, ny and nz coordinates")
    dummy = torch.empty((B, 0, 3), device=dev, dtype=torch.float32)
    s, keep, _, _, Cp, Cv = _fill_sample_params(dummy, volume, planes, voxel_size, origin, padding)
    Q = nx * ny * nz
    out = torch.empty((B, Q, weights.w.d_out), device=dev, dtype=torch.float32) if want_out else None
    tsdf = torch.empty((B, nx, ny, nz), device=dev, dtype=torch.float32)
    packed = weights.tc_image(precision)
    if (mode == "image" or weights.wide or (mode == "auto" and B * Q >= (1 << 16) and not _lib.get_option("GNB_QUERY_FUSED"))) and \
            lib().gnb_decoder_image_kchunks(C.byref(weights.w)) > 0:
        # many grid points: the two-kernel query (query_image) is faster than the fused kernel's in-kernel sampling.  The
        # points of a chunk are generated on the device from the axes (same values, so the same bits) and never exist
        # for more than one chunk (48 MB per 4 Mi points).
        axd = (ax[:nx], ax[nx:nx + ny], ax[nx + ny:])
        tflat = tsdf.view(B, Q, 1)
        for q0 in range(0, Q, IMAGE_CHUNK):
            q1 = min(q0 + IMAGE_CHUNK, Q)
            idx = torch.arange(q0, q1, device=dev)
            xyz = torch.stack((axd[0][idx // (ny * nz)], axd[1][(idx // nz) % ny], axd[2][idx % nz]), dim=-1).unsqueeze(0)
            for b in range(B):
                scratch_out = out[b:b + 1, q0:q1] if out is not None else None
                query_image(weights, xyz, volume[b:b + 1] if volume is not None else None,
                            {k: (v[b:b + 1] if v is not None else None) for k, v in planes.items()} if planes else None,
                            voxel_size=voxel_size, origin=origin, padding=padding, precision=precision,
                            out=scratch_out, tsdf=tflat[b:b + 1, q0:q1], want_out=out is not None)
        return tsdf, out
    g3 = (C.c_int32 * 3)(nx, ny, nz)
    with torch.cuda.device(dev):
        check(lib().gnb_query_grid_fused_tc(C.byref(s), g3, ax.data_ptr(), C.byref(weights.w), packed.data_ptr(),
                                              out.data_ptr() if out is not None else None, tsdf.data_ptr(), _stream()),
              "gnb_query_grid_fused_tc")
    return tsdf, out


def fused_query_applies(volume=None, planes=None):
    """True when gnb_query_fused_tc can read these tensors (16-byte aligned channels-last fp32 storage with channel
    counts % 4 == 0); otherwise callers use sample_features + decode."""
    ts = ([volume] if volume is not None else []) + [p for p in (planes or {}).values() if p is not None]
    for t in ts:
        if t.dtype != torch.float32 or t.shape[1] % 4 or t.stride(1) != 1 or t.data_ptr() % 16:
            return False
        if any(st % 4 for i, st in enumerate(t.stride()) if i != 1):
            return False
    if planes:
        ps = [p for p in planes.values() if p is not None]
        if any(p.stride() != ps[0].stride() or p.shape != ps[0].shape for p in ps):
            return False
    return True


def positional_encoding(x, num_freqs, freq_factor, include_input=True):
    """PositionalEncoding.forward (reference positional_encoding.py:28-40): (n,3) -> (n,d_code)."""
    _need_cuda(x)
    x = _f32(x).contiguous()
    n = x.shape[0]
    d = (3 if include_input else 0) + 6 * int(num_freqs)
    out = torch.empty((n, d), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().gnb_positional_encoding(x.data_ptr(), n, int(num_freqs), float(freq_factor), int(include_input),
                                            out.data_ptr(), _stream()), "gnb_positional_encoding")
    return out


def tsdf_head(feat_geo, weight, bias):
    """TSDFHeadSimple.forward (reference heads3d.py:36-50): (...,d_geo) -> (...,1) = tanh(fc(x))."""
    _need_cuda(feat_geo, weight, bias)
    lead = feat_geo.shape[:-1]
    d_geo = feat_geo.shape[-1]
    g = _f32(feat_geo)
    if g.stride(-1) != 1 or not g.reshape(-1, d_geo).is_contiguous():
        # a column slice of a row-major (n, d_out) tensor is fine: pass its row stride
        if g.dim() >= 2 and g.stride(-1) == 1 and all(g.stride(i) == g.stride(i + 1) * g.shape[i + 1] for i in range(g.dim() - 2)):
            pass
        else:
            g = g.contiguous()
    n = 1
    for s in lead:
        n *= s
    stride = g.stride(-2) if g.dim() >= 2 else d_geo
    w = _f32(weight).reshape(-1).contiguous()
    b = _f32(bias).reshape(-1).contiguous()
    out = torch.empty((n,), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        check(lib().gnb_tsdf_head(g.data_ptr(), n, d_geo, stride, w.data_ptr(), b.data_ptr(), out.data_ptr(), _stream()),
              "gnb_tsdf_head")
    return out.reshape(*lead, 1)


# ------------------------------------------------------------------------------------------
# backward passes -- raw ops; gennerf_b200.autograd wires them into torch.autograd
# ------------------------------------------------------------------------------------------
@_nvtx
def backproject_frames_bwd(voxel_dim, voxel_size, origin, projections, grad_volume, feat_shape, n_frames, *,
                           nhwc=False, mean=False, count=None, x_range=None):
    """grad_volume (B,C,nx,ny,nz) logical (any dense strides) -> list of T gradient maps (B,C,H,W) logical, in
    channels_last memory when `nhwc` else NCHW-contiguous."""
    nx, ny, nz = (int(d) for d in voxel_dim)
    B, Cc, H, W = feat_shape
    _need_cuda(grad_volume)
    gv = _f32(grad_volume)
    sb, sc, sx, sy, sz = gv.stride()
    if not (sz * nz == sy and sy * ny == sx):
        gv = gv.contiguous()
        sb, sc, sx, sy, sz = gv.stride()
    P = torch.as_tensor(projections).detach().to("cpu", torch.float32).reshape(-1, n_frames, 3, 4).contiguous()
    dev = gv.device
    grads = []
    with torch.cuda.device(dev):
        for t0 in range(0, n_frames, _lib.GNB_MAX_FRAMES):
            n = min(_lib.GNB_MAX_FRAMES, n_frames - t0)
            if nhwc:
                chunk = [torch.empty((B, H, W, Cc), device=dev, dtype=torch.float32).permute(0, 3, 1, 2) for _ in range(n)]
            else:
                chunk = [torch.empty((B, Cc, H, W), device=dev, dtype=torch.float32) for _ in range(n)]
            p = GnbLiftParams()
            p.nx, p.ny, p.nz = nx, ny, nz
            p.voxel_size = float(voxel_size)
            p.origin[:] = _origin3(origin)
            p.batch, p.n_frames, p.C, p.H, p.W = B, n, Cc, H, W
            p.feat_layout = _lib.LAYOUT_NHWC if nhwc else _lib.LAYOUT_NCHW
            Pc = P[:, t0:t0 + n].contiguous()
            p.h_projection = Pc.data_ptr()
            scratch = None
            if not nhwc:
                scratch = torch.empty(n * B * H * W * Cc, device=dev, dtype=torch.float32)
                p.scratch = scratch.data_ptr()
            p.vol_stride_b, p.vol_stride_v, p.vol_stride_c = sb, sz, sc
            p.mean = int(bool(mean))
            if mean:
                p.count = count.data_ptr()
            if x_range is not None:
                p.x_begin, p.x_end = int(x_range[0]), int(x_range[1])
            ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in chunk])
            check(lib().gnb_backproject_frames_bwd(C.byref(p), gv.data_ptr(), ptrs, _stream()), "gnb_backproject_frames_bwd")
            grads.extend(chunk)
    return grads


@_nvtx
def sample_features_bwd(grad_out, xyz, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1,
                        need_volume=True, need_planes=True, need_xyz=True):
    """Backward of sample_features: returns (grad_xyz | None, grad_volume | None, {plane: grad} | None) with the
    strides of the forward inputs."""
    s, keep, B, Q, Cp, Cv = _fill_sample_params(xyz, volume, planes, voxel_size, origin, padding)
    go = _f32(grad_out).contiguous()
    gvol = torch.zeros_like(volume, memory_format=torch.preserve_format) if (volume is not None and need_volume) else None
    if gvol is not None and gvol.stride() != volume.stride():
        raise RuntimeError("gradient volume must share the forward strides")
    gpl, ptrs = None, None
    if planes and need_planes:
        gpl, ptrs = {}, (C.c_void_p * 3)()
        fwd = {t.data_ptr(): t for t in keep}                     # the tensors the forward pointers refer to
        for k, name in enumerate(PLANES):
            if planes.get(name) is not None:
                g = torch.zeros_like(fwd[s.plane[k]], memory_format=torch.preserve_format)
                gpl[name] = g
                ptrs[k] = g.data_ptr()
    gxyz = torch.empty((B, Q, 3), device=go.device, dtype=torch.float32) if need_xyz else None
    with torch.cuda.device(go.device):
        check(lib().gnb_sample_features_bwd(C.byref(s), go.data_ptr(), go.shape[-1], gvol.data_ptr() if gvol is not None else None,
                                            ptrs, gxyz.data_ptr() if gxyz is not None else None, _stream()),
              "gnb_sample_features_bwd")
    return gxyz, gvol, gpl


@_nvtx
def sample_features_bwd2(grad_out, gg_xyz, xyz, volume=None, planes=None, *, voxel_size=0.04, origin=None, padding=0.1,
                         need_grad_out=True, need_volume=True, need_planes=True, need_xyz=True):
    """Double backward of sample_features (gnb_sample_features_bwd2): the gradients of <gg_xyz, grad_xyz> where grad_xyz
    is sample_features_bwd's coordinate gradient (eikonal / gradient losses, reference utils.py:636-649).
    Returns (g_grad_out | None, g_volume | None, {plane: g} | None, g_xyz | None)."""
    s, keep, B, Q, Cp, Cv = _fill_sample_params(xyz, volume, planes, voxel_size, origin, padding)
    go, gg = _f32(grad_out).contiguous(), _f32(gg_xyz).contiguous()
    _need_cuda(go, gg)
    ggo = torch.empty((B, Q, Cp + Cv), device=go.device, dtype=torch.float32) if need_grad_out else None
    gvol = torch.zeros_like(volume, memory_format=torch.preserve_format) if (volume is not None and need_volume) else None
    if gvol is not None and gvol.stride() != volume.stride():
        raise RuntimeError("gradient volume must share the forward strides")
    gpl, ptrs = None, None
    if planes and need_planes:
        gpl, ptrs = {}, (C.c_void_p * 3)()
        fwd = {t.data_ptr(): t for t in keep}
        for k, name in enumerate(PLANES):
            if planes.get(name) is not None:
                g = torch.zeros_like(fwd[s.plane[k]], memory_format=torch.preserve_format)
                gpl[name] = g
                ptrs[k] = g.data_ptr()
    gx = torch.empty((B, Q, 3), device=go.device, dtype=torch.float32) if need_xyz else None
    with torch.cuda.device(go.device):
        check(lib().gnb_sample_features_bwd2(C.byref(s), go.data_ptr(), go.shape[-1], gg.data_ptr(),
                                             ggo.data_ptr() if ggo is not None else None, Cp + Cv,
                                             gvol.data_ptr() if gvol is not None else None, ptrs,
                                             gx.data_ptr() if gx is not None else None, _stream()), "gnb_sample_features_bwd2")
    return ggo, gvol, gpl, gx


@_nvtx
def scatter_mean_planes_bwd(p, grad_planes, count, padding=0.1):
    """grad_planes (3,B,C_p,R,R) logical -> grad_c (B,N,C_p)."""
    _need_cuda(p, grad_planes, count)
    p = _f32(p).contiguous()
    B, N, _ = p.shape
    store = _f32(grad_planes).permute(0, 1, 3, 4, 2).contiguous()                 # (3,B,R,R,C_p)
    Cp, R = store.shape[-1], store.shape[2]
    gc = torch.empty((B, N, Cp), device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        check(lib().gnb_scatter_mean_planes_bwd(p.data_ptr(), store.data_ptr(), count.data_ptr(), B, N, Cp, R, float(padding),
                                                gc.data_ptr(), _stream()), "gnb_scatter_mean_planes_bwd")
    return gc


@_nvtx
def pool_local_fwd_keep(p, c, reso, padding=0.1, scatter_type="max"):
    """pool_local that also returns the scratch buffer the backward needs."""
    _need_cuda(p, c)
    p, c = _f32(p).contiguous(), _f32(c).contiguous()
    B, N, _ = p.shape
    Hd = c.shape[2]
    t = {"max": _lib.POOL_MAX, "mean": _lib.POOL_MEAN}[scatter_type]
    out = torch.empty_like(c)
    nbytes = lib().gnb_pool_scratch_bytes(B, N, Hd, int(reso))
    scratch = torch.empty(nbytes, device=p.device, dtype=torch.uint8)
    with torch.cuda.device(p.device):
        check(lib().gnb_pool_local(p.data_ptr(), c.data_ptr(), B, N, Hd, int(reso), float(padding), t,
                                   out.data_ptr(), scratch.data_ptr(), nbytes, _stream()), "gnb_pool_local")
    return out, scratch


@_nvtx
def pool_local_bwd(p, c, grad_out, fwd_scratch, reso, padding=0.1, scatter_type="max"):
    p, c, go = _f32(p).contiguous(), _f32(c).contiguous(), _f32(grad_out).contiguous()
    B, N, _ = p.shape
    Hd = c.shape[2]
    t = {"max": _lib.POOL_MAX, "mean": _lib.POOL_MEAN}[scatter_type]
    gc = torch.empty_like(c)
    nbytes = lib().gnb_pool_bwd_scratch_bytes(B, N, Hd, int(reso))
    scratch = torch.empty(nbytes, device=p.device, dtype=torch.uint8)
    with torch.cuda.device(p.device):
        check(lib().gnb_pool_local_bwd(p.data_ptr(), c.data_ptr(), go.data_ptr(), B, N, Hd, int(reso), float(padding), t,
                                       fwd_scratch.data_ptr(), gc.data_ptr(), scratch.data_ptr(), nbytes, _stream()),
              "gnb_pool_local_bwd")
    return gc


# ------------------------------------------------------------------------------------------
# front end of the triplane branch
# ------------------------------------------------------------------------------------------
@_nvtx
def get_3d_points(depth_map, projection):
    """get_3d_points (reference utils.py:120-175): depth (B,H,W), projection (B,3,4) -> (B,H,W,3)."""
    _need_cuda(depth_map)
    d = _f32(depth_map).contiguous()
    B, H, W = d.shape
    P = torch.as_tensor(projection).detach().to("cpu", torch.float32).reshape(B, 3, 4).contiguous()
    out = torch.empty((B, H, W, 3), device=d.device, dtype=torch.float32)
    with torch.cuda.device(d.device):
        check(lib().gnb_get_3d_points(d.data_ptr(), P.data_ptr(), B, H, W, out.data_ptr(), _stream()), "gnb_get_3d_points")
    return out


@_nvtx
def farthest_point_sample(xyz, npoint, start=None):
    """farthest_point_sample (reference utils.py:178-202): xyz (B,N,3) -> sampled (B,npoint,3), indices (B,npoint).
    `start` (B,) int64 is the first index; the reference draws it with torch.randint(0, N, (B,))."""
    _need_cuda(xyz)
    x = _f32(xyz).contiguous()
    B, N, _ = x.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long, device=x.device)
    start = start.to(device=x.device, dtype=torch.long).contiguous()
    idx = torch.empty((B, npoint), device=x.device, dtype=torch.long)
    out = torch.empty((B, npoint, 3), device=x.device, dtype=torch.float32)
    scratch = torch.empty((B, N), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().gnb_farthest_point_sample(x.data_ptr(), B, N, int(npoint), start.data_ptr(), scratch.data_ptr(),
                                              idx.data_ptr(), out.data_ptr(), _stream()), "gnb_farthest_point_sample")
    return out, idx


# ------------------------------------------------------------------------------------------
# TSDF fusion
# ------------------------------------------------------------------------------------------
@_nvtx
def tsdf_fusion_integrate(voxel_dim, voxel_size, origin, trunc_margin, projections, depths, tsdf_vol, weight_vol,
                          colors=None, color_vol=None, labels=None, label_vol=None, depth_culling=True):
    """TSDFFusion.integrate (reference src/data/tsdf.py:369-418) for T frames in one launch per 64 frames.

    projections (T,3,4); depths (T,H,W) CUDA fp32; colors (T,3,H,W) fp32 / labels (T,H,W) int32 optional.
    tsdf_vol, weight_vol (V) fp32, color_vol (3,V) fp32, label_vol (V) int32 are updated IN PLACE, frames in
    order: bit-identical to T sequential integrate() calls of the reference."""
    nx, ny, nz = (int(d) for d in voxel_dim)
    _need_cuda(depths, tsdf_vol, weight_vol, colors, color_vol, labels, label_vol)
    V = nx * ny * nz
    d = _f32(depths).contiguous()
    T, H, W = d.shape
    P = torch.as_tensor(projections).detach().to("cpu", torch.float32).reshape(T, 3, 4).contiguous()
    for name, t, shape, dt in (("tsdf_vol", tsdf_vol, (V,), torch.float32), ("weight_vol", weight_vol, (V,), torch.float32),
                               ("color_vol", color_vol, (3, V), torch.float32), ("label_vol", label_vol, (V,), torch.int32)):
        if t is not None and (tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous()):
            raise ValueError(f"tsdf_fusion_integrate: {name} must be a contiguous {dt} tensor of shape {shape}")
    if (colors is None) != (color_vol is None) or (labels is None) != (label_vol is None):
        raise ValueError("tsdf_fusion_integrate: colour / label frames and volumes go together")
    c = _f32(colors).contiguous() if colors is not None else None
    lb = labels.to(torch.int32).contiguous() if labels is not None else None
    if c is not None and tuple(c.shape) != (T, 3, H, W):
        raise ValueError("tsdf_fusion_integrate: colors must be (T,3,H,W)")
    if lb is not None and tuple(lb.shape) != (T, H, W):
        raise ValueError("tsdf_fusion_integrate: labels must be (T,H,W)")
    org = _origin3(origin)
    with torch.cuda.device(d.device):
        for t0 in range(0, T, GNB_MAX_FRAMES):
            n = min(GNB_MAX_FRAMES, T - t0)
            q = GnbFusionParams()
            q.nx, q.ny, q.nz = nx, ny, nz
            q.voxel_size = float(voxel_size)
            q.origin[0], q.origin[1], q.origin[2] = org
            q.trunc_margin = float(trunc_margin)
            q.n_frames, q.H, q.W = n, H, W
            q.h_projection = P[t0:t0 + n].data_ptr()
            q.depth = d[t0:t0 + n].data_ptr()
            q.color = c[t0:t0 + n].data_ptr() if c is not None else None
            q.label = lb[t0:t0 + n].data_ptr() if lb is not None else None
            q.tsdf_vol, q.weight_vol = tsdf_vol.data_ptr(), weight_vol.data_ptr()
            q.color_vol = color_vol.data_ptr() if color_vol is not None else None
            q.label_vol = label_vol.data_ptr() if label_vol is not None else None
            if depth_culling:
                nbytes = lib().gnb_tsdf_fusion_scratch_bytes(n, H, W)
                scratch = torch.empty(nbytes, dtype=torch.uint8, device=d.device)
                q.scratch, q.scratch_bytes = scratch.data_ptr(), nbytes
            check(lib().gnb_tsdf_fusion_integrate(C.byref(q), _stream()), "gnb_tsdf_fusion_integrate")
    return tsdf_vol, weight_vol


@_nvtx
def tsdf_fusion_finalize(tsdf_vol, weight_vol, color_vol=None):
    """The normalisation of TSDFFusion.get_tsdf (reference tsdf.py:426-434): vol / weight where weight > 0."""
    _need_cuda(tsdf_vol, weight_vol, color_vol)
    V = weight_vol.numel()
    out = torch.empty_like(tsdf_vol)
    cout = torch.empty_like(color_vol) if color_vol is not None else None
    with torch.cuda.device(tsdf_vol.device):
        check(lib().gnb_tsdf_fusion_finalize(tsdf_vol.data_ptr(), weight_vol.data_ptr(),
                                             color_vol.data_ptr() if color_vol is not None else None, V, out.data_ptr(),
                                             cout.data_ptr() if cout is not None else None, _stream()), "gnb_tsdf_fusion_finalize")
    return out, cout


# ------------------------------------------------------------------------------------------
# training-time ray sampler
# ------------------------------------------------------------------------------------------
@_nvtx
def sample_points_on_rays(h_idxs, w_idxs, depths, intrinsics, poses, N, M, delta, min_dist, sigma, gaussian_depths=None):
    """sample_points_on_rays (reference utils.py:458-540) in one launch: xyz_world (B,S,1+N+M,3), z (B,S,1+N+M).
    The M gaussian depths per ray are drawn here exactly as the reference draws them on its device (one
    torch.normal(D, sigma) call per camera, utils.py:496-498) unless `gaussian_depths` (B,S,M) is given."""
    _need_cuda(h_idxs, w_idxs, depths, intrinsics, poses, gaussian_depths)
    dev = depths.device
    B, S = depths.shape
    d = _f32(depths).contiguous()
    if gaussian_depths is None:
        gaussian_depths = torch.stack([torch.normal(d[b].unsqueeze(-1).expand(S, M), sigma * torch.ones((S, M), device=dev))
                                       for b in range(B)]) if B > 0 else torch.empty((0, S, M), device=dev)
    gd = _f32(gaussian_depths).contiguous()
    K = 1 + int(N) + int(M)
    xyz = torch.empty((B, S, K, 3), device=dev, dtype=torch.float32)
    z = torch.empty((B, S, K), device=dev, dtype=torch.float32)
    h = h_idxs.to(torch.long).contiguous()
    w = w_idxs.to(torch.long).contiguous()
    ki = _f32(intrinsics).contiguous()
    po = _f32(poses).contiguous()
    with torch.cuda.device(dev):
        check(lib().gnb_sample_points_on_rays(h.data_ptr(), w.data_ptr(), d.data_ptr(), ki.data_ptr(), po.data_ptr(), gd.data_ptr(),
                                              B, S, int(N), int(M), float(delta), float(min_dist), xyz.data_ptr(), z.data_ptr(),
                                              _stream()), "gnb_sample_points_on_rays")
    return xyz, z


@_nvtx
def sample_valid_depth_pixels(depth, num_samples):
    """sample_valid_depth_pixels (reference utils.py:340-363): depth (B,H,W) -> b_idxs (B,1), h_idxs, w_idxs (B,S) int64.
    The ranks are drawn as the reference draws them (one torch.randperm(n_valid[b], device)[:S] per map, in map order, so
    the same generator state selects the same pixels); the argwhere list itself is never materialised."""
    _need_cuda(depth)
    d = _f32(depth).contiguous()
    B, H, W = d.shape
    dev = d.device
    prefix = torch.empty((B, H), device=dev, dtype=torch.int32)
    nvalid = torch.empty((B,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib().gnb_valid_pixel_count(d.data_ptr(), B, H, W, prefix.data_ptr(), nvalid.data_ptr(), _stream()),
              "gnb_valid_pixel_count")
        counts = nvalid.tolist()                             # host sync, as torch.argwhere has in the reference
        if any(c < num_samples for c in counts):
            raise ValueError("Not enough non-zero depth pixels to sample from.")
        rank = torch.stack([torch.randperm(c, device=dev)[:num_samples] for c in counts]) if B > 0 else \
            torch.empty((0, num_samples), device=dev, dtype=torch.long)
        h = torch.empty((B, num_samples), device=dev, dtype=torch.long)
        w = torch.empty((B, num_samples), device=dev, dtype=torch.long)
        check(lib().gnb_valid_pixel_select(d.data_ptr(), B, H, W, prefix.data_ptr(), nvalid.data_ptr(), rank.contiguous().data_ptr(),
                                           int(num_samples), h.data_ptr(), w.data_ptr(), _stream()), "gnb_valid_pixel_select")
    return torch.arange(B, device=dev).unsqueeze(1), h, w

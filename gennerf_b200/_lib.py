"""ctypes binding of libgennerf_b200.so (include/gennerf_b200.h).

The library is the product's only compute path.  There is no CPU or eager-PyTorch
fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GNB_LIB_PATH selects another build of the same library (e.g. the tracing variant, `python -m gennerf_b200.build --trace`)
LIB_PATH = os.environ.get("GNB_LIB_PATH") or os.path.join(HERE, "libgennerf_b200.so")

GNB_MAX_FRAMES = 64
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
SCATTER_ATOMIC, SCATTER_DETERMINISTIC, SCATTER_ATOMIC_SUM = 0, 1, 2
POOL_MAX, POOL_MEAN = 0, 1
TC_FP16, TC_BF16 = 0, 1

c_float_p = C.POINTER(C.c_float)


class GnbLiftParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("voxel_size", C.c_float), ("origin", C.c_float * 3),
        ("batch", C.c_int32), ("n_frames", C.c_int32),
        ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("feat_layout", C.c_int32),
        ("features", C.c_void_p * GNB_MAX_FRAMES),
        ("h_projection", C.c_void_p),
        ("scratch", C.c_void_p),
        ("volume", C.c_void_p),
        ("vol_stride_b", C.c_int64), ("vol_stride_v", C.c_int64), ("vol_stride_c", C.c_int64),
        ("count", C.c_void_p), ("valid", C.c_void_p),
        ("accumulate", C.c_int32), ("mean", C.c_int32),
        ("x_begin", C.c_int32), ("x_end", C.c_int32),
    ]


class GnbSampleParams(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_query", C.c_int64), ("xyz", C.c_void_p),
        ("volume", C.c_void_p),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("C", C.c_int32),
        ("vol_stride_b", C.c_int64), ("vol_stride_x", C.c_int64), ("vol_stride_y", C.c_int64),
        ("vol_stride_z", C.c_int64), ("vol_stride_c", C.c_int64),
        ("voxel_size", C.c_float), ("origin", C.c_float * 3),
        ("plane", C.c_void_p * 3),
        ("R", C.c_int32), ("Cp", C.c_int32),
        ("pl_stride_b", C.c_int64), ("pl_stride_h", C.c_int64), ("pl_stride_w", C.c_int64),
        ("pl_stride_c", C.c_int64),
        ("padding", C.c_double),
        ("out", C.c_void_p), ("out_stride", C.c_int64),
        ("image", C.c_void_p), ("image_kchunks", C.c_int32), ("image_dtype", C.c_int32), ("image_status", C.c_void_p),
    ]


class GnbDecoderWeights(C.Structure):
    _fields_ = [
        ("d_feat", C.c_int32), ("d_code", C.c_int32), ("d_hidden", C.c_int32), ("n_blocks", C.c_int32),
        ("d_out", C.c_int32), ("d_geo", C.c_int32), ("alpha", C.c_float),
        ("use_code", C.c_int32), ("num_freqs", C.c_int32), ("freq_factor", C.c_float),
        ("include_input", C.c_int32),
        ("lin_in_w", C.c_void_p), ("lin_in_b", C.c_void_p),
        ("lin_z_w", C.c_void_p * 8), ("lin_z_b", C.c_void_p * 8),
        ("fc0_w", C.c_void_p * 8), ("fc0_b", C.c_void_p * 8),
        ("fc1_w", C.c_void_p * 8), ("fc1_b", C.c_void_p * 8),
        ("lin_out_w", C.c_void_p), ("lin_out_b", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("tc_dtype", C.c_int32),
        ("status", C.c_void_p),
        ("alpha_dev", C.c_void_p),
    ]


class GnbDecoderGrads(C.Structure):
    _fields_ = [
        ("lin_in_w", C.c_void_p), ("lin_in_b", C.c_void_p),
        ("lin_z_w", C.c_void_p * 8), ("lin_z_b", C.c_void_p * 8),
        ("fc0_w", C.c_void_p * 8), ("fc0_b", C.c_void_p * 8),
        ("fc1_w", C.c_void_p * 8), ("fc1_b", C.c_void_p * 8),
        ("lin_out_w", C.c_void_p), ("lin_out_b", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("alpha", C.c_void_p),
        ("g_code", C.c_void_p), ("g_feat", C.c_void_p),
    ]


class GnbFusionParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("voxel_size", C.c_float), ("origin", C.c_float * 3), ("trunc_margin", C.c_float),
        ("n_frames", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("h_projection", C.c_void_p), ("depth", C.c_void_p), ("color", C.c_void_p), ("label", C.c_void_p),
        ("tsdf_vol", C.c_void_p), ("weight_vol", C.c_void_p), ("color_vol", C.c_void_p), ("label_vol", C.c_void_p),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/gennerf_b200.h declares
SIGNATURES = {
    "gnb_version": (C.c_int, []),
    "gnb_last_error": (C.c_char_p, []),
    "gnb_struct_size": (C.c_int, [C.c_int]),
    "gnb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "gnb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "gnb_nchw_to_nhwc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "gnb_backproject_frames": (C.c_int, [C.POINTER(GnbLiftParams), C.c_void_p]),
    "gnb_project_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_sample_features": (C.c_int, [C.POINTER(GnbSampleParams), C.c_void_p]),
    "gnb_sample_binned_scratch_bytes": (C.c_int64, [C.POINTER(GnbSampleParams)]),
    "gnb_sample_features_binned": (C.c_int, [C.POINTER(GnbSampleParams), C.c_void_p, C.c_int64, C.c_void_p]),
    "gnb_plane_coords": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_scatter_scratch_bytes": (C.c_int64, [C.c_int, C.c_int64, C.c_int, C.c_int]),
    "gnb_scatter_mean_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gnb_scatter_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "gnb_backproject_frames_bwd": (C.c_int, [C.POINTER(GnbLiftParams), C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "gnb_sample_features_bwd": (C.c_int, [C.POINTER(GnbSampleParams), C.c_void_p, C.c_int64, C.c_void_p,
                                          C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "gnb_sample_features_bwd2": (C.c_int, [C.POINTER(GnbSampleParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "gnb_scatter_mean_planes_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                              C.c_double, C.c_void_p, C.c_void_p]),
    "gnb_pool_bwd_scratch_bytes": (C.c_int64, [C.c_int, C.c_int64, C.c_int, C.c_int]),
    "gnb_pool_local_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gnb_pool_scratch_bytes": (C.c_int64, [C.c_int, C.c_int64, C.c_int, C.c_int]),
    "gnb_pool_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gnb_get_3d_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gnb_farthest_point_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "gnb_sample_points_on_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_valid_pixel_count": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_valid_pixel_select": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_tsdf_fusion_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "gnb_tsdf_fusion_integrate": (C.c_int, [C.POINTER(GnbFusionParams), C.c_void_p]),
    "gnb_tsdf_fusion_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_positional_encoding": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "gnb_tsdf_head": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "gnb_decode_fp32": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "gnb_decoder_packed_bytes": (C.c_int64, [C.POINTER(GnbDecoderWeights)]),
    "gnb_decoder_pack_tc": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p]),
    "gnb_decoder_image_kchunks": (C.c_int, [C.POINTER(GnbDecoderWeights)]),
    "gnb_decode_image_tc": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "gnb_features_to_image": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_decode_tc_save": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_decode_tc": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_mlp_grad_link": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "gnb_mlp_grad_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_decode_train_bwd_workspace_bytes": (C.c_int64, [C.POINTER(GnbDecoderWeights), C.c_int64]),
    "gnb_decode_train_bwd": (C.c_int, [C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(GnbDecoderGrads), C.c_void_p, C.c_int64, C.c_int,
                                       C.c_void_p]),
    "gnb_query_fused_tc": (C.c_int, [C.POINTER(GnbSampleParams), C.POINTER(GnbDecoderWeights), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnb_query_fused_sorted_scratch_bytes": (C.c_int64, [C.POINTER(GnbSampleParams)]),
    "gnb_query_fused_sorted_tc": (C.c_int, [C.POINTER(GnbSampleParams), C.POINTER(GnbDecoderWeights), C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "gnb_query_grid_fused_tc": (C.c_int, [C.POINTER(GnbSampleParams), C.POINTER(C.c_int32), C.c_void_p,
                                            C.POINTER(GnbDecoderWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"gennerf_b200: {LIB_PATH} is missing -- build it with `python -m gennerf_b200.build` "
                "(there is no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        for which, st in enumerate((GnbLiftParams, GnbSampleParams, GnbDecoderWeights, GnbFusionParams, GnbDecoderGrads)):
            if L.gnb_struct_size(which) != C.sizeof(st):
                raise RuntimeError(f"gennerf_b200: ABI mismatch for {st.__name__}: library "
                                   f"{L.gnb_struct_size(which)} bytes, binding {C.sizeof(st)} bytes")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().gnb_last_error().decode(errors="replace")
        raise RuntimeError(f"gennerf_b200: {what} failed (code {rc}): {msg}")


def last_error():
    return lib().gnb_last_error().decode(errors="replace")


def get_option(name):
    v = C.c_int(0)
    check(lib().gnb_get_option(name.encode(), C.byref(v)), "gnb_get_option")
    return v.value


def set_option(name, value):
    """Tuning / debugging switch of the library (see gnb_set_option in include/gennerf_b200.h); returns the old value."""
    old = C.c_int(0)
    check(lib().gnb_get_option(name.encode(), C.byref(old)), "gnb_get_option")
    check(lib().gnb_set_option(name.encode(), int(value)), "gnb_set_option")
    return old.value

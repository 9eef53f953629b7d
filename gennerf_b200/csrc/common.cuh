// Shared helpers for the gennerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gennerf_b200.h"

namespace gnb {

void set_error(const char* fmt, ...);

// tuning / debugging options (error.cu): defaults read from the environment once, changed with gnb_set_option
enum Opt : int { OPT_TC_TWO_CTA = 0, OPT_TC_NO_EARLY, OPT_DEBUG_MAX_CLUSTERS, OPT_DEBUG_PRINT, OPT_LIFT_NVW, OPT_SCATTER_SCALAR,
                 OPT_FPS_SINGLE_CTA, OPT_FPS_CLUSTER, OPT_SAMPLE_GENERIC, OPT_BIN_UNIT, OPT_BIN_ROWCOPY, OPT_SCATTER_TILED,
                 OPT_BIN_PRESORTED, OPT_TC_NO_STG, OPT_TC_PAIR, OPT_DEBUG_NO_WCOPY, OPT_QUERY_FUSED, OPT_FPS_GRID, OPT_COUNT };
int opt(int which);

#define GNB_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            gnb::set_error(__VA_ARGS__);         \
            return GNB_E_INVALID;                \
        }                                        \
    } while (0)

#define GNB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            gnb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

#define GNB_LAUNCH_CHECK() GNB_CUDA(cudaGetLastError())

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 128-bit store (default policy: the volume is re-read from L2 by the sampler right after)
__device__ __forceinline__ void stcs4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
// Coordinate arithmetic shared by the sampler, the scatter kernels and the fused decoder.
// Every step is a separately rounded fp32 operation, as in the reference's eager PyTorch.
// ---------------------------------------------------------------------------------------

// normalize_coordinate (reference src/models/utils.py:88-97): u = p/den + 0.5, >=1 -> 1-1e-5, <0 -> 0
__device__ __forceinline__ float plane_unit(float p, float den) {
    float u = __fadd_rn(__fdiv_rn(p, den), 0.5f);
    // python: 1 - 10e-6 evaluated in double, stored into an fp32 tensor
    if (u >= 1.0f) u = (float)(1.0 - 10e-6);
    if (u < 0.0f) u = 0.0f;
    return u;
}

// grid_sample unnormalise (align_corners=True) + border clip, ATen GridSampler.h
__device__ __forceinline__ float unnorm_clip(float g, int size) {
    float x = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    return fminf((float)(size - 1), fmaxf(x, 0.0f));
}

// trilinear_interpolation's normalisation (reference src/models/utils.py:1018-1020)
__device__ __forceinline__ float query_grid(float x, float origin, float extent) {
    float t = __fdiv_rn(__fsub_rn(x, origin), extent);
    return __fsub_rn(__fmul_rn(2.0f, t), 1.0f);
}

}  // namespace gnb

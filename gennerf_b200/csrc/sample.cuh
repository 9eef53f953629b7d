// Point-query sampling arithmetic shared by the stand-alone sampler (sample.cu) and the fused
// sampler + decoder (decoder_tc.cu).
//
// Restates, for one query point,
//   trilinear_interpolation()          reference src/models/utils.py:999-1042
//       = F.grid_sample 3-D, bilinear, border, align_corners=True (ATen GridSampler.cpp)
//   GenNerf.sample_plane_feature() x3  reference src/models/model.py:153-161
//       = normalize_coordinate (utils.py:75-98) + F.grid_sample 2-D (ATen GridSamplerKernel.cpp)
// with the same operation order; each coordinate step is a separately rounded fp32 op.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gnb {

struct SampleKP {
    const float* xyz;
    long long Q;                 // queries per scene
    long long total;             // B * Q
    // volume
    const float* volume;
    int nx, ny, nz, C;
    long long vsb, vsx, vsy, vsz, vsc;
    float ext_x, ext_y, ext_z;   // fl32(fl32(n) * fl32(voxel_size))   (utils.py:1019)
    float ox, oy, oz;
    // planes (xz, xy, yz)
    const float* plane[3];
    int R, Cp;
    long long psb, psh, psw, psc;
    float den;                   // fl32(1 + padding + 10e-6)          (utils.py:88)
    // output
    float* out;
    long long out_stride;
    // optional 16-bit operand image of the tcgen05 decoder (GnbSampleParams.image)
    unsigned char* img;
    int img_kf, img_bf16;
    int* img_status;
};

// Stores the 4 feature columns [c, c+4) (c % 4 == 0; plane columns first) of flat query q: fp32 row and / or operand image.
__device__ __forceinline__ void store_feat4(const SampleKP& p, long long q, int c, float4 r) {
    if (p.out) *reinterpret_cast<float4*>(p.out + q * p.out_stride + c) = r;
    if (p.img) {
        uint2 pk;
        if (p.img_bf16) {
            __nv_bfloat162 a = __floats2bfloat162_rn(r.x, r.y), b = __floats2bfloat162_rn(r.z, r.w);
            pk.x = *reinterpret_cast<uint32_t*>(&a), pk.y = *reinterpret_cast<uint32_t*>(&b);
        } else {
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(r.y), "f"(r.x));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(r.w), "f"(r.z));
            // +-65504 exactly when satfinite clipped the value (decoder_tc.cu: sat_probe)
            if (p.img_status && ((((pk.x & 0x7FFF7FFFu) + 0x04010401u) | ((pk.y & 0x7FFF7FFFu) + 0x04010401u)) & 0x80008000u))
                atomicOr(p.img_status, 1);
        }
        const int row = (int)(q & 127), u = (c & 63) >> 3;
        const long long off = ((q >> 7) * p.img_kf + (c >> 6)) * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((u ^ (row & 7)) << 4) + ((c & 4) ? 8 : 0);
        *reinterpret_cast<uint2*>(p.img + off) = pk;
    }
}

struct TriCorners {              // trilinear: base offsets and weights of the 8 corners
    long long off[8];
    float w[8];
};

// Corner order = ATen's accumulation order: tnw, tne, tsw, tse, bnw, bne, bsw, bse with
// w/e = x0/x1, n/s = y0/y1, t/b = z0/z1.  Corners outside the grid (index == n, which only
// happens with weight 0 after the border clip) get weight 0 and a clamped, safe offset.
__device__ __forceinline__ void trilinear_setup(const SampleKP& p, float x, float y, float z, TriCorners& tc) {
    float ix = unnorm_clip(query_grid(x, p.ox, p.ext_x), p.nx);
    float iy = unnorm_clip(query_grid(y, p.oy, p.ext_y), p.ny);
    float iz = unnorm_clip(query_grid(z, p.oz, p.ext_z), p.nz);
    float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
    float x1 = x0 + 1.0f, y1 = y0 + 1.0f, z1 = z0 + 1.0f;
    float wx0 = x1 - ix, wx1 = ix - x0;
    float wy0 = y1 - iy, wy1 = iy - y0;
    float wz0 = z1 - iz, wz1 = iz - z0;
    int xi0 = (int)x0, yi0 = (int)y0, zi0 = (int)z0;
    bool bx = xi0 + 1 <= p.nx - 1, by = yi0 + 1 <= p.ny - 1, bz = zi0 + 1 <= p.nz - 1;
    long long ox0 = xi0 * p.vsx, ox1 = (bx ? xi0 + 1 : xi0) * p.vsx;
    long long oy0 = yi0 * p.vsy, oy1 = (by ? yi0 + 1 : yi0) * p.vsy;
    long long oz0 = zi0 * p.vsz, oz1 = (bz ? zi0 + 1 : zi0) * p.vsz;
    tc.off[0] = ox0 + oy0 + oz0; tc.w[0] = __fmul_rn(__fmul_rn(wx0, wy0), wz0);
    tc.off[1] = ox1 + oy0 + oz0; tc.w[1] = bx ? __fmul_rn(__fmul_rn(wx1, wy0), wz0) : 0.0f;
    tc.off[2] = ox0 + oy1 + oz0; tc.w[2] = by ? __fmul_rn(__fmul_rn(wx0, wy1), wz0) : 0.0f;
    tc.off[3] = ox1 + oy1 + oz0; tc.w[3] = (bx && by) ? __fmul_rn(__fmul_rn(wx1, wy1), wz0) : 0.0f;
    tc.off[4] = ox0 + oy0 + oz1; tc.w[4] = bz ? __fmul_rn(__fmul_rn(wx0, wy0), wz1) : 0.0f;
    tc.off[5] = ox1 + oy0 + oz1; tc.w[5] = (bx && bz) ? __fmul_rn(__fmul_rn(wx1, wy0), wz1) : 0.0f;
    tc.off[6] = ox0 + oy1 + oz1; tc.w[6] = (by && bz) ? __fmul_rn(__fmul_rn(wx0, wy1), wz1) : 0.0f;
    tc.off[7] = ox1 + oy1 + oz1; tc.w[7] = (bx && by && bz) ? __fmul_rn(__fmul_rn(wx1, wy1), wz1) : 0.0f;
}

struct BiCorners {               // bilinear on one plane: nw, ne, sw, se
    long long off[4];
    float w[4];
};

// u0 -> W axis (last), u1 -> H axis of the (B,C_p,R,R) plane (grid x indexes W).
__device__ __forceinline__ void bilinear_setup(const SampleKP& p, float u0, float u1, BiCorners& bc) {
    float ix = unnorm_clip(__fsub_rn(__fmul_rn(2.0f, u0), 1.0f), p.R);
    float iy = unnorm_clip(__fsub_rn(__fmul_rn(2.0f, u1), 1.0f), p.R);
    float x0 = floorf(ix), y0 = floorf(iy);
    float w = ix - x0, e = 1.0f - w;            // ATen 2-D CPU kernel: w = x - x_w, e = 1 - w
    float n = iy - y0, s = 1.0f - n;
    int xi0 = (int)x0, yi0 = (int)y0;
    bool bx = xi0 + 1 <= p.R - 1, by = yi0 + 1 <= p.R - 1;
    long long ox0 = xi0 * p.psw, ox1 = (bx ? xi0 + 1 : xi0) * p.psw;
    long long oy0 = yi0 * p.psh, oy1 = (by ? yi0 + 1 : yi0) * p.psh;
    bc.off[0] = oy0 + ox0; bc.w[0] = __fmul_rn(e, s);
    bc.off[1] = oy0 + ox1; bc.w[1] = bx ? __fmul_rn(w, s) : 0.0f;
    bc.off[2] = oy1 + ox0; bc.w[2] = by ? __fmul_rn(e, n) : 0.0f;
    bc.off[3] = oy1 + ox1; bc.w[3] = (bx && by) ? __fmul_rn(w, n) : 0.0f;
}

template <int VEC>
struct Vals {
    float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Vals<VEC> load_vals(const float* base, long long cstride) {
    Vals<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = ldg4(base);          // only instantiated for cstride == 1
        r.v[0] = t.x, r.v[1] = t.y, r.v[2] = t.z, r.v[3] = t.w;
    } else {
        r.v[0] = __ldg(base);
    }
    return r;
}

// Volume part for channels [c, c+VEC) of query (x,y,z) in scene b.
template <int VEC>
__device__ __forceinline__ Vals<VEC> sample_volume(const SampleKP& p, const TriCorners& tc, int b, int c) {
    const float* base = p.volume + b * p.vsb + c * p.vsc;
    Vals<VEC> val[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] = load_vals<VEC>(base + tc.off[k], p.vsc);
    Vals<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float a = __fmul_rn(val[0].v[i], tc.w[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) a = fmaf(val[k].v[i], tc.w[k], a);
        r.v[i] = a;
    }
    return r;
}

// Plane part: ((xz + xy) + yz), reference model.py:185-190.
template <int VEC>
__device__ __forceinline__ Vals<VEC> sample_planes(const SampleKP& p, const BiCorners bc[3], int b, int c) {
    Vals<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (p.plane[k] == nullptr) continue;
        const float* base = p.plane[k] + b * p.psb + c * p.psc;
        Vals<VEC> val[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) val[j] = load_vals<VEC>(base + bc[k].off[j], p.psc);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float a = __fmul_rn(val[0].v[i], bc[k].w[0]);
#pragma unroll
            for (int j = 1; j < 4; ++j) a = fmaf(val[j].v[i], bc[k].w[j], a);
            r.v[i] = __fadd_rn(r.v[i], a);
        }
    }
    return r;
}

__device__ __forceinline__ void planes_setup(const SampleKP& p, float x, float y, float z, BiCorners bc[3]) {
    float ux = plane_unit(x, p.den), uy = plane_unit(y, p.den), uz = plane_unit(z, p.den);
    bilinear_setup(p, ux, uz, bc[0]);      // 'xz': p[:, :, [0, 2]]
    bilinear_setup(p, ux, uy, bc[1]);      // 'xy'
    bilinear_setup(p, uy, uz, bc[2]);      // 'yz'
}

int fill_sample_kp(const GnbSampleParams* s, SampleKP& kp);

// counting sort of the queries by voxel brick (sample_binned.cu); 0 bytes = the parameters do not allow it
int64_t bin_sort_scratch_bytes(const GnbSampleParams* sp);
int bin_sort(const GnbSampleParams* sp, void* scratch, int64_t scratch_bytes, void* stream, const float4** sorted);

}  // namespace gnb

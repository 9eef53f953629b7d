// Stand-alone point-query sampler: xyz -> [plane features | volume features]   (sm_100a)
// Replaces GenNerf.map_features (reference src/models/model.py:163-204).  HBM/L2-bound gather:
// G lanes share a query, each lane one float4 of channels, so every corner fetch of a
// channels-last volume/plane is one contiguous C*4-byte run.
#include <stdlib.h>

#include "sample.cuh"

namespace gnb {

template <int VEC>
__global__ void __launch_bounds__(256) sample_kernel(const __grid_constant__ SampleKP p, int G) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = tid / G;
    const int sub = (int)(tid % G);
    if (q >= p.total) return;
    const int b = (int)(q / p.Q);
    const float x = __ldg(p.xyz + q * 3 + 0), y = __ldg(p.xyz + q * 3 + 1), z = __ldg(p.xyz + q * 3 + 2);
    float* __restrict__ out = p.out + q * p.out_stride;          // (VEC == 1 only: the image needs the float4 layout)
    int c_off = 0;
    if (p.Cp > 0) {
        BiCorners bc[3];
        planes_setup(p, x, y, z, bc);
        for (int c = sub * VEC; c < p.Cp; c += G * VEC) {
            Vals<VEC> r = sample_planes<VEC>(p, bc, b, c);
            if constexpr (VEC == 4) store_feat4(p, q, c, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
            else out[c] = r.v[0];
        }
        c_off = p.Cp;
    }
    if (p.volume) {
        TriCorners tc;
        trilinear_setup(p, x, y, z, tc);
        for (int c = sub * VEC; c < p.C; c += G * VEC) {
            Vals<VEC> r = sample_volume<VEC>(p, tc, b, c);
            if constexpr (VEC == 4) store_feat4(p, q, c_off + c, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
            else out[c_off + c] = r.v[0];
        }
    }
}

// Fast path (channels-last, C % 4 == 0, < 2^31 elements): a warp takes 32 queries.  Lane i sets up
// query i ONCE (normalisation, clip, floor, 8 + 12 corner offsets and weights) and parks the result
// in shared memory; then the warp walks the 32 queries with G lanes per query, each lane one
// float4 of channels, reading the corner table with broadcast LDS.128.  The setup arithmetic is
// the same device functions as the generic kernel, so both produce identical bits.
constexpr int ST_WORDS = 20;     // per query: 8 offsets + 8 weights (+4 pad: conflict-free LDS.128)

// volume part of 32 staged queries: tab = [32][ST_WORDS] corner table of this warp
__device__ __forceinline__ void staged_volume(const SampleKP& p, float* __restrict__ tab, int lane, int G, long long q0,
                                              long long nq, int c_off) {
    const int sub = lane % G, qpi = 32 / G;
    const int Cn = p.C;
    for (int it = 0; it < G; ++it) {
        const int ql = it * qpi + lane / G;
        const long long q = q0 + ql;
        if (q >= nq) continue;
        const float* e = tab + ql * ST_WORDS;
        const int4 o0 = *reinterpret_cast<const int4*>(e), o1 = *reinterpret_cast<const int4*>(e + 4);
        const float4 w0 = *reinterpret_cast<const float4*>(e + 8), w1 = *reinterpret_cast<const float4*>(e + 12);
        const int b = (int)(q / p.Q);
        for (int c = sub * 4; c < Cn; c += G * 4) {
            {
                const float* base = p.volume + b * p.vsb + c;
                const float4 v0 = ldg4(base + o0.x), v1 = ldg4(base + o0.y), v2 = ldg4(base + o0.z), v3 = ldg4(base + o0.w);
                const float4 v4 = ldg4(base + o1.x), v5 = ldg4(base + o1.y), v6 = ldg4(base + o1.z), v7 = ldg4(base + o1.w);
                float4 r = make_float4(__fmul_rn(v0.x, w0.x), __fmul_rn(v0.y, w0.x), __fmul_rn(v0.z, w0.x), __fmul_rn(v0.w, w0.x));
                r.x = fmaf(v1.x, w0.y, r.x), r.y = fmaf(v1.y, w0.y, r.y), r.z = fmaf(v1.z, w0.y, r.z), r.w = fmaf(v1.w, w0.y, r.w);
                r.x = fmaf(v2.x, w0.z, r.x), r.y = fmaf(v2.y, w0.z, r.y), r.z = fmaf(v2.z, w0.z, r.z), r.w = fmaf(v2.w, w0.z, r.w);
                r.x = fmaf(v3.x, w0.w, r.x), r.y = fmaf(v3.y, w0.w, r.y), r.z = fmaf(v3.z, w0.w, r.z), r.w = fmaf(v3.w, w0.w, r.w);
                r.x = fmaf(v4.x, w1.x, r.x), r.y = fmaf(v4.y, w1.x, r.y), r.z = fmaf(v4.z, w1.x, r.z), r.w = fmaf(v4.w, w1.x, r.w);
                r.x = fmaf(v5.x, w1.y, r.x), r.y = fmaf(v5.y, w1.y, r.y), r.z = fmaf(v5.z, w1.y, r.z), r.w = fmaf(v5.w, w1.y, r.w);
                r.x = fmaf(v6.x, w1.z, r.x), r.y = fmaf(v6.y, w1.z, r.y), r.z = fmaf(v6.z, w1.z, r.z), r.w = fmaf(v6.w, w1.z, r.w);
                r.x = fmaf(v7.x, w1.w, r.x), r.y = fmaf(v7.y, w1.w, r.y), r.z = fmaf(v7.z, w1.w, r.z), r.w = fmaf(v7.w, w1.w, r.w);
                store_feat4(p, q, c_off + c, r);
            }
        }
    }
}

constexpr int PT_WORDS = 28;     // per query: 3 planes x (4 offsets + 4 weights) (+4 pad)
constexpr int ST_WARPS = 4;      // warps per block of the staged kernel

// one plane's 4 corners: words 0-3 offsets, 4-7 weights
__device__ __forceinline__ float4 staged_plane(const SampleKP& p, const float* __restrict__ e, const float* __restrict__ base) {
    const int4 o = *reinterpret_cast<const int4*>(e);
    const float4 w = *reinterpret_cast<const float4*>(e + 4);
    const float4 v0 = ldg4(base + o.x), v1 = ldg4(base + o.y), v2 = ldg4(base + o.z), v3 = ldg4(base + o.w);
    float4 r = make_float4(__fmul_rn(v0.x, w.x), __fmul_rn(v0.y, w.x), __fmul_rn(v0.z, w.x), __fmul_rn(v0.w, w.x));
    r.x = fmaf(v1.x, w.y, r.x), r.y = fmaf(v1.y, w.y, r.y), r.z = fmaf(v1.z, w.y, r.z), r.w = fmaf(v1.w, w.y, r.w);
    r.x = fmaf(v2.x, w.z, r.x), r.y = fmaf(v2.y, w.z, r.y), r.z = fmaf(v2.z, w.z, r.z), r.w = fmaf(v2.w, w.z, r.w);
    r.x = fmaf(v3.x, w.w, r.x), r.y = fmaf(v3.y, w.w, r.y), r.z = fmaf(v3.z, w.w, r.z), r.w = fmaf(v3.w, w.w, r.w);
    return r;
}

// 9 resident blocks per SM = 56 registers: measured sweet spot (1 Mi queries, config-2 volume: 9 -> 108 us; 10 blocks /
// 48 registers split the 8 corner loads into two batches -> 160 us; 5 blocks / 93 registers -> 125 us).
__global__ void __launch_bounds__(32 * ST_WARPS, 9) sample_staged_kernel(const __grid_constant__ SampleKP p, int Gv, int Gp) {
    // per warp: volume corner table, then (only when there are planes) the plane corner tables -- dynamic shared memory, so
    // that a volume-only launch is not limited to 8 blocks per SM by 24.5 KB of tables it does not use
    extern __shared__ __align__(16) float s_staged[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tv = s_staged + warp * 32 * ST_WORDS;
    float* tp = s_staged + ST_WARPS * 32 * ST_WORDS + warp * 32 * PT_WORDS;
    const long long ngroups = (p.total + 31) / 32;
    for (long long grp = (long long)blockIdx.x * ST_WARPS + warp; grp < ngroups; grp += (long long)gridDim.x * ST_WARPS) {
        const long long q0 = grp * 32, q = q0 + lane;
        // ---- setup: one lane per query --------------------------------------------------
        if (q < p.total) {
            const float x = __ldg(p.xyz + q * 3), y = __ldg(p.xyz + q * 3 + 1), z = __ldg(p.xyz + q * 3 + 2);
            if (p.volume) {
                TriCorners tc;
                trilinear_setup(p, x, y, z, tc);
                float* e = tv + lane * ST_WORDS;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    reinterpret_cast<int*>(e)[k] = (int)tc.off[k];
                    e[8 + k] = tc.w[k];
                }
            }
            if (p.Cp > 0) {
                BiCorners bc[3];
                planes_setup(p, x, y, z, bc);
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    float* e = tp + lane * PT_WORDS + pl * 8;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        reinterpret_cast<int*>(e)[k] = (int)bc[pl].off[k];
                        e[4 + k] = bc[pl].w[k];
                    }
                }
            }
        }
        __syncwarp();
        // ---- gather: G lanes per query, one float4 of channels per lane ---------------------
        if (p.Cp > 0) {
            const int sub = lane % Gp, qpi = 32 / Gp;
            for (int it = 0; it < Gp; ++it) {
                const int ql = it * qpi + lane / Gp;
                const long long qq = q0 + ql;
                if (qq >= p.total) continue;
                const int b = (int)(qq / p.Q);
                for (int c = sub * 4; c < p.Cp; c += Gp * 4) {
                    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        if (p.plane[pl] == nullptr) continue;
                        const float4 a = staged_plane(p, tp + ql * PT_WORDS + pl * 8, p.plane[pl] + b * p.psb + c);
                        r.x = __fadd_rn(r.x, a.x), r.y = __fadd_rn(r.y, a.y), r.z = __fadd_rn(r.z, a.z), r.w = __fadd_rn(r.w, a.w);
                    }
                    store_feat4(p, qq, c, r);
                }
            }
        }
        if (p.volume) staged_volume(p, tv, lane, Gv, q0, p.total, p.Cp);
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------
// Backward of the sampler (ATen grid_sampler_{3d,2d}_backward with bilinear / border /
// align_corners=True, chained through the reference's coordinate normalisations):
//   grad_volume[corner_k, :] += w_k * grad_out[q, C_p:]        (8 corners, vector reductions)
//   grad_plane_p[corner_k,:] += w_k * grad_out[q, :C_p]        (4 corners x 3 planes)
//   grad_xyz[q] = sum_c grad_out[q,c] * d feat_c / d xyz       (optional)
// The coordinate gradient is zero where the border clip is active (ATen
// clip_coordinates_set_grad) and where normalize_coordinate clamps (masked assignment,
// utils.py:94-97).  One lane group of G lanes per query, as in the forward.
// ---------------------------------------------------------------------------------------
struct SampleBwdKP {
    SampleKP s;                    // forward parameters (volume / planes are read only for grad_xyz)
    const float* gout;             // (B,Q,gout_stride): [planes C_p | volume C]
    long long gout_stride;
    float* gvolume;                // same strides as the forward volume, or null
    float* gplane[3];              // same strides as the forward planes, or null
    float* gxyz;                   // (B,Q,3) or null
};

// d(unnormalised, clipped coordinate)/d(grid coordinate g): (size-1)/2 inside the open interval, else 0
__device__ __forceinline__ float unnorm_clip_grad(float g, int size) {
    const float x = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    return (x <= 0.0f || x >= (float)(size - 1)) ? 0.0f : 0.5f * (float)(size - 1);
}

__global__ void __launch_bounds__(256) sample_bwd_kernel(const __grid_constant__ SampleBwdKP p, int G) {
    const SampleKP& s = p.s;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = tid / G;
    const int sub = (int)(tid % G);
    const bool live = q < s.total;                        // keep every lane for the group shuffles
    const long long qq = live ? q : 0;
    const int b = (int)(qq / s.Q);
    const float x = __ldg(s.xyz + qq * 3 + 0), y = __ldg(s.xyz + qq * 3 + 1), z = __ldg(s.xyz + qq * 3 + 2);
    const float* __restrict__ go = p.gout + qq * p.gout_stride;
    float gx = 0.f, gy = 0.f, gz = 0.f;                   // d loss / d (x,y,z), partial over this lane's channels
    if (s.Cp > 0 && live) {
        BiCorners bc[3];
        planes_setup(s, x, y, z, bc);
        const float ux = plane_unit(x, s.den), uy = plane_unit(y, s.den), uz = plane_unit(z, s.den);
        const float u0[3] = {ux, ux, uy}, u1[3] = {uz, uy, uz};
        float gu[3] = {0.f, 0.f, 0.f};                    // gradient w.r.t. the unit coordinates ux, uy, uz
        const int a0[3] = {0, 0, 1}, a1[3] = {2, 1, 2};
        for (int c = sub * 4; c < s.Cp; c += G * 4) {
            float g4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) g4[e] = (c + e < s.Cp) ? __ldg(go + c + e) : 0.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (s.plane[k] == nullptr) continue;
                if (p.gplane[k]) {
                    float* gb = p.gplane[k] + b * s.psb + c * s.psc;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (bc[k].w[j] == 0.0f) continue;
                        if (s.psc == 1 && (c + 3 < s.Cp) && ((reinterpret_cast<uintptr_t>(gb + bc[k].off[j]) & 15) == 0)) {
                            atomicAdd(reinterpret_cast<float4*>(gb + bc[k].off[j]),        // one 16-byte reduction
                                      make_float4(bc[k].w[j] * g4[0], bc[k].w[j] * g4[1], bc[k].w[j] * g4[2], bc[k].w[j] * g4[3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c + e < s.Cp) atomicAdd(gb + bc[k].off[j] + e * s.psc, bc[k].w[j] * g4[e]);
                        }
                    }
                }
                if (p.gxyz) {
                    // weights: nw = e*s, ne = w*s, sw = e*n, se = w*n with w = ix-x0, e = 1-w, n = iy-y0, s = 1-n
                    const float ix = unnorm_clip(__fsub_rn(__fmul_rn(2.0f, u0[k]), 1.0f), s.R);
                    const float iy = unnorm_clip(__fsub_rn(__fmul_rn(2.0f, u1[k]), 1.0f), s.R);
                    const float w = ix - floorf(ix), e = 1.0f - w, n = iy - floorf(iy), sth = 1.0f - n;
                    const float* pb = s.plane[k] + b * s.psb + c * s.psc;
                    float dix = 0.f, diy = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        if (c + ch >= s.Cp) continue;
                        const float v0 = bc[k].w[0] != 0.f || true ? __ldg(pb + bc[k].off[0] + ch * s.psc) : 0.f;
                        const float v1 = __ldg(pb + bc[k].off[1] + ch * s.psc), v2 = __ldg(pb + bc[k].off[2] + ch * s.psc),
                                    v3 = __ldg(pb + bc[k].off[3] + ch * s.psc);
                        // corners beyond the border carry weight 0 in the forward and no gradient here
                        const float m1 = bc[k].off[1] != bc[k].off[0] ? 1.f : 0.f, m2 = bc[k].off[2] != bc[k].off[0] ? 1.f : 0.f;
                        const float m3 = (m1 != 0.f && m2 != 0.f) ? 1.f : 0.f;
                        dix += g4[ch] * (-v0 * sth + m1 * v1 * sth - m2 * v2 * n + m3 * v3 * n);
                        diy += g4[ch] * (-v0 * e - m1 * v1 * w + m2 * v2 * e + m3 * v3 * w);
                    }
                    // chain: vgrid = 2u - 1 (x2), unnormalise + clip
                    gu[a0[k]] += dix * unnorm_clip_grad(__fsub_rn(__fmul_rn(2.0f, u0[k]), 1.0f), s.R) * 2.0f;
                    gu[a1[k]] += diy * unnorm_clip_grad(__fsub_rn(__fmul_rn(2.0f, u1[k]), 1.0f), s.R) * 2.0f;
                }
            }
        }
        if (p.gxyz) {
            // u = p/den + 0.5, masked to constants where u >= 1 or u < 0
            const float raw[3] = {__fadd_rn(__fdiv_rn(x, s.den), 0.5f), __fadd_rn(__fdiv_rn(y, s.den), 0.5f),
                                  __fadd_rn(__fdiv_rn(z, s.den), 0.5f)};
            gx += (raw[0] >= 1.0f || raw[0] < 0.0f) ? 0.0f : gu[0] / s.den;
            gy += (raw[1] >= 1.0f || raw[1] < 0.0f) ? 0.0f : gu[1] / s.den;
            gz += (raw[2] >= 1.0f || raw[2] < 0.0f) ? 0.0f : gu[2] / s.den;
        }
    }
    if (s.volume && live) {
        TriCorners tc;
        trilinear_setup(s, x, y, z, tc);
        const float gxn = query_grid(x, s.ox, s.ext_x), gyn = query_grid(y, s.oy, s.ext_y), gzn = query_grid(z, s.oz, s.ext_z);
        const float ix = unnorm_clip(gxn, s.nx), iy = unnorm_clip(gyn, s.ny), iz = unnorm_clip(gzn, s.nz);
        const float fx = ix - floorf(ix), fy = iy - floorf(iy), fz = iz - floorf(iz);
        const float wx[2] = {1.0f - fx, fx}, wy[2] = {1.0f - fy, fy}, wz[2] = {1.0f - fz, fz};
        float dix = 0.f, diy = 0.f, diz = 0.f;
        for (int c = sub * 4; c < s.C; c += G * 4) {
            float g4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) g4[e] = (c + e < s.C) ? __ldg(go + s.Cp + c + e) : 0.0f;
            if (p.gvolume) {
                float* gb = p.gvolume + b * s.vsb + c * s.vsc;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (tc.w[k] == 0.0f) continue;
                    if (s.vsc == 1 && (c + 3 < s.C) && ((reinterpret_cast<uintptr_t>(gb + tc.off[k]) & 15) == 0)) {
                        atomicAdd(reinterpret_cast<float4*>(gb + tc.off[k]),
                                  make_float4(tc.w[k] * g4[0], tc.w[k] * g4[1], tc.w[k] * g4[2], tc.w[k] * g4[3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c + e < s.C) atomicAdd(gb + tc.off[k] + e * s.vsc, tc.w[k] * g4[e]);
                    }
                }
            }
            if (p.gxyz) {
                const float* vb = s.volume + b * s.vsb + c * s.vsc;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int bx = k & 1, byy = (k >> 1) & 1, bzz = (k >> 2) & 1;
                    // a corner beyond the border (clamped offset) has no value in the forward
                    const bool inb = (!bx || tc.off[k] != tc.off[k & ~1]) && (!byy || tc.off[k] != tc.off[k & ~2]) &&
                                     (!bzz || tc.off[k] != tc.off[k & ~4]);
                    if (!inb) continue;
                    float dot = 0.f;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c + e < s.C) dot += g4[e] * __ldg(vb + tc.off[k] + e * s.vsc);
                    dix += dot * (bx ? 1.0f : -1.0f) * wy[byy] * wz[bzz];
                    diy += dot * (byy ? 1.0f : -1.0f) * wx[bx] * wz[bzz];
                    diz += dot * (bzz ? 1.0f : -1.0f) * wx[bx] * wy[byy];
                }
            }
        }
        if (p.gxyz) {
            // chain: g = 2*((x - o)/ext) - 1, then unnormalise + clip
            gx += __fdiv_rn(2.0f * dix * unnorm_clip_grad(gxn, s.nx), s.ext_x);
            gy += __fdiv_rn(2.0f * diy * unnorm_clip_grad(gyn, s.ny), s.ext_y);
            gz += __fdiv_rn(2.0f * diz * unnorm_clip_grad(gzn, s.nz), s.ext_z);
        }
    }
    if (p.gxyz) {
        // reduce over the G lanes of the query (G is a power of two, groups are lane-aligned)
        for (int d = G >> 1; d > 0; d >>= 1) {
            gx += __shfl_xor_sync(FULL, gx, d);
            gy += __shfl_xor_sync(FULL, gy, d);
            gz += __shfl_xor_sync(FULL, gz, d);
        }
        if (live && sub == 0) {
            p.gxyz[q * 3 + 0] = gx, p.gxyz[q * 3 + 1] = gy, p.gxyz[q * 3 + 2] = gz;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Double backward of the sampler: what torch.autograd.grad(tsdf, xyz, create_graph=True) followed by
// loss.backward() needs (the eikonal / gradient losses, reference utils.py:636-649, model.py:385-400).
// The reference gets it for the planes from its pure-PyTorch grid_sample_2d (utils.py:1117-1174, model.py:157-158);
// for the volume ATen has no grid_sampler_3d double backward at all.
// sample_bwd_kernel is linear in grad_out.  With a = d L / d grad_xyz (B,Q,3), s_d = d i_d / d xyz_d (the chain through
// the normalisation and the border clip: piecewise constant), w_k the corner weights in pixel coordinates i:
//   g_gout[q,c]    = sum_k Dw_k V[k,c],          Dw_k = sum_d a_d s_d dw_k/di_d
//   g_volume[k,:] += Dw_k * grad_out[q,:]        (same for the planes)
//   g_xyz[q,e]     = s_e sum_{d != e} a_d s_d sum_k d2w_k/(di_d di_e) <grad_out[q,:], V[k,:]>
// (d2w/di_d^2 = 0: the weights are multilinear).  Same lane-group layout as sample_bwd_kernel.
// ---------------------------------------------------------------------------------------
struct SampleBwd2KP {
    SampleKP s;
    const float* gout;             // (B,Q,gout_stride): [planes C_p | volume C]
    long long gout_stride;
    const float* ggxyz;            // (B,Q,3): gradient w.r.t. the first backward's grad_xyz
    float* g_gout;                 // (B,Q,g_gout_stride) or null: overwritten
    long long g_gout_stride;
    float* g_volume;               // forward strides, accumulated into, or null
    float* g_plane[3];
    float* g_xyz;                  // (B,Q,3) or null: overwritten
};

__global__ void __launch_bounds__(256) sample_bwd2_kernel(const __grid_constant__ SampleBwd2KP p, int G) {
    const SampleKP& s = p.s;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = tid / G;
    const int sub = (int)(tid % G);
    const bool live = q < s.total;
    const long long qq = live ? q : 0;
    const int b = (int)(qq / s.Q);
    const float x = __ldg(s.xyz + qq * 3 + 0), y = __ldg(s.xyz + qq * 3 + 1), z = __ldg(s.xyz + qq * 3 + 2);
    const float av[3] = {__ldg(p.ggxyz + qq * 3 + 0), __ldg(p.ggxyz + qq * 3 + 1), __ldg(p.ggxyz + qq * 3 + 2)};
    const float* __restrict__ go = p.gout + qq * p.gout_stride;
    float* __restrict__ ggo = p.g_gout ? p.g_gout + qq * p.g_gout_stride : nullptr;
    float h[3] = {0.f, 0.f, 0.f};                         // g_xyz, partial over this lane's channels
    if (s.Cp > 0 && live) {
        BiCorners bc[3];
        planes_setup(s, x, y, z, bc);
        const float xyz3[3] = {x, y, z};
        float u[3], sc[3];                                // unit coordinates and d u / d p (0 where normalize_coordinate clamps)
        for (int d = 0; d < 3; ++d) {
            u[d] = plane_unit(xyz3[d], s.den);
            const float raw = __fadd_rn(__fdiv_rn(xyz3[d], s.den), 0.5f);
            sc[d] = (raw >= 1.0f || raw < 0.0f) ? 0.0f : __fdiv_rn(1.0f, s.den);
        }
        const int a0[3] = {0, 0, 1}, a1[3] = {2, 1, 2};
        float dw[3][4], cr[3][4], s0[3], s1[3], A0[3], A1[3], cross[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float vg0 = __fsub_rn(__fmul_rn(2.0f, u[a0[k]]), 1.0f), vg1 = __fsub_rn(__fmul_rn(2.0f, u[a1[k]]), 1.0f);
            const float ix = unnorm_clip(vg0, s.R), iy = unnorm_clip(vg1, s.R);
            const float w = ix - floorf(ix), e = 1.0f - w, n = iy - floorf(iy), sth = 1.0f - n;
            s0[k] = 2.0f * unnorm_clip_grad(vg0, s.R) * sc[a0[k]];
            s1[k] = 2.0f * unnorm_clip_grad(vg1, s.R) * sc[a1[k]];
            A0[k] = av[a0[k]] * s0[k], A1[k] = av[a1[k]] * s1[k];
            const float m1 = bc[k].off[1] != bc[k].off[0] ? 1.f : 0.f, m2 = bc[k].off[2] != bc[k].off[0] ? 1.f : 0.f;
            const float m3 = (m1 != 0.f && m2 != 0.f) ? 1.f : 0.f;
            // nw = e*s, ne = w*s, sw = e*n, se = w*n
            dw[k][0] = -A0[k] * sth - A1[k] * e;
            dw[k][1] = m1 * (A0[k] * sth - A1[k] * w);
            dw[k][2] = m2 * (-A0[k] * n + A1[k] * e);
            dw[k][3] = m3 * (A0[k] * n + A1[k] * w);
            cr[k][0] = 1.0f, cr[k][1] = -m1, cr[k][2] = -m2, cr[k][3] = m3;
        }
        for (int c = sub * 4; c < s.Cp; c += G * 4) {
            float g4[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) g4[e] = (c + e < s.Cp) ? __ldg(go + c + e) : 0.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (s.plane[k] == nullptr) continue;
                const float* pb = s.plane[k] + b * s.psb + c * s.psc;
                float* gb = p.g_plane[k] ? p.g_plane[k] + b * s.psb + c * s.psc : nullptr;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (cr[k][j] == 0.0f) continue;       // corner beyond the border: no value in the forward
                    float dot = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        if (c + ch >= s.Cp) continue;
                        const float v = __ldg(pb + bc[k].off[j] + ch * s.psc);
                        acc[ch] = fmaf(dw[k][j], v, acc[ch]);
                        dot = fmaf(g4[ch], v, dot);
                    }
                    cross[k] = fmaf(cr[k][j], dot, cross[k]);
                    if (gb && dw[k][j] != 0.0f) {
                        if (s.psc == 1 && (c + 3 < s.Cp) && ((reinterpret_cast<uintptr_t>(gb + bc[k].off[j]) & 15) == 0)) {
                            atomicAdd(reinterpret_cast<float4*>(gb + bc[k].off[j]),
                                      make_float4(dw[k][j] * g4[0], dw[k][j] * g4[1], dw[k][j] * g4[2], dw[k][j] * g4[3]));
                        } else {
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch)
                                if (c + ch < s.Cp) atomicAdd(gb + bc[k].off[j] + ch * s.psc, dw[k][j] * g4[ch]);
                        }
                    }
                }
            }
            if (ggo) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    if (c + ch < s.Cp) ggo[c + ch] = acc[ch];
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            h[a0[k]] = fmaf(s0[k] * A1[k], cross[k], h[a0[k]]);
            h[a1[k]] = fmaf(s1[k] * A0[k], cross[k], h[a1[k]]);
        }
    }
    if (s.volume && live) {
        TriCorners tc;
        trilinear_setup(s, x, y, z, tc);
        const float gxn = query_grid(x, s.ox, s.ext_x), gyn = query_grid(y, s.oy, s.ext_y), gzn = query_grid(z, s.oz, s.ext_z);
        const float ix = unnorm_clip(gxn, s.nx), iy = unnorm_clip(gyn, s.ny), iz = unnorm_clip(gzn, s.nz);
        const float fx = ix - floorf(ix), fy = iy - floorf(iy), fz = iz - floorf(iz);
        const float wx[2] = {1.0f - fx, fx}, wy[2] = {1.0f - fy, fy}, wz[2] = {1.0f - fz, fz};
        const float sx = __fdiv_rn(2.0f * unnorm_clip_grad(gxn, s.nx), s.ext_x), sy = __fdiv_rn(2.0f * unnorm_clip_grad(gyn, s.ny), s.ext_y),
                    sz = __fdiv_rn(2.0f * unnorm_clip_grad(gzn, s.nz), s.ext_z);
        const float Ax = av[0] * sx, Ay = av[1] * sy, Az = av[2] * sz;
        float dw[8], cxy[8], cxz[8], cyz[8];
        bool inb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int bx = k & 1, byy = (k >> 1) & 1, bzz = (k >> 2) & 1;
            inb[k] = (!bx || tc.off[k] != tc.off[k & ~1]) && (!byy || tc.off[k] != tc.off[k & ~2]) && (!bzz || tc.off[k] != tc.off[k & ~4]);
            const float gx = bx ? 1.0f : -1.0f, gy = byy ? 1.0f : -1.0f, gz = bzz ? 1.0f : -1.0f;
            dw[k] = inb[k] ? Ax * gx * wy[byy] * wz[bzz] + Ay * gy * wx[bx] * wz[bzz] + Az * gz * wx[bx] * wy[byy] : 0.0f;
            cxy[k] = gx * gy * wz[bzz], cxz[k] = gx * gz * wy[byy], cyz[k] = gy * gz * wx[bx];
        }
        float Sxy = 0.f, Sxz = 0.f, Syz = 0.f;
        for (int c = sub * 4; c < s.C; c += G * 4) {
            float g4[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) g4[e] = (c + e < s.C) ? __ldg(go + s.Cp + c + e) : 0.0f;
            const float* vb = s.volume + b * s.vsb + c * s.vsc;
            float* gb = p.g_volume ? p.g_volume + b * s.vsb + c * s.vsc : nullptr;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (!inb[k]) continue;
                float dot = 0.f;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    if (c + ch >= s.C) continue;
                    const float v = __ldg(vb + tc.off[k] + ch * s.vsc);
                    acc[ch] = fmaf(dw[k], v, acc[ch]);
                    dot = fmaf(g4[ch], v, dot);
                }
                Sxy = fmaf(cxy[k], dot, Sxy), Sxz = fmaf(cxz[k], dot, Sxz), Syz = fmaf(cyz[k], dot, Syz);
                if (gb && dw[k] != 0.0f) {
                    if (s.vsc == 1 && (c + 3 < s.C) && ((reinterpret_cast<uintptr_t>(gb + tc.off[k]) & 15) == 0)) {
                        atomicAdd(reinterpret_cast<float4*>(gb + tc.off[k]), make_float4(dw[k] * g4[0], dw[k] * g4[1], dw[k] * g4[2], dw[k] * g4[3]));
                    } else {
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch)
                            if (c + ch < s.C) atomicAdd(gb + tc.off[k] + ch * s.vsc, dw[k] * g4[ch]);
                    }
                }
            }
            if (ggo) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    if (c + ch < s.C) ggo[s.Cp + c + ch] = acc[ch];
            }
        }
        h[0] += sx * (Ay * Sxy + Az * Sxz);
        h[1] += sy * (Ax * Sxy + Az * Syz);
        h[2] += sz * (Ax * Sxz + Ay * Syz);
    }
    if (p.g_xyz) {
        for (int d = G >> 1; d > 0; d >>= 1) {
            h[0] += __shfl_xor_sync(FULL, h[0], d);
            h[1] += __shfl_xor_sync(FULL, h[1], d);
            h[2] += __shfl_xor_sync(FULL, h[2], d);
        }
        if (live && sub == 0) p.g_xyz[q * 3 + 0] = h[0], p.g_xyz[q * 3 + 1] = h[1], p.g_xyz[q * 3 + 2] = h[2];
    }
}

// fp32 feature rows -> the tcgen05 decoder's 16-bit lin_in operand image (the layout store_feat4 writes): every column of
// every tile is written, columns >= d_feat and the rows past n of the last tile as zeros (they meet zero weights, but
// 0 * NaN of an uninitialised buffer would still poison the accumulator).
__global__ void __launch_bounds__(256) features_to_image_kernel(const __grid_constant__ SampleKP p, const float* __restrict__ feat, long long n,
                                                                int d_feat, long long stride, int vec) {
    const int groups = p.img_kf * 16;                        // 4-column groups per row
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = idx / groups;
    const int c = (int)(idx % groups) * 4;
    if (q >= ((n + 127) >> 7 << 7)) return;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < n && c < d_feat) {
        const float* src = feat + q * stride + c;
        if (vec && c + 3 < d_feat) r = ldg4(src);
        else {
            r.x = __ldg(src);
            if (c + 1 < d_feat) r.y = __ldg(src + 1);
            if (c + 2 < d_feat) r.z = __ldg(src + 2);
            if (c + 3 < d_feat) r.w = __ldg(src + 3);
        }
    }
    store_feat4(p, q, c, r);
}

int fill_sample_kp(const GnbSampleParams* s, SampleKP& kp) {
    GNB_CHECK_ARG(s, "sample: null params");
    GNB_CHECK_ARG(s->batch >= 1 && s->n_query >= 0 && (s->xyz || s->n_query == 0), "sample: bad batch / n_query / xyz");
    bool has_planes = s->plane[0] || s->plane[1] || s->plane[2];
    GNB_CHECK_ARG(s->volume || has_planes, "sample: neither a volume nor planes given");
    kp.xyz = s->xyz;
    kp.Q = s->n_query;
    kp.total = (long long)s->batch * s->n_query;
    kp.volume = s->volume;
    kp.nx = kp.ny = kp.nz = 1;
    kp.C = 0;
    if (s->volume) {
        GNB_CHECK_ARG(s->nx > 0 && s->ny > 0 && s->nz > 0 && s->C > 0, "sample: bad volume shape");
        kp.nx = s->nx, kp.ny = s->ny, kp.nz = s->nz, kp.C = s->C;
        kp.vsb = s->vol_stride_b, kp.vsx = s->vol_stride_x, kp.vsy = s->vol_stride_y, kp.vsz = s->vol_stride_z;
        kp.vsc = s->vol_stride_c;
        // torch.tensor([nx,ny,nz]) * voxel_size: int64 tensor x python float -> fp32 product
        kp.ext_x = (float)s->nx * s->voxel_size;
        kp.ext_y = (float)s->ny * s->voxel_size;
        kp.ext_z = (float)s->nz * s->voxel_size;
        kp.ox = s->origin[0], kp.oy = s->origin[1], kp.oz = s->origin[2];
    }
    kp.R = 1, kp.Cp = 0;
    for (int k = 0; k < 3; ++k) kp.plane[k] = s->plane[k];
    if (has_planes) {
        GNB_CHECK_ARG(s->R > 0 && s->Cp > 0, "sample: bad plane shape");
        kp.R = s->R, kp.Cp = s->Cp;
        kp.psb = s->pl_stride_b, kp.psh = s->pl_stride_h, kp.psw = s->pl_stride_w, kp.psc = s->pl_stride_c;
        kp.den = (float)(1.0 + s->padding + 10e-6);
    }
    kp.out = s->out;
    kp.out_stride = s->out_stride;
    kp.img = (unsigned char*)s->image;
    kp.img_kf = s->image_kchunks, kp.img_bf16 = s->image_dtype == GNB_TC_BF16, kp.img_status = s->image_status;
    if (kp.img) {
        GNB_CHECK_ARG((reinterpret_cast<uintptr_t>(kp.img) & 15) == 0 && kp.img_kf >= 1 && 64 * kp.img_kf >= kp.C + kp.Cp,
                      "sample: bad operand image (16-byte aligned, 64 * image_kchunks >= C_p + C)");
        GNB_CHECK_ARG(s->image_dtype == GNB_TC_FP16 || s->image_dtype == GNB_TC_BF16, "sample: bad image_dtype");
    }
    return 0;
}

}  // namespace gnb

using namespace gnb;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int gnb_sample_features(const GnbSampleParams* s, void* stream) {
    SampleKP kp;
    int rc = fill_sample_kp(s, kp);
    if (rc) return rc;
    if (kp.total == 0) return 0;
    GNB_CHECK_ARG((s->out && s->out_stride >= kp.C + kp.Cp) || (!s->out && s->image), "sample: bad output");
    // float4 path: unit channel stride, channel counts and every base/stride a multiple of 4
    bool vec = !s->out || ((s->out_stride % 4 == 0) && aligned16(s->out));
    if (kp.volume)
        vec = vec && kp.vsc == 1 && kp.C % 4 == 0 && aligned16(kp.volume) && kp.vsb % 4 == 0 && kp.vsx % 4 == 0 &&
              kp.vsy % 4 == 0 && kp.vsz % 4 == 0;
    if (kp.Cp > 0) {
        vec = vec && kp.psc == 1 && kp.Cp % 4 == 0 && kp.psb % 4 == 0 && kp.psh % 4 == 0 && kp.psw % 4 == 0;
        for (int k = 0; k < 3; ++k) vec = vec && aligned16(kp.plane[k]);
    }
    // staged fast path: float4 channels-last and every element offset fits int32
    auto pow2_lanes = [](int n) { int g = 1; while (g < n && g < 32) g <<= 1; return g; };
    bool small = true;
    if (kp.volume) small = small && ((long long)kp.nx * kp.vsx + (long long)kp.ny * kp.vsy + (long long)kp.nz * kp.vsz < 0x7fffffffLL);
    if (kp.Cp > 0) small = small && ((long long)kp.R * kp.psh + (long long)kp.R * kp.psw < 0x7fffffffLL);
    if (vec && small && !opt(OPT_SAMPLE_GENERIC)) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        long long groups = (kp.total + 31) / 32;
        long long want = (groups + ST_WARPS - 1) / ST_WARPS;
        unsigned blocks = (unsigned)(want < (long long)sms * 16 ? want : (long long)sms * 16);
        const size_t smem = (size_t)ST_WARPS * 32 * (ST_WORDS + (kp.Cp > 0 ? PT_WORDS : 0)) * sizeof(float);
        sample_staged_kernel<<<blocks, 32 * ST_WARPS, smem, (cudaStream_t)stream>>>(kp, pow2_lanes(kp.C / 4 > 0 ? kp.C / 4 : 1),
                                                                        pow2_lanes(kp.Cp / 4 > 0 ? kp.Cp / 4 : 1));
        GNB_LAUNCH_CHECK();
        return 0;
    }
    if (kp.img && !vec) {
        set_error("sample: the operand image needs unit channel strides, channel counts % 4 == 0 and 16-byte aligned bases");
        return GNB_E_UNSUPPORTED;
    }
    int cmax = kp.C > kp.Cp ? kp.C : kp.Cp;
    int lanes = vec ? cmax / 4 : cmax;
    int G = 1;
    while (G < lanes && G < 32) G <<= 1;
    long long threads = kp.total * G;
    unsigned blocks = (unsigned)((threads + 255) / 256);
    if (vec)
        sample_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(kp, G);
    else
        sample_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(kp, G);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_sample_features_bwd(const GnbSampleParams* s, const float* grad_out, int64_t grad_out_stride,
                                       float* grad_volume, float* const* h_grad_planes3, float* grad_xyz, void* stream) {
    SampleBwdKP kp;
    int rc = fill_sample_kp(s, kp.s);
    if (rc) return rc;
    GNB_CHECK_ARG(grad_out && grad_out_stride >= kp.s.C + kp.s.Cp, "gnb_sample_features_bwd: bad grad_out");
    GNB_CHECK_ARG(grad_volume || h_grad_planes3 || grad_xyz, "gnb_sample_features_bwd: nothing to compute");
    if (kp.s.total == 0) return 0;
    kp.gout = grad_out, kp.gout_stride = grad_out_stride;
    kp.gvolume = kp.s.volume ? grad_volume : nullptr;
    for (int k = 0; k < 3; ++k) kp.gplane[k] = (h_grad_planes3 && kp.s.plane[k]) ? h_grad_planes3[k] : nullptr;
    kp.gxyz = grad_xyz;
    int cmax = kp.s.C > kp.s.Cp ? kp.s.C : kp.s.Cp;
    int lanes = (cmax + 3) / 4, G = 1;
    while (G < lanes && G < 32) G <<= 1;
    long long threads = kp.s.total * G;
    sample_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kp, G);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_sample_features_bwd2(const GnbSampleParams* s, const float* grad_out, int64_t grad_out_stride, const float* gg_xyz,
                                        float* g_grad_out, int64_t g_grad_out_stride, float* g_volume, float* const* h_g_planes3,
                                        float* g_xyz, void* stream) {
    SampleBwd2KP kp;
    int rc = fill_sample_kp(s, kp.s);
    if (rc) return rc;
    GNB_CHECK_ARG(grad_out && grad_out_stride >= kp.s.C + kp.s.Cp && gg_xyz, "gnb_sample_features_bwd2: bad grad_out / gg_xyz");
    GNB_CHECK_ARG(g_grad_out || g_volume || h_g_planes3 || g_xyz, "gnb_sample_features_bwd2: nothing to compute");
    GNB_CHECK_ARG(!g_grad_out || g_grad_out_stride >= kp.s.C + kp.s.Cp, "gnb_sample_features_bwd2: bad g_grad_out stride");
    if (kp.s.total == 0) return 0;
    kp.gout = grad_out, kp.gout_stride = grad_out_stride, kp.ggxyz = gg_xyz;
    kp.g_gout = g_grad_out, kp.g_gout_stride = g_grad_out_stride;
    kp.g_volume = kp.s.volume ? g_volume : nullptr;
    for (int k = 0; k < 3; ++k) kp.g_plane[k] = (h_g_planes3 && kp.s.plane[k]) ? h_g_planes3[k] : nullptr;
    kp.g_xyz = g_xyz;
    int cmax = kp.s.C > kp.s.Cp ? kp.s.C : kp.s.Cp;
    int lanes = (cmax + 3) / 4, G = 1;
    while (G < lanes && G < 32) G <<= 1;
    long long threads = kp.s.total * G;
    sample_bwd2_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kp, G);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_features_to_image(const float* feat, int64_t n_rows, int d_feat, int64_t feat_stride, int image_kchunks, int image_dtype,
                                     void* image, int32_t* status, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0 && d_feat >= 1 && feat_stride >= d_feat && image_kchunks >= 1 && 64 * image_kchunks >= d_feat,
                  "gnb_features_to_image: bad shape (64 * image_kchunks >= d_feat, feat_stride >= d_feat)");
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(feat && image && (reinterpret_cast<uintptr_t>(image) & 15) == 0, "gnb_features_to_image: null / unaligned pointer");
    GNB_CHECK_ARG(image_dtype == GNB_TC_FP16 || image_dtype == GNB_TC_BF16, "gnb_features_to_image: bad image_dtype");
    SampleKP kp = {};
    kp.img = (unsigned char*)image, kp.img_kf = image_kchunks, kp.img_bf16 = image_dtype == GNB_TC_BF16, kp.img_status = status;
    const int vec = (feat_stride % 4 == 0) && aligned16(feat);
    const long long rows = (n_rows + 127) / 128 * 128;
    const long long threads = rows * image_kchunks * 16;
    features_to_image_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kp, feat, n_rows, d_feat, feat_stride, vec);
    GNB_LAUNCH_CHECK();
    return 0;
}

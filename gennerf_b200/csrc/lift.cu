// Fused volumetric back-projection ("lift") for sm_100a.
//
// Replaces, per scene and for all T frames in ONE pass over the voxels:
//   coordinates()            reference src/data/tsdf.py:25-40       (never materialised)
//   backproject()            reference src/models/utils.py:948-996
//   accumulation in encode() reference src/models/model.py:121-127
//   normalisation            reference src/models/model.py:195-199  (identity on the sum, trap T2)
//
// Data layout: features are channels-last (H,W,C): a voxel-frame reads one contiguous
// C*4-byte run (one 128 B line at C=32).  G = min(32, C/4) lanes share a voxel, each lane one
// float4.  A warp owns NVW consecutive voxels (consecutive z); lane i projects voxel i ONCE
// per frame and the pixel offset is handed to the G lanes that fetch it with __shfl_sync, so
// the projection arithmetic is not repeated per channel group.  Sums stay in registers over
// the T frames (frame order == the reference's summation order => bit-exact) and are written
// once with streaming stores.  HBM-bound: algorithmic bytes = features read once + volume
// written once.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace gnb {

struct LiftKP {
    const float* feat[GNB_MAX_FRAMES];   // per frame, this scene, NHWC
    float P[GNB_MAX_FRAMES][12];
    int T, H, W, C;
    int nx, ny, nz;
    long long V;
    float vs, ox, oy, oz;
    float* volume;
    long long stride_v, stride_c;
    int* count;
    unsigned char* valid;
    int accumulate, mean;
    int x_begin, x_end;                  // slab of the grid this launch computes
};
static_assert(sizeof(LiftKP) <= 4096, "kernel parameter block must stay below 4 KB");

// Projection of one voxel by one 3x4 matrix with the reference's exact arithmetic
// (SURVEY trap T3): torch.bmm(P, [w;1]) on the CPU is the FMA chain
//   fma(P3, 1, fma(P2, wz, fma(P1, wy, P0*wx)));  px = rint(cx/cz) (round-half-even, T1).
// Returns the pixel offset py*W+px, or -1 when the frame does not see the voxel.
__device__ __forceinline__ int project_voxel(const float* __restrict__ P, float wx, float wy, float wz,
                                             int H, int W, float* fx_out = nullptr, float* fy_out = nullptr) {
    float cx = __fadd_rn(__fmaf_rn(P[2], wz, __fmaf_rn(P[1], wy, __fmul_rn(P[0], wx))), P[3]);
    float cy = __fadd_rn(__fmaf_rn(P[6], wz, __fmaf_rn(P[5], wy, __fmul_rn(P[4], wx))), P[7]);
    float cz = __fadd_rn(__fmaf_rn(P[10], wz, __fmaf_rn(P[9], wy, __fmul_rn(P[8], wx))), P[11]);
    // Cheap rejection before the two IEEE divisions (most projections that survive the brick culling still miss the
    // image): behind the camera, or the exact quotient is more than a quarter pixel outside the range that can round
    // to a valid pixel -- the correctly rounded quotient then rounds outside too, so the result is unchanged.
    if (!fx_out) {
        if (!(cz > 0.0f)) return -1;
        if (cx < -0.75f * cz || cx > ((float)W - 0.25f) * cz || cy < -0.75f * cz || cy > ((float)H - 0.25f) * cz) return -1;
    }
    float fx, fy;
    if (fx_out) {
        fx = rintf(__fdiv_rn(cx, cz));
        fy = rintf(__fdiv_rn(cy, cz));
        *fx_out = fx, *fy_out = fy;
    } else {
        // Only the ROUNDED pixel matters, so the two IEEE divisions are replaced by one reciprocal: qa = c * rcp(cz) is within
        // a few ulp of the exact quotient (MUFU.RCP: 1 ulp, two roundings), hence within 2^-20 |qa| of the correctly rounded
        // one; if qa is farther than that from a half-integer, both round to the same integer.  Near a tie (and for
        // NaN / inf / |qa| >= 2^19, where the test below is false) the IEEE division decides, as in the reference.
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(cz));
        const float qx = __fmul_rn(cx, r), qy = __fmul_rn(cy, r);
        fx = rintf(qx), fy = rintf(qy);
        const float mx = 0.5f - fabsf(qx - fx), my = 0.5f - fabsf(qy - fy);       // margins to the nearest tie
        const bool safe = (mx > fabsf(qx) * 9.5367431640625e-7f + 1e-30f) && (my > fabsf(qy) * 9.5367431640625e-7f + 1e-30f);
        if (!safe) {
            fx = rintf(__fdiv_rn(cx, cz));
            fy = rintf(__fdiv_rn(cy, cz));
        }
    }
    // float comparisons == the reference's int64 comparisons for every finite value; NaN/inf
    // (cz == 0) compare false here and convert to INT64_MIN (invalid) there.
    bool ok = (fx >= 0.0f) && (fy >= 0.0f) && (fx < (float)W) && (fy < (float)H) && (cz > 0.0f);
    return ok ? (int)fy * W + (int)fx : -1;
}

__device__ __forceinline__ void voxel_world(long long v, int ny, int nz, float vs, float ox, float oy, float oz,
                                            float& wx, float& wy, float& wz) {
    int iz = (int)(v % nz);
    long long r = v / nz;
    int iy = (int)(r % ny);
    int ix = (int)(r / ny);
    // world = fl(i) * voxel_size + origin: two separately rounded operations (utils.py:974)
    wx = __fadd_rn(__fmul_rn((float)ix, vs), ox);
    wy = __fadd_rn(__fmul_rn((float)iy, vs), oy);
    wz = __fadd_rn(__fmul_rn((float)iz, vs), oz);
}

// G lanes per voxel, VEC floats per lane (4: float4 path, needs C % 4 == 0; 1: scalar path).
// A block owns a compact BX x BY x 16 brick of voxels (8 warps x NVW voxels); grid.y = channel
// chunks of G*VEC.  CL: channels-last volume (stride_c == 1); otherwise the reference's (C,V) layout.
//
// Frame culling: before the per-voxel work the block projects the 8 corner voxels of its brick by
// every frame.  A frame whose 8 corners all lie behind the camera, or all lie (with cz > 0) more
// than a pixel outside the same image border, cannot see any voxel of the brick (voxel centres
// are convex combinations of the corner centres and x/z is linear-fractional), so it is skipped
// for the whole block.  Frames are still visited in ascending order, so the sum order is kept.
// Resident blocks per SM: the kernel is latency / issue bound, so occupancy pays more than registers -- 6 blocks (40 registers, a
// few spilled words) for up to 8 lanes per voxel (C <= 32: config 4 514 -> 414 us), 5 blocks (48 registers) above (C = 128: 80 -> 70 us).
template <int G, int VEC, bool CL, int NVW>
__global__ void __launch_bounds__(256, (G <= 8 ? 6 : 5)) lift_kernel(const __grid_constant__ LiftKP p) {
    // NVW = voxels per warp (power of two, 32/G <= NVW <= 32)
    constexpr int ITER = NVW * G / 32;                     // gathers per lane per frame
    constexpr int VPI = 32 / G;                            // voxels per gather instruction
    constexpr int NVB = 8 * NVW;                           // voxels per block: 256 / 128 / 64
    constexpr int BZ = NVB >= 64 ? 16 : 8, BY = NVB >= 256 ? 4 : 2, BX = NVB / (BZ * BY);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbz = (p.nz + BZ - 1) / BZ, nby = (p.ny + BY - 1) / BY;
    const int bz = blockIdx.x % nbz, by = (blockIdx.x / nbz) % nby, bx = blockIdx.x / (nbz * nby);
    const int x0 = p.x_begin + bx * BX, y0 = by * BY, z0 = bz * BZ;

    // ---- which frames can see this brick ---------------------------------------------------
    __shared__ unsigned long long s_vis;
    if (threadIdx.x == 0) s_vis = 0ull;
    __syncthreads();
    {
        const int x1 = min(x0 + BX, p.x_end) - 1, y1 = min(y0 + BY, p.ny) - 1, z1 = min(z0 + BZ, p.nz) - 1;
        for (int base = 0; base < p.T * 8; base += 256) {
            const int task = base + threadIdx.x;
            const bool active = task < p.T * 8;               // whole 8-lane groups are active or not
            const int t = active ? (task >> 3) : 0, corner = task & 7;
            const float cwx = __fadd_rn(__fmul_rn((float)((corner & 1) ? x1 : x0), p.vs), p.ox);
            const float cwy = __fadd_rn(__fmul_rn((float)((corner & 2) ? y1 : y0), p.vs), p.oy);
            const float cwz = __fadd_rn(__fmul_rn((float)((corner & 4) ? z1 : z0), p.vs), p.oz);
            const float* P = p.P[t];
            const float cx = fmaf(P[2], cwz, fmaf(P[1], cwy, P[0] * cwx)) + P[3];
            const float cy = fmaf(P[6], cwz, fmaf(P[5], cwy, P[4] * cwx)) + P[7];
            const float cz = fmaf(P[10], cwz, fmaf(P[9], cwy, P[8] * cwx)) + P[11];
            const bool front = cz > 1e-3f;
            const float u = cx / cz, v = cy / cz;
            const unsigned sh = lane & ~7u;
            const bool all_behind = ((__ballot_sync(FULL, cz < -1e-3f) >> sh) & 0xffu) == 0xffu;
            const bool all_left = ((__ballot_sync(FULL, front && u < -2.0f) >> sh) & 0xffu) == 0xffu;
            const bool all_right = ((__ballot_sync(FULL, front && u > (float)p.W + 1.0f) >> sh) & 0xffu) == 0xffu;
            const bool all_top = ((__ballot_sync(FULL, front && v < -2.0f) >> sh) & 0xffu) == 0xffu;
            const bool all_bottom = ((__ballot_sync(FULL, front && v > (float)p.H + 1.0f) >> sh) & 0xffu) == 0xffu;
            if (active && corner == 0 && !(all_behind || all_left || all_right || all_top || all_bottom))
                atomicOr(&s_vis, 1ull << t);
        }
    }
    __syncthreads();
    unsigned long long vis = s_vis;

    const int sub = lane % G;
    const int c0 = (blockIdx.y * G + sub) * VEC;           // first channel of this lane
    const bool c_ok = c0 < p.C;

    // lane i (< NVW) owns voxel i of this warp's part of the brick for the projection; with NVW <= 16 the lanes
    // [NVW, 2 NVW) project the same voxels by the trip's SECOND frame (SPLIT), so a trip of two frames costs the issue
    // slots of one projection
    constexpr bool SPLIT = NVW <= 16;
    const int pl = lane & (NVW - 1), ph = lane / NVW;      // projection voxel / which frame of the trip
    const int iv = warp * NVW + pl;
    const int vx = x0 + iv / (BZ * BY), vy = y0 + (iv / BZ) % BY, vz = z0 + iv % BZ;
    const bool proj = (SPLIT ? ph < 2 : true) && vx < p.x_end && vy < p.ny && vz < p.nz;
    const bool own = (lane < NVW) && proj;
    const int v_own = own ? (vx * p.ny + vy) * p.nz + vz : -1;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    if (proj) {
        // world = fl(i) * voxel_size + origin: two separately rounded operations (utils.py:974)
        wx = __fadd_rn(__fmul_rn((float)vx, p.vs), p.ox);
        wy = __fadd_rn(__fmul_rn((float)vy, p.vs), p.oy);
        wz = __fadd_rn(__fmul_rn((float)vz, p.vs), p.oz);
    }
    int vj[ITER];                                          // voxel handled by gather slot j of this lane
#pragma unroll
    for (int j = 0; j < ITER; ++j) vj[j] = __shfl_sync(FULL, v_own, j * VPI + lane / G);

    float acc[ITER][VEC];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < ITER; ++j) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[j][k] = 0.0f;
    }
    if (p.accumulate) {
        if (own && p.count) cnt = p.count[v_own];
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
            if (vj[j] >= 0 && c_ok) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[j][k] = p.volume[(long long)vj[j] * p.stride_v + (c0 + k) * p.stride_c];
            }
        }
    }

    while (vis) {                                          // ascending frame order == reference summation order
        // two frames per trip: both projections, then all gathers of both frames in flight together
        const int t0 = __ffsll((long long)vis) - 1;
        vis &= vis - 1;
        const int t1 = vis ? __ffsll((long long)vis) - 1 : -1;
        if (t1 >= 0) vis &= vis - 1;
        int off0 = -1, off1 = -1;
        if constexpr (SPLIT) {
            const int t = ph == 0 ? t0 : t1;
            int off = -1;
            if (proj && t >= 0) off = project_voxel(p.P[t], wx, wy, wz, p.H, p.W);
            off0 = __shfl_sync(FULL, off, pl);
            off1 = __shfl_sync(FULL, off, pl + NVW);
        } else if (own) {
            off0 = project_voxel(p.P[t0], wx, wy, wz, p.H, p.W);
            if (t1 >= 0) off1 = project_voxel(p.P[t1], wx, wy, wz, p.H, p.W);
        }
        cnt += (off0 >= 0) + (off1 >= 0);
        if (__ballot_sync(FULL, (off0 & off1) >= 0 || off0 >= 0 || off1 >= 0) == 0u) continue;      // warp-uniform
        const float* __restrict__ f0 = p.feat[t0];
        const float* __restrict__ f1 = p.feat[t1 >= 0 ? t1 : t0];
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
            const int o0 = __shfl_sync(FULL, off0, j * VPI + lane / G);
            const int o1 = __shfl_sync(FULL, off1, j * VPI + lane / G);
            float x0[VEC], x1[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) x0[k] = x1[k] = 0.0f;       // + 0.0f leaves the sum bit-identical
            if (o0 >= 0 && c_ok) {
                const float* src = f0 + (long long)o0 * p.C + c0;
                if constexpr (VEC == 4) { const float4 v = ldg4(src); x0[0] = v.x, x0[1] = v.y, x0[2] = v.z, x0[3] = v.w; }
                else x0[0] = __ldg(src);
            }
            if (o1 >= 0 && c_ok) {
                const float* src = f1 + (long long)o1 * p.C + c0;
                if constexpr (VEC == 4) { const float4 v = ldg4(src); x1[0] = v.x, x1[1] = v.y, x1[2] = v.z, x1[3] = v.w; }
                else x1[0] = __ldg(src);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[j][k] = __fadd_rn(__fadd_rn(acc[j][k], x0[k]), x1[k]);
        }
    }

    if (p.mean) {
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
            int n = __shfl_sync(FULL, cnt, j * VPI + lane / G);
            float d = (float)(n > 0 ? n : 1);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[j][k] = __fdiv_rn(acc[j][k], d);
        }
    }

    // ---- write once ------------------------------------------------------------------
    if (own && blockIdx.y == 0) {
        if (p.count) p.count[v_own] = cnt;
        if (p.valid) p.valid[v_own] = (unsigned char)(cnt > 0);
    }
    if constexpr (CL) {
        // channels-last volume: every voxel is one contiguous C*4-byte run
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
            if (vj[j] >= 0 && c_ok) {
                float* dst = p.volume + (long long)vj[j] * p.stride_v + c0;
                if constexpr (VEC == 4) {
                    stcs4(dst, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
                } else {
                    dst[0] = acc[j][0];
                }
            }
        }
    } else {
        // reference layout (C,V): transpose the warp's NVW x (G*VEC) tile through shared memory so
        // that each channel row is written as runs of consecutive z
        constexpr int NC = G * VEC;
        __shared__ float tile[8][NC][NVW + 1];
        float(*tw)[NVW + 1] = tile[warp];
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) tw[sub * VEC + k][j * VPI + lane / G] = acc[j][k];
        }
        __syncwarp();
        const int cbase = blockIdx.y * NC;
#pragma unroll 4
        for (int idx = lane; idx < NC * NVW; idx += 32) {           // NC * NVW is a multiple of 32
            const int c = idx / NVW, vv = idx % NVW;
            const int v = __shfl_sync(FULL, v_own, vv);
            if (cbase + c < p.C && v >= 0) p.volume[(long long)v * p.stride_v + (long long)(cbase + c) * p.stride_c] = tw[c][vv];
        }
    }
}

// ---- NCHW -> NHWC ---------------------------------------------------------------------
struct TransposeKP {
    const float* src[GNB_MAX_FRAMES];
    float* dst;
    int T, B, C;
    long long HW;
};

// One (b,t) image per blockIdx.z; a block moves a 32-channel x 128-pixel tile through shared memory:
// float4 loads along the pixel axis (NCHW rows), float4 stores along the channel axis (NHWC rows).  The tile row
// stride of 129 floats keeps the transposed reads conflict-free.
constexpr int TP_PX = 128;
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const __grid_constant__ TransposeKP p) {
    __shared__ float tile[32][TP_PX + 1];
    const int t = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const float* __restrict__ src = p.src[t] + (long long)b * p.C * p.HW;
    float* __restrict__ dst = p.dst + ((long long)t * p.B + b) * p.HW * p.C;
    const long long px0 = (long long)blockIdx.x * TP_PX;
    const int c0 = blockIdx.y * 32;
    const bool vec_in = (p.HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    // ---- load: 32 rows (channels) x 32 float4 (128 pixels) = 1024 float4, 4 per thread ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int r = idx >> 5, q = idx & 31;              // channel row, float4 column
        const int c = c0 + r;
        const long long px = px0 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.C) {
            const float* s = src + (long long)c * p.HW + px;
            if (vec_in && px + 3 < p.HW) v = ldg4(s);
            else {
                if (px < p.HW) v.x = __ldg(s);
                if (px + 1 < p.HW) v.y = __ldg(s + 1);
                if (px + 2 < p.HW) v.z = __ldg(s + 2);
                if (px + 3 < p.HW) v.w = __ldg(s + 3);
            }
        }
        tile[r][q * 4 + 0] = v.x, tile[r][q * 4 + 1] = v.y, tile[r][q * 4 + 2] = v.z, tile[r][q * 4 + 3] = v.w;
    }
    __syncthreads();
    // ---- store: 128 pixels x 8 float4 (32 channels) = 1024 float4, 4 per thread ----
    const bool vec_out = (p.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int pl = idx >> 3, cq = idx & 7;             // pixel in tile, channel quad
        const long long px = px0 + pl;
        const int c = c0 + cq * 4;
        if (px >= p.HW || c >= p.C) continue;
        float* d = dst + px * p.C + c;
        const float v0 = tile[cq * 4 + 0][pl], v1 = tile[cq * 4 + 1][pl], v2 = tile[cq * 4 + 2][pl], v3 = tile[cq * 4 + 3][pl];
        if (vec_out && c + 3 < p.C) *reinterpret_cast<float4*>(d) = make_float4(v0, v1, v2, v3);
        else {
            d[0] = v0;
            if (c + 1 < p.C) d[1] = v1;
            if (c + 2 < p.C) d[2] = v2;
            if (c + 3 < p.C) d[3] = v3;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Backward of the fused lift (autograd of utils.py:991 `volume[b,:,valid] = features[b,:,py,px]`,
// i.e. index_put_(accumulate=True) into the feature maps, summed over what encode() accumulates):
//   grad_features_t[b, py, px, :] += grad_volume[b, v, :]   for every frame t that sees voxel v.
// Same brick / culling / projection code as the forward, so exactly the forward's voxel-frame
// pairs receive gradient.  Many voxels along a ray hit the same pixel: the adds are 128-bit vector
// reductions (red.global.add.v4.f32) into channels-last gradient maps; their order is not
// deterministic (as in the reference's CUDA index_put_).
// ---------------------------------------------------------------------------------------
struct LiftBwdKP {
    float* gfeat[GNB_MAX_FRAMES];        // per frame, this scene, NHWC, zero-initialised by the caller
    float P[GNB_MAX_FRAMES][12];
    int T, H, W, C;
    int nx, ny, nz;
    float vs, ox, oy, oz;
    const float* gvol;
    long long stride_v, stride_c;
    const int* count;                    // mean mode: divide the incoming gradient by count
    int mean;
    int x_begin, x_end;
};
static_assert(sizeof(LiftBwdKP) <= 4096, "kernel parameter block must stay below 4 KB");

template <int G, int VEC>
__global__ void __launch_bounds__(256) lift_bwd_kernel(const __grid_constant__ LiftBwdKP p) {
    constexpr int NVW = (256 / G) < 16 ? (256 / G) : 16;
    constexpr int ITER = NVW * G / 32, VPI = 32 / G, NVB = 8 * NVW;
    constexpr int BZ = NVB >= 64 ? 16 : 8, BY = NVB >= 256 ? 4 : 2, BX = NVB / (BZ * BY);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbz = (p.nz + BZ - 1) / BZ, nby = (p.ny + BY - 1) / BY;
    const int bz = blockIdx.x % nbz, by = (blockIdx.x / nbz) % nby, bx = blockIdx.x / (nbz * nby);
    const int x0 = p.x_begin + bx * BX, y0 = by * BY, z0 = bz * BZ;
    __shared__ unsigned long long s_vis;
    if (threadIdx.x == 0) s_vis = 0ull;
    __syncthreads();
    {
        const int x1 = min(x0 + BX, p.x_end) - 1, y1 = min(y0 + BY, p.ny) - 1, z1 = min(z0 + BZ, p.nz) - 1;
        for (int base = 0; base < p.T * 8; base += 256) {
            const int task = base + threadIdx.x;
            const bool active = task < p.T * 8;
            const int t = active ? (task >> 3) : 0, corner = task & 7;
            const float cwx = __fadd_rn(__fmul_rn((float)((corner & 1) ? x1 : x0), p.vs), p.ox);
            const float cwy = __fadd_rn(__fmul_rn((float)((corner & 2) ? y1 : y0), p.vs), p.oy);
            const float cwz = __fadd_rn(__fmul_rn((float)((corner & 4) ? z1 : z0), p.vs), p.oz);
            const float* P = p.P[t];
            const float cx = fmaf(P[2], cwz, fmaf(P[1], cwy, P[0] * cwx)) + P[3];
            const float cy = fmaf(P[6], cwz, fmaf(P[5], cwy, P[4] * cwx)) + P[7];
            const float cz = fmaf(P[10], cwz, fmaf(P[9], cwy, P[8] * cwx)) + P[11];
            const bool front = cz > 1e-3f;
            const float u = cx / cz, v = cy / cz;
            const unsigned sh = lane & ~7u;
            const bool c0 = ((__ballot_sync(FULL, cz < -1e-3f) >> sh) & 0xffu) == 0xffu;
            const bool c1 = ((__ballot_sync(FULL, front && u < -2.0f) >> sh) & 0xffu) == 0xffu;
            const bool c2 = ((__ballot_sync(FULL, front && u > (float)p.W + 1.0f) >> sh) & 0xffu) == 0xffu;
            const bool c3 = ((__ballot_sync(FULL, front && v < -2.0f) >> sh) & 0xffu) == 0xffu;
            const bool c4 = ((__ballot_sync(FULL, front && v > (float)p.H + 1.0f) >> sh) & 0xffu) == 0xffu;
            if (active && corner == 0 && !(c0 || c1 || c2 || c3 || c4)) atomicOr(&s_vis, 1ull << t);
        }
    }
    __syncthreads();
    unsigned long long vis = s_vis;
    if (vis == 0ull) return;
    const int sub = lane % G;
    const int c0 = (blockIdx.y * G + sub) * VEC;
    const bool c_ok = c0 < p.C;
    const int iv = warp * NVW + lane;
    const int vx = x0 + iv / (BZ * BY), vy = y0 + (iv / BZ) % BY, vz = z0 + iv % BZ;
    const bool own = (lane < NVW) && vx < p.x_end && vy < p.ny && vz < p.nz;
    const int v_own = own ? (vx * p.ny + vy) * p.nz + vz : -1;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    if (own) {
        wx = __fadd_rn(__fmul_rn((float)vx, p.vs), p.ox);
        wy = __fadd_rn(__fmul_rn((float)vy, p.vs), p.oy);
        wz = __fadd_rn(__fmul_rn((float)vz, p.vs), p.oz);
    }
    float g[ITER][VEC];
#pragma unroll
    for (int j = 0; j < ITER; ++j) {
        const int v = __shfl_sync(FULL, v_own, j * VPI + lane / G);
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[j][k] = 0.0f;
        if (v >= 0 && c_ok) {
            float sc = 1.0f;
            if (p.mean) { const int n = p.count[v]; sc = 1.0f / (float)(n > 0 ? n : 1); }
#pragma unroll
            for (int k = 0; k < VEC; ++k) g[j][k] = sc * __ldg(p.gvol + (long long)v * p.stride_v + (c0 + k) * p.stride_c);
        }
    }
    while (vis) {
        const int t = __ffsll((long long)vis) - 1;
        vis &= vis - 1;
        int off = -1;
        if (own) off = project_voxel(p.P[t], wx, wy, wz, p.H, p.W);
        if (__ballot_sync(FULL, off >= 0) == 0u) continue;
        float* __restrict__ f = p.gfeat[t];
#pragma unroll
        for (int j = 0; j < ITER; ++j) {
            const int o = __shfl_sync(FULL, off, j * VPI + lane / G);
            if (o >= 0 && c_ok) {
                float* dst = f + (long long)o * p.C + c0;
                if constexpr (VEC == 4) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(g[j][0], g[j][1], g[j][2], g[j][3]));
                else atomicAdd(dst, g[j][0]);
            }
        }
    }
}

template <int G, int VEC>
static int launch_lift_bwd(const LiftBwdKP& kp, cudaStream_t st) {
    constexpr int NVW = (256 / G) < 16 ? (256 / G) : 16;
    constexpr int NVB = 8 * NVW, BZ = NVB >= 64 ? 16 : 8, BY = NVB >= 256 ? 4 : 2, BX = NVB / (BZ * BY);
    long long bricks = (long long)ceil_div(kp.x_end - kp.x_begin, BX) * ceil_div(kp.ny, BY) * ceil_div(kp.nz, BZ);
    dim3 grid((unsigned)bricks, (unsigned)ceil_div(kp.C, G * VEC));
    lift_bwd_kernel<G, VEC><<<grid, 256, 0, st>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

static int dispatch_lift_bwd(const LiftBwdKP& kp, cudaStream_t st) {
    const int C = kp.C;
    if (C % 4 == 0) {
        int g = C / 4;
        if (g >= 32) return launch_lift_bwd<32, 4>(kp, st);
        if (g > 8) return launch_lift_bwd<16, 4>(kp, st);
        if (g > 4) return launch_lift_bwd<8, 4>(kp, st);
        if (g > 2) return launch_lift_bwd<4, 4>(kp, st);
        return launch_lift_bwd<2, 4>(kp, st);
    }
    if (C >= 32) return launch_lift_bwd<32, 1>(kp, st);
    if (C > 8) return launch_lift_bwd<16, 1>(kp, st);
    if (C > 4) return launch_lift_bwd<8, 1>(kp, st);
    if (C > 2) return launch_lift_bwd<4, 1>(kp, st);
    return launch_lift_bwd<2, 1>(kp, st);
}

// NHWC -> NCHW (gradient maps back to the reference layout)
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const __grid_constant__ TransposeKP p) {
    __shared__ float tile[32][33];
    const int t = blockIdx.z / p.B, b = blockIdx.z % p.B;
    float* __restrict__ dst = const_cast<float*>(p.src[t]) + (long long)b * p.C * p.HW;      // (C,HW) output
    const float* __restrict__ src = p.dst + ((long long)t * p.B + b) * p.HW * p.C;           // (HW,C) input
    const long long px0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        long long px = px0 + r;
        int c = c0 + tx;
        tile[r][tx] = (c < p.C && px < p.HW) ? __ldg(src + px * p.C + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int c = c0 + r;
        long long px = px0 + tx;
        if (c < p.C && px < p.HW) dst[(long long)c * p.HW + px] = tile[tx][r];
    }
}

__global__ void project_indices_kernel(int nx, int ny, int nz, float vs, float ox, float oy, float oz,
                                       const __grid_constant__ LiftKP p, int* px, int* py, unsigned char* valid) {
    long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.V) return;
    float wx, wy, wz, fx, fy;
    voxel_world(v, ny, nz, vs, ox, oy, oz, wx, wy, wz);
    int off = project_voxel(p.P[0], wx, wy, wz, p.H, p.W, &fx, &fy);
    // (long) of a non-finite / out-of-range float is INT64_MIN on the CPU; report INT32_MIN
    px[v] = (fabsf(fx) < 2147483520.0f) ? (int)fx : INT_MIN;
    py[v] = (fabsf(fy) < 2147483520.0f) ? (int)fy : INT_MIN;
    valid[v] = (unsigned char)(off >= 0);
}

template <int G, int VEC, int NVW>
static int launch_lift_nvw(const LiftKP& kp, int C, cudaStream_t st) {
    constexpr int NVB = 8 * NVW, BZ = NVB >= 64 ? 16 : 8, BY = NVB >= 256 ? 4 : 2, BX = NVB / (BZ * BY);
    long long bricks = (long long)ceil_div(kp.x_end - kp.x_begin, BX) * ceil_div(kp.ny, BY) * ceil_div(kp.nz, BZ);
    dim3 grid((unsigned)bricks, (unsigned)ceil_div(C, G * VEC));
    if (kp.stride_c == 1)
        lift_kernel<G, VEC, true, NVW><<<grid, 256, 0, st>>>(kp);
    else
        lift_kernel<G, VEC, false, NVW><<<grid, 256, 0, st>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

// voxels per warp: fewer voxels per warp = fewer accumulator registers, more resident warps and a
// finer culling brick; GNB_LIFT_NVW overrides the default for tuning
template <int G, int VEC>
static int launch_lift(const LiftKP& kp, int C, cudaStream_t st) {
    constexpr int NVMAX = (256 / G) < 32 ? (256 / G) : 32;
    constexpr int NVMIN = (32 / G) > 4 ? (32 / G) : 4;      // bricks are at least 2 x 2 x 8
    int nvw = NVMAX >= 16 ? 16 : NVMAX;
    if (const int e = opt(OPT_LIFT_NVW)) nvw = e;
    if (nvw <= NVMIN) return launch_lift_nvw<G, VEC, NVMIN>(kp, C, st);
    if constexpr (NVMAX >= 2 * NVMIN) { if (nvw <= 2 * NVMIN || NVMAX == 2 * NVMIN) return launch_lift_nvw<G, VEC, 2 * NVMIN>(kp, C, st); }
    if constexpr (NVMAX >= 4 * NVMIN) { if (nvw <= 4 * NVMIN || NVMAX == 4 * NVMIN) return launch_lift_nvw<G, VEC, 4 * NVMIN>(kp, C, st); }
    return launch_lift_nvw<G, VEC, NVMAX>(kp, C, st);
}

static int dispatch_lift(const LiftKP& kp, cudaStream_t st) {
    const int C = kp.C;
    if (C % 4 == 0) {
        int g = C / 4;
        if (g >= 32) return launch_lift<32, 4>(kp, C, st);
        if (g > 8) return launch_lift<16, 4>(kp, C, st);
        if (g > 4) return launch_lift<8, 4>(kp, C, st);
        if (g > 2) return launch_lift<4, 4>(kp, C, st);
        if (g > 1) return launch_lift<2, 4>(kp, C, st);
        return launch_lift<1, 4>(kp, C, st);
    }
    if (C >= 32) return launch_lift<32, 1>(kp, C, st);
    if (C > 8) return launch_lift<16, 1>(kp, C, st);
    if (C > 4) return launch_lift<8, 1>(kp, C, st);
    if (C > 2) return launch_lift<4, 1>(kp, C, st);
    if (C > 1) return launch_lift<2, 1>(kp, C, st);
    return launch_lift<1, 1>(kp, C, st);
}

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_nchw_to_nhwc(const float* const* h_src, int n_frames, float* dst, int B, int C, int H, int W,
                                void* stream) {
    GNB_CHECK_ARG(h_src && dst, "gnb_nchw_to_nhwc: null pointer");
    GNB_CHECK_ARG(n_frames >= 1 && n_frames <= GNB_MAX_FRAMES, "gnb_nchw_to_nhwc: n_frames %d not in [1,%d]", n_frames,
                  GNB_MAX_FRAMES);
    GNB_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1, "gnb_nchw_to_nhwc: bad shape");
    GNB_CHECK_ARG((long long)n_frames * B <= 65535, "gnb_nchw_to_nhwc: T*B too large");
    TransposeKP kp;
    for (int t = 0; t < n_frames; ++t) kp.src[t] = h_src[t];
    kp.dst = dst;
    kp.T = n_frames, kp.B = B, kp.C = C, kp.HW = (long long)H * W;
    dim3 grid((unsigned)ceil_div(kp.HW, TP_PX), (unsigned)ceil_div(C, 32), (unsigned)(n_frames * B));
    nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_backproject_frames(const GnbLiftParams* p, void* stream) {
    GNB_CHECK_ARG(p, "gnb_backproject_frames: null params");
    GNB_CHECK_ARG(p->nx > 0 && p->ny > 0 && p->nz > 0, "gnb_backproject_frames: bad voxel grid %dx%dx%d", p->nx, p->ny, p->nz);
    GNB_CHECK_ARG(p->n_frames >= 1 && p->n_frames <= GNB_MAX_FRAMES, "gnb_backproject_frames: n_frames %d not in [1,%d]",
                  p->n_frames, GNB_MAX_FRAMES);
    GNB_CHECK_ARG(p->batch >= 1 && p->C >= 1 && p->H >= 1 && p->W >= 1, "gnb_backproject_frames: bad feature shape");
    GNB_CHECK_ARG((long long)p->H * p->W < INT_MAX, "gnb_backproject_frames: image too large");
    GNB_CHECK_ARG((long long)p->nx * p->ny * p->nz < INT_MAX, "gnb_backproject_frames: voxel grid too large");
    GNB_CHECK_ARG(p->volume && p->h_projection, "gnb_backproject_frames: null volume/projection");
    GNB_CHECK_ARG(p->x_begin >= 0 && p->x_end <= p->nx && (p->x_end == 0 || p->x_end >= p->x_begin), "gnb_backproject_frames: bad x range");
    GNB_CHECK_ARG(p->feat_layout == GNB_LAYOUT_NHWC || p->feat_layout == GNB_LAYOUT_NCHW, "gnb_backproject_frames: bad layout");
    for (int t = 0; t < p->n_frames; ++t) GNB_CHECK_ARG(p->features[t], "gnb_backproject_frames: features[%d] is null", t);
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)p->nx * p->ny * p->nz;
    const long long img = (long long)p->H * p->W * p->C;

    const float* base[GNB_MAX_FRAMES];
    long long frame_stride_b = img;       // elements between scenes inside one frame tensor
    if (p->feat_layout == GNB_LAYOUT_NCHW) {
        GNB_CHECK_ARG(p->scratch, "gnb_backproject_frames: NCHW features need a scratch buffer of T*B*H*W*C floats");
        int rc = gnb_nchw_to_nhwc(p->features, p->n_frames, p->scratch, p->batch, p->C, p->H, p->W, stream);
        if (rc) return rc;
        for (int t = 0; t < p->n_frames; ++t) base[t] = p->scratch + (long long)t * p->batch * img;
    } else {
        for (int t = 0; t < p->n_frames; ++t) base[t] = p->features[t];
    }
    for (int b = 0; b < p->batch; ++b) {
        LiftKP kp;
        for (int t = 0; t < p->n_frames; ++t) {
            kp.feat[t] = base[t] + (long long)b * frame_stride_b;
            for (int k = 0; k < 12; ++k) kp.P[t][k] = p->h_projection[((long long)b * p->n_frames + t) * 12 + k];
        }
        kp.T = p->n_frames, kp.H = p->H, kp.W = p->W, kp.C = p->C;
        kp.nx = p->nx, kp.ny = p->ny, kp.nz = p->nz, kp.V = V;
        kp.vs = p->voxel_size, kp.ox = p->origin[0], kp.oy = p->origin[1], kp.oz = p->origin[2];
        kp.volume = p->volume + (long long)b * p->vol_stride_b;
        kp.stride_v = p->vol_stride_v, kp.stride_c = p->vol_stride_c;
        kp.count = p->count ? p->count + (long long)b * V : nullptr;
        kp.valid = p->valid ? p->valid + (long long)b * V : nullptr;
        kp.accumulate = p->accumulate, kp.mean = p->mean;
        kp.x_begin = p->x_begin, kp.x_end = (p->x_end > 0) ? p->x_end : p->nx;
        if (kp.x_end <= kp.x_begin) continue;
        int rc = dispatch_lift(kp, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int gnb_project_indices(int nx, int ny, int nz, float voxel_size, const float* h_origin3,
                                   const float* h_projection12, int H, int W, int32_t* px, int32_t* py, uint8_t* valid,
                                   void* stream) {
    GNB_CHECK_ARG(nx > 0 && ny > 0 && nz > 0 && H > 0 && W > 0, "gnb_project_indices: bad shape");
    GNB_CHECK_ARG(h_origin3 && h_projection12 && px && py && valid, "gnb_project_indices: null pointer");
    LiftKP kp;
    for (int k = 0; k < 12; ++k) kp.P[0][k] = h_projection12[k];
    kp.T = 1, kp.H = H, kp.W = W, kp.C = 1;
    kp.V = (long long)nx * ny * nz;
    project_indices_kernel<<<ceil_div(kp.V, 256), 256, 0, (cudaStream_t)stream>>>(
        nx, ny, nz, voxel_size, h_origin3[0], h_origin3[1], h_origin3[2], kp, px, py, valid);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_backproject_frames_bwd(const GnbLiftParams* p, const float* grad_volume, float* const* h_grad_features,
                                          void* stream) {
    GNB_CHECK_ARG(p && grad_volume && h_grad_features, "gnb_backproject_frames_bwd: null argument");
    GNB_CHECK_ARG(p->nx > 0 && p->ny > 0 && p->nz > 0 && p->batch >= 1 && p->C >= 1 && p->H >= 1 && p->W >= 1,
                  "gnb_backproject_frames_bwd: bad shape");
    GNB_CHECK_ARG(p->n_frames >= 1 && p->n_frames <= GNB_MAX_FRAMES, "gnb_backproject_frames_bwd: n_frames %d not in [1,%d]",
                  p->n_frames, GNB_MAX_FRAMES);
    GNB_CHECK_ARG(p->h_projection, "gnb_backproject_frames_bwd: null projection");
    GNB_CHECK_ARG(!p->mean || p->count, "gnb_backproject_frames_bwd: mean mode needs the forward's count");
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)p->nx * p->ny * p->nz;
    const long long img = (long long)p->H * p->W * p->C;
    // gradient maps are accumulated channels-last; for the reference layout they are built in `scratch`
    float* base[GNB_MAX_FRAMES];
    const bool nchw = p->feat_layout == GNB_LAYOUT_NCHW;
    GNB_CHECK_ARG(!nchw || p->scratch, "gnb_backproject_frames_bwd: NCHW gradients need a scratch buffer of T*B*H*W*C floats");
    for (int t = 0; t < p->n_frames; ++t) {
        GNB_CHECK_ARG(h_grad_features[t], "gnb_backproject_frames_bwd: grad_features[%d] is null", t);
        base[t] = nchw ? p->scratch + (long long)t * p->batch * img : h_grad_features[t];
        GNB_CUDA(cudaMemsetAsync(base[t], 0, sizeof(float) * p->batch * img, st));
    }
    for (int b = 0; b < p->batch; ++b) {
        LiftBwdKP kp;
        for (int t = 0; t < p->n_frames; ++t) {
            kp.gfeat[t] = base[t] + (long long)b * img;
            for (int k = 0; k < 12; ++k) kp.P[t][k] = p->h_projection[((long long)b * p->n_frames + t) * 12 + k];
        }
        kp.T = p->n_frames, kp.H = p->H, kp.W = p->W, kp.C = p->C;
        kp.nx = p->nx, kp.ny = p->ny, kp.nz = p->nz;
        kp.vs = p->voxel_size, kp.ox = p->origin[0], kp.oy = p->origin[1], kp.oz = p->origin[2];
        kp.gvol = grad_volume + (long long)b * p->vol_stride_b;
        kp.stride_v = p->vol_stride_v, kp.stride_c = p->vol_stride_c;
        kp.count = p->count ? p->count + (long long)b * V : nullptr;
        kp.mean = p->mean;
        kp.x_begin = p->x_begin, kp.x_end = (p->x_end > 0) ? p->x_end : p->nx;
        if (kp.x_end <= kp.x_begin) continue;
        int rc = dispatch_lift_bwd(kp, st);
        if (rc) return rc;
    }
    if (nchw) {
        TransposeKP kp;
        for (int t = 0; t < p->n_frames; ++t) kp.src[t] = h_grad_features[t];     // destination of the inverse transpose
        kp.dst = p->scratch;
        kp.T = p->n_frames, kp.B = p->batch, kp.C = p->C, kp.HW = (long long)p->H * p->W;
        dim3 grid((unsigned)ceil_div(kp.HW, 32), (unsigned)ceil_div(p->C, 32), (unsigned)(p->n_frames * p->batch));
        nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(kp);
        GNB_LAUNCH_CHECK();
    }
    return 0;
}

// TSDF fusion of posed depth maps (sm_100a).  SURVEY 8(f) rank 3.
//
// Replaces TSDFFusion.integrate (reference src/data/tsdf.py:369-418) for a whole batch of frames and the
// normalisation of TSDFFusion.get_tsdf (tsdf.py:420-440).  The reference integrates one frame per call with
// ~25 full-volume PyTorch ops (boolean masks, masked gathers / scatters); here one thread owns one voxel, keeps
// its running (tsdf, weight, colour, label) in registers over all frames of the launch -- visited in frame
// order, so the result is bit-identical to T sequential integrate() calls -- and reads / writes the volumes once.
//
// Projection arithmetic is the lift's (csrc/lift.cu project_voxel): world = fl(i)*voxel_size + origin (two rounded
// ops, tsdf.py:344), camera = FMA chain of the 3x4 . 4-vector product (what ATen's CPU matmul computes for K = 4,
// pinned in tests/test_oracle_pinning.py), pixel = round-half-even(cx/cz), true fp32 divisions.
//
// A block owns a 4 x 4 x 16 brick of voxels and first drops the frames that cannot see the brick (all 8 corner
// voxels behind the camera or beyond the same image border; exact, because voxel centres are convex combinations of
// the corners and x/z is linear-fractional) -- the same culling as the lift.  With a scratch buffer it also culls by
// DEPTH: a first kernel reduces every 16x16 pixel tile of every depth map to (min, max); per brick and frame the tiles
// under the brick's pixel bounding box give a depth range, and
//   * min corner depth - max measured depth >= trunc  ->  every voxel is beyond the truncation band: frame skipped;
//   * the box has no zero-depth pixel, lies inside the image, and max corner depth - min measured depth <= -trunc
//       ->  every voxel is free space (clamped distance -1): `if (weight == 0) tsdf = -1` without projecting;
// both exact (the per-voxel tests could only come out that way; small margins cover the fp32 rounding).
#include "common.cuh"

namespace gnb {

constexpr int FUSE_BX = 4, FUSE_BY = 4, FUSE_BZ = 16;

struct FuseKP {
    int nx, ny, nz;
    float vs, ox, oy, oz;
    float trunc;
    int T, H, W;
    float P[GNB_MAX_FRAMES][12];
    const float* depth;      // (T,H,W)
    const float* color;      // (T,3,H,W) or null
    const int* label;        // (T,H,W) or null
    float* tsdf;             // (V)
    float* weight;           // (V)
    float* color_vol;        // (3,V) or null
    int* label_vol;          // (V) or null
    const float2* tiles;     // (T, tiles_y, tiles_x) (min, max) of every 16x16 depth tile, or null
    int tiles_x, tiles_y;
};
static_assert(sizeof(FuseKP) <= 4096, "kernel parameter block must stay below 4 KB");

// EARLY: reject before the two IEEE divisions when the voxel is behind the camera or the exact quotient lies more than
// a quarter pixel outside the range that can round to a valid pixel (the rounded quotient is then outside as well)
template <bool EARLY>
__device__ __forceinline__ bool fuse_project(const float* __restrict__ P, float wx, float wy, float wz, float& fx, float& fy, float& cz,
                                             int H = 0, int W = 0) {
    const float cx = __fadd_rn(__fmaf_rn(P[2], wz, __fmaf_rn(P[1], wy, __fmul_rn(P[0], wx))), P[3]);
    const float cy = __fadd_rn(__fmaf_rn(P[6], wz, __fmaf_rn(P[5], wy, __fmul_rn(P[4], wx))), P[7]);
    cz = __fadd_rn(__fmaf_rn(P[10], wz, __fmaf_rn(P[9], wy, __fmul_rn(P[8], wx))), P[11]);
    if constexpr (EARLY) {
        if (!(cz > 0.0f)) return false;
        if (cx < -0.75f * cz || cx > ((float)W - 0.25f) * cz || cy < -0.75f * cz || cy > ((float)H - 0.25f) * cz) return false;
    }
    fx = rintf(__fdiv_rn(cx, cz));
    fy = rintf(__fdiv_rn(cy, cz));
    return true;
}

constexpr int FUSE_TILE = 16;
constexpr int F_SKIP = 0, F_FULL = 1, F_FREE = 2;

// (min, max) of every 16x16 tile of every depth map; zero-depth ("no measurement") pixels take part as zeros.
// A block reduces a 64x16 pixel strip (four tiles): 16 threads x float4 per image row.
__global__ void __launch_bounds__(256) depth_tiles_kernel(const float* __restrict__ depth, int H, int W, int tiles_x, int tiles_y,
                                                          float2* __restrict__ tiles) {
    __shared__ float s_min[16][4], s_max[16][4];
    const int strips_x = (tiles_x + 3) / 4;
    const int sx = blockIdx.x % strips_x, ty = blockIdx.x / strips_x, f = blockIdx.y;
    const int r = threadIdx.x >> 4, q = threadIdx.x & 15;             // image row inside the tile, float4 inside the strip
    const int py = ty * FUSE_TILE + r, px = sx * 64 + q * 4;
    float lo = 3.0e38f, hi = -3.0e38f;
    if (py < H) {
        const float* row = depth + ((long long)f * H + py) * W;
        if ((W & 3) == 0 && px + 3 < W) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + px));
            lo = fminf(fminf(v.x, v.y), fminf(v.z, v.w)), hi = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        } else {
            for (int k = 0; k < 4; ++k)
                if (px + k < W) { const float v = __ldg(row + px + k); lo = fminf(lo, v), hi = fmaxf(hi, v); }
        }
    }
    // the 4 threads of one tile in this row
    lo = fminf(lo, __shfl_xor_sync(FULL, lo, 1)), hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, 1));
    lo = fminf(lo, __shfl_xor_sync(FULL, lo, 2)), hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, 2));
    if ((q & 3) == 0) { s_min[r][q >> 2] = lo; s_max[r][q >> 2] = hi; }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int tx = sx * 4 + threadIdx.x;
        if (tx < tiles_x) {
            lo = s_min[0][threadIdx.x], hi = s_max[0][threadIdx.x];
            for (int i = 1; i < 16; ++i) { lo = fminf(lo, s_min[i][threadIdx.x]); hi = fmaxf(hi, s_max[i][threadIdx.x]); }
            tiles[((long long)f * tiles_y + ty) * tiles_x + tx] = make_float2(lo, hi);
        }
    }
}

// (8 resident blocks per SM = 32 registers: the kernel is issue / latency bound, full occupancy is worth ~11 %)
__global__ void __launch_bounds__(256, 8) fuse_kernel(const __grid_constant__ FuseKP p) {
    __shared__ unsigned char keep[GNB_MAX_FRAMES];
    __shared__ int kept[GNB_MAX_FRAMES];
    __shared__ int n_kept;
    const int gbz = (p.nz + FUSE_BZ - 1) / FUSE_BZ, gby = (p.ny + FUSE_BY - 1) / FUSE_BY;
    const int bz = blockIdx.x % gbz, by = (blockIdx.x / gbz) % gby, bx = blockIdx.x / (gbz * gby);
    const int x0 = bx * FUSE_BX, y0 = by * FUSE_BY, z0 = bz * FUSE_BZ;

    // ---- frame culling: 8 lanes per frame, one brick corner each (all 256 threads busy for 32 frames) ----
    {
        const int x1 = min(x0 + FUSE_BX, p.nx) - 1, y1 = min(y0 + FUSE_BY, p.ny) - 1, z1 = min(z0 + FUSE_BZ, p.nz) - 1;
        const int lane = threadIdx.x & 31;
        const unsigned sh = lane & ~7u;
        for (int base = 0; base < p.T * 8; base += 256) {
            const int task = base + threadIdx.x;
            const bool active = task < p.T * 8;                 // whole 8-lane groups are active or not
            const int t = active ? (task >> 3) : 0, c = task & 7;
            const float* P = p.P[t];
            const float wx = __fadd_rn(__fmul_rn((float)((c & 1) ? x1 : x0), p.vs), p.ox);
            const float wy = __fadd_rn(__fmul_rn((float)((c & 2) ? y1 : y0), p.vs), p.oy);
            const float wz = __fadd_rn(__fmul_rn((float)((c & 4) ? z1 : z0), p.vs), p.oz);
            float fx, fy, cz;
            fuse_project<false>(P, wx, wy, wz, fx, fy, cz);
            const bool front = cz > 0.0f;
            auto all8 = [&](bool v) { return ((__ballot_sync(FULL, v) >> sh) & 0xffu) == 0xffu; };
            // one pixel of slack covers the rounding of the per-voxel arithmetic
            const bool behind = all8(!front), left = all8(front && fx < -1.0f), right = all8(front && fx > (float)p.W);
            const bool above = all8(front && fy < -1.0f), below = all8(front && fy > (float)p.H);
            const bool all_front = all8(cz > 1e-3f);
            float cz_min = cz, cz_max = cz, fx_min = fx, fx_max = fx, fy_min = fy, fy_max = fy;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                cz_min = fminf(cz_min, __shfl_xor_sync(FULL, cz_min, o)), cz_max = fmaxf(cz_max, __shfl_xor_sync(FULL, cz_max, o));
                fx_min = fminf(fx_min, __shfl_xor_sync(FULL, fx_min, o)), fx_max = fmaxf(fx_max, __shfl_xor_sync(FULL, fx_max, o));
                fy_min = fminf(fy_min, __shfl_xor_sync(FULL, fy_min, o)), fy_max = fmaxf(fy_max, __shfl_xor_sync(FULL, fy_max, o));
            }
            if (!active || c != 0) continue;
            int cls = (behind || left || right || above || below) ? F_SKIP : F_FULL;
            if (cls == F_FULL && p.tiles && all_front && fx_min > -1.0e6f && fx_max < 1.0e6f && fy_min > -1.0e6f && fy_max < 1.0e6f) {
                // pixel bounding box of the brick (every voxel centre projects inside the corners' box; +-1 px for rounding),
                // clipped to the image: voxels that project outside the image are not updated anyway
                const bool inside = fx_min >= 1.0f && fy_min >= 1.0f && fx_max <= (float)(p.W - 2) && fy_max <= (float)(p.H - 2);
                const int bx0 = max(0, (int)fx_min - 1), bx1 = min(p.W - 1, (int)fx_max + 1);
                const int by0 = max(0, (int)fy_min - 1), by1 = min(p.H - 1, (int)fy_max + 1);
                const int tx0 = bx0 / FUSE_TILE, tx1 = bx1 / FUSE_TILE, ty0 = by0 / FUSE_TILE, ty1 = by1 / FUSE_TILE;
                if (bx0 <= bx1 && by0 <= by1 && (tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= 64) {
                    float dmin = 3.0e38f, dmax = -3.0e38f;
                    const float2* tl = p.tiles + (long long)t * p.tiles_y * p.tiles_x;
                    for (int ty = ty0; ty <= ty1; ++ty)
                        for (int tx = tx0; tx <= tx1; ++tx) {
                            const float2 mm = __ldg(tl + ty * p.tiles_x + tx);
                            dmin = fminf(dmin, mm.x), dmax = fmaxf(dmax, mm.y);
                        }
                    const float band = p.trunc * 1.0001f + 1.0e-5f;
                    if (!(dmax > 0.0f)) cls = F_SKIP;                                  // no measurement under the brick
                    else if (cz_min - dmax >= band) cls = F_SKIP;                       // dist >= 1 for every voxel (tsdf.py:398)
                    else if (inside && dmin > 0.0f && cz_max - dmin <= -band) cls = F_FREE;    // dist clamps to -1 for every voxel
                }
            }
            keep[t] = (unsigned char)cls;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int f = 0; f < p.T; ++f)
            if (keep[f] != F_SKIP) kept[n++] = f | ((int)keep[f] << 8);      // ascending: the reference's frame order
        n_kept = n;
    }
    __syncthreads();
    const int nk = n_kept;
    if (nk == 0) return;                                 // nothing to update: the volumes keep their values

    const int lz = threadIdx.x % FUSE_BZ, ly = (threadIdx.x / FUSE_BZ) % FUSE_BY, lx = threadIdx.x / (FUSE_BZ * FUSE_BY);
    const int ix = x0 + lx, iy = y0 + ly, iz = z0 + lz;
    if (ix >= p.nx || iy >= p.ny || iz >= p.nz) return;
    const long long V = (long long)p.nx * p.ny * p.nz;
    const long long v = ((long long)ix * p.ny + iy) * p.nz + iz;
    const float wx = __fadd_rn(__fmul_rn((float)ix, p.vs), p.ox);
    const float wy = __fadd_rn(__fmul_rn((float)iy, p.vs), p.oy);
    const float wz = __fadd_rn(__fmul_rn((float)iz, p.vs), p.oz);

    float tsdf = p.tsdf[v], weight = p.weight[v];
    float cr = 0.f, cg = 0.f, cb = 0.f;
    int lab = 0;
    const bool has_color = p.color_vol != nullptr, has_label = p.label_vol != nullptr;
    if (has_color) { cr = p.color_vol[v], cg = p.color_vol[V + v], cb = p.color_vol[2 * V + v]; }
    if (has_label) lab = p.label_vol[v];
    bool touched = false;
    const long long HW = (long long)p.H * p.W;

    for (int k = 0; k < nk; ++k) {
        const int f = kept[k] & 0xff;
        if ((kept[k] >> 8) == F_FREE) {
            // free space for the whole brick: valid pixel, depth > 0, distance clamped to -1 -> only the first-observation
            // copy of tsdf.py:405 can happen
            if (weight == 0.0f) { tsdf = -1.0f; touched = true; }
            continue;
        }
        float fx, fy, cz;
        if (!fuse_project<true>(p.P[f], wx, wy, wz, fx, fy, cz, p.H, p.W)) continue;
        // float comparisons == the reference's int64 comparisons for every finite value; NaN / inf compare false here
        // and convert to INT64_MIN (invalid) there                                        (tsdf.py:387)
        if (!((fx >= 0.0f) && (fy >= 0.0f) && (fx < (float)p.W) && (fy < (float)p.H) && (cz > 0.0f))) continue;
        const long long off = (long long)f * HW + (long long)(int)fy * p.W + (int)fx;
        const float d = __ldg(p.depth + off);
        if (!(d > 0.0f)) continue;                                                       // tsdf.py:391
        float dist = __fdiv_rn(__fsub_rn(cz, d), p.trunc);                               // tsdf.py:394-395
        dist = dist < -1.0f ? -1.0f : dist;                                              // clamp(min=-1)
        if (!(dist < 1.0f)) continue;                                                    // tsdf.py:398
        const bool first = weight == 0.0f;                                               // tsdf.py:404
        const bool near = dist > -1.0f;                                                  // tsdf.py:409
        if (first) tsdf = dist;                              // copied even when it is the clamped -1 (tsdf.py:405)
        else if (near) tsdf = __fadd_rn(tsdf, dist);                                     // tsdf.py:412
        touched = touched || first || near;
        if (near) {
            weight = __fadd_rn(weight, 1.0f);
            if (has_color) {
                const float* c = p.color + (long long)f * 3 * HW + (off - (long long)f * HW);
                cr = __fadd_rn(cr, __ldg(c)), cg = __fadd_rn(cg, __ldg(c + HW)), cb = __fadd_rn(cb, __ldg(c + 2 * HW));
            }
            if (has_label) lab = __ldg(p.label + off);                                    // newest label wins
        }
    }
    if (!touched) return;
    p.tsdf[v] = tsdf;
    p.weight[v] = weight;
    if (has_color) { p.color_vol[v] = cr, p.color_vol[V + v] = cg, p.color_vol[2 * V + v] = cb; }
    if (has_label) p.label_vol[v] = lab;
}

// get_tsdf's normalisation (tsdf.py:426-434): value / weight where weight > 0, else unchanged
__global__ void fuse_finalize_kernel(const float* __restrict__ tsdf, const float* __restrict__ weight, const float* __restrict__ color,
                                     long long V, float* __restrict__ tsdf_out, float* __restrict__ color_out) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float w = weight[v];
    const bool seen = w > 0.0f;
    if (tsdf_out) tsdf_out[v] = seen ? __fdiv_rn(tsdf[v], w) : tsdf[v];
    if (color_out) {
#pragma unroll
        for (int c = 0; c < 3; ++c) color_out[c * V + v] = seen ? __fdiv_rn(color[c * V + v], w) : color[c * V + v];
    }
}

}  // namespace gnb

using namespace gnb;

extern "C" int64_t gnb_tsdf_fusion_scratch_bytes(int n_frames, int H, int W) {
    return (int64_t)n_frames * ceil_div(H, FUSE_TILE) * ceil_div(W, FUSE_TILE) * (int64_t)sizeof(float2);
}

extern "C" int gnb_tsdf_fusion_integrate(const GnbFusionParams* q, void* stream) {
    GNB_CHECK_ARG(q, "gnb_tsdf_fusion_integrate: null parameters");
    GNB_CHECK_ARG(q->nx > 0 && q->ny > 0 && q->nz > 0 && q->voxel_size > 0.0f, "gnb_tsdf_fusion_integrate: bad grid");
    GNB_CHECK_ARG(q->n_frames >= 0 && q->n_frames <= GNB_MAX_FRAMES, "gnb_tsdf_fusion_integrate: n_frames %d not in [0,%d]",
                  q->n_frames, GNB_MAX_FRAMES);
    if (q->n_frames == 0) return 0;
    GNB_CHECK_ARG(q->H > 0 && q->W > 0 && q->h_projection && q->depth, "gnb_tsdf_fusion_integrate: bad frames");
    GNB_CHECK_ARG(q->tsdf_vol && q->weight_vol, "gnb_tsdf_fusion_integrate: null volumes");
    GNB_CHECK_ARG(!q->color_vol == !q->color, "gnb_tsdf_fusion_integrate: colour frames and colour volume go together");
    GNB_CHECK_ARG(!q->label_vol == !q->label, "gnb_tsdf_fusion_integrate: label frames and label volume go together");
    GNB_CHECK_ARG(q->trunc_margin > 0.0f, "gnb_tsdf_fusion_integrate: bad truncation margin");
    FuseKP kp;
    kp.nx = q->nx, kp.ny = q->ny, kp.nz = q->nz;
    kp.vs = q->voxel_size, kp.ox = q->origin[0], kp.oy = q->origin[1], kp.oz = q->origin[2];
    kp.trunc = q->trunc_margin;
    kp.T = q->n_frames, kp.H = q->H, kp.W = q->W;
    for (int f = 0; f < q->n_frames; ++f)
        for (int i = 0; i < 12; ++i) kp.P[f][i] = q->h_projection[f * 12 + i];
    kp.depth = q->depth, kp.color = q->color, kp.label = q->label;
    kp.tsdf = q->tsdf_vol, kp.weight = q->weight_vol, kp.color_vol = q->color_vol, kp.label_vol = q->label_vol;
    const long long bricks = (long long)ceil_div(q->nx, FUSE_BX) * ceil_div(q->ny, FUSE_BY) * ceil_div(q->nz, FUSE_BZ);
    GNB_CHECK_ARG(bricks < (1ll << 31), "gnb_tsdf_fusion_integrate: grid too large");
    kp.tiles = nullptr, kp.tiles_x = ceil_div(q->W, FUSE_TILE), kp.tiles_y = ceil_div(q->H, FUSE_TILE);
    if (q->scratch && q->scratch_bytes >= gnb_tsdf_fusion_scratch_bytes(q->n_frames, q->H, q->W)) {
        kp.tiles = (const float2*)q->scratch;
        depth_tiles_kernel<<<dim3((unsigned)(((kp.tiles_x + 3) / 4) * kp.tiles_y), (unsigned)q->n_frames), 256, 0, (cudaStream_t)stream>>>(
            q->depth, q->H, q->W, kp.tiles_x, kp.tiles_y, (float2*)q->scratch);
        GNB_LAUNCH_CHECK();
    }
    fuse_kernel<<<(unsigned)bricks, 256, 0, (cudaStream_t)stream>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_tsdf_fusion_finalize(const float* tsdf_vol, const float* weight_vol, const float* color_vol, int64_t n_voxels,
                                        float* tsdf_out, float* color_out, void* stream) {
    GNB_CHECK_ARG(n_voxels >= 0, "gnb_tsdf_fusion_finalize: bad size");
    if (n_voxels == 0) return 0;
    GNB_CHECK_ARG(weight_vol && (tsdf_out || color_out), "gnb_tsdf_fusion_finalize: bad arguments");
    GNB_CHECK_ARG(!tsdf_out || tsdf_vol, "gnb_tsdf_fusion_finalize: null tsdf volume");
    GNB_CHECK_ARG(!color_out || color_vol, "gnb_tsdf_fusion_finalize: null colour volume");
    fuse_finalize_kernel<<<ceil_div(n_voxels, 256), 256, 0, (cudaStream_t)stream>>>(tsdf_vol, weight_vol, color_vol, n_voxels, tsdf_out,
                                                                                color_out);
    GNB_LAUNCH_CHECK();
    return 0;
}

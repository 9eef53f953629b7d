// Brick-binned point-query sampler (sm_100a): the stand-alone sampler for MANY queries per voxel.
//
// Replaces the same reference code as sample.cu -- trilinear_interpolation() (src/models/utils.py:999-1042),
// GenNerf.sample_plane_feature() x3 (src/models/model.py:153-161), GenNerf.map_features (model.py:163-204) --
// and produces the same bits as sample_staged_kernel (same setup functions, same accumulation order).
//
// Why: a uniformly random query touches 8 distinct 128 B lines of the channels-last volume, so the staged
// kernel moves ~1 KB of L2->SM sectors per query and is bound by L2 (160 us per Mi queries at C = 32, 0.19 of the
// HBM roofline on algorithmic bytes).  Here the queries are first counting-sorted by the brick of voxels that holds
// their base cell (order inside a bin is irrelevant: queries are independent); then a CTA stages a brick's corner
// voxels ((bx+1)(by+1)(bz+1) x C floats, <= 93 KB) into shared memory with ONE tensor-map TMA copy (a 5-D box of the
// channels-last volume; fallback: a 1-D bulk copy per contiguous z-row) and serves the brick's queries with LDS.128 gathers.  HBM/L2 see every voxel ~1.4x
// (halo) instead of every query 8x.  Bricks with few queries are gathered straight from global memory.
//
//   bin_count_kernel   -> bid[q], count[brick]    (shared-memory histogram per block, one global add per non-empty bin)
//   bin_scan_kernel    -> start[brick]            (exclusive scan, one block)
//   bin_scatter_kernel -> sorted[rank] = (x,y,z,q)(ranges reserved per block and bin, ranks by shared-memory atomics)
//   sample_binned_kernel                          (one persistent CTA per SM: a producer warp claims bricks from an atomic
//                                                  counter and keeps two tile buffers filled; 15 consumer warps gather)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "binsort.cuh"
#include "sample.cuh"

namespace gnb {

constexpr int BIN_WARPS = 15;                // consumer warps of the gather kernel (+ 1 producer warp = 4 per scheduler), one CTA per SM
constexpr int BIN_TAB = 20;                  // words per query in a warp's corner table
constexpr int BIN_TILE_BYTES = 93312;        // one tile buffer: 9*9*9 voxels x 32 channels x 4 B
constexpr int BIN_HIST_MAX = 40960;          // bins a block can hold in shared memory (160 KB)

struct BinKP {
    SampleKP s;
    int bx, by, bz;                          // brick size in cells
    float inv_bx, inv_by, inv_bz;
    int nbx, nby, nbz, nb;                   // bricks per axis, per scene
    int nbricks;                             // batch * nb
    int lsx, lsy;                            // tile strides in floats (z stride = C)
    int tile_min;                            // bins with fewer queries gather from global memory
    int use_tmap;                            // tiles arrive as ONE tensor-map copy (else: a bulk copy per z-row)
    unsigned* count;                         // [nbricks]   } zeroed by the launcher
    unsigned* cursor;                        // [nbricks]   }
    unsigned* work;                          // [1]         }
    unsigned* start;                         // [nbricks + 1] first sorted query of every bin
    unsigned* ustart;                        // [nbricks + 1] first work unit of every bin (a unit = up to unit_max queries of one bin)
    int unit_max;
    uint2* units;                            // [n_units] (bin, part): written by the scatter kernel's prologue
    unsigned* bid;                           // [total] brick of every query
    unsigned* hmat;                          // [nbricks][hstride] per-block histograms (bin-major), then per-block offsets inside every bin
    int hstride;
    float4* sorted;                          // [total]: x, y, z, query index (int bits)
    long long chunk;                         // queries per block in the count / scatter kernels
};

// cell / brick size for cell < 2^15: (cell + 0.5) / b is at least 0.5 / b away from an integer, far above the rounding error
__device__ __forceinline__ int brick_coord(float x, float o, float ext, int n, float inv_b) {
    const float cell = floorf(unnorm_clip(query_grid(x, o, ext), n));       // base cell, as in trilinear_setup
    return __float2int_rz((cell + 0.5f) * inv_b);
}

__device__ __forceinline__ unsigned brick_of(const BinKP& p, float x, float y, float z, unsigned b) {
    const SampleKP& s = p.s;
    const int cx = brick_coord(x, s.ox, s.ext_x, s.nx, p.inv_bx), cy = brick_coord(y, s.oy, s.ext_y, s.ny, p.inv_by),
              cz = brick_coord(z, s.oz, s.ext_z, s.nz, p.inv_bz);
    return ((b * p.nbx + cx) * p.nby + cy) * p.nbz + cz;
}

template <bool SMEM>
__global__ void __launch_bounds__(1024) bin_count_kernel(const __grid_constant__ BinKP p) {
    extern __shared__ unsigned s_hist[];
    const unsigned q0 = (unsigned)(blockIdx.x * p.chunk), q1 = (unsigned)min((long long)q0 + p.chunk, p.s.total);
    const unsigned Q = (unsigned)min(p.s.Q, p.s.total);
    const bool one_scene = p.s.Q >= p.s.total;
    if constexpr (SMEM) {
        for (int i = threadIdx.x; i < p.nbricks; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
#pragma unroll 4
    for (unsigned q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const float* src = p.s.xyz + (size_t)q * 3;
        const unsigned br = brick_of(p, __ldg(src), __ldg(src + 1), __ldg(src + 2), one_scene ? 0u : q / Q);
        p.bid[q] = br;
        if constexpr (SMEM) atomicAdd(&s_hist[br], 1u);
        else atomicAdd(&p.count[br], 1u);
    }
    if constexpr (SMEM) {
        __syncthreads();
        unsigned* col = p.hmat + blockIdx.x;
        for (int i = threadIdx.x; i < p.nbricks; i += blockDim.x) col[(size_t)i * p.hstride] = s_hist[i];
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(1024) bin_scatter_kernel(const __grid_constant__ BinKP p) {
    extern __shared__ unsigned s_hist[];
    const unsigned q0 = (unsigned)(blockIdx.x * p.chunk), q1 = (unsigned)min((long long)q0 + p.chunk, p.s.total);
    // the work-unit list of the gather kernel: every bin's parts, in bin order (this block's slice of the bins)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nbricks; i += gridDim.x * blockDim.x) {
        const unsigned u0 = p.ustart[i], u1 = p.ustart[i + 1];
        for (unsigned u = u0; u < u1; ++u) p.units[u] = make_uint2((unsigned)i, u - u0);
    }
    if constexpr (SMEM) {
        // this block's write position in every bin: bin start + queries of the bin in earlier blocks
        const unsigned* col = p.hmat + blockIdx.x;
        for (int i = threadIdx.x; i < p.nbricks; i += blockDim.x) s_hist[i] = p.start[i] + col[(size_t)i * p.hstride];
        __syncthreads();
    }
#pragma unroll 4
    for (unsigned q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const unsigned br = __ldg(p.bid + q);
        const float* src = p.s.xyz + (size_t)q * 3;
        const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
        unsigned rank;
        if constexpr (SMEM) rank = atomicAdd(&s_hist[br], 1u);
        else rank = p.start[br] + atomicAdd(&p.cursor[br], 1u);
        p.sorted[rank] = make_float4(x, y, z, __int_as_float((int)q));
    }
}

// ---- mbarrier / bulk-copy helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// Weighted sum of the 8 corners for one float4 of channels, in ATen's corner order (sample.cuh: sample_volume).
__device__ __forceinline__ float4 corners8(const float4 v[8], const float w[8]) {
    float4 r = make_float4(__fmul_rn(v[0].x, w[0]), __fmul_rn(v[0].y, w[0]), __fmul_rn(v[0].z, w[0]), __fmul_rn(v[0].w, w[0]));
#pragma unroll
    for (int k = 1; k < 8; ++k) r.x = fmaf(v[k].x, w[k], r.x), r.y = fmaf(v[k].y, w[k], r.y), r.z = fmaf(v[k].z, w[k], r.z), r.w = fmaf(v[k].w, w[k], r.w);
    return r;
}

struct Unit {                                // a claimed work unit: queries [beg, end) of one bin, decoded
    unsigned beg, end;
    int b, x0, y0, z0;
    bool use_tile;
};
__device__ __forceinline__ Unit decode_unit(const BinKP& p, int brick, int part) {
    Unit r;
    const unsigned b0 = p.start[brick], n = p.start[brick + 1] - b0;
    const unsigned parts = (n + p.unit_max - 1) / p.unit_max, per = (n + parts - 1) / parts;      // even split
    r.beg = b0 + part * per, r.end = min(r.beg + per, b0 + n);
    int t = brick;
    const int cz = t % p.nbz; t /= p.nbz;
    const int cy = t % p.nby; t /= p.nby;
    const int cx = t % p.nbx;
    r.b = t / p.nbx;
    r.x0 = cx * p.bx, r.y0 = cy * p.by, r.z0 = cz * p.bz;
    r.use_tile = (int)(r.end - r.beg) >= p.tile_min;
    return r;
}

// Volume part of up to 32 staged queries (tab): G lanes per query, NV float4s of channels per lane.
// TILE: corners come from the shared-memory tile (LDS.128), else from global memory; the corner steps are the per-query
// ones of the table (0 at the border).
template <int NV, bool TILE>
__device__ __forceinline__ void gather_volume(const SampleKP& s, const float* __restrict__ src, const float* __restrict__ tab, int nq, int lane,
                                              int lgG, int dx, int dy, int dz) {
    const int G = 1 << lgG, sub = lane & (G - 1), grp = lane >> lgG, qpi = 32 >> lgG, CV = s.C >> 2;
    const bool swap = NV == 2 && G == 4 && (grp & 1);    // 64 B per query and load: odd queries start with the other half row
    const int fa0 = sub + (swap ? G : 0), fb0 = sub + (swap ? 0 : G);
    for (int vit = 0; vit < G; ++vit) {
        const int ql = vit * qpi + grp;
        if (ql >= nq) continue;
        const float* e = tab + ql * BIN_TAB;
        const int4 hd = *reinterpret_cast<const int4*>(e);                   // base, x / y / z step (0 at the border)
        const float4 wa = *reinterpret_cast<const float4*>(e + 4), wb = *reinterpret_cast<const float4*>(e + 8);
        const int ox = hd.y, oy = hd.z, oz = hd.w;          // (per-query steps also on the tile path: measured faster than the uniform ones at C = 128)
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const long long q = *reinterpret_cast<const long long*>(e + 12);
        const float* base = src + hd.x;
        for (int f0 = 0; f0 < CV; f0 += G * NV) {
            const int fa = f0 + fa0, fb = f0 + fb0;
            const bool ha = fa < CV, hb = NV == 2 && fb < CV;
            if constexpr (TILE) {
                float4 va[8], vb[8];                       // all 16 loads in flight
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int o = ((k & 1) ? ox : 0) + ((k & 2) ? oy : 0) + ((k & 4) ? oz : 0);
                    if (ha) va[k] = *reinterpret_cast<const float4*>(base + o + fa * 4);
                    if (hb) vb[k] = *reinterpret_cast<const float4*>(base + o + fb * 4);
                }
                if (ha) store_feat4(s, q, s.Cp + fa * 4, corners8(va, w));
                if (hb) store_feat4(s, q, s.Cp + fb * 4, corners8(vb, w));
            } else {
#pragma unroll 1
                for (int h = 0; h < NV; ++h) {             // sparse bins: one float4 of channels at a time (64-bit addresses)
                    const int f = h ? fb : fa;
                    if (f >= CV) continue;
                    float4 v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = ldg4(base + ((k & 1) ? ox : 0) + ((k & 2) ? oy : 0) + ((k & 4) ? oz : 0) + f * 4);
                    store_feat4(s, q, s.Cp + f * 4, corners8(v, w));
                }
            }
        }
    }
}

// The tile path at C = 32 (9 x 9 x 9 voxels of 128 B): 4 lanes per query, two float4 each, every corner step a compile-time
// immediate of the LDS.128.
constexpr int T32_Z = 32, T32_Y = 9 * 32, T32_X = 81 * 32;
__device__ __forceinline__ void gather_tile_c32(const SampleKP& s, const float* __restrict__ tile, const float* __restrict__ tab, int nq, int lane) {
    const int sub = lane & 3, grp = lane >> 2;
    const int fa = (sub + ((grp & 1) ? 4 : 0)) * 4, fb = fa ^ 16;       // odd queries start with the other half row (bank halves)
#pragma unroll 1
    for (int vit = 0; vit < 4; ++vit) {
        const int ql = vit * 8 + grp;
        if (ql >= nq) continue;
        const float* e = tab + ql * BIN_TAB;
        const int base = *reinterpret_cast<const int*>(e);
        const float4 wa = *reinterpret_cast<const float4*>(e + 4), wb = *reinterpret_cast<const float4*>(e + 8);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const long long q = *reinterpret_cast<const long long*>(e + 12);
        const float* pa = tile + base + fa;
        const float* pb = tile + base + fb;
        float4 va[8], vb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr int dummy = 0;
            const int o = ((k & 1) ? T32_X : dummy) + ((k & 2) ? T32_Y : dummy) + ((k & 4) ? T32_Z : dummy);
            va[k] = *reinterpret_cast<const float4*>(pa + o);
            vb[k] = *reinterpret_cast<const float4*>(pb + o);
        }
        store_feat4(s, q, s.Cp + fa, corners8(va, w));
        store_feat4(s, q, s.Cp + fb, corners8(vb, w));
    }
}

// NV = float4s of channels per lane in the volume gather (2: C = 32 takes 4 lanes per query, 8 queries per trip)
template <int NV>
__global__ void __launch_bounds__(32 * (BIN_WARPS + 1), 1) sample_binned_kernel(const __grid_constant__ BinKP p, const __grid_constant__ CUtensorMap tmap, int lgGv, int Gp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[2], s_empty[2];
    __shared__ Unit s_unit[2];               // the producer's decoded claim per stage (b < 0: no more work)
    const SampleKP& s = p.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) bar_init(smem_u32(&s_full[i]), 1), bar_init(smem_u32(&s_empty[i]), BIN_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == BIN_WARPS) {
        // ---------------- producer: claim the next work unit, wait for its buffer, stage the tile --------------------
        const int n_units = (int)p.ustart[p.nbricks];
        for (int it = 0;; ++it) {
            const int st = it & 1;
            int brick = -1, part = 0;
            if (lane == 0) {
                const int u = (int)atomicAdd(p.work, 1u);
                if (u < n_units) {
                    const uint2 w = p.units[u];
                    brick = (int)w.x, part = (int)w.y;
                }
            }
            brick = __shfl_sync(FULL, brick, 0), part = __shfl_sync(FULL, part, 0);
            const uint32_t full = smem_u32(&s_full[st]);
            bar_wait(smem_u32(&s_empty[st]), ((it >> 1) & 1) ^ 1);           // every consumer warp has left this buffer
            Unit un;
            un.b = -1;
            if (brick >= 0) un = decode_unit(p, brick, part);
            if (lane == 0) s_unit[st] = un;
            if (brick < 0) {
                if (lane == 0) bar_arrive(full);
                break;
            }
            if (!un.use_tile) {
                if (lane == 0) bar_arrive(full);
                continue;
            }
            if (p.use_tmap) {
                // the whole (bx+1) x (by+1) x (bz+1) x C box in one request; what lies beyond the grid arrives as zeros and
                // is never read (border bricks take the gather with per-query corner steps, 0 at the border)
                if (lane == 0) {
                    const uint32_t box_bytes = (uint32_t)((p.bx + 1) * (p.by + 1) * (p.bz + 1) * s.C * 4);
                    bar_expect_tx(full, box_bytes);
                    asm volatile(
                        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                            smem_u32(smem_raw + st * BIN_TILE_BYTES)),
                        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(un.z0), "r"(un.y0), "r"(un.x0), "r"(un.b), "r"(full)
                        : "memory");
                }
                continue;
            }
            // the tile always holds (bx+1) x (by+1) x (bz+1) corner voxels: where the grid ends inside it, the last voxel of
            // the axis is replicated, which is exactly the clamped corner (weight 0) the reference-order sum reads there
            const int tx = min(p.bx + 1, s.nx - un.x0), ty = min(p.by + 1, s.ny - un.y0), tz = min(p.bz + 1, s.nz - un.z0);
            const int txr = min(p.bx + 1, tx + 1), tyr = min(p.by + 1, ty + 1);
            const bool zrep = tz < p.bz + 1;
            const uint32_t row_bytes = (uint32_t)(tz * s.C * 4), vox_bytes = (uint32_t)(s.C * 4);
            if (lane == 0) bar_expect_tx(full, (row_bytes + (zrep ? vox_bytes : 0u)) * (uint32_t)(txr * tyr));
            __syncwarp();
            float* tile = reinterpret_cast<float*>(smem_raw + st * BIN_TILE_BYTES);
            const float* src = s.volume + (long long)un.b * s.vsb + un.x0 * s.vsx + un.y0 * s.vsy + un.z0 * s.vsz;
            for (int r = lane; r < txr * tyr; r += 32) {
                const int xi = r / tyr, yi = r - xi * tyr;
                const float* row = src + min(xi, tx - 1) * s.vsx + min(yi, ty - 1) * s.vsy;
                float* dst = tile + xi * p.lsx + yi * p.lsy;
                bulk_g2s(smem_u32(dst), row, row_bytes, full);
                if (zrep) bulk_g2s(smem_u32(dst + tz * s.C), row + (tz - 1) * s.vsz, vox_bytes, full);
            }
        }
        return;
    }
    // -------------------- consumers: a warp takes 32 queries of the unit at a time ----------------------------------
    float* tab = reinterpret_cast<float*>(smem_raw + 2 * BIN_TILE_BYTES) + warp * 32 * BIN_TAB;
    for (int it = 0;; ++it) {
        const int st = it & 1;
        bar_wait(smem_u32(&s_full[st]), (it >> 1) & 1);
        const Unit un = s_unit[st];
        if (un.b < 0) break;
        const int n = (int)(un.end - un.beg), b = un.b;
        const float* vol = s.volume + (long long)b * s.vsb;
        const float* tile = reinterpret_cast<const float*>(smem_raw + st * BIN_TILE_BYTES);
        const bool use_tile = un.use_tile;
        const bool interior = un.x0 + p.bx + 1 <= s.nx && un.y0 + p.by + 1 <= s.ny && un.z0 + p.bz + 1 <= s.nz;   // no replicated / zero halo
        const int dx = use_tile ? p.lsx : (int)s.vsx, dy = use_tile ? p.lsy : (int)s.vsy, dz = use_tile ? s.C : (int)s.vsz;   // gather strides
        const long long origin_off = use_tile ? (long long)un.x0 * dx + (long long)un.y0 * dy + (long long)un.z0 * dz : 0;
        float4 pt_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp * 32 + lane < n) pt_next = __ldg(p.sorted + un.beg + warp * 32 + lane);
        for (int g = warp; g * 32 < n; g += BIN_WARPS) {
            const bool live = g * 32 + lane < n;
            const int nq = min(32, n - g * 32);
            const float4 pt = pt_next;
            {                                               // the next group's points travel while this group is gathered
                const int nxt = (g + BIN_WARPS) * 32 + lane;
                if (nxt < n) pt_next = __ldg(p.sorted + un.beg + nxt);
            }
            const int qidx = __float_as_int(pt.w);
            // ---- plane part: straight from global memory (L1/L2; the bin keeps the accesses local) ----------------
            if (s.Cp > 0) {
                if (live) {
                    BiCorners bc[3];
                    planes_setup(s, pt.x, pt.y, pt.z, bc);
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        float* e = tab + lane * BIN_TAB + pl * 6;
                        const int fl = (bc[pl].off[1] != bc[pl].off[0] ? 1 : 0) | (bc[pl].off[2] != bc[pl].off[0] ? 2 : 0);
                        reinterpret_cast<int*>(e)[0] = (int)bc[pl].off[0];
                        reinterpret_cast<int*>(e)[1] = fl;
                        e[2] = bc[pl].w[0], e[3] = bc[pl].w[1], e[4] = bc[pl].w[2], e[5] = bc[pl].w[3];
                    }
                    reinterpret_cast<int*>(tab + lane * BIN_TAB)[18] = qidx;
                }
                __syncwarp();
                const int psub = lane % Gp, pqpi = 32 / Gp;
                for (int pit = 0; pit < Gp; ++pit) {
                    const int ql = pit * pqpi + lane / Gp;
                    if (ql >= nq) continue;
                    const float* e = tab + ql * BIN_TAB;
                    const long long q = (long long)reinterpret_cast<const int*>(e)[18];
                    for (int c = psub * 4; c < s.Cp; c += Gp * 4) {
                        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) {
                            if (s.plane[pl] == nullptr) continue;
                            const float* ee = e + pl * 6;
                            const int2 of = *reinterpret_cast<const int2*>(ee);
                            const float2 w01 = *reinterpret_cast<const float2*>(ee + 2), w23 = *reinterpret_cast<const float2*>(ee + 4);
                            const bool fx = of.y & 1, fy = of.y & 2;
                            const float* base = s.plane[pl] + (long long)b * s.psb + c + of.x;
                            const int ox = fx ? (int)s.psw : 0, oy = fy ? (int)s.psh : 0;
                            const float4 v0 = ldg4(base), v1 = ldg4(base + ox), v2 = ldg4(base + oy), v3 = ldg4(base + oy + ox);
                            const float w0 = w01.x, w1 = w01.y, w2 = w23.x, w3 = w23.y;
                            float4 a = make_float4(__fmul_rn(v0.x, w0), __fmul_rn(v0.y, w0), __fmul_rn(v0.z, w0), __fmul_rn(v0.w, w0));
                            a.x = fmaf(v1.x, w1, a.x), a.y = fmaf(v1.y, w1, a.y), a.z = fmaf(v1.z, w1, a.z), a.w = fmaf(v1.w, w1, a.w);
                            a.x = fmaf(v2.x, w2, a.x), a.y = fmaf(v2.y, w2, a.y), a.z = fmaf(v2.z, w2, a.z), a.w = fmaf(v2.w, w2, a.w);
                            a.x = fmaf(v3.x, w3, a.x), a.y = fmaf(v3.y, w3, a.y), a.z = fmaf(v3.z, w3, a.z), a.w = fmaf(v3.w, w3, a.w);
                            r.x = __fadd_rn(r.x, a.x), r.y = __fadd_rn(r.y, a.y), r.z = __fadd_rn(r.z, a.z), r.w = __fadd_rn(r.w, a.w);
                        }
                        store_feat4(s, q, c, r);
                    }
                }
                __syncwarp();
            }
            // ---- volume part: base offset, 3 border flags and the 8 corner weights per query ----------------------
            if (live) {
                TriCorners tc;                              // offsets in gather strides: relative to the tile / the scene's volume
                SampleKP ls = s;
                ls.vsx = dx, ls.vsy = dy, ls.vsz = dz;
                trilinear_setup(ls, pt.x, pt.y, pt.z, tc);
                float* e = tab + lane * BIN_TAB;
                reinterpret_cast<int*>(e)[0] = (int)(tc.off[0] - origin_off);
                reinterpret_cast<int*>(e)[1] = (int)(tc.off[1] - tc.off[0]);   // 0 where the +1 corner is beyond the border
                reinterpret_cast<int*>(e)[2] = (int)(tc.off[2] - tc.off[0]);
                reinterpret_cast<int*>(e)[3] = (int)(tc.off[4] - tc.off[0]);
                *reinterpret_cast<long long*>(e + 12) = (long long)qidx;
#pragma unroll
                for (int k = 0; k < 8; ++k) e[4 + k] = tc.w[k];
            }
            __syncwarp();
            if (use_tile) {
                if (NV == 2 && s.C == 32 && dx == T32_X && dy == T32_Y && (!p.use_tmap || interior)) gather_tile_c32(s, tile, tab, nq, lane);
                else gather_volume<NV, true>(s, tile, tab, nq, lane, lgGv, dx, dy, dz);
            } else {
                gather_volume<NV, false>(s, vol, tab, nq, lane, lgGv, dx, dy, dz);
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) bar_arrive(smem_u32(&s_empty[st]));                  // this warp is done with the buffer
    }
}

struct BinPlan {
    BinKP kp;
    int NV, lgGv, Gp;
    size_t zero_bytes, total_bytes, hist_bytes, o_hmat;
    bool smem_hist;
    int sort_blocks;
};

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Fills the plan; returns false when the binned path does not apply to these parameters.
static bool plan_binned(const GnbSampleParams* sp, BinPlan& pl) {
    SampleKP& s = pl.kp.s;
    if (fill_sample_kp(sp, s)) return false;
    if (!s.volume || s.total <= 0 || s.total >= 0x7fffffffLL) return false;
    // channels-last volume whose z-rows are contiguous, float4 everywhere, int32 offsets
    if (s.vsc != 1 || s.vsz != s.C || s.C % 4 || s.C * 4 * 8 > BIN_TILE_BYTES) return false;
    if (s.vsx % 4 || s.vsy % 4 || s.vsb % 4 || !aligned16(s.volume)) return false;
    if ((long long)s.nx * s.vsx + (long long)s.ny * s.vsy + (long long)s.nz * s.vsz >= 0x7fffffffLL) return false;
    if (s.Cp > 0) {
        if (s.psc != 1 || s.Cp % 4 || s.psb % 4 || s.psh % 4 || s.psw % 4) return false;
        if ((long long)s.R * s.psh + (long long)s.R * s.psw >= 0x7fffffffLL) return false;
        for (int k = 0; k < 3; ++k)
            if (s.plane[k] && !aligned16(s.plane[k])) return false;
    }
    if (sp->out && (!aligned16(sp->out) || sp->out_stride % 4)) return false;
    // brick = the largest near-cube whose corner voxels fit the tile
    const int maxvox = BIN_TILE_BYTES / (s.C * 4);
    int e = 2;
    while ((e + 1) * (e + 1) * (e + 1) <= maxvox) ++e;       // e = corner voxels per axis
    int ez = maxvox / (e * e);
    BinKP& k = pl.kp;
    k.bx = k.by = e - 1, k.bz = ez - 1;
    if (k.bx < 1 || k.bz < 1 || s.nx >= 32768 || s.ny >= 32768 || s.nz >= 32768) return false;
    k.inv_bx = 1.0f / (float)k.bx, k.inv_by = 1.0f / (float)k.by, k.inv_bz = 1.0f / (float)k.bz;
    k.nbx = (s.nx + k.bx - 1) / k.bx, k.nby = (s.ny + k.by - 1) / k.by, k.nbz = (s.nz + k.bz - 1) / k.bz;
    const long long nb = (long long)k.nbx * k.nby * k.nbz, nbricks = nb * sp->batch;
    if (nbricks >= (1LL << 24)) return false;
    k.nb = (int)nb, k.nbricks = (int)nbricks;
    k.lsy = ez * s.C, k.lsx = e * k.lsy;
    k.tile_min = e * e * ez / 8;
    auto pow2_lanes = [](int n) { int g = 1; while (g < n && g < 32) g <<= 1; return g; };
    const int CV = s.C / 4;
    pl.NV = (CV >= 8 && CV % 2 == 0) ? 2 : 1;                 // float4s per lane; lanes per query = pow2 >= CV / NV
    const int Gv = pow2_lanes(CV / pl.NV);
    pl.lgGv = 0;
    while ((1 << pl.lgGv) < Gv) ++pl.lgGv;
    pl.Gp = pow2_lanes(s.Cp / 4 > 0 ? s.Cp / 4 : 1);
    pl.smem_hist = k.nbricks <= BIN_HIST_MAX;
    pl.hist_bytes = pl.smem_hist ? (size_t)k.nbricks * 4 : 0;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_count = 0, o_cursor = up((size_t)k.nbricks * 4), o_work = o_cursor + up((size_t)k.nbricks * 4);
    pl.zero_bytes = o_work + 256;
    const size_t o_start = pl.zero_bytes, o_ustart = o_start + up(((size_t)k.nbricks + 1) * 4);
    k.unit_max = 1536;
    if (const int e = opt(OPT_BIN_UNIT)) k.unit_max = e > 31 ? e : 1536;       // tuning aid
    const size_t max_units = (size_t)k.nbricks + (size_t)(s.total / k.unit_max) + 1;
    const size_t o_units = o_ustart + up(((size_t)k.nbricks + 1) * 4), o_bid = o_units + up(max_units * 8), o_sorted = o_bid + up((size_t)s.total * 4);
    k.units = reinterpret_cast<uint2*>(o_units);
    pl.o_hmat = o_sorted + up((size_t)s.total * 16);
    k.ustart = reinterpret_cast<unsigned*>(o_ustart);
    pl.total_bytes = pl.o_hmat + (pl.smem_hist ? (size_t)((BIN_MAXB + 31) / 32 * 32) * k.nbricks * 4 : 0);
    k.bid = reinterpret_cast<unsigned*>(o_bid);
    k.count = reinterpret_cast<unsigned*>(o_count), k.cursor = reinterpret_cast<unsigned*>(o_cursor);     // offsets, rebased by the launcher
    k.work = reinterpret_cast<unsigned*>(o_work), k.start = reinterpret_cast<unsigned*>(o_start);
    k.sorted = reinterpret_cast<float4*>(o_sorted);
    return true;
}

}  // namespace gnb

using namespace gnb;

extern "C" int64_t gnb_sample_binned_scratch_bytes(const GnbSampleParams* sp) {
    BinPlan pl;
    if (!sp || !plan_binned(sp, pl)) return 0;
    return (int64_t)pl.total_bytes;
}

// The counting sort of the queries by brick (count -> reduce -> scan -> scatter): fills k.sorted / k.start / k.units in `scratch`.
static int launch_bin_sort(BinPlan& pl, void* scratch, cudaStream_t stream, int sms) {
    BinKP& k = pl.kp;
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    k.count = reinterpret_cast<unsigned*>(base + (size_t)k.count), k.cursor = reinterpret_cast<unsigned*>(base + (size_t)k.cursor);
    k.work = reinterpret_cast<unsigned*>(base + (size_t)k.work), k.start = reinterpret_cast<unsigned*>(base + (size_t)k.start);
    k.sorted = reinterpret_cast<float4*>(base + (size_t)k.sorted), k.bid = reinterpret_cast<unsigned*>(base + (size_t)k.bid);
    k.ustart = reinterpret_cast<unsigned*>(base + (size_t)k.ustart), k.units = reinterpret_cast<uint2*>(base + (size_t)k.units);
    k.hmat = reinterpret_cast<unsigned*>(base + pl.o_hmat);
    k.hstride = (BIN_MAXB + 31) / 32 * 32;
    // count / scatter: contiguous chunks of queries per block (a multiple of the block size), the same in both kernels
    long long blocks = (k.s.total + 2047) / 2048;
    const long long max_blocks = 2LL * sms < BIN_MAXB ? 2LL * sms : BIN_MAXB;
    if (blocks > max_blocks) blocks = max_blocks;
    k.chunk = ((k.s.total + blocks - 1) / blocks + 31) / 32 * 32;          // equal chunks: every SM gets the same number of blocks
    blocks = (k.s.total + k.chunk - 1) / k.chunk;
    if (pl.smem_hist) {
        GNB_CUDA(cudaFuncSetAttribute(bin_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.hist_bytes));
        GNB_CUDA(cudaFuncSetAttribute(bin_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.hist_bytes));
        bin_count_kernel<true><<<(unsigned)blocks, 1024, pl.hist_bytes, stream>>>(k);
        GNB_LAUNCH_CHECK();
        bin_reduce_kernel<<<(unsigned)((k.nbricks + 31) / 32), 1024, 0, stream>>>(k.hmat, k.count, k.nbricks, (int)blocks, k.hstride);
    } else {
        GNB_CUDA(cudaMemsetAsync(base, 0, pl.zero_bytes, stream));       // count and cursor are accumulated with global atomics
        bin_count_kernel<false><<<(unsigned)blocks, 1024, 0, stream>>>(k);
    }
    GNB_LAUNCH_CHECK();
    bin_scan_kernel<<<1, 1024, 0, stream>>>(k.count, k.start, k.ustart, k.nbricks, k.unit_max, k.work);
    GNB_LAUNCH_CHECK();
    if (pl.smem_hist) bin_scatter_kernel<true><<<(unsigned)blocks, 1024, pl.hist_bytes, stream>>>(k);
    else bin_scatter_kernel<false><<<(unsigned)blocks, 1024, 0, stream>>>(k);
    GNB_LAUNCH_CHECK();
    return 0;
}

// Sort only (the fused decoder's prologue then walks the queries brick by brick: its gathers hit L1 / L2 instead of DRAM).
// *sorted: [total] records (x, y, z, query index as int bits) inside `scratch`.
namespace gnb {
int64_t bin_sort_scratch_bytes(const GnbSampleParams* sp) {
    BinPlan pl;
    if (!sp || !plan_binned(sp, pl)) return 0;
    return (int64_t)pl.total_bytes;
}
int bin_sort(const GnbSampleParams* sp, void* scratch, int64_t scratch_bytes, void* stream_, const float4** sorted) {
    BinPlan pl;
    GNB_CHECK_ARG(sp && sorted, "bin_sort: null argument");
    GNB_CHECK_ARG(plan_binned(sp, pl), "bin_sort: needs a channels-last fp32 volume with C %% 4 == 0");
    GNB_CHECK_ARG(scratch && aligned16(scratch) && scratch_bytes >= (int64_t)pl.total_bytes, "bin_sort: scratch too small");
    int dev = 0, sms = 148;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int rc = launch_bin_sort(pl, scratch, (cudaStream_t)stream_, sms);
    if (rc) return rc;
    *sorted = pl.kp.sorted;
    return 0;
}
}  // namespace gnb

extern "C" int gnb_sample_features_binned(const GnbSampleParams* sp, void* scratch, int64_t scratch_bytes, void* stream_) {
    BinPlan pl;
    GNB_CHECK_ARG(sp, "sample_binned: null params");
    {
        SampleKP chk;
        int rc = fill_sample_kp(sp, chk);
        if (rc) return rc;
        if (chk.total == 0) return 0;
    }
    GNB_CHECK_ARG(plan_binned(sp, pl), "sample_binned: needs a channels-last fp32 volume with C %% 4 == 0 (use gnb_sample_features)");
    GNB_CHECK_ARG((sp->out && sp->out_stride >= pl.kp.s.C + pl.kp.s.Cp) || (!sp->out && sp->image), "sample_binned: bad output");
    GNB_CHECK_ARG(scratch && aligned16(scratch) && scratch_bytes >= (int64_t)pl.total_bytes, "sample_binned: scratch too small (gnb_sample_binned_scratch_bytes)");
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 148;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    {
        int rc = launch_bin_sort(pl, scratch, stream, sms);
        if (rc) return rc;
    }
    BinKP& k = pl.kp;
    // tensor map of the channels-last volume (C, z, y, x, scene): a brick's tile is one TMA box
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    k.use_tmap = 0;
    if (k.s.C <= 256 && !opt(OPT_BIN_ROWCOPY)) {
        const int e = k.bx + 1, ez = k.bz + 1;
        cuuint64_t gdim[5] = {(cuuint64_t)k.s.C, (cuuint64_t)k.s.nz, (cuuint64_t)k.s.ny, (cuuint64_t)k.s.nx, (cuuint64_t)sp->batch};
        cuuint64_t gstr[4] = {(cuuint64_t)k.s.vsz * 4, (cuuint64_t)k.s.vsy * 4, (cuuint64_t)k.s.vsx * 4, (cuuint64_t)k.s.vsb * 4};
        cuuint32_t box[5] = {(cuuint32_t)k.s.C, (cuuint32_t)ez, (cuuint32_t)e, (cuuint32_t)e, 1u};
        cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
        if (sp->batch == 1) gstr[3] = gstr[2] * (cuuint64_t)k.s.nx;        // unused extent-1 dimension: any valid stride
        // the driver entry point comes through the runtime, so the library does not link against libcuda
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
            const CUresult r = reinterpret_cast<EncodeTiled>(fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(k.s.volume), gdim, gstr, box,
                                                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            k.use_tmap = r == CUDA_SUCCESS ? 1 : 0;
        } else {
            (void)cudaGetLastError();
        }
    }
    const size_t smem = 2 * (size_t)BIN_TILE_BYTES + (size_t)BIN_WARPS * 32 * BIN_TAB * 4;
    const unsigned grid = (unsigned)(sms < k.nbricks ? sms : k.nbricks), threads = 32 * (BIN_WARPS + 1);
    if (pl.NV == 2) {
        GNB_CUDA(cudaFuncSetAttribute(sample_binned_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_binned_kernel<2><<<grid, threads, smem, stream>>>(k, tmap, pl.lgGv, pl.Gp);
    } else {
        GNB_CUDA(cudaFuncSetAttribute(sample_binned_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_binned_kernel<1><<<grid, threads, smem, stream>>>(k, tmap, pl.lgGv, pl.Gp);
    }
    GNB_LAUNCH_CHECK();
    return 0;
}

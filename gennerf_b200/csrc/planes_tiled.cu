// Tiled triplane scatter-mean (sm_100a): the shared-memory path for MANY points.
//
// Replaces LocalPoolPointnet.generate_plane_features() (reference src/models/components/pointnet.py:72-89,
// = torch_scatter.scatter_mean: scatter_add_ of features and of ones, clamp(count, 1), divide) for the three planes at
// once, like the atomic mode of planes.cu -- but without a single global floating-point atomic.
//
// Why: 614 400 points x 3 planes x 32 channels are 59 M fp32 reductions; the L2 processes them element by element
// (~0.4 T/s), which pins the atomic kernel at ~155 us whatever the access pattern, and the clamped border cells of the
// reference's metric coordinates (SURVEY trap T6) serialise on top of that.  Here the (point, plane) pairs are
// counting-sorted by the TS x TS tile of plane cells they fall into (binsort.cuh); a block then takes one work unit (up
// to UNIT_MAX points of one tile), buckets its points by cell in shared memory, lets every warp add the rows of the cells
// it owns in registers and writes the MEANS and the counts with plain stores: no zero-fill, no finalize pass.  A tile
// with more points than one unit is split: every part stores its partial tile, the part that arrives last adds the
// partials in part order, divides and writes.  Counts are integers and exact; sums are within fp32 rounding of the reference (same bar as the
// atomic mode); empty cells are written as 0.
#include "binsort.cuh"

namespace gnb {

constexpr int TILE_BYTES = 32 * 1024;        // a tile's cells x channels (sets the tile side: 16 x 16 cells at C_p = 32)
constexpr int TILE_UNIT_MAX = 4096;          // points per work unit
constexpr int TILE_HIST_MAX = 40960;

struct TileKP {
    const float* p;
    const float* c;
    int B;
    long long N;                             // points per scene
    long long total;                         // B * N
    int Cp, R;
    float den;
    int TS, nt;                              // tile side in cells, tiles per plane axis
    int nbins;                               // 3 * B * nt * nt, bin = ((plane * B + b) * nt + ty) * nt + tx
    float* planes;                           // (3, B, R, R, Cp)
    int* count;                              // (3, B, R, R)
    unsigned *hmat, *binsize, *start, *ustart, *work, *arrive;
    int hstride;
    uint2* units;                            // (bin, part)
    uint2* sorted;                           // (global point index, cell inside the tile), grouped by bin
    float* partial;                          // [unit][TS*TS*(Cp+1)] partial tiles of split bins
    long long chunk;
};

// plane k uses point coordinates (A0[k], A1[k]): xz -> (0,2), xy -> (0,1), yz -> (1,2); cell = i0 + R * i1 (utils.py:67-69)
__device__ __forceinline__ void tile_bins(const TileKP& q, unsigned gidx, unsigned bin[3], unsigned local[3]) {
    const float* pp = q.p + (size_t)gidx * 3;
    const float x = __ldg(pp), y = __ldg(pp + 1), z = __ldg(pp + 2);
    const unsigned b = q.total == q.N ? 0u : gidx / (unsigned)q.N;
    int i[3];
    i[0] = (int)__fmul_rn(plane_unit(x, q.den), (float)q.R);
    i[1] = (int)__fmul_rn(plane_unit(y, q.den), (float)q.R);
    i[2] = (int)__fmul_rn(plane_unit(z, q.den), (float)q.R);
    const int a0[3] = {0, 0, 1}, a1[3] = {2, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int cx = i[a0[k]], cy = i[a1[k]];
        const int tx = cx / q.TS, ty = cy / q.TS;
        bin[k] = (((unsigned)k * q.B + b) * q.nt + ty) * q.nt + tx;
        local[k] = (unsigned)((cy - ty * q.TS) * q.TS + (cx - tx * q.TS));
    }
}

__global__ void __launch_bounds__(1024) tile_count_kernel(const __grid_constant__ TileKP q) {
    extern __shared__ unsigned s_hist[];
    const unsigned g0 = (unsigned)(blockIdx.x * q.chunk), g1 = (unsigned)min((long long)g0 + q.chunk, q.total);
    for (int i = threadIdx.x; i < q.nbins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
#pragma unroll 2
    for (unsigned g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
        unsigned bin[3], local[3];
        tile_bins(q, g, bin, local);
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(&s_hist[bin[k]], 1u);
    }
    __syncthreads();
    unsigned* col = q.hmat + blockIdx.x;
    for (int i = threadIdx.x; i < q.nbins; i += blockDim.x) col[(size_t)i * q.hstride] = s_hist[i];
}

__global__ void __launch_bounds__(1024) tile_scatter_kernel(const __grid_constant__ TileKP q) {
    extern __shared__ unsigned s_hist[];
    const unsigned g0 = (unsigned)(blockIdx.x * q.chunk), g1 = (unsigned)min((long long)g0 + q.chunk, q.total);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.nbins; i += gridDim.x * blockDim.x) {       // work-unit list
        const unsigned u0 = q.ustart[i], u1 = q.ustart[i + 1];
        for (unsigned u = u0; u < u1; ++u) q.units[u] = make_uint2((unsigned)i, u - u0);
        q.arrive[i] = 0u;
    }
    const unsigned* col = q.hmat + blockIdx.x;
    for (int i = threadIdx.x; i < q.nbins; i += blockDim.x) s_hist[i] = q.start[i] + col[(size_t)i * q.hstride];
    __syncthreads();
#pragma unroll 2
    for (unsigned g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
        unsigned bin[3], local[3];
        tile_bins(q, g, bin, local);
#pragma unroll
        for (int k = 0; k < 3; ++k) q.sorted[atomicAdd(&s_hist[bin[k]], 1u)] = make_uint2(g, local[k]);
    }
}

// One work unit per block trip.  NCH = 32-channel groups per point row.
// Shared memory has no native fp32 reduction (atomicAdd on shared floats is a compare-and-swap loop), so the unit's points
// are first bucketed by cell with integer shared-memory atomics (histogram, scan, scatter of the point indices); then every
// cell is OWNED by one warp, which adds the cell's feature rows in registers (lanes = channels, the next cell's rows
// already in flight) and writes the result once.  Cells with very many points (the clamped cells of metric coordinates)
// are summed by all warps in slices that are combined in a fixed order.
constexpr int TILE_THREADS = 512, TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_HOT = 256;                // points per cell from which all warps share the cell
constexpr int TILE_PF = 8;                   // rows of a cell fetched ahead

template <int NCH>
__global__ void __launch_bounds__(TILE_THREADS) tile_accum_kernel(const __grid_constant__ TileKP q) {
    extern __shared__ __align__(16) unsigned s_mem[];            // cnt[cells] | off[cells + 1] | cur[cells] | pts[UNIT_MAX] | slices | hot list
    __shared__ int s_last, s_nhot;
    __shared__ unsigned s_scan[TILE_WARPS];
    const int cells = q.TS * q.TS;
    unsigned* s_cnt = s_mem;
    unsigned* s_off = s_cnt + cells;
    unsigned* s_cur = s_off + cells + 1;
    unsigned* s_pts = s_cur + cells;
    float* s_slice = reinterpret_cast<float*>(s_pts + TILE_UNIT_MAX);          // [TILE_WARPS][NCH * 32]
    unsigned* s_hot = reinterpret_cast<unsigned*>(s_slice + TILE_WARPS * NCH * 32);   // [TILE_UNIT_MAX / TILE_HOT + 1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n_units = q.ustart[q.nbins];
    const long long RR = (long long)q.R * q.R;
    for (unsigned u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint2 un = q.units[u];
        const unsigned bin = un.x, part = un.y;
        const unsigned b0 = q.start[bin], n = q.start[bin + 1] - b0;
        const unsigned parts = (n + TILE_UNIT_MAX - 1) / TILE_UNIT_MAX, per = (n + parts - 1) / parts;
        const unsigned beg = b0 + part * per, end = min(beg + per, b0 + n);
        // ---- bucket the unit's points by cell -----------------------------------------------------------------------
        for (int i = threadIdx.x; i < cells; i += blockDim.x) s_cnt[i] = 0;
        if (threadIdx.x == 0) s_nhot = 0;
        __syncthreads();
        // (lanes that share a cell go through ONE shared-memory atomic: metric coordinates put whole units into one cell)
        for (unsigned e0 = beg + warp * 32; e0 < end; e0 += blockDim.x) {
            const unsigned e = e0 + lane;
            const unsigned cell = e < end ? __ldg(q.sorted + e).y : 0xffffffffu;
            const unsigned peers = __match_any_sync(FULL, cell);
            if (cell != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[cell], (unsigned)__popc(peers));
        }
        __syncthreads();
        {   // exclusive scan of cnt -> off: contiguous runs per thread, warp scan, scan of the warp totals
            const int per_t = (cells + blockDim.x - 1) / blockDim.x, i0 = threadIdx.x * per_t, i1 = min(i0 + per_t, cells);
            unsigned sum = 0;
            for (int i = i0; i < i1; ++i) sum += s_cnt[i];
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            unsigned base = 0;
            for (int w = 0; w < warp; ++w) base += s_scan[w];
            unsigned run = base + inc - sum;
            for (int i = i0; i < i1; ++i) {
                const unsigned k = s_cnt[i];
                s_off[i] = run, s_cur[i] = run;
                if (k >= TILE_HOT) s_hot[atomicAdd(&s_nhot, 1)] = (unsigned)i;
                run += k;
            }
            if (threadIdx.x == 0) s_off[cells] = end - beg;
        }
        __syncthreads();
        for (unsigned e0 = beg + warp * 32; e0 < end; e0 += blockDim.x) {
            const unsigned e = e0 + lane;
            const uint2 en = e < end ? __ldg(q.sorted + e) : make_uint2(0u, 0xffffffffu);
            const unsigned peers = __match_any_sync(FULL, en.y);
            const int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if (en.y != 0xffffffffu && lane == leader) base = atomicAdd(&s_cur[en.y], (unsigned)__popc(peers));
            base = __shfl_sync(FULL, base, leader);
            if (en.y != 0xffffffffu) s_pts[base + __popc(peers & ((1u << lane) - 1))] = en.x;
        }
        __syncthreads();
        // ---- where the tile lies in its plane, where the results go ---------------------------------------------------
        unsigned t = bin;
        const int tx = t % q.nt; t /= q.nt;
        const int ty = t % q.nt; t /= q.nt;                      // t = plane * B + b
        const int x0 = tx * q.TS, y0 = ty * q.TS;
        const int w = min(q.TS, q.R - x0), hgt = min(q.TS, q.R - y0);
        float* pl = q.planes + (size_t)t * RR * q.Cp;
        int* cn = q.count + (size_t)t * RR;
        const size_t psize = (size_t)cells * (q.Cp + 1);
        float* mine = q.partial + (size_t)u * psize;
        // a cell's result: the mean straight into the plane (whole tile in this unit) or the sum into this part's partial tile
        auto emit = [&](int cell, const float (&acc)[NCH], unsigned k) {
            const int ry = cell / q.TS, rx = cell - ry * q.TS;
            if (parts == 1) {
                if (ry >= hgt || rx >= w) return;
                const size_t gc = (size_t)(y0 + ry) * q.R + x0 + rx;
#pragma unroll
                for (int h = 0; h < NCH; ++h)
                    if (h * 32 + lane < q.Cp) pl[gc * q.Cp + h * 32 + lane] = k > 1 ? __fdiv_rn(acc[h], (float)k) : acc[h];
                if (lane == 0) cn[gc] = (int)k;
            } else {
#pragma unroll
                for (int h = 0; h < NCH; ++h)
                    if (h * 32 + lane < q.Cp) mine[(size_t)cell * q.Cp + h * 32 + lane] = acc[h];
                if (lane == 0) mine[(size_t)cells * q.Cp + cell] = __int_as_float((int)k);
            }
        };
        // ---- cells owned by one warp: the first rows of the next cell are fetched while this one is added -----------
        float nx[TILE_PF][NCH];
        auto fetch = [&](int cell) {
            if (cell >= cells) return;
            const unsigned k = s_cnt[cell], o = s_off[cell];
            if (k >= TILE_HOT) return;
#pragma unroll
            for (int j = 0; j < TILE_PF; ++j) {
                if ((unsigned)j < k) {
                    const float* row = q.c + (size_t)s_pts[o + j] * q.Cp;
#pragma unroll
                    for (int h = 0; h < NCH; ++h) nx[j][h] = (h * 32 + lane < q.Cp) ? __ldg(row + h * 32 + lane) : 0.0f;
                }
            }
        };
        fetch(warp);
        for (int cell = warp; cell < cells; cell += TILE_WARPS) {
            const unsigned k = s_cnt[cell], o = s_off[cell];
            float cur[TILE_PF][NCH];
#pragma unroll
            for (int j = 0; j < TILE_PF; ++j)
#pragma unroll
                for (int h = 0; h < NCH; ++h) cur[j][h] = nx[j][h];
            fetch(cell + TILE_WARPS);
            if (k >= TILE_HOT) continue;
            float acc[NCH];
#pragma unroll
            for (int h = 0; h < NCH; ++h) acc[h] = 0.0f;
#pragma unroll
            for (int j = 0; j < TILE_PF; ++j)
                if ((unsigned)j < k) {
#pragma unroll
                    for (int h = 0; h < NCH; ++h) acc[h] += cur[j][h];
                }
            for (unsigned j = TILE_PF; j < k; j += 4) {           // longer cells: four rows in flight
                float r[4][NCH];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const bool ok = j + jj < k;
                    const float* row = q.c + (size_t)s_pts[o + (ok ? j + jj : 0)] * q.Cp;
#pragma unroll
                    for (int h = 0; h < NCH; ++h) r[jj][h] = (ok && h * 32 + lane < q.Cp) ? __ldg(row + h * 32 + lane) : 0.0f;
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int h = 0; h < NCH; ++h) acc[h] += r[jj][h];
            }
            emit(cell, acc, k);
        }
        // ---- hot cells: every warp adds a slice, warp 0 adds the slices in order -------------------------------------
        const int nhot = s_nhot;                                 // (written before the barriers above)
        for (int hi = 0; hi < nhot; ++hi) {
            const int cell = (int)s_hot[hi];
            const unsigned k = s_cnt[cell], o = s_off[cell];
            const unsigned sl = (k + TILE_WARPS - 1) / TILE_WARPS, j0 = warp * sl, j1 = min(j0 + sl, k);
            float acc[NCH];
#pragma unroll
            for (int h = 0; h < NCH; ++h) acc[h] = 0.0f;
            for (unsigned j = j0; j < j1; j += 8) {
                float r[8][NCH];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const bool ok = j + jj < j1;
                    const float* row = q.c + (size_t)s_pts[o + (ok ? j + jj : j0)] * q.Cp;
#pragma unroll
                    for (int h = 0; h < NCH; ++h) r[jj][h] = (ok && h * 32 + lane < q.Cp) ? __ldg(row + h * 32 + lane) : 0.0f;
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                    for (int h = 0; h < NCH; ++h) acc[h] += r[jj][h];
            }
#pragma unroll
            for (int h = 0; h < NCH; ++h) s_slice[warp * NCH * 32 + h * 32 + lane] = acc[h];
            __syncthreads();
            if (warp == 0) {
                float tot[NCH];
#pragma unroll
                for (int h = 0; h < NCH; ++h) {
                    tot[h] = 0.0f;
                    for (int ww = 0; ww < TILE_WARPS; ++ww) tot[h] += s_slice[ww * NCH * 32 + h * 32 + lane];
                }
                emit(cell, tot, k);
            }
            __syncthreads();
        }
        // ---- split tiles: the part that arrives last adds the parts in order ----------------------------------------
        if (parts > 1) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = atomicAdd(q.arrive + bin, 1u) == parts - 1;
            __syncthreads();
            if (s_last) {
                __threadfence();
                const float* first = q.partial + (size_t)(u - part) * psize;                     // the bin's units are consecutive
                for (int i = threadIdx.x; i < hgt * w * q.Cp; i += blockDim.x) {
                    const int ry = i / (w * q.Cp), rem = i - ry * (w * q.Cp), rx = rem / q.Cp, ch = rem - rx * q.Cp;
                    const int lc = ry * q.TS + rx;
                    float sum = 0.0f;
                    int k = 0;
                    for (unsigned pt = 0; pt < parts; ++pt) {
                        const float* pp = first + pt * psize;
                        sum += __ldcg(pp + (size_t)lc * q.Cp + ch);
                        k += __float_as_int(__ldcg(pp + (size_t)cells * q.Cp + lc));
                    }
                    pl[((size_t)(y0 + ry) * q.R + x0) * q.Cp + rem] = k > 1 ? __fdiv_rn(sum, (float)k) : sum;
                    if (ch == 0) cn[(size_t)(y0 + ry) * q.R + x0 + rx] = k;
                }
            }
        }
        __syncthreads();
    }
}

// bins nobody fell into: their tiles are zeros
__global__ void __launch_bounds__(256) tile_empty_kernel(const __grid_constant__ TileKP q) {
    const long long RR = (long long)q.R * q.R;
    for (int bin = blockIdx.x; bin < q.nbins; bin += gridDim.x) {
        if (q.start[bin + 1] != q.start[bin]) continue;
        unsigned t = bin;
        const int tx = t % q.nt; t /= q.nt;
        const int ty = t % q.nt; t /= q.nt;
        const int x0 = tx * q.TS, y0 = ty * q.TS;
        const int w = min(q.TS, q.R - x0), hgt = min(q.TS, q.R - y0);
        float* pl = q.planes + (size_t)t * RR * q.Cp;
        int* cn = q.count + (size_t)t * RR;
        for (int i = threadIdx.x; i < hgt * w * q.Cp; i += blockDim.x) {
            const int ry = i / (w * q.Cp), rem = i - ry * (w * q.Cp);
            pl[((size_t)(y0 + ry) * q.R + x0) * q.Cp + rem] = 0.0f;
        }
        for (int i = threadIdx.x; i < hgt * w; i += blockDim.x) cn[(size_t)(y0 + i / w) * q.R + x0 + i % w] = 0;
    }
}

struct TilePlan {
    TileKP kp;
    size_t bytes, o_hmat, o_binsize, o_start, o_ustart, o_work, o_arrive, o_units, o_sorted, o_partial;
    size_t smem_tile;
};

static bool plan_tiled(int B, long long N, int Cp, int R, TilePlan& pl) {
    if (B < 1 || N < 1 || Cp < 1 || Cp > 256 || R < 1 || R > 4096 || (long long)B * N >= 0x7fffffffLL) return false;
    TileKP& k = pl.kp;
    k.B = B, k.N = N, k.total = (long long)B * N, k.Cp = Cp, k.R = R;
    int ts = 1;
    while ((ts + 1) * (ts + 1) * Cp * 4 <= TILE_BYTES) ++ts;
    if (ts > R) ts = R;
    k.TS = ts, k.nt = (R + ts - 1) / ts;
    const long long nbins = 3LL * B * k.nt * k.nt;
    if (nbins > TILE_HIST_MAX) return false;
    k.nbins = (int)nbins;
    k.hstride = (BIN_MAXB + 31) / 32 * 32;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t max_units = (size_t)nbins + (size_t)(3 * k.total / TILE_UNIT_MAX) + 1;
    size_t o = 0;
    pl.o_hmat = o, o += up((size_t)nbins * k.hstride * 4);
    pl.o_binsize = o, o += up((size_t)nbins * 4);
    pl.o_start = o, o += up(((size_t)nbins + 1) * 4);
    pl.o_ustart = o, o += up(((size_t)nbins + 1) * 4);
    pl.o_work = o, o += 256;
    pl.o_arrive = o, o += up((size_t)nbins * 4);
    pl.o_units = o, o += up(max_units * 8);
    pl.o_sorted = o, o += up((size_t)3 * k.total * 8);
    pl.o_partial = o, o += up(max_units * (size_t)ts * ts * (Cp + 1) * 4);
    pl.bytes = o;
    const int nch = Cp <= 32 ? 1 : (Cp <= 64 ? 2 : (Cp <= 128 ? 4 : 8));
    pl.smem_tile = ((size_t)3 * ts * ts + 1 + TILE_UNIT_MAX + (size_t)TILE_WARPS * nch * 32 + TILE_UNIT_MAX / TILE_HOT + 1) * 4;
    return true;
}

}  // namespace gnb

using namespace gnb;

extern "C" int64_t gnb_scatter_tiled_scratch_bytes(int B, int64_t N, int Cp, int R) {
    TilePlan pl;
    if (!plan_tiled(B, N, Cp, R, pl)) return 0;
    return (int64_t)pl.bytes;
}

extern "C" int gnb_scatter_mean_planes_tiled(const float* p, const float* c, int B, int64_t N, int Cp, int R, double padding,
                                             float* planes, int32_t* count, void* scratch, int64_t scratch_bytes, void* stream) {
    GNB_CHECK_ARG(p && c && planes && count, "gnb_scatter_mean_planes_tiled: null pointer");
    TilePlan pl;
    GNB_CHECK_ARG(plan_tiled(B, N, Cp, R, pl), "gnb_scatter_mean_planes_tiled: shape not supported (use gnb_scatter_mean_planes)");
    GNB_CHECK_ARG(scratch && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && scratch_bytes >= (int64_t)pl.bytes,
                  "gnb_scatter_mean_planes_tiled: scratch too small (gnb_scatter_tiled_scratch_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    TileKP& k = pl.kp;
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    k.p = p, k.c = c, k.planes = planes, k.count = count;
    k.den = (float)(1.0 + padding + 10e-6);
    k.hmat = reinterpret_cast<unsigned*>(base + pl.o_hmat), k.binsize = reinterpret_cast<unsigned*>(base + pl.o_binsize);
    k.start = reinterpret_cast<unsigned*>(base + pl.o_start), k.ustart = reinterpret_cast<unsigned*>(base + pl.o_ustart);
    k.work = reinterpret_cast<unsigned*>(base + pl.o_work), k.arrive = reinterpret_cast<unsigned*>(base + pl.o_arrive);
    k.units = reinterpret_cast<uint2*>(base + pl.o_units), k.sorted = reinterpret_cast<uint2*>(base + pl.o_sorted);
    k.partial = reinterpret_cast<float*>(base + pl.o_partial);
    int dev = 0, sms = 148;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (k.total + 2047) / 2048;
    const long long max_blocks = 2LL * sms < BIN_MAXB ? 2LL * sms : BIN_MAXB;
    if (blocks > max_blocks) blocks = max_blocks;
    k.chunk = ((k.total + blocks - 1) / blocks + 1023) / 1024 * 1024;
    blocks = (k.total + k.chunk - 1) / k.chunk;
    const size_t hist_bytes = (size_t)k.nbins * 4;
    GNB_CUDA(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
    GNB_CUDA(cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
    tile_count_kernel<<<(unsigned)blocks, 1024, hist_bytes, st>>>(k);
    GNB_LAUNCH_CHECK();
    bin_reduce_kernel<<<(unsigned)((k.nbins + 31) / 32), 1024, 0, st>>>(k.hmat, k.binsize, k.nbins, (int)blocks, k.hstride);
    GNB_LAUNCH_CHECK();
    bin_scan_kernel<<<1, 1024, 0, st>>>(k.binsize, k.start, k.ustart, k.nbins, TILE_UNIT_MAX, k.work);
    GNB_LAUNCH_CHECK();
    tile_scatter_kernel<<<(unsigned)blocks, 1024, hist_bytes, st>>>(k);
    GNB_LAUNCH_CHECK();
    tile_empty_kernel<<<(unsigned)(k.nbins < 4 * sms ? k.nbins : 4 * sms), 256, 0, st>>>(k);
    GNB_LAUNCH_CHECK();
    const size_t max_units = (size_t)k.nbins + (size_t)(3 * k.total / TILE_UNIT_MAX) + 1;
    const unsigned grid = (unsigned)(max_units < (size_t)4 * sms ? max_units : (size_t)4 * sms);
#define GNB_TILE_ACCUM(NCH)                                                                                                  \
    do {                                                                                                                     \
        GNB_CUDA(cudaFuncSetAttribute(tile_accum_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_tile)); \
        tile_accum_kernel<NCH><<<grid, TILE_THREADS, pl.smem_tile, st>>>(k);                                                          \
    } while (0)
    if (Cp <= 32) GNB_TILE_ACCUM(1);
    else if (Cp <= 64) GNB_TILE_ACCUM(2);
    else if (Cp <= 128) GNB_TILE_ACCUM(4);
    else GNB_TILE_ACCUM(8);
#undef GNB_TILE_ACCUM
    GNB_LAUNCH_CHECK();
    return 0;
}

// Triplane projection for sm_100a: plane coordinates / cell indices, scatter-mean onto the three
// planes (atomic and deterministic modes) and local pooling.
//
// Replaces
//   normalize_coordinate(), coordinate2index()       reference src/models/utils.py:57-98
//   LocalPoolPointnet.generate_plane_features()      reference src/models/components/pointnet.py:72-89
//       (= torch_scatter.scatter_mean: scatter_add_ of features and of ones, clamp(count,1), divide)
//   LocalPoolPointnet.pool_local()                   reference src/models/components/pointnet.py:105-121
//
// Planes are written channels-last (R,R,C_p) so that the sampler reads one contiguous run per
// corner; the logical shape (B,C_p,R,R) is kept by the Python layer through strides.
#include <stdlib.h>

#include "common.cuh"

namespace gnb {

// plane k uses point coordinates (A0[k], A1[k]): xz -> (0,2), xy -> (0,1), yz -> (1,2)
__device__ __forceinline__ void plane_cells(float x, float y, float z, float den, int R, int cell[3], float* uu = nullptr) {
    float u[3] = {plane_unit(x, den), plane_unit(y, den), plane_unit(z, den)};
    // coordinate2index: (u * reso).long() truncates; index = x0 + reso * x1   (utils.py:67-69)
    int i[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) i[d] = (int)__fmul_rn(u[d], (float)R);
    cell[0] = i[0] + R * i[2];
    cell[1] = i[0] + R * i[1];
    cell[2] = i[1] + R * i[2];
    if (uu) { uu[0] = u[0], uu[1] = u[1], uu[2] = u[2]; }
}

__global__ void plane_coords_kernel(const float* __restrict__ p, long long n, float den, int R, float* __restrict__ coord,
                                    long long* __restrict__ index) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u[3];
    int cell[3];
    plane_cells(p[i * 3], p[i * 3 + 1], p[i * 3 + 2], den, R, cell, u);
    const int a0[3] = {0, 0, 1}, a1[3] = {2, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (coord) {
            coord[((long long)k * n + i) * 2 + 0] = u[a0[k]];
            coord[((long long)k * n + i) * 2 + 1] = u[a1[k]];
        }
        if (index) index[(long long)k * n + i] = cell[k];
    }
}

// ---------------------------------------------------------------------------------------
// Atomic mode.  One warp takes 32 consecutive points of one scene; lane i computes the three
// cell indices of point i.  The 32 feature rows are loaded ONCE into registers with lanes =
// channels (coalesced C_p*4-byte rows, all loads in flight together) and serve the three planes.
// Consecutive points that fall into the same cell are summed in registers first (run-length
// warp aggregation) and leave as ONE vector of reductions.  In the reference's real usage
// (metric coordinates, SURVEY trap T6) most points clamp into the last row / column and, above
// all, into the corner cell (R-1, R-1): that cell has its own register accumulator per warp,
// is combined across the block in shared memory and reaches global memory once per block.
// ---------------------------------------------------------------------------------------
template <int NCH>   // 32-channel groups = ceil(C_p / 32)
__global__ void __launch_bounds__(256) scatter_atomic_kernel(const float* __restrict__ p, const float* __restrict__ c,
                                                             int B, long long N, int Cp, int R, float den,
                                                             float* __restrict__ planes, int* __restrict__ count) {
    __shared__ float s_corner[NCH * 32];
    __shared__ int s_corner_n;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y, k = blockIdx.z;                   // scene, plane (the reductions of a warp are issued one after
                                                                // the other, so the three planes go to different warps)
    const long long group = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // 32-point group inside the scene
    const long long n0 = group * 32;
    const long long RR = (long long)R * R;
    const int corner = (R - 1) + R * (R - 1);
    for (int i = threadIdx.x; i < NCH * 32; i += blockDim.x) s_corner[i] = 0.0f;
    if (threadIdx.x == 0) s_corner_n = 0;
    __syncthreads();
    float* __restrict__ pl = planes + ((long long)k * B + b) * RR * Cp;
    int* __restrict__ cn = count + ((long long)k * B + b) * RR;
    if (n0 < N) {
        int cellk = -1;
        if (n0 + lane < N) {
            const float* pp = p + ((long long)b * N + n0 + lane) * 3;
            int cell[3];
            plane_cells(pp[0], pp[1], pp[2], den, R, cell);
            cellk = k == 0 ? cell[0] : (k == 1 ? cell[1] : cell[2]);
        }
        const int npts = (int)min((long long)32, N - n0);
        const float* __restrict__ cb = c + ((long long)b * N + n0) * Cp;
#pragma unroll 1
        for (int h = 0; h < NCH; ++h) {
            const int ch = h * 32 + lane;
            const bool has = ch < Cp;
            float cr[32];                                   // this lane's channel of the 32 points
#pragma unroll
            for (int j = 0; j < 32; ++j) cr[j] = (has && j < npts) ? __ldg(cb + (long long)j * Cp + ch) : 0.0f;
            int run_cell = -1, run_n = 0, corner_n = 0;
            float acc = 0.0f, corner_acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int cj = __shfl_sync(FULL, cellk, j);
                if (j < npts) {
                    if (cj == corner) {
                        corner_acc += cr[j], ++corner_n;
                    } else {
                        if (cj != run_cell) {
                            if (run_cell >= 0) {
                                if (has) atomicAdd(pl + (long long)run_cell * Cp + ch, acc);
                                if (lane == 0 && h == 0) atomicAdd(cn + run_cell, run_n);
                            }
                            run_cell = cj, run_n = 0, acc = 0.0f;
                        }
                        acc += cr[j], ++run_n;
                    }
                }
            }
            if (run_cell >= 0) {
                if (has) atomicAdd(pl + (long long)run_cell * Cp + ch, acc);
                if (lane == 0 && h == 0) atomicAdd(cn + run_cell, run_n);
            }
            if (corner_n) {
                atomicAdd(&s_corner[h * 32 + lane], corner_acc);
                if (lane == 0 && h == 0) atomicAdd(&s_corner_n, corner_n);
            }
        }
    }
    __syncthreads();
    if (s_corner_n > 0) {
        for (int ch = threadIdx.x; ch < Cp; ch += blockDim.x) atomicAdd(pl + (long long)corner * Cp + ch, s_corner[ch]);
        if (threadIdx.x == 0) atomicAdd(cn + corner, s_corner_n);
    }
}

// Vector variant for C_p in {4, 8, ..., 128} (a power of two times 4): LPP = C_p / 4 lanes share a point, each lane one
// float4 of channels, so a warp works on 32 / LPP points at once and every reduction is a 16-byte `red.global.add.v4.f32`
// (sm_90+): the L2 retires a quarter of the operations.  Slot g of the warp owns the LPP consecutive points
// [g * LPP, (g + 1) * LPP) of the warp's 32 and run-length-aggregates them like the scalar kernel; the corner cell is
// collected per lane and combined across the block.
// gridDim.z = 3: one plane per warp (few points: the reductions a warp issues one after the other are what takes time);
// gridDim.z = 1: a warp serves all three planes from the feature rows it loaded once (many points: the rows are not read
// three times through the L2 that is busy with the reductions).
template <int LPP>
__global__ void __launch_bounds__(256) scatter_atomic_v4_kernel(const float* __restrict__ p, const float* __restrict__ c,
                                                                int B, long long N, int R, float den,
                                                                float* __restrict__ planes, int* __restrict__ count) {
    constexpr int Cp = LPP * 4, PPS = LPP;                      // 32 / LPP slots of LPP points each
    __shared__ float s_corner[3][8][Cp];                        // per plane and warp
    __shared__ int s_corner_n[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane / LPP, sub = lane % LPP;
    const int b = blockIdx.y;
    const int k0 = gridDim.z == 3 ? (int)blockIdx.z : 0, k1 = gridDim.z == 3 ? k0 + 1 : 3;
    const long long group = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long n0 = group * 32;
    const long long RR = (long long)R * R;
    const int corner = (R - 1) + R * (R - 1);
    int cell[3] = {-1, -1, -1};
    const int npts = n0 < N ? (int)min((long long)32, N - n0) : 0;
    float4 cr[PPS];                                             // this lane's float4 of the slot's points
    if (npts > 0) {
        if (n0 + lane < N) {
            const float* pp = p + ((long long)b * N + n0 + lane) * 3;
            plane_cells(pp[0], pp[1], pp[2], den, R, cell);
        }
        const float* __restrict__ cb = c + ((long long)b * N + n0) * Cp;
#pragma unroll
        for (int j = 0; j < PPS; ++j) {
            const int pt = slot * PPS + j;
            cr[j] = pt < npts ? ldg4(cb + (long long)pt * Cp + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < k0 || k >= k1) continue;
        float* __restrict__ pl = planes + ((long long)k * B + b) * RR * Cp;
        int* __restrict__ cn = count + ((long long)k * B + b) * RR;
        float4 corner_acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int corner_n = 0;
        if (npts > 0) {
            int run_cell = -1, run_n = 0;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            auto flush = [&]() {
                if (run_cell >= 0) {
                    atomicAdd(reinterpret_cast<float4*>(pl + (long long)run_cell * Cp + sub * 4), acc);
                    if (sub == 0) atomicAdd(cn + run_cell, run_n);
                }
            };
#pragma unroll
            for (int j = 0; j < PPS; ++j) {
                const int pt = slot * PPS + j;
                const int cj = __shfl_sync(FULL, cell[k], pt);
                if (pt < npts) {
                    if (cj == corner) {
                        corner_acc.x += cr[j].x, corner_acc.y += cr[j].y, corner_acc.z += cr[j].z, corner_acc.w += cr[j].w;
                        ++corner_n;
                    } else {
                        if (cj != run_cell) {
                            flush();
                            run_cell = cj, run_n = 0, acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        acc.x += cr[j].x, acc.y += cr[j].y, acc.z += cr[j].z, acc.w += cr[j].w;
                        ++run_n;
                    }
                }
            }
            flush();
        }
        // corner cell: slots of the warp by shuffles, warps of the block through shared memory, one reduction per block
#pragma unroll
        for (int o = LPP; o < 32; o <<= 1) {
            corner_acc.x += __shfl_xor_sync(FULL, corner_acc.x, o), corner_acc.y += __shfl_xor_sync(FULL, corner_acc.y, o);
            corner_acc.z += __shfl_xor_sync(FULL, corner_acc.z, o), corner_acc.w += __shfl_xor_sync(FULL, corner_acc.w, o);
            corner_n += __shfl_xor_sync(FULL, corner_n, o);
        }
        if (slot == 0) *reinterpret_cast<float4*>(&s_corner[k][warp][sub * 4]) = corner_acc;
        if (lane == 0) s_corner_n[k][warp] = corner_n;
    }
    __syncthreads();
    for (int k = k0; k < k1; ++k) {
        if (threadIdx.x < Cp) {
            float t = 0.0f;
            int nn = 0;
            for (int w = 0; w < 8; ++w) t += s_corner[k][w][threadIdx.x], nn += s_corner_n[k][w];
            if (nn > 0) {
                atomicAdd(planes + (((long long)k * B + b) * RR + corner) * Cp + threadIdx.x, t);
                if (threadIdx.x == 0) atomicAdd(count + ((long long)k * B + b) * RR + corner, nn);
            }
        }
    }
}

// mean = sum / max(count, 1) (torch_scatter.scatter_mean), in place over (3*B*R*R, C_p).  Cells with at most one point are
// left alone (not even read).  VEC = 4: one thread per float4 of a cell, lanes_per_cell = C_p / 4.
template <int VEC>
__global__ void __launch_bounds__(256) scatter_finalize_kernel(float* __restrict__ planes, const int* __restrict__ count, long long cells,
                                                               int Cp, int lanes_per_cell, int lg_lanes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells * lanes_per_cell) return;
    const long long cell = lg_lanes >= 0 ? (i >> lg_lanes) : i / lanes_per_cell;
    const int n = __ldg(count + cell);
    if (n <= 1) return;
    const float d = (float)n;
    if constexpr (VEC == 4) {
        float4* q = reinterpret_cast<float4*>(planes) + i;
        float4 v = *q;
        v.x = __fdiv_rn(v.x, d), v.y = __fdiv_rn(v.y, d), v.z = __fdiv_rn(v.z, d), v.w = __fdiv_rn(v.w, d);
        *q = v;
    } else {
        planes[i] = __fdiv_rn(planes[i], d);
    }
}

static int launch_finalize(float* planes, const int* count, long long cells, int Cp, cudaStream_t st) {
    const bool vec = Cp % 4 == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0;
    const int lanes = vec ? Cp / 4 : Cp;
    int lg = 0;
    while ((1 << lg) < lanes) ++lg;
    if ((1 << lg) != lanes) lg = -1;
    const unsigned blocks = (unsigned)((cells * lanes + 255) / 256);
    if (vec) scatter_finalize_kernel<4><<<blocks, 256, 0, st>>>(planes, count, cells, Cp, lanes, lg);
    else scatter_finalize_kernel<1><<<blocks, 256, 0, st>>>(planes, count, cells, Cp, lanes, lg);
    GNB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------
// Deterministic mode: stable LSD radix sort of (cell, point) per (plane, scene) segment, then
// one warp per cell sums its points in ascending point index -- the summation order of the CPU
// scatter_add_, hence bit-identical sums; counts are integers and exact.
// ---------------------------------------------------------------------------------------
constexpr int SORT_TILE = 2048;    // items per block and radix pass (256 threads x 8)

__global__ void det_keys_kernel(const float* __restrict__ p, int B, long long N, int R, float den,
                                unsigned* __restrict__ keys, unsigned* __restrict__ vals, unsigned* __restrict__ long_count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *long_count = 0u;
    if (i >= (long long)B * N) return;
    int b = (int)(i / N);
    long long n = i % N;
    int cell[3];
    plane_cells(p[i * 3], p[i * 3 + 1], p[i * 3 + 2], den, R, cell);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        long long seg = (long long)k * B + b;
        keys[seg * N + n] = (unsigned)cell[k];
        vals[seg * N + n] = (unsigned)n;
    }
}

// hist[seg][digit][tile]
__global__ void __launch_bounds__(256) radix_hist_kernel(const unsigned* __restrict__ keys, long long N, int shift, int tiles,
                                                         unsigned* __restrict__ hist) {
    __shared__ unsigned h[256];
    const int seg = blockIdx.y, tile = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned* k = keys + (long long)seg * N;
    long long i0 = (long long)tile * SORT_TILE;
    for (int j = threadIdx.x; j < SORT_TILE; j += 256) {
        long long i = i0 + j;
        if (i < N) atomicAdd(&h[(k[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[((long long)seg * 256 + threadIdx.x) * tiles + tile] = h[threadIdx.x];
}

// exclusive scan of each segment's 256*tiles counters (digit-major), one block per segment
__global__ void __launch_bounds__(1024) radix_scan_kernel(unsigned* __restrict__ hist, int tiles) {
    __shared__ unsigned part[1024];
    unsigned* h = hist + (long long)blockIdx.x * 256 * tiles;
    const int E = 256 * tiles;
    const int per = (E + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(E, lo + per);
    unsigned s = 0;
    for (int i = lo; i < hi; ++i) s += h[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {       // Hillis-Steele inclusive scan
        unsigned v = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        unsigned v = h[i];
        h[i] = run;
        run += v;
    }
}

// stable scatter of one tile: rank inside the tile = (items of the same digit held by earlier
// warps) + (same digit earlier in this warp's contiguous 256-item run, via __match_any_sync)
__global__ void __launch_bounds__(256) radix_scatter_kernel(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                            unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                            long long N, int shift, int tiles, const unsigned* __restrict__ hist) {
    __shared__ unsigned wh[8][256];
    const int seg = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&wh[0][0])[i] = 0;
    __syncthreads();
    const long long base = (long long)seg * N;
    const long long i0 = (long long)tile * SORT_TILE + w * 256;     // this warp's contiguous run
    unsigned key[8], val[8], rank[8];
    bool ok[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        long long i = i0 + r * 32 + lane;
        ok[r] = i < N;
        key[r] = ok[r] ? keys_in[base + i] : 0xffffffffu;
        val[r] = ok[r] ? vals_in[base + i] : 0u;
        unsigned d = ok[r] ? ((key[r] >> shift) & 255u) : 256u + lane;    // inactive lanes match nobody
        unsigned m = __match_any_sync(FULL, d);
        unsigned before = __popc(m & ((1u << lane) - 1u));
        unsigned prior = ok[r] ? wh[w][d & 255u] : 0u;
        __syncwarp();
        if (ok[r] && before == 0) wh[w][d] = prior + __popc(m);            // group leader updates
        __syncwarp();
        rank[r] = prior + before;
    }
    __syncthreads();
    // exclusive prefix over the 8 warps, per digit
    {
        unsigned run = 0;
        const int d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) {
            unsigned v = wh[ww][d];
            wh[ww][d] = run;
            run += v;
        }
    }
    __syncthreads();
    const unsigned* hs = hist + (long long)seg * 256 * tiles;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (!ok[r]) continue;
        unsigned d = (key[r] >> shift) & 255u;
        long long dst = base + hs[(long long)d * tiles + tile] + wh[w][d] + rank[r];
        keys_out[dst] = key[r];
        vals_out[dst] = val[r];
    }
}

// start[cell] = first sorted position of the cell, count[cell] = one past its last (per segment; count is zero-filled before,
// so it stays 0 for cells without points).  det_reduce turns count into the number of points.  No atomics: the counts are
// exact by construction.
__global__ void det_starts_kernel(const unsigned* __restrict__ keys, long long N, long long RR, unsigned* __restrict__ start,
                                  int* __restrict__ count) {
    const int seg = blockIdx.y;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned* k = keys + (long long)seg * N;
    const unsigned ki = k[i];
    if (i == 0 || k[i - 1] != ki) start[(long long)seg * RR + ki] = (unsigned)i;
    if (i == N - 1 || k[i + 1] != ki) count[(long long)seg * RR + ki] = (int)(i + 1);
}

constexpr int DET_LONG = 1024;     // cells with at least this many points go to det_reduce_long_kernel

// one warp per cell: sequential fp32 sum over the cell's points in ascending point index
__global__ void __launch_bounds__(256) det_reduce_kernel(const float* __restrict__ c, const unsigned* __restrict__ vals,
                                                         const unsigned* __restrict__ start, int* __restrict__ count,
                                                         int B, long long N, int Cp, long long RR, float* __restrict__ planes,
                                                         unsigned* __restrict__ long_list, unsigned* __restrict__ long_count) {
    const int lane = threadIdx.x & 31;
    const long long cell_g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // over 3*B*RR
    if (cell_g >= 3LL * B * RR) return;
    const long long seg = cell_g / RR;
    const int b = (int)(seg % B);
    const int end = count[cell_g];
    const int n = end > 0 ? end - (int)start[cell_g] : 0;
    __syncwarp();
    if (lane == 0 && end > 0) count[cell_g] = n;
    if (n >= DET_LONG) {
        if (lane == 0) long_list[atomicAdd(long_count, 1u)] = (unsigned)cell_g;
        return;
    }
    const unsigned* v = vals + seg * N + (n > 0 ? start[cell_g] : 0u);
    const float* __restrict__ cb = c + (long long)b * N * Cp;
    for (int ch = lane; ch < Cp; ch += 32) {
        float acc = 0.0f;
        int i = 0;
        for (; i + 8 <= n; i += 8) {                        // 8 gathers in flight, added in ascending point order
            float r[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) r[u] = __ldg(cb + (long long)__ldg(v + i + u) * Cp + ch);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, r[u]);
        }
        for (; i < n; ++i) acc = __fadd_rn(acc, __ldg(cb + (long long)v[i] * Cp + ch));
        if (n > 1) acc = __fdiv_rn(acc, (float)n);
        planes[cell_g * Cp + ch] = acc;
    }
}

// Cells with very many points (the clamped border cells of SURVEY trap T6): the sum stays ONE sequential fp32 chain per
// channel (that is what makes it bit-identical to the CPU scatter_add_), but the gathers are taken off the chain: all 8
// warps of a block stage the next batch of feature rows in shared memory (double buffered) while warp 0 adds the current one.
constexpr int DET_LONG_SMEM = 64 * 1024;           // bytes per batch buffer
template <int NH>   // 32-channel groups = ceil(C_p / 32)
__global__ void __launch_bounds__(256) det_reduce_long_kernel(const float* __restrict__ c, const unsigned* __restrict__ vals,
                                                              const unsigned* __restrict__ start, const int* __restrict__ count,
                                                              int B, long long N, int Cp, long long RR, float* __restrict__ planes,
                                                              const unsigned* __restrict__ long_list, const unsigned* __restrict__ long_count) {
    extern __shared__ __align__(16) float s_rows[];             // [2][rows_per_batch][Cp]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = DET_LONG_SMEM / (Cp * 4);
    const unsigned n_long = *long_count;
    for (unsigned e = blockIdx.x; e < n_long; e += gridDim.x) {
        const long long cell_g = long_list[e];
        const long long seg = cell_g / RR;
        const int b = (int)(seg % B), n = count[cell_g];
        const unsigned* v = vals + seg * N + start[cell_g];
        const float* __restrict__ cb = c + (long long)b * N * Cp;
        const int batches = (n + rows - 1) / rows;
        float acc[NH];                                         // warp 0: channel lane + 32*h
#pragma unroll
        for (int h = 0; h < NH; ++h) acc[h] = 0.0f;
        auto stage = [&](int bt, int tid, int nthr) {                             // every thread: 8 gathers in flight, coalesced along the channels
            float* dst = s_rows + (size_t)(bt & 1) * rows * Cp;
            const int r0 = bt * rows, nr = min(rows, n - r0);
            if ((Cp & 3) == 0) {
                const int c4 = Cp >> 2, total = nr * c4;
                for (int base = 0; base < total; base += 8 * nthr) {
                    float4 t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = base + u * nthr + tid;
                        if (idx < total) {
                            const int r = idx / c4, q = idx - r * c4;
                            t[u] = ldg4(cb + (long long)__ldg(v + r0 + r) * Cp + q * 4);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = base + u * nthr + tid;
                        if (idx < total) reinterpret_cast<float4*>(dst)[idx] = t[u];
                    }
                }
            } else {
                for (int idx = tid; idx < nr * Cp; idx += nthr) {
                    const int r = idx / Cp, ch = idx - r * Cp;
                    dst[idx] = __ldg(cb + (long long)__ldg(v + r0 + r) * Cp + ch);
                }
            }
        };
        stage(0, threadIdx.x, blockDim.x);
        __syncthreads();
        for (int bt = 0; bt < batches; ++bt) {
            if (warp != 0) {
                if (bt + 1 < batches) stage(bt + 1, threadIdx.x - 32, blockDim.x - 32);      // warp 0 is adding
            } else {
                const float* src = s_rows + (size_t)(bt & 1) * rows * Cp;
                const int nr = min(rows, n - bt * rows);
#pragma unroll 8
                for (int r = 0; r < nr; ++r) {
#pragma unroll
                    for (int h = 0; h < NH; ++h)
                        if (NH == 1 || h * 32 + lane < Cp) acc[h] = __fadd_rn(acc[h], src[r * Cp + h * 32 + (NH == 1 ? min(lane, Cp - 1) : lane)]);
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
#pragma unroll
            for (int h = 0; h < NH; ++h)
                if (h * 32 + lane < Cp) planes[cell_g * Cp + h * 32 + lane] = __fdiv_rn(acc[h], (float)n);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// pool_local: only the cells touched by points are initialised, reduced and read back.
// max: order-preserving integer encoding + atomicMax (exact, order independent).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_max(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_max(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// phase 0: init touched cells; 1: reduce; 2: gather + sum over planes (xz, xy, yz order)
template <int PHASE>
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ p, const float* __restrict__ c, int B, long long N,
                                                   int Hd, int R, float den, int pool_type, unsigned* __restrict__ cellbuf,
                                                   int* __restrict__ cnt, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // over B*N
    if (pt >= (long long)B * N) return;
    const int b = (int)(pt / N);
    const long long RR = (long long)R * R;
    int cell[3];
    plane_cells(__ldg(p + pt * 3), __ldg(p + pt * 3 + 1), __ldg(p + pt * 3 + 2), den, R, cell);
    for (int h = lane; h < Hd; h += 32) {
        float o = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            long long cg = ((long long)k * B + b) * RR + cell[k];
            unsigned* slot = cellbuf + cg * Hd + h;
            if (PHASE == 0) {
                *slot = (pool_type == GNB_POOL_MAX) ? enc_max(-INFINITY) : 0u;
                if (h == 0) cnt[cg] = 0;
            } else if (PHASE == 1) {
                float v = __ldg(c + pt * Hd + h);
                if (pool_type == GNB_POOL_MAX) atomicMax(slot, enc_max(v));
                else atomicAdd(reinterpret_cast<float*>(slot), v);
                if (h == 0 && pool_type != GNB_POOL_MAX) atomicAdd(cnt + cg, 1);
            } else {
                float v = (pool_type == GNB_POOL_MAX) ? dec_max(*slot) : __fdiv_rn(__uint_as_float(*slot), (float)cnt[cg]);
                o = __fadd_rn(o, v);
            }
        }
        if (PHASE == 2) out[pt * Hd + h] = o;
    }
}


// ---------------------------------------------------------------------------------------
// Backward passes.
// scatter_mean (pointnet.py:82): grad_c[b,n,:] = sum over planes of grad_plane[cell(n),:] / max(count,1).
// pool_local  (pointnet.py:113-119): out[n] = sum_k fea_k[cell_k(n)], fea_k = scatter_(max|mean)(c):
//   gsum_k[cell] = sum of grad_out over the points of the cell (backward of the gather);
//   mean: grad_c[n] = sum_k gsum_k[cell_k(n)] / count;   max: gsum goes to the ONE point that attains
//   the maximum (torch_scatter's arg); exact ties are broken towards the smallest point index.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_mean_bwd_kernel(const float* __restrict__ p, const float* __restrict__ gplanes,
                                                               const int* __restrict__ count, int B, long long N, int Cp, int R,
                                                               float den, long long psb_unused, float* __restrict__ gc) {
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)B * N) return;
    const int b = (int)(pt / N);
    const long long RR = (long long)R * R;
    int cell[3];
    plane_cells(__ldg(p + pt * 3), __ldg(p + pt * 3 + 1), __ldg(p + pt * 3 + 2), den, R, cell);
    for (int ch = lane; ch < Cp; ch += 32) {
        float g = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const long long cg = ((long long)k * B + b) * RR + cell[k];
            const int n = count[cg];
            g += __ldg(gplanes + cg * Cp + ch) / (float)(n > 0 ? n : 1);
        }
        gc[pt * Cp + ch] = g;
    }
}

// phases: 0 init touched cells (gsum = 0, arg = INT_MAX); 1 gsum += grad_out and (max) arg = min index
// among the points whose value equals the cell maximum; 2 route to the points
template <int PHASE>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ p, const float* __restrict__ c,
                                                       const float* __restrict__ gout, int B, long long N, int Hd, int R, float den,
                                                       int pool_type, const unsigned* __restrict__ cellmax, const int* __restrict__ cnt,
                                                       float* __restrict__ gsum, int* __restrict__ arg, float* __restrict__ gc) {
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)B * N) return;
    const int b = (int)(pt / N);
    const long long RR = (long long)R * R;
    int cell[3];
    plane_cells(__ldg(p + pt * 3), __ldg(p + pt * 3 + 1), __ldg(p + pt * 3 + 2), den, R, cell);
    for (int h = lane; h < Hd; h += 32) {
        float o = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const long long cg = ((long long)k * B + b) * RR + cell[k];
            const long long slot = cg * Hd + h;
            if (PHASE == 0) {
                gsum[slot] = 0.0f;
                if (pool_type == GNB_POOL_MAX) arg[slot] = 0x7fffffff;
            } else if (PHASE == 1) {
                atomicAdd(gsum + slot, __ldg(gout + pt * Hd + h));
                if (pool_type == GNB_POOL_MAX && enc_max(__ldg(c + pt * Hd + h)) == cellmax[slot]) atomicMin(arg + slot, (int)(pt % N));
            } else {
                if (pool_type == GNB_POOL_MAX) o += (arg[slot] == (int)(pt % N)) ? gsum[slot] : 0.0f;
                else o += gsum[slot] / (float)cnt[cg];
            }
        }
        if (PHASE == 2) gc[pt * Hd + h] = o;
    }
}

static int radix_bits(long long RR) {
    int bits = 1;
    while ((1LL << bits) < RR) ++bits;
    return bits;
}

struct DetScratch {
    unsigned *keys[2], *vals[2], *hist, *start, *long_list, *long_count;
    long long bytes;
};

static DetScratch det_layout(void* base, int B, long long N, int R) {
    DetScratch s;
    const long long segs = 3LL * B, items = segs * N, RR = (long long)R * R;
    const long long tiles = (N + SORT_TILE - 1) / SORT_TILE;
    char* q = (char*)base;
    auto take = [&](long long n) { char* r = q; q += (n * 4 + 255) / 256 * 256; return (unsigned*)r; };
    s.keys[0] = take(items), s.keys[1] = take(items);
    s.vals[0] = take(items), s.vals[1] = take(items);
    s.hist = take(segs * 256 * (tiles > 0 ? tiles : 1));
    s.start = take(segs * RR);
    s.long_list = take(items / DET_LONG + 1);
    s.long_count = take(1);
    s.bytes = q - (char*)base;
    return s;
}

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_plane_coords(const float* p, int64_t n, double padding, int R, float* coord, int64_t* index, void* stream) {
    GNB_CHECK_ARG((p || n == 0) && n >= 0 && R > 0 && (coord || index || n == 0), "gnb_plane_coords: bad arguments");
    if (n == 0) return 0;
    float den = (float)(1.0 + padding + 10e-6);
    plane_coords_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, n, den, R, coord, (long long*)index);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t gnb_scatter_scratch_bytes(int B, int64_t N, int R, int mode) {
    if (mode != GNB_SCATTER_DETERMINISTIC) return 0;
    return det_layout(nullptr, B, N, R).bytes;
}

extern "C" int gnb_scatter_mean_planes(const float* p, const float* c, int B, int64_t N, int Cp, int R, double padding,
                                       int mode, float* planes, int32_t* count, void* scratch, int64_t scratch_bytes,
                                       void* stream) {
    GNB_CHECK_ARG(((p && c) || N == 0) && planes && count, "gnb_scatter_mean_planes: null pointer");
    GNB_CHECK_ARG(B >= 1 && N >= 0 && Cp >= 1 && R >= 1 && R <= 4096, "gnb_scatter_mean_planes: bad shape");
    GNB_CHECK_ARG((long long)B * N < 0x7fffffffLL, "gnb_scatter_mean_planes: too many points");
    cudaStream_t st = (cudaStream_t)stream;
    const float den = (float)(1.0 + padding + 10e-6);
    const long long RR = (long long)R * R, cells = 3LL * B * RR;
    GNB_CUDA(cudaMemsetAsync(count, 0, cells * sizeof(int), st));
    if (mode == GNB_SCATTER_ATOMIC || mode == GNB_SCATTER_ATOMIC_SUM) {
        GNB_CHECK_ARG(Cp <= 256, "gnb_scatter_mean_planes: C_p %d > 256 not supported", Cp);
        GNB_CUDA(cudaMemsetAsync(planes, 0, cells * Cp * sizeof(float), st));
        if (N > 0) {
            const dim3 blocks((unsigned)(((N + 31) / 32 + 7) / 8), (unsigned)B, 3);       // 8 groups of 32 points per block; grid rows = scenes, layers = planes
            const dim3 blocks4(blocks.x, blocks.y, N >= 32768 ? 1 : 3);                      // v4 kernels: see the kernel comment
            const bool v4 = (reinterpret_cast<uintptr_t>(c) & 15) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0 && !opt(OPT_SCATTER_SCALAR);
            if (v4 && Cp == 4) scatter_atomic_v4_kernel<1><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (v4 && Cp == 8) scatter_atomic_v4_kernel<2><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (v4 && Cp == 16) scatter_atomic_v4_kernel<4><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (v4 && Cp == 32) scatter_atomic_v4_kernel<8><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (v4 && Cp == 64) scatter_atomic_v4_kernel<16><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (v4 && Cp == 128) scatter_atomic_v4_kernel<32><<<blocks4, 256, 0, st>>>(p, c, B, N, R, den, planes, count);
            else if (Cp <= 32) scatter_atomic_kernel<1><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else if (Cp <= 64) scatter_atomic_kernel<2><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else if (Cp <= 128) scatter_atomic_kernel<4><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else scatter_atomic_kernel<8><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            GNB_LAUNCH_CHECK();
            if (mode == GNB_SCATTER_ATOMIC) {
                int rc = launch_finalize(planes, count, cells, Cp, st);
                if (rc) return rc;
            }
        }
        return 0;
    }
    GNB_CHECK_ARG(mode == GNB_SCATTER_DETERMINISTIC, "gnb_scatter_mean_planes: unknown mode %d", mode);
    DetScratch s = det_layout(scratch, B, N, R);
    GNB_CHECK_ARG(scratch && scratch_bytes >= s.bytes, "gnb_scatter_mean_planes: scratch too small (%lld < %lld)",
                  (long long)scratch_bytes, s.bytes);
    const int segs = 3 * B;
    int cur = 0;
    if (N > 0) {
        det_keys_kernel<<<ceil_div((long long)B * N, 256), 256, 0, st>>>(p, B, N, R, den, s.keys[0], s.vals[0], s.long_count);
        GNB_LAUNCH_CHECK();
        const int tiles = (int)((N + SORT_TILE - 1) / SORT_TILE);
        const int bits = radix_bits(RR);
        for (int shift = 0; shift < bits; shift += 8) {
            radix_hist_kernel<<<dim3(tiles, segs), 256, 0, st>>>(s.keys[cur], N, shift, tiles, s.hist);
            GNB_LAUNCH_CHECK();
            radix_scan_kernel<<<segs, 1024, 0, st>>>(s.hist, tiles);
            GNB_LAUNCH_CHECK();
            radix_scatter_kernel<<<dim3(tiles, segs), 256, 0, st>>>(s.keys[cur], s.vals[cur], s.keys[cur ^ 1], s.vals[cur ^ 1], N,
                                                                     shift, tiles, s.hist);
            GNB_LAUNCH_CHECK();
            cur ^= 1;
        }
        det_starts_kernel<<<dim3(ceil_div(N, 256), segs), 256, 0, st>>>(s.keys[cur], N, RR, s.start, count);
        GNB_LAUNCH_CHECK();
    }
    if (N == 0) GNB_CUDA(cudaMemsetAsync(s.long_count, 0, 4, st));
    det_reduce_kernel<<<ceil_div(cells, 8), 256, 0, st>>>(c, s.vals[cur], s.start, count, B, N, Cp, RR, planes, s.long_list, s.long_count);
    GNB_LAUNCH_CHECK();
    if (N >= DET_LONG) {
        GNB_CHECK_ARG(Cp <= 256, "gnb_scatter_mean_planes: C_p %d > 256 not supported", Cp);
        const long long cap = (long long)segs * N / DET_LONG + 1;
        const unsigned grid = (unsigned)(cap < 296 ? cap : 296);
#define GNB_DET_LONG(NH)                                                                                                               \
    do {                                                                                                                               \
        GNB_CUDA(cudaFuncSetAttribute(det_reduce_long_kernel<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * DET_LONG_SMEM));    \
        det_reduce_long_kernel<NH><<<grid, 256, 2 * DET_LONG_SMEM, st>>>(c, s.vals[cur], s.start, count, B, N, Cp, RR, planes,         \
                                                                         s.long_list, s.long_count);                                   \
    } while (0)
        if (Cp <= 32) GNB_DET_LONG(1);
        else if (Cp <= 64) GNB_DET_LONG(2);
        else if (Cp <= 128) GNB_DET_LONG(4);
        else GNB_DET_LONG(8);
#undef GNB_DET_LONG
        GNB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int gnb_scatter_finalize(float* planes, const int32_t* count, int64_t n_cells, int Cp, void* stream) {
    GNB_CHECK_ARG(planes && count && n_cells >= 0 && Cp >= 1, "gnb_scatter_finalize: bad arguments");
    if (n_cells == 0) return 0;
    return launch_finalize(planes, count, n_cells, Cp, (cudaStream_t)stream);
}

extern "C" int64_t gnb_pool_scratch_bytes(int B, int64_t N, int Hd, int R) {
    (void)N;
    const long long cells = 3LL * B * R * R;
    return (cells * Hd * 4 + 255) / 256 * 256 + cells * 4;
}

extern "C" int gnb_pool_local(const float* p, const float* c, int B, int64_t N, int Hd, int R, double padding, int pool_type,
                              float* out, void* scratch, int64_t scratch_bytes, void* stream) {
    GNB_CHECK_ARG((p && c && out) || N == 0, "gnb_pool_local: null pointer");
    GNB_CHECK_ARG(B >= 1 && N >= 0 && Hd >= 1 && R >= 1, "gnb_pool_local: bad shape");
    GNB_CHECK_ARG(pool_type == GNB_POOL_MAX || pool_type == GNB_POOL_MEAN, "gnb_pool_local: unknown pool type %d", pool_type);
    GNB_CHECK_ARG(scratch && scratch_bytes >= gnb_pool_scratch_bytes(B, N, Hd, R), "gnb_pool_local: scratch too small");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float den = (float)(1.0 + padding + 10e-6);
    const long long cells = 3LL * B * R * R;
    unsigned* cellbuf = (unsigned*)scratch;
    int* cnt = (int*)((char*)scratch + (cells * Hd * 4 + 255) / 256 * 256);
    const unsigned blocks = (unsigned)(((long long)B * N + 7) / 8);
    pool_kernel<0><<<blocks, 256, 0, st>>>(p, c, B, N, Hd, R, den, pool_type, cellbuf, cnt, out);
    GNB_LAUNCH_CHECK();
    pool_kernel<1><<<blocks, 256, 0, st>>>(p, c, B, N, Hd, R, den, pool_type, cellbuf, cnt, out);
    GNB_LAUNCH_CHECK();
    pool_kernel<2><<<blocks, 256, 0, st>>>(p, c, B, N, Hd, R, den, pool_type, cellbuf, cnt, out);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_scatter_mean_planes_bwd(const float* p, const float* grad_planes, const int32_t* count, int B, int64_t N,
                                           int Cp, int R, double padding, float* grad_c, void* stream) {
    GNB_CHECK_ARG(p && grad_planes && count && grad_c, "gnb_scatter_mean_planes_bwd: null pointer");
    GNB_CHECK_ARG(B >= 1 && N >= 0 && Cp >= 1 && R >= 1, "gnb_scatter_mean_planes_bwd: bad shape");
    if (N == 0) return 0;
    const float den = (float)(1.0 + padding + 10e-6);
    scatter_mean_bwd_kernel<<<(unsigned)(((long long)B * N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p, grad_planes, count, B, N, Cp,
                                                                                                      R, den, 0, grad_c);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t gnb_pool_bwd_scratch_bytes(int B, int64_t N, int Hd, int R) {
    (void)N;
    const long long cells = 3LL * B * R * R;
    return 2 * ((cells * Hd * 4 + 255) / 256 * 256);
}

extern "C" int gnb_pool_local_bwd(const float* p, const float* c, const float* grad_out, int B, int64_t N, int Hd, int R,
                                  double padding, int pool_type, const void* fwd_scratch, float* grad_c, void* scratch,
                                  int64_t scratch_bytes, void* stream) {
    GNB_CHECK_ARG(p && c && grad_out && grad_c && fwd_scratch, "gnb_pool_local_bwd: null pointer");
    GNB_CHECK_ARG(B >= 1 && N >= 0 && Hd >= 1 && R >= 1, "gnb_pool_local_bwd: bad shape");
    GNB_CHECK_ARG(pool_type == GNB_POOL_MAX || pool_type == GNB_POOL_MEAN, "gnb_pool_local_bwd: unknown pool type %d", pool_type);
    GNB_CHECK_ARG(scratch && scratch_bytes >= gnb_pool_bwd_scratch_bytes(B, N, Hd, R), "gnb_pool_local_bwd: scratch too small");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float den = (float)(1.0 + padding + 10e-6);
    const long long cells = 3LL * B * R * R;
    const long long half = (cells * Hd * 4 + 255) / 256 * 256;
    const unsigned* cellmax = (const unsigned*)fwd_scratch;                          // layout of gnb_pool_local's scratch
    const int* cnt = (const int*)((const char*)fwd_scratch + half);
    float* gsum = (float*)scratch;
    int* arg = (int*)((char*)scratch + half);
    const unsigned blocks = (unsigned)(((long long)B * N + 7) / 8);
    pool_bwd_kernel<0><<<blocks, 256, 0, st>>>(p, c, grad_out, B, N, Hd, R, den, pool_type, cellmax, cnt, gsum, arg, grad_c);
    GNB_LAUNCH_CHECK();
    pool_bwd_kernel<1><<<blocks, 256, 0, st>>>(p, c, grad_out, B, N, Hd, R, den, pool_type, cellmax, cnt, gsum, arg, grad_c);
    GNB_LAUNCH_CHECK();
    pool_bwd_kernel<2><<<blocks, 256, 0, st>>>(p, c, grad_out, B, N, Hd, R, den, pool_type, cellmax, cnt, gsum, arg, grad_c);
    GNB_LAUNCH_CHECK();
    return 0;
}

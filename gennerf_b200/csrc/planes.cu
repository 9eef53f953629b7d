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
// cell indices of point i.  The features of the 32 points are then streamed with lanes =
// channels (coalesced C_p*4-byte rows).  Consecutive points that fall into the same cell are
// summed in registers first (run-length warp aggregation) and leave as ONE vector of
// reductions: in the reference's real usage (metric coordinates, SURVEY trap T6) most points
// clamp into the last row/column, and this removes the hot-address serialisation.
// ---------------------------------------------------------------------------------------
template <int NCH>   // channels per lane = ceil(C_p / 32)
__global__ void __launch_bounds__(256) scatter_atomic_kernel(const float* __restrict__ p, const float* __restrict__ c,
                                                             int B, long long N, int Cp, int R, float den,
                                                             float* __restrict__ planes, int* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long warps_per_scene = (N + 31) / 32;
    if (warp >= warps_per_scene * B) return;
    const int b = (int)(warp / warps_per_scene);
    const long long n0 = (warp % warps_per_scene) * 32;
    const long long RR = (long long)R * R;
    int cell[3] = {-1, -1, -1};
    if (n0 + lane < N) {
        const float* pp = p + ((long long)b * N + n0 + lane) * 3;
        plane_cells(pp[0], pp[1], pp[2], den, R, cell);
    }
    const int npts = (int)min((long long)32, N - n0);
    const float* __restrict__ cb = c + ((long long)b * N + n0) * Cp;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float* __restrict__ pl = planes + ((long long)k * B + b) * RR * Cp;
        int* __restrict__ cn = count + ((long long)k * B + b) * RR;
        int run_cell = -1, run_n = 0;
        float acc[NCH];
#pragma unroll
        for (int h = 0; h < NCH; ++h) acc[h] = 0.0f;
        for (int j = 0; j <= npts; ++j) {
            int cj = (j < npts) ? __shfl_sync(FULL, cell[k], j) : -2;
            if (cj != run_cell) {
                if (run_cell >= 0) {
#pragma unroll
                    for (int h = 0; h < NCH; ++h) {
                        int ch = h * 32 + lane;
                        if (ch < Cp) atomicAdd(pl + (long long)run_cell * Cp + ch, acc[h]);
                        acc[h] = 0.0f;
                    }
                    if (lane == 0) atomicAdd(cn + run_cell, run_n);
                }
                run_cell = cj, run_n = 0;
            }
            if (j < npts) {
#pragma unroll
                for (int h = 0; h < NCH; ++h) {
                    int ch = h * 32 + lane;
                    if (ch < Cp) acc[h] += __ldg(cb + (long long)j * Cp + ch);
                }
                ++run_n;
            }
        }
    }
}

// mean = sum / max(count, 1) (torch_scatter.scatter_mean), in place over (3*B*R*R, C_p)
__global__ void scatter_finalize_kernel(float* __restrict__ planes, const int* __restrict__ count, long long cells, int Cp) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells * Cp) return;
    int n = count[i / Cp];
    if (n > 1) planes[i] = __fdiv_rn(planes[i], (float)n);
}

// ---------------------------------------------------------------------------------------
// Deterministic mode: stable LSD radix sort of (cell, point) per (plane, scene) segment, then
// one warp per cell sums its points in ascending point index -- the summation order of the CPU
// scatter_add_, hence bit-identical sums; counts are integers and exact.
// ---------------------------------------------------------------------------------------
constexpr int SORT_TILE = 2048;    // items per block and radix pass (256 threads x 8)

__global__ void det_keys_kernel(const float* __restrict__ p, int B, long long N, int R, float den,
                                unsigned* __restrict__ keys, unsigned* __restrict__ vals, int* __restrict__ count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    int b = (int)(i / N);
    long long n = i % N;
    int cell[3];
    plane_cells(p[i * 3], p[i * 3 + 1], p[i * 3 + 2], den, R, cell);
    const long long RR = (long long)R * R;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        long long seg = (long long)k * B + b;
        keys[seg * N + n] = (unsigned)cell[k];
        vals[seg * N + n] = (unsigned)n;
        atomicAdd(count + seg * RR + cell[k], 1);
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

// start[cell] = first sorted position of the cell (per segment)
__global__ void det_starts_kernel(const unsigned* __restrict__ keys, long long N, long long RR, unsigned* __restrict__ start) {
    const int seg = blockIdx.y;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned* k = keys + (long long)seg * N;
    if (i == 0 || k[i - 1] != k[i]) start[(long long)seg * RR + k[i]] = (unsigned)i;
}

// one warp per cell: sequential fp32 sum over the cell's points in ascending point index
__global__ void __launch_bounds__(256) det_reduce_kernel(const float* __restrict__ c, const unsigned* __restrict__ vals,
                                                         const unsigned* __restrict__ start, const int* __restrict__ count,
                                                         int B, long long N, int Cp, long long RR, float* __restrict__ planes) {
    const int lane = threadIdx.x & 31;
    const long long cell_g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // over 3*B*RR
    if (cell_g >= 3LL * B * RR) return;
    const long long seg = cell_g / RR;
    const int b = (int)(seg % B);
    const int n = count[cell_g];
    const unsigned* v = vals + seg * N + (n > 0 ? start[cell_g] : 0u);
    const float* __restrict__ cb = c + (long long)b * N * Cp;
    for (int ch = lane; ch < Cp; ch += 32) {
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, __ldg(cb + (long long)v[i] * Cp + ch));
        if (n > 1) acc = __fdiv_rn(acc, (float)n);
        planes[cell_g * Cp + ch] = acc;
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
    unsigned *keys[2], *vals[2], *hist, *start;
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
            long long warps = (N + 31) / 32 * B;
            unsigned blocks = (unsigned)((warps + 7) / 8);
            if (Cp <= 32) scatter_atomic_kernel<1><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else if (Cp <= 64) scatter_atomic_kernel<2><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else if (Cp <= 128) scatter_atomic_kernel<4><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            else scatter_atomic_kernel<8><<<blocks, 256, 0, st>>>(p, c, B, N, Cp, R, den, planes, count);
            GNB_LAUNCH_CHECK();
            if (mode == GNB_SCATTER_ATOMIC) {
                scatter_finalize_kernel<<<ceil_div(cells * Cp, 256), 256, 0, st>>>(planes, count, cells, Cp);
                GNB_LAUNCH_CHECK();
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
        det_keys_kernel<<<ceil_div((long long)B * N, 256), 256, 0, st>>>(p, B, N, R, den, s.keys[0], s.vals[0], count);
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
        det_starts_kernel<<<dim3(ceil_div(N, 256), segs), 256, 0, st>>>(s.keys[cur], N, RR, s.start);
        GNB_LAUNCH_CHECK();
    }
    det_reduce_kernel<<<ceil_div(cells, 8), 256, 0, st>>>(c, s.vals[cur], s.start, count, B, N, Cp, RR, planes);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_scatter_finalize(float* planes, const int32_t* count, int64_t n_cells, int Cp, void* stream) {
    GNB_CHECK_ARG(planes && count && n_cells >= 0 && Cp >= 1, "gnb_scatter_finalize: bad arguments");
    if (n_cells == 0) return 0;
    scatter_finalize_kernel<<<ceil_div(n_cells * Cp, 256), 256, 0, (cudaStream_t)stream>>>(planes, count, n_cells, Cp);
    GNB_LAUNCH_CHECK();
    return 0;
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

// Pieces of the block-histogram counting sort of the brick-binned sampler (sample_binned.cu; also used by the tiled
// triplane scatter experiment, tools/experiments/planes_tiled_scatter.cu.txt): every block of a count kernel leaves its histogram in hmat[bin][block]; the kernels
// below turn that into per-block write offsets, bin sizes / starts and the list of work units -- without a single atomic,
// so the sorted order is the same in every run.
#pragma once
#include "common.cuh"

namespace gnb {

constexpr int BIN_MAXB = 296;                // blocks of the count / scatter kernels (columns of the histogram matrix)

// Column-wise exclusive scan of the histogram matrix (no atomics, so the sorted order is deterministic per block):
// hmat[i][j] <- queries of bin i in blocks < j, count[i] <- size of bin i.  One warp per bin, lanes over the blocks.
static __global__ void __launch_bounds__(1024) bin_reduce_kernel(unsigned* __restrict__ hmat, unsigned* __restrict__ count, int nbricks, int blocks, int hstride) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 32 + (threadIdx.x >> 5);
    if (i >= nbricks) return;
    constexpr int NCH = (BIN_MAXB + 31) / 32;
    unsigned h[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {                        // every load in flight before the first scan
        const int j = c * 32 + lane;
        h[c] = j < blocks ? hmat[(size_t)i * hstride + j] : 0u;
    }
    unsigned carry = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j = c * 32 + lane;
        unsigned inc = h[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (j < blocks) hmat[(size_t)i * hstride + j] = carry + inc - h[c];
        carry += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) count[i] = carry;
}

// exclusive scans of the bin sizes (start) and of the bins' work-unit counts (ustart), packed into one 64-bit scan
static __global__ void __launch_bounds__(1024) bin_scan_kernel(const unsigned* __restrict__ count, unsigned* __restrict__ start,
                                                        unsigned* __restrict__ ustart, int n, int unit_max, unsigned* __restrict__ work) {
    typedef unsigned long long u64;
    __shared__ u64 s_warp[32];
    __shared__ u64 s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned c = i < n ? count[i] : 0u;
        const u64 v = (u64)c | ((u64)((c + unit_max - 1) / unit_max) << 32);
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u64 w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(FULL, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;                       // exclusive prefix of the warp totals
        }
        __syncthreads();
        const u64 carry = s_carry;
        if (i < n) {
            const u64 ex = carry + s_warp[warp] + inc - v;
            start[i] = (unsigned)ex, ustart[i] = (unsigned)(ex >> 32);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = (unsigned)s_carry, ustart[n] = (unsigned)(s_carry >> 32), *work = 0u;
}


}  // namespace gnb

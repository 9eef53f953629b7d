// Microbenchmark: how fast can SMs pull an L2-resident buffer into shared memory with 1-D bulk async copies
// (the decoder's weight stream)?  Answers two design questions for decoder_tc.cu:
//   * the per-SM ingest rate (ring depth x stage size vs latency), and
//   * the chip-wide L2 -> SM throughput cap when every SM streams the same few MB.
// Also times tcgen05.ld of a 128-lane x N-column fp32 accumulator (epilogue floor).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one thread per CTA: keep `nstage` copies of `stage_bytes` in flight, walking a `buf_bytes` buffer `reps` times
__global__ void __launch_bounds__(128, 1) stream_kernel(const unsigned char* buf, long long buf_bytes, int stage_bytes, int nstage, int reps,
                                                         long long* cycles, int nissue) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars_all[32];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 32; ++s) mbar_init(smem_u32(&bars_all[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // `nissue` warps each run an independent ring (lane 0 issues)
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nissue) {
        const int iw = threadIdx.x >> 5;
        uint64_t* bars = bars_all + iw * 8;
        sm += (size_t)iw * nstage * stage_bytes;
        reps /= nissue;
        const long long per_rep = buf_bytes / stage_bytes;
        const long long total = per_rep * reps;
        // stagger the starting offset per CTA so that the SMs do not all hit the same L2 slice at once
        long long pos = ((long long)blockIdx.x * 7 + iw * 31) % per_rep;
        const long long t0 = clock64();
        long long issued = 0, done = 0;
        uint32_t phase_bits = 0;
        for (; issued < nstage && issued < total; ++issued) {
            const int s = (int)(issued % nstage);
            mbar_expect_tx(smem_u32(&bars[s]), stage_bytes);
            bulk_g2s(smem_u32(sm + (size_t)s * stage_bytes), buf + pos * stage_bytes, stage_bytes, smem_u32(&bars[s]));
            if (++pos == per_rep) pos = 0;
        }
        for (; done < total; ++done) {
            const int s = (int)(done % nstage);
            mbar_wait(smem_u32(&bars[s]), (phase_bits >> s) & 1);
            phase_bits ^= 1u << s;
            if (issued < total) {
                mbar_expect_tx(smem_u32(&bars[s]), stage_bytes);
                bulk_g2s(smem_u32(sm + (size_t)s * stage_bytes), buf + pos * stage_bytes, stage_bytes, smem_u32(&bars[s]));
                if (++pos == per_rep) pos = 0;
                ++issued;
            }
        }
        if (iw == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

// tcgen05.ld throughput: 4 or 8 warps read a 128 x ncols fp32 accumulator `reps` times
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__global__ void __launch_bounds__(256, 1) tmem_ld_kernel(int ncols, int reps, int nwarps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    uint32_t acc = 0;
    long long t0 = 0;
    if (warp < nwarps) {
        const int q = warp & 3, g = warp >> 2, ng = nwarps >> 2;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int c = g * 32; c < ncols; c += 32 * ng) {
                uint32_t v[32];
                tmem_ld32(tl + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 32; ++e) acc ^= v[e];
            }
        }
        t0 = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t0;
    if (acc == 0x12345678) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    int sms = 0, clk_khz = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("SMs %d, max SM clock %.0f MHz\n", sms, clk_khz / 1e3);
    const long long buf_bytes = 5 * 1024 * 1024 + 512 * 1024;     // 5.5 MB: the decoder's weight image, L2 resident
    unsigned char* buf;
    long long* cyc;
    uint32_t* sink;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 1, buf_bytes));
    CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
    CK(cudaMalloc(&sink, 1024 * sizeof(uint32_t)));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("%6s %8s %7s %6s | %10s %12s %10s\n", "grid", "stageKB", "nstage", "reps", "ms", "chip GB/s", "B/clk/SM");
    const int grids[] = {8, 37, 74, 148};
    const int stage_kb[] = {8, 16, 32, 64};
    const int nstages[] = {1, 2, 3, 4};
    for (int gi = 3; gi < 4; ++gi)
        for (int si = 0; si < 4; ++si)
            for (int ni = 0; ni < 4; ++ni) {
              for (int nissue = 1; nissue <= 2; ++nissue) {
                const int grid = grids[gi], sb = stage_kb[si] * 1024, ns = nstages[ni];
                if ((long long)sb * ns * nissue > 198 * 1024) continue;
                const int reps = 40;
                stream_kernel<<<grid, 128, (size_t)sb * ns * nissue + 1024>>>(buf, buf_bytes, sb, ns, 2, cyc, nissue);      // warm-up (fills L2)
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                stream_kernel<<<grid, 128, (size_t)sb * ns * nissue + 1024>>>(buf, buf_bytes, sb, ns, reps, cyc, nissue);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                long long h[148];
                CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                double mean_cyc = 0;
                for (int i = 0; i < grid; ++i) mean_cyc += (double)h[i];
                mean_cyc /= grid;
                const double bytes_per_cta = (double)(buf_bytes / sb) * sb * reps;
                printf("%6d %8d %7d %6d | %10.3f %12.1f %10.2f  (issuing warps %d)\n", grid, stage_kb[si], ns, reps, ms, bytes_per_cta * grid / (ms * 1e-3) / 1e9,
                       bytes_per_cta / mean_cyc, nissue);
              }
            }
    printf("\ntcgen05.ld 32x32b.x32 throughput (one CTA per SM, 128 lanes x ncols fp32)\n");
    printf("%6s %6s %7s | %12s %12s\n", "grid", "warps", "ncols", "cyc/pass", "B/clk/SM");
    for (int nw = 4; nw <= 8; nw += 4)
        for (int ncols = 128; ncols <= 512; ncols *= 2) {
            const int reps = 200;
            tmem_ld_kernel<<<sms, 256>>>(ncols, reps, nw, cyc, sink);
            CK(cudaDeviceSynchronize());
            long long h[148];
            CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
            double mean_cyc = 0;
            for (int i = 0; i < sms; ++i) mean_cyc += (double)h[i];
            mean_cyc /= sms;
            printf("%6d %6d %7d | %12.1f %12.1f\n", sms, nw, ncols, mean_cyc / reps, 128.0 * ncols * 4 * reps / mean_cyc);
        }
    return 0;
}

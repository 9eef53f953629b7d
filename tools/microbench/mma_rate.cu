// Microbenchmark: sustained tcgen05.mma rate of ONE CTA per SM in the decoder's operand configuration
// (kind::f16, M128, N = 128/256, K16, both operands in 128B-swizzled shared memory), alone and with a
// concurrent global -> shared bulk-copy stream (the weight ring) competing for shared-memory bandwidth.
// Answers: is the decoder's k-step (4 MMAs, nominal 512 cycles at N=256) limited by operand bandwidth?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
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
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t umma_idesc(int n, int m) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// warp 1 lane 0: `ksteps` k-steps of (256/N) x 4 MMAs, a commit every `commit_every` k-steps (waiting so that at most
//                 `depth` commits are outstanding -- the issue queue is never the limit)
// warp 0 lane 0 and warp 2 lane 0: bulk-copy streams of `copy_kb` KB copies until the MMA thread raises `stop`
__global__ void __launch_bounds__(128, 1) mma_kernel(const unsigned char* buf, long long buf_bytes, int N, int ksteps, int copy_kb, int ncopiers,
                                                      long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    __shared__ long long copied[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // operands: A 16 KB at 0, B 32 KB at 16 KB; copy ring from 64 KB
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;   // fp16 0.0625
    if (threadIdx.x == 0) {
        for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc(N, 128);
        const uint64_t da = umma_desc(smem_u32(sm)), db = umma_desc(smem_u32(sm + 16 * 1024));
        const int ntile = 256 / N;
        const long long t0 = clock64();
        uint32_t committed = 0, waited = 0;
        for (int k = 0; k < ksteps; ++k) {
            for (int nt = 0; nt < ntile; ++nt) {
                const uint64_t dbn = db + (uint64_t)((nt * N * 128) >> 4);
                umma_f16(tmem + nt * N, da, dbn, idesc, k > 0);
                umma_f16(tmem + nt * N, da + 2, dbn + 2, idesc, 1u);
                umma_f16(tmem + nt * N, da + 4, dbn + 4, idesc, 1u);
                umma_f16(tmem + nt * N, da + 6, dbn + 6, idesc, 1u);
            }
            umma_commit(smem_u32(&bars[8 + (committed & 3)]));
            ++committed;
            if (committed - waited == 4) {       // keep at most 4 k-steps in flight
                mbar_wait(smem_u32(&bars[8 + (waited & 3)]), (waited >> 2) & 1);
                ++waited;
            }
        }
        for (; waited < committed; ++waited) mbar_wait(smem_u32(&bars[8 + (waited & 3)]), (waited >> 2) & 1);
        out[blockIdx.x * 4] = clock64() - t0;
        stop = 1;
    } else if ((warp == 0 || warp == 2) && lane == 0 && (warp >> 1) < ncopiers && copy_kb > 0) {
        const int iw = warp >> 1;
        const int bytes = copy_kb * 1024, nst = 2;
        const long long per = buf_bytes / bytes;
        long long pos = (blockIdx.x * 7 + iw * 13) % per, n = 0;
        unsigned char* ring = sm + 64 * 1024 + (size_t)iw * nst * bytes;
        uint32_t ph = 0;
        for (int s = 0; s < nst; ++s) {
            mbar_expect_tx(smem_u32(&bars[iw * 4 + s]), bytes);
            bulk_g2s(smem_u32(ring + (size_t)s * bytes), buf + pos * bytes, bytes, smem_u32(&bars[iw * 4 + s]));
            if (++pos == per) pos = 0;
        }
        while (!stop) {
            const int s = (int)(n % nst);
            mbar_wait(smem_u32(&bars[iw * 4 + s]), (ph >> s) & 1);
            ph ^= 1u << s;
            ++n;
            mbar_expect_tx(smem_u32(&bars[iw * 4 + s]), bytes);
            bulk_g2s(smem_u32(ring + (size_t)s * bytes), buf + pos * bytes, bytes, smem_u32(&bars[iw * 4 + s]));
            if (++pos == per) pos = 0;
        }
        for (int s = 0; s < nst; ++s) {          // drain
            const int ss = (int)((n + s) % nst);
            mbar_wait(smem_u32(&bars[iw * 4 + ss]), (ph >> ss) & 1);
            ph ^= 1u << ss;
        }
        copied[iw] = n * bytes;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long c = 0;
        for (int i = 0; i < ncopiers && copy_kb > 0; ++i) c += copied[i];
        out[blockIdx.x * 4 + 1] = c;
    }
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// The decoder's issue loop in isolation: a whole warp walks `ksteps` steps; per step `nwaits` waits on barriers that are
// already complete, then one elected lane issues 4 MMAs (N columns) + `ncommits` commits.  mode 0: every lane polls
// (try_wait loop), 1: lane 0 polls and a vote makes the result uniform, 2: single thread (lane 0) runs the whole loop.
__global__ void __launch_bounds__(128, 1) issue_kernel(int N, int ksteps, int nwaits, int ncommits, int mode, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && (mode != 2 || lane == 0)) {
        const uint32_t idesc = umma_idesc(N, 128);
        const long long t0 = clock64();
        uint32_t committed = 0, waited = 0;
        for (int k = 0; k < ksteps; ++k) {
            // waits on barriers 0..nwaits-1: never armed, so waiting for parity 1 ("previous phase") completes at once
            for (int wv = 0; wv < nwaits; ++wv) {
                const uint32_t bar = smem_u32(&bars[wv]);
                if (mode == 0 || mode == 2) {
                    while (!mbar_try(bar, 1)) { }
                } else {
                    uint32_t done = 0;
                    do {
                        if (lane == 0) done = mbar_try(bar, 1);
                        done = __any_sync(0xffffffffu, done);
                    } while (!done);
                }
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = umma_desc(smem_u32(sm) + (k & 1) * 0), db = umma_desc(smem_u32(sm + 16 * 1024));
            if (mode == 2 || elect_one()) {
                umma_f16(tmem, da, db, idesc, k > 0);
                umma_f16(tmem, da + 2, db + 2, idesc, 1u);
                umma_f16(tmem, da + 4, db + 4, idesc, 1u);
                umma_f16(tmem, da + 6, db + 6, idesc, 1u);
                for (int c = 1; c < ncommits; ++c) umma_commit(smem_u32(&bars[12 + (c & 3)]));      // nobody waits on these
                umma_commit(smem_u32(&bars[8 + (committed & 3)]));
            }
            if (mode != 2) __syncwarp();
            ++committed;
            if (committed - waited == 4) {
                if (mode == 2 || lane == 0) mbar_wait(smem_u32(&bars[8 + (waited & 3)]), (waited >> 2) & 1);
                if (mode != 2) __syncwarp();
                ++waited;
            }
        }
        if (mode == 2 || lane == 0) {
            for (; waited < committed; ++waited) mbar_wait(smem_u32(&bars[8 + (waited & 3)]), (waited >> 2) & 1);
            out[blockIdx.x * 4] = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// Per-instruction cost on the issuing thread: a single thread runs `ksteps` steps of
//   [nwaits x try_wait on a complete barrier] [fence x tcgen05.fence::after_thread_sync] nmma x MMA [ncommits x commit]
// and waits for the commits only every 8 steps (so the wait for completion is amortised away).
__global__ void __launch_bounds__(128, 1) cost_kernel(int N, int ksteps, int nmma, int nwaits, int nfence, int ncommits, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c002c00u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc(N, 128);
        const uint64_t da = umma_desc(smem_u32(sm)), db = umma_desc(smem_u32(sm + 16 * 1024));
        const long long t0 = clock64();
        uint32_t big = 0;
        for (int k = 0; k < ksteps; ++k) {
            for (int wv = 0; wv < nwaits; ++wv) while (!mbar_try(smem_u32(&bars[wv]), 1)) { }
            for (int f = 0; f < nfence; ++f) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int m = 0; m < nmma; ++m) umma_f16(tmem, da + 2 * (m & 3), db + 2 * (m & 3), idesc, (k | m) > 0);
            for (int c = 0; c < ncommits; ++c) umma_commit(smem_u32(&bars[12 + (c & 3)]));      // nobody waits on these
            if ((k & 7) == 7) {
                umma_commit(smem_u32(&bars[8]));
                mbar_wait(smem_u32(&bars[8]), big & 1);
                ++big;
            }
        }
        out[blockIdx.x * 4] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long buf_bytes = 5 * 1024 * 1024 + 512 * 1024;
    unsigned char* buf;
    long long* out;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 0x2c, buf_bytes));
    CK(cudaMalloc(&out, 4 * 1024 * sizeof(long long)));
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    printf("%4s %7s %8s %8s | %12s %10s %12s\n", "N", "ksteps", "copyKB", "copiers", "cyc/kstep", "MMA eff", "copy B/clk");
    const int Ns[] = {256, 128};
    const int copies[][2] = {{0, 0}, {16, 1}, {32, 1}, {32, 2}, {64, 1}};
    for (int ni = 0; ni < 2; ++ni)
        for (int ci = 0; ci < 5; ++ci) {
            const int N = Ns[ni], ksteps = 4000, ckb = copies[ci][0], ncp = copies[ci][1];
            for (int rep = 0; rep < 2; ++rep) {
                mma_kernel<<<sms, 128, 194 * 1024>>>(buf, buf_bytes, N, ksteps, ckb, ncp, out);
                CK(cudaDeviceSynchronize());
            }
            long long h[4 * 148];
            CK(cudaMemcpy(h, out, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
            double cyc = 0, cp = 0;
            for (int i = 0; i < sms; ++i) cyc += (double)h[i * 4], cp += (double)h[i * 4 + 1];
            cyc /= sms, cp /= sms;
            printf("%4d %7d %8d %8d | %12.1f %10.3f %12.1f\n", N, ksteps, ckb, ncp, cyc / ksteps, 512.0 / (cyc / ksteps), cp / cyc);
        }
    printf("\nper-instruction cost on the issuing thread (single thread; completion waited for every 8 steps)\n");
    printf("%4s %5s %7s %7s %8s | %10s\n", "N", "nmma", "nwaits", "nfence", "commits", "cyc/step");
    CK(cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    {
        const int cfg[][5] = {{32, 4, 0, 0, 0}, {32, 8, 0, 0, 0}, {32, 16, 0, 0, 0}, {32, 4, 1, 0, 0}, {32, 4, 2, 0, 0}, {32, 4, 0, 1, 0}, {32, 4, 0, 2, 0},
                              {32, 4, 0, 0, 1}, {32, 4, 0, 0, 2}, {32, 4, 1, 0, 1}, {32, 4, 1, 1, 1}, {64, 4, 0, 0, 0}, {128, 4, 0, 0, 0}, {128, 8, 0, 0, 0},
                              {128, 16, 0, 0, 0}, {256, 4, 0, 0, 0}, {256, 8, 0, 0, 0}, {256, 16, 0, 0, 0}, {256, 4, 1, 0, 1}, {256, 4, 2, 0, 2}, {128, 4, 1, 0, 1}};
        for (auto& c : cfg) {
            for (int rep = 0; rep < 2; ++rep) {
                cost_kernel<<<sms, 128, 64 * 1024>>>(c[0], 2048, c[1], c[2], c[3], c[4], out);
                CK(cudaDeviceSynchronize());
            }
            long long h[4 * 148];
            CK(cudaMemcpy(h, out, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
            double cyc = 0;
            for (int i = 0; i < sms; ++i) cyc += (double)h[i * 4];
            printf("%4d %5d %7d %7d %8d | %10.1f\n", c[0], c[1], c[2], c[3], c[4], cyc / sms / 2048);
        }
    }
    printf("\nissue loop: cycles per step (4 MMAs, nominal %d / %d cycles at N=128 / 256)\n", 256, 512);
    printf("%4s %7s %8s %6s | %10s\n", "N", "nwaits", "commits", "mode", "cyc/step");
    CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int N = 256; N <= 256; N *= 2)
        for (int mode = 0; mode < 3; ++mode)
            for (int nw = 0; nw <= 2; ++nw)
                for (int nc = 1; nc <= 3; nc += 2) {
                    for (int rep = 0; rep < 2; ++rep) {
                        issue_kernel<<<sms, 128, 64 * 1024>>>(N, 2000, nw, nc, mode, out);
                        CK(cudaDeviceSynchronize());
                    }
                    long long h[4 * 148];
                    CK(cudaMemcpy(h, out, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
                    double cyc = 0;
                    for (int i = 0; i < sms; ++i) cyc += (double)h[i * 4];
                    printf("%4d %7d %8d %6d | %10.1f\n", N, nw, nc, mode, cyc / sms / 2000);
                }
    return 0;
}

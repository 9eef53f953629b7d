// Microtest + microbenchmark for the "transposed pair" decoder design (tools/experiments/README.md):
//   D[hidden (M = 256 over a CTA pair), queries (N = 128)] += W[hidden, k] (A, K-major) x Act[k, queries] (B, MN-major)
// issued as tcgen05.mma.cta_group::2: each CTA supplies its own 128 rows of A and HALF of B's N (its 64 queries), the
// hardware shares B across the pair; each CTA's TMEM receives its 128 hidden units x all 128 queries.
// B rows (one k each, 128 bytes = 64 queries, 128B-swizzled in 8-row atoms) are written by generic stores, the half that
// lives in the partner CTA through st.shared::cluster, followed by fence.proxy.async and a (remote) mbarrier arrive.
// Part 1 checks D against a host reference; part 2 times the leader's issue loop (8 MMAs + commit per 32 KB "stage").
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_pair_test umma_pair_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory"); }
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr_cluster, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 128B-swizzled tile of 128-byte rows in 8-row atoms of 1024 B: K-major A (row = m, 64 k per row) and MN-major B (row = k, 64 n per row)
__device__ __host__ inline uint32_t tile_off(int r, int u) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)); }
// start>>4 | LBO | SBO = 1024 B | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16, D = f32, A/B = f16, A K-major, B MN-major (bit 16), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int m, int n, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_2cta(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_mc_2cta(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
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

// Warp-converged issue: ALL 32 lanes execute the instruction with warp-uniform operands; tcgen05.mma / commit are
// uniform-datapath instructions (one issue per warp), so ptxas emits a bare UTCHMMA instead of the per-lane
// "ELECT ... BRA.U.ANY" serialisation loop it needs inside a divergent `if (lane == 0)` region (~80 cycles per MMA).
__device__ __forceinline__ void umma_2cta_u(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_mc_2cta_u(uint32_t bar, uint16_t mask) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int KTOT = 64;     // one 64-wide k-chunk
// W (256 x 64) and Act (64 x 128) as small integers (exact in fp16 and in the fp32 accumulator)
__device__ __host__ inline float w_val(int m, int k) { return (float)(((m * 7 + k * 3) % 11) - 5); }
__device__ __host__ inline float a_val(int k, int n) { return (float)(((k * 5 + n * 13) % 9) - 4); }

// mode 0: correctness; mode 1: issue-loop timing (nstages stages of 8 MMAs + commit, `nwaits` waits on a complete barrier each)
__global__ void __launch_bounds__(128, 1) pair_kernel(float* out, int mode, int nstages, int nwaits, long long* cycles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const uint32_t sbase = smem_u32(sm);
    const uint32_t a_off = 0;              // A: 128 rows x 64 k fp16, K-major SW128: 16 KB
    const uint32_t b_off = 16 * 1024;      // B: 64 k-rows x 64 own queries fp16, MN-major SW128: 8 KB
    const uint32_t in_ready = smem_u32(&bars[0]), done = smem_u32(&bars[1]), dummy = smem_u32(&bars[2]);
    if (threadIdx.x == 0) {
        mbar_init(in_ready, 2 * 128);      // every thread of both CTAs arrives on the LEADER's barrier
        mbar_init(done, 1);
        mbar_init(dummy, 1);
        for (int i = 4; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    // ---- operands ----
    // A: this CTA's 128 hidden rows; thread t writes row t (8 units of 8 k)
    {
        const int r = threadIdx.x, m = rank * 128 + r;
        for (int u = 0; u < 8; ++u) {
            __half2 h[4];
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(w_val(m, u * 8 + 2 * e), w_val(m, u * 8 + 2 * e + 1));
            *reinterpret_cast<uint4*>(sm + a_off + tile_off(r, u)) = *reinterpret_cast<uint4*>(h);
        }
    }
    // B: k-row k is "produced" by CTA (k / 32) & 1 (as an epilogue would): threads 0..31 of that CTA write row k = 32*j + lane ... here
    //    thread t < 64 handles (k-row = (t & 31) + 32 * rank, half = t >> 5): queries [64*half, 64*half + 64) go to CTA `half`
    if (threadIdx.x < 64) {
        const int k = (threadIdx.x & 31) + 32 * (int)rank, half = threadIdx.x >> 5;
        const uint32_t dst_base = map_to_cta(sbase + b_off, (uint32_t)half);
        for (int u = 0; u < 8; ++u) {
            __half2 h[4];
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(a_val(k, 64 * half + u * 8 + 2 * e), a_val(k, 64 * half + u * 8 + 2 * e + 1));
            st_cluster_v4(dst_base + tile_off(k, u), *reinterpret_cast<uint4*>(h));
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy stores (local and remote) -> async proxy (tensor core)
    if (rank == 0) mbar_arrive(in_ready);
    else mbar_arrive_remote(map_to_cta(in_ready, 0));

    if (rank == 0 && threadIdx.x == 32) {
        mbar_wait(in_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = umma_idesc(256, 128, 1);
        const uint64_t da = umma_desc(sbase + a_off, 16), db = umma_desc(sbase + b_off, 16);
        if (mode >= 2) {
            if (mode == 4) umma_commit_mc_2cta(done, 3);
        } else if (mode == 0) {
            for (int s = 0; s < KTOT / 16; ++s)               // K16 steps: A +32 bytes inside the swizzled row, B +2 atoms (16 k-rows)
                umma_2cta(tmem, da + 2 * s, db + (uint64_t)((2048 * s) >> 4), idesc, s > 0);
            umma_commit_mc_2cta(done, 3);
        } else {
            const long long t0 = clock64();
            uint32_t committed = 0, waited = 0;
            for (int st = 0; st < nstages; ++st) {
                for (int wv = 0; wv < nwaits; ++wv) while (!mbar_try(dummy, 1)) { }
                for (int half = 0; half < 2; ++half)
                    for (int s = 0; s < 4; ++s) umma_2cta(tmem + half * 128, da + 2 * s, db + (uint64_t)((2048 * s) >> 4), idesc, 1u);
                umma_commit_mc_2cta(smem_u32(&bars[4 + (committed & 3)]), 1);
                ++committed;
                if (committed - waited == 4) { mbar_wait(smem_u32(&bars[4 + (waited & 3)]), (waited >> 2) & 1); ++waited; }
            }
            for (; waited < committed; ++waited) mbar_wait(smem_u32(&bars[4 + (waited & 3)]), (waited >> 2) & 1);
            cycles[blockIdx.x / 2] = clock64() - t0;
            umma_commit_mc_2cta(done, 3);
        }
    }
    if (mode == 4) {
        // DSMEM store rate: every thread of BOTH CTAs writes `nstages` x 8 x 16 B into the partner's shared memory (the
        // epilogue's remote half: 8 units of a 128-byte row per thread), rows spread like the real tiles
        const uint32_t dst = map_to_cta(sbase + 24 * 1024, rank ^ 1u);
        const long long t0 = clock64();
        for (int it = 0; it < nstages; ++it) {
            const int r = (threadIdx.x + it * 128) & 255;
#pragma unroll
            for (int u = 0; u < 8; ++u) st_cluster_v4(dst + tile_off(r, u), make_uint4(it, u, r, 7));
        }
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0 && rank == 0) cycles[blockIdx.x / 2] = clock64() - t0;
    }
    if (rank == 0 && warp == 2 && (mode == 2 || mode == 3)) {
        // the whole warp runs the loop, converged
        const uint32_t idesc = umma_idesc(256, 128, 1);
        const uint64_t da = umma_desc(sbase + a_off, 16), db = umma_desc(sbase + b_off, 16);
        const long long t0 = clock64();
        uint32_t committed = 0, waited = 0;
        for (int st = 0; st < nstages; ++st) {
            for (int wv = 0; wv < nwaits; ++wv) {
                if (mode == 2) { while (!mbar_try(dummy, 1)) { } }
                else {
                    uint32_t ok = 0;
                    do { if (lane == 0) ok = mbar_try(dummy, 1); ok = __shfl_sync(0xffffffffu, ok, 0); } while (!ok);
                }
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
                for (int s = 0; s < 4; ++s) umma_2cta_u(tmem + half * 128, da + 2 * s, db + (uint64_t)((2048 * s) >> 4), idesc, 1u);
            umma_commit_mc_2cta_u(smem_u32(&bars[4 + (committed & 3)]), 1);
            ++committed;
            if (committed - waited == 4) {
                if (mode == 2) mbar_wait(smem_u32(&bars[4 + (waited & 3)]), (waited >> 2) & 1);
                else { if (lane == 0) mbar_wait(smem_u32(&bars[4 + (waited & 3)]), (waited >> 2) & 1); __syncwarp(); }
                ++waited;
            }
        }
        for (; waited < committed; ++waited) mbar_wait(smem_u32(&bars[4 + (waited & 3)]), (waited >> 2) & 1);
        if (lane == 0) cycles[blockIdx.x / 2] = clock64() - t0;
        umma_commit_mc_2cta_u(done, 3);
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
        // D: lane = local hidden row (warp w -> lanes 32w..32w+31), columns = 128 queries
        const int m = rank * 128 + warp * 32 + lane;
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 32; ++e) out[(long long)(blockIdx.x / 2) * 256 * 128 + m * 128 + c0 + e] = __uint_as_float(v[e]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

static void launch(int clusters, float* out, int mode, int nstages, int nwaits, long long* cyc) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 64 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, pair_kernel, out, mode, nstages, nwaits, cyc));
    CK(cudaDeviceSynchronize());
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * 256 * 128 * 2));
    CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    CK(cudaMemset(out, 0xff, sizeof(float) * 256 * 128 * 2));
    launch(2, out, 0, 0, 0, cyc);
    float* h = (float*)malloc(sizeof(float) * 256 * 128 * 2);
    CK(cudaMemcpy(h, out, sizeof(float) * 256 * 128 * 2, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < 2; ++c)
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < 128; ++n) {
                float ref = 0;
                for (int k = 0; k < KTOT; ++k) ref += w_val(m, k) * a_val(k, n);
                const float got = h[c * 256 * 128 + m * 128 + n];
                if (got != ref) {
                    if (bad < 10) printf("mismatch cluster %d m %d n %d: got %g want %g\n", c, m, n, got, ref);
                    ++bad;
                }
            }
    printf("correctness: %s (%d mismatches of %d)\n", bad ? "FAILED" : "ok", bad, 2 * 256 * 128);
    printf("\nleader issue loop, cta_group::2 M256 N128 K16, 8 MMAs + 1 commit per stage (nominal 542 cycles of MMA per stage per SM)\n");
    printf("%8s %7s | %10s\n", "clusters", "nwaits", "cyc/stage");
    for (int md = 1; md <= 3; ++md)
    for (int nw = 0; nw <= 2; ++nw) {
        const int clusters = sms / 2, nst = 2000;
        launch(clusters, out, md, nst, nw, cyc);
        launch(clusters, out, md, nst, nw, cyc);
        long long hc[256];
        CK(cudaMemcpy(hc, cyc, sizeof(long long) * clusters, cudaMemcpyDeviceToHost));
        double s = 0;
        for (int i = 0; i < clusters; ++i) s += (double)hc[i];
        printf("%8d %7d | %10.1f   mode %d (%s)\n", clusters, nw, s / clusters / nst, md,
               md == 1 ? "single thread" : md == 2 ? "converged warp, every lane polls" : "converged warp, lane 0 polls + shfl");
    }
    {
        const int clusters = sms / 2, nst = 2000;
        launch(clusters, out, 4, nst, 0, cyc);
        launch(clusters, out, 4, nst, 0, cyc);
        long long hc[256];
        CK(cudaMemcpy(hc, cyc, sizeof(long long) * clusters, cudaMemcpyDeviceToHost));
        double s = 0;
        for (int i = 0; i < clusters; ++i) s += (double)hc[i];
        printf("\nDSMEM st.shared::cluster.v4, 128 threads per CTA, both directions at once: %.1f cycles per 16 KB per direction = %.1f B/clk per direction\n",
               s / clusters / nst, 16384.0 / (s / clusters / nst));
    }
    return bad ? 1 : 0;
}

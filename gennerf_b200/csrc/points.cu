// Front end of the triplane branch ("next" row 1):
//   get_3d_points()          reference src/models/utils.py:120-175  depth map -> world points
//   farthest_point_sample()  reference src/models/utils.py:178-202  512 sequential arg-max steps per frame
// The reference runs FPS as npoint Python iterations of ~6 small kernels each (308 ms per 240x320 frame on the
// CPU); here one CTA per cloud keeps the whole iteration on chip: distance update, running arg-max and the
// block-wide reduction with two barriers per step.
#include <stdlib.h>

#include "common.cuh"

namespace gnb {

// out[b,h,w,:] = M_b . [u*d, v*d, d, 1]  with M_b = the first three rows of inverse([P_b; 0 0 0 1]) (computed on the
// host in double precision); the fourth homogeneous coordinate is exactly 1, so the reference's final division
// (utils.py:170) is the identity.
struct UnprojectKP {
    float M[8][12];
};
__global__ void unproject_kernel(const float* __restrict__ depth, const __grid_constant__ UnprojectKP kp, int b0, int nb, int H,
                                 int W, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)H * W;
    if (i >= per * nb) return;
    const int bl = (int)(i / per);
    const long long r = i % per;
    const float u = (float)(r % W), v = (float)(r / W);
    const long long gi = (long long)(b0 + bl) * per + r;
    const float d = __ldg(depth + gi);
    const float a0 = __fmul_rn(u, d), a1 = __fmul_rn(v, d);
    const float* M = kp.M[bl];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out[gi * 3 + k] = __fadd_rn(fmaf(M[k * 4 + 2], d, fmaf(M[k * 4 + 1], a1, __fmul_rn(M[k * 4 + 0], a0))), M[k * 4 + 3]);
}

// One CTA per cloud.  dist lives in global scratch (L2-resident: N * 4 bytes per cloud).
// Arithmetic of the reference: dist = ((dx*dx + dy*dy) + dz*dz); distance = min(distance, dist) through the
// `dist < distance` mask; farthest = FIRST index of the maximum (torch.max).
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ xyz, long long N, int npoint,
                                                   const long long* __restrict__ start, float* __restrict__ dist,
                                                   long long* __restrict__ out_idx, float* __restrict__ out_xyz) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_far;
    const int b = blockIdx.x;
    const float* __restrict__ p = xyz + (long long)b * N * 3;
    float* __restrict__ dd = dist + (long long)b * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long i = threadIdx.x; i < N; i += blockDim.x) dd[i] = 1e10f;
    int far = (int)start[b];
    __syncthreads();
    for (int it = 0; it < npoint; ++it) {
        if (threadIdx.x == 0) {
            out_idx[(long long)b * npoint + it] = far;
            out_xyz[((long long)b * npoint + it) * 3 + 0] = p[far * 3LL + 0];
            out_xyz[((long long)b * npoint + it) * 3 + 1] = p[far * 3LL + 1];
            out_xyz[((long long)b * npoint + it) * 3 + 2] = p[far * 3LL + 2];
        }
        const float cx = __ldg(p + far * 3LL), cy = __ldg(p + far * 3LL + 1), cz = __ldg(p + far * 3LL + 2);
        float best = -1.0f;
        int besti = 0x7fffffff;
        for (long long i = threadIdx.x; i < N; i += blockDim.x) {
            const float dx = __fsub_rn(p[i * 3], cx), dy = __fsub_rn(p[i * 3 + 1], cy), dz = __fsub_rn(p[i * 3 + 2], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float cur = dd[i];
            if (d < cur) { cur = d; dd[i] = d; }
            if (cur > best) { best = cur; besti = (int)i; }       // ascending i: keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = besti; }
        __syncthreads();
        if (warp == 0) {
            best = lane < (int)(blockDim.x >> 5) ? s_val[lane] : -1.0f;
            besti = lane < (int)(blockDim.x >> 5) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, besti, o);
                if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
            }
            if (lane == 0) s_far = besti;
        }
        __syncthreads();
        far = s_far;
    }
}

// Cluster version: CS CTAs (a thread-block cluster, up to 16) share one cloud.  CTA r keeps the coordinates of its
// contiguous slice of the points in shared memory (SoA) and their running distances in registers, so an iteration
// touches no global memory at all: distance update + local arg-max, one candidate (value, index, xyz) per CTA written
// into every CTA's shared memory through DSMEM followed by a remote mbarrier arrive (a hardware cluster barrier per
// iteration cost ~2.5 us), then warp 0 of every CTA reduces the CS candidates.  Candidates and barriers are
// double-buffered by iteration parity (a buffer is rewritten only after its reader has arrived for the next iteration).  Same
// arithmetic and tie rule as above (lowest index among equal maxima, slices ascend with the CTA rank), so the selected
// indices are bit-identical to the reference's.  307 200 points (480x640) x 512 samples: 30 ms -> ~0.3 ms per cloud,
// and the clouds of a batch run in parallel (one cluster per GPC).
constexpr int FPS_PPT = 20;                 // points per thread (registers)
constexpr int FPS_THREADS = 1024;
struct FpsCand { float val; int idx; float x, y, z; };

__device__ __forceinline__ uint32_t fps_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__global__ void __launch_bounds__(FPS_THREADS, 1) fps_cluster_kernel(const float* __restrict__ xyz, long long N, int npoint,
                                                                     const long long* __restrict__ start, int cs, int chunk,
                                                                     long long* __restrict__ out_idx, float* __restrict__ out_xyz) {
    extern __shared__ float fps_sm[];
    __shared__ FpsCand cand[2][16];
    __shared__ FpsCand s_res;
    __shared__ unsigned long long xbar[2];      // mbarriers: CS candidate arrivals per iteration parity
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    float* sx = fps_sm, *sy = fps_sm + chunk, *sz = fps_sm + 2 * chunk;
    const int b = blockIdx.x / cs;
    const int rank = (int)fps_cluster_rank();
    const float* __restrict__ p = xyz + (long long)b * N * 3;
    const long long lo = (long long)rank * chunk;
    const int n_local = (int)max(0LL, min((long long)chunk, N - lo));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_local; i += FPS_THREADS) {
        sx[i] = p[(lo + i) * 3], sy[i] = p[(lo + i) * 3 + 1], sz[i] = p[(lo + i) * 3 + 2];
    }
    float dist[FPS_PPT];
#pragma unroll
    for (int j = 0; j < FPS_PPT; ++j) dist[j] = 1e10f;
    const int jmax = (n_local + FPS_THREADS - 1) / FPS_THREADS;
    int far = (int)start[b];
    float cx = __ldg(p + far * 3LL), cy = __ldg(p + far * 3LL + 1), cz = __ldg(p + far * 3LL + 2);
    if (threadIdx.x == 0) {
        for (int q = 0; q < 2; ++q)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&xbar[q])), "r"(cs) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // every CTA's barriers exist before any peer arrives on them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && threadIdx.x == 0) {
            out_idx[(long long)b * npoint + it] = far;
            out_xyz[((long long)b * npoint + it) * 3 + 0] = cx;
            out_xyz[((long long)b * npoint + it) * 3 + 1] = cy;
            out_xyz[((long long)b * npoint + it) * 3 + 2] = cz;
        }
        float best = -1.0f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < FPS_PPT; ++j) {
            if (j >= jmax) break;                 // block-uniform: predicated-off trips would still cost their issue slots
            const int i = j * FPS_THREADS + threadIdx.x;
            if (i < n_local) {
                const float dx = __fsub_rn(sx[i], cx), dy = __fsub_rn(sy[i], cy), dz = __fsub_rn(sz[i], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[j]) dist[j] = d;
                if (dist[j] > best) { best = dist[j]; besti = i; }       // ascending i: keeps the first maximum
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = besti; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane];
            besti = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, besti, o);
                if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
            }
            // lane r hands this CTA's candidate to CTA r and arrives on that CTA's barrier (release: the stores first)
            if (lane < cs) {
                FpsCand c;
                c.val = best;
                c.idx = besti == 0x7fffffff ? 0x7fffffff : (int)(lo + besti);
                const int li = besti == 0x7fffffff ? 0 : besti;
                c.x = n_local > 0 ? sx[li] : 0.f, c.y = n_local > 0 ? sy[li] : 0.f, c.z = n_local > 0 ? sz[li] : 0.f;
                const uint32_t local = (uint32_t)__cvta_generic_to_shared(&cand[it & 1][rank]);
                const uint32_t lbar = (uint32_t)__cvta_generic_to_shared(&xbar[it & 1]);
                uint32_t remote, rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(lane));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lbar), "r"(lane));
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(c.val) : "memory");
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + 4), "r"(c.idx) : "memory");
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote + 8), "f"(c.x) : "memory");
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote + 12), "f"(c.y) : "memory");
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote + 16), "f"(c.z) : "memory");
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
            }
            // wait for the CS candidates of this iteration (this barrier is used every second iteration)
            {
                const uint32_t lbar = (uint32_t)__cvta_generic_to_shared(&xbar[it & 1]);
                const uint32_t parity = (uint32_t)(it >> 1) & 1u;
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(lbar), "r"(parity) : "memory");
            }
            FpsCand c;
            c.val = -2.0f, c.idx = 0x7fffffff, c.x = c.y = c.z = 0.f;
            if (lane < cs) c = cand[it & 1][lane];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                FpsCand q;
                q.val = __shfl_xor_sync(FULL, c.val, o), q.idx = __shfl_xor_sync(FULL, c.idx, o);
                q.x = __shfl_xor_sync(FULL, c.x, o), q.y = __shfl_xor_sync(FULL, c.y, o), q.z = __shfl_xor_sync(FULL, c.z, o);
                if (q.val > c.val || (q.val == c.val && q.idx < c.idx)) c = q;
            }
            if (lane == 0) s_res = c;
        }
        __syncthreads();
        far = s_res.idx, cx = s_res.x, cy = s_res.y, cz = s_res.z;
    }
    // no CTA may exit while a peer can still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Grid version: CPC co-resident CTAs ANYWHERE on the GPU share one cloud and exchange their candidates through global memory
// (L2), launched cooperatively.  Thread-block clusters must sit inside one GPC: at most seven 16-CTA clusters are co-resident
// on a B200, so the 8 clouds of a scene (480x640 points each: 16 CTAs' worth of shared memory) took two waves of the cluster
// kernel.  Here 8 clouds x 18 CTAs = 144 of the 148 SMs run in ONE wave.  Per iteration every CTA publishes {value, index, xyz}
// into its slot of the cloud's exchange buffer as (word, iteration number) pairs -- 8-byte units are written and read atomically,
// so a reader that sees the iteration number next to every word has the whole record, without a fence on either side (the
// release store / acquire poll version cost 5.8 us per iteration) --.  Every slot has ONE writer and ONE reader (32 CTAs polling the same lines cost 5 us per iteration): CTA r sends
// its candidate to the cloud's leader (rank 0), whose lane r polls it; the leader reduces and sends the result back, one slot per CTA.  Slots are double-buffered by iteration parity: a CTA overwrites its parity-p slot two iterations later,
// after it has seen every peer's record of the iteration in between, which those peers published only after reading parity p.
// Same arithmetic, same slices in ascending rank order and the same tie rule as the cluster kernel: identical indices.
constexpr int FPS_XWORDS = 8;               // words per slot (one 32-byte sector): value, index | tag, 6 pad
constexpr int FPS_MAX_CPC = 32;
// A record travels as (word, sequence number) pairs in three 16-byte stores: 8-byte units are written and read atomically, so a
// reader that finds the iteration's number next to every word has the whole record -- no fence on either side.
__device__ __forceinline__ void fps_ll_store(unsigned* s, const FpsCand& c, unsigned seq) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %2};" ::"l"(s), "r"(__float_as_uint(c.val)), "r"(seq), "r"((unsigned)c.idx) : "memory");
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %2};" ::"l"(s + 4), "r"(__float_as_uint(c.x)), "r"(seq), "r"(__float_as_uint(c.y)) : "memory");
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %2};" ::"l"(s + 8), "r"(__float_as_uint(c.z)), "r"(seq), "r"(0u) : "memory");
}
__device__ __forceinline__ FpsCand fps_ll_poll(const unsigned* s, unsigned seq) {
    unsigned a0, f0, a1, f1, b0, g0, b1, g1, c0, h0, c1, h1;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(f0), "=r"(a1), "=r"(f1) : "l"(s) : "memory");
        asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(g0), "=r"(b1), "=r"(g1) : "l"(s + 4) : "memory");
        asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(c0), "=r"(h0), "=r"(c1), "=r"(h1) : "l"(s + 8) : "memory");
        if (f0 == seq && f1 == seq && g0 == seq && g1 == seq && h0 == seq && h1 == seq) break;
        if (spin > (1u << 24)) __trap();                     // a peer that never arrives must not hang the GPU
    }
    FpsCand c;
    c.val = __uint_as_float(a0), c.idx = (int)a1, c.x = __uint_as_float(b0), c.y = __uint_as_float(b1), c.z = __uint_as_float(c0);
    return c;
}
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_grid_kernel(const float* __restrict__ xyz, long long N, int npoint,
                                                                  const long long* __restrict__ start, int cpc, int chunk, int b0,
                                                                  unsigned* __restrict__ xbuf, long long* __restrict__ out_idx,
                                                                  float* __restrict__ out_xyz) {
    extern __shared__ float fps_sm[];
    __shared__ FpsCand s_res;
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    float* sx = fps_sm, *sy = fps_sm + chunk, *sz = fps_sm + 2 * chunk;
    const int b = b0 + blockIdx.x / cpc, rank = blockIdx.x % cpc;
    const float* __restrict__ p = xyz + (long long)b * N * 3;
    unsigned* xb = xbuf + (long long)b * 2 * cpc * FPS_XWORDS;
    const long long lo = (long long)rank * chunk;
    const int n_local = (int)max(0LL, min((long long)chunk, N - lo));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_local; i += FPS_THREADS) {
        sx[i] = p[(lo + i) * 3], sy[i] = p[(lo + i) * 3 + 1], sz[i] = p[(lo + i) * 3 + 2];
    }
    float dist[FPS_PPT];
#pragma unroll
    for (int j = 0; j < FPS_PPT; ++j) dist[j] = 1e10f;
    const int jmax = (n_local + FPS_THREADS - 1) / FPS_THREADS;
    int far = (int)start[b];
    float cx = __ldg(p + far * 3LL), cy = __ldg(p + far * 3LL + 1), cz = __ldg(p + far * 3LL + 2);
    __syncthreads();
    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && threadIdx.x == 0) {
            out_idx[(long long)b * npoint + it] = far;
            out_xyz[((long long)b * npoint + it) * 3 + 0] = cx;
            out_xyz[((long long)b * npoint + it) * 3 + 1] = cy;
            out_xyz[((long long)b * npoint + it) * 3 + 2] = cz;
        }
        float best = -1.0f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < FPS_PPT; ++j) {
            if (j >= jmax) break;                 // block-uniform: predicated-off trips would still cost their issue slots
            const int i = j * FPS_THREADS + threadIdx.x;
            if (i < n_local) {
                const float dx = __fsub_rn(sx[i], cx), dy = __fsub_rn(sy[i], cy), dz = __fsub_rn(sz[i], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[j]) dist[j] = d;
                if (dist[j] > best) { best = dist[j]; besti = i; }       // ascending i: keeps the first maximum
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = besti; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane];
            besti = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, besti, o);
                if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
            }
            const unsigned seq = (unsigned)it + 1u;
            // this CTA's candidate: ONE 8-byte record {value, index | iteration number in the top byte} -- 8-byte accesses are
            // single-copy atomic, so there is no fence and no flag on either side; slot r is written by CTA r and polled by
            // lane r of every CTA of the cloud
            const unsigned tag = (seq & 0xffu) << 24;
            if (lane == 0) {
                const unsigned gi = besti == 0x7fffffff ? 0x00ffffffu : (unsigned)(lo + besti);
                asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(xb + ((it & 1) * cpc + rank) * FPS_XWORDS),
                             "r"(__float_as_uint(best)), "r"(gi | tag) : "memory");
            }
            float cv = -2.0f;
            unsigned ci = 0x00ffffffu;
            if (lane < cpc) {
                const unsigned* s = xb + ((it & 1) * cpc + lane) * FPS_XWORDS;
                unsigned v0, v1;
                for (unsigned spin = 0;; ++spin) {
                    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "l"(s) : "memory");
                    if ((v1 & 0xff000000u) == tag) break;
                    if (spin > (1u << 24)) __trap();         // a peer that never arrives must not hang the GPU
                }
                cv = __uint_as_float(v0), ci = v1 & 0x00ffffffu;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, cv, o);
                const unsigned oi = __shfl_xor_sync(FULL, ci, o);
                if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
            }
            FpsCand c;
            c.val = cv, c.idx = (int)ci;
            c.x = __ldg(p + ci * 3LL), c.y = __ldg(p + ci * 3LL + 1), c.z = __ldg(p + ci * 3LL + 2);
            if (lane == 0) s_res = c;
        }
        __syncthreads();
        far = s_res.idx, cx = s_res.x, cy = s_res.y, cz = s_res.z;
    }
}

// Training-time ray sampler: sample_points_on_rays(), reference src/models/utils.py:458-540 (iSDF).
// One thread per (camera, ray, sample k): z = depth (k = 0) | linspace(min_dist, depth + delta, N)[k-1] | the k-th
// gaussian depth (drawn by the caller, utils.py:496-498); camera point ((w-cx)/fx*z, (h-cy)/fy*z, z); world point =
// pose . [x y z 1] as the FMA chain ATen's bmm computes for K = 4, divided by the homogeneous coordinate.  The
// reference builds the stratified depths with one torch.linspace call per ray in a Python double loop.
struct RayKP {
    const long long* h_idx;    // (B,S)
    const long long* w_idx;    // (B,S)
    const float* depth;        // (B,S)
    const float* intr;         // (B,3,3)
    const float* pose;         // (B,4,4)
    const float* gauss;        // (B,S,M)
    int B, S, N, M;
    float delta, min_dist;
    float* xyz;                // (B,S,1+N+M,3)
    float* z;                  // (B,S,1+N+M)
};
__global__ void ray_points_kernel(const RayKP p) {
    const int K = 1 + p.N + p.M;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.B * p.S * K) return;
    const int k = (int)(i % K);
    const long long bs = i / K;
    const int b = (int)(bs / p.S);
    const float D = p.depth[bs];
    float zz;
    if (k == 0) zz = D;
    else if (k <= p.N) {
        // torch.linspace(min_dist, D + delta, N): step = (end - start) / (N - 1); symmetric evaluation about the midpoint
        const float end = __fadd_rn(D, p.delta), start = p.min_dist;
        const float step = __fdiv_rn(__fsub_rn(end, start), (float)(p.N - 1));
        const int j = k - 1;
        zz = j < p.N / 2 ? __fadd_rn(start, __fmul_rn(step, (float)j)) : __fsub_rn(end, __fmul_rn(step, (float)(p.N - 1 - j)));
    } else zz = p.gauss[bs * p.M + (k - 1 - p.N)];
    const float* Kc = p.intr + (long long)b * 9;
    const float wn = __fdiv_rn(__fsub_rn((float)p.w_idx[bs], Kc[2]), Kc[0]);
    const float hn = __fdiv_rn(__fsub_rn((float)p.h_idx[bs], Kc[5]), Kc[4]);
    const float x = __fmul_rn(wn, zz), y = __fmul_rn(hn, zz);
    const float* T = p.pose + (long long)b * 16;
    float o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r] = __fadd_rn(__fmaf_rn(T[r * 4 + 2], zz, __fmaf_rn(T[r * 4 + 1], y, __fmul_rn(T[r * 4 + 0], x))), T[r * 4 + 3]);
    p.xyz[i * 3 + 0] = __fdiv_rn(o[0], o[3]);
    p.xyz[i * 3 + 1] = __fdiv_rn(o[1], o[3]);
    p.xyz[i * 3 + 2] = __fdiv_rn(o[2], o[3]);
    p.z[i] = zz;
}

// 4x4 inverse of [P; 0 0 0 1] in double precision (Gauss-Jordan with partial pivoting); returns the first 3 rows
static bool inverse_rows(const float* P12, float* M12) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = r < 3 ? (double)P12[r * 4 + c] : (c == 3 ? 1.0 : 0.0);
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if ((a[r][c] < 0 ? -a[r][c] : a[r][c]) > (a[piv][c] < 0 ? -a[piv][c] : a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < 8; ++k) { double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
        const double inv = 1.0 / a[c][c];
        for (int k = 0; k < 8; ++k) a[c][k] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                const double f = a[r][c];
                if (f != 0.0)
                    for (int k = 0; k < 8; ++k) a[r][k] -= f * a[c][k];
            }
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) M12[r * 4 + c] = (float)a[r][4 + c];
    return true;
}

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_get_3d_points(const float* depth, const float* h_projection, int B, int H, int W, float* out, void* stream) {
    GNB_CHECK_ARG(depth && h_projection && out && B >= 1 && H >= 1 && W >= 1, "gnb_get_3d_points: bad arguments");
    for (int b0 = 0; b0 < B; b0 += 8) {
        const int nb = B - b0 < 8 ? B - b0 : 8;
        UnprojectKP kp;
        for (int b = 0; b < nb; ++b)
            GNB_CHECK_ARG(inverse_rows(h_projection + (long long)(b0 + b) * 12, kp.M[b]), "gnb_get_3d_points: singular projection %d", b0 + b);
        const long long n = (long long)nb * H * W;
        unproject_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(depth, kp, b0, nb, H, W, out);
        GNB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int gnb_farthest_point_sample(const float* xyz, int B, int64_t N, int npoint, const int64_t* start, float* scratch,
                                         int64_t* out_idx, float* out_xyz, void* stream) {
    GNB_CHECK_ARG(xyz && start && scratch && out_idx && out_xyz, "gnb_farthest_point_sample: null pointer");
    GNB_CHECK_ARG(B >= 1 && N >= 1 && N < 0x7fffffffLL / 3 && npoint >= 1, "gnb_farthest_point_sample: bad shape");
    // cluster kernel when a cloud fits the shared memory + registers of at most 16 CTAs, else one CTA per cloud.
    // Cluster size: the iteration time is ~2.1 us of exchange / reduction latency + ~0.063 ns per point of a CTA's slice
    // (measured: 2.41 us at 4 800, 3.32 us at 19 200 points per CTA), and only a few 16-CTA clusters are co-resident
    // (~6 on a B200), so with many clouds a smaller cluster that lets every cloud run in the first wave wins.
    if (!opt(OPT_FPS_SINGLE_CTA)) {
        GNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int best_cs = 0;
        double best_t = 1e30;
        // grid kernel (co-operative launch, exchange through L2): as many clouds per launch as fit the GPU with the fewest CTAs a
        // cloud needs, the CTAs of a wave spread evenly over its clouds; several launches when the batch needs more than one wave
        int grid_cpc = 0, grid_per_wave = 0;
        double grid_t = 1e30;
        {
            int dev = 0, sms = 0, coop = 0;
            GNB_CUDA(cudaGetDevice(&dev));
            GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            GNB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
            const long long by_smem = (N * 12 + 225 * 1024 - 1) / (225 * 1024), by_regs = (N + (long long)FPS_PPT * FPS_THREADS - 1) / ((long long)FPS_PPT * FPS_THREADS);
            long long cpc_min = by_smem > by_regs ? by_smem : by_regs;
            if (cpc_min < 2) cpc_min = 2;
            const int forced = opt(OPT_FPS_GRID);             // tuning aid: > 0 forces the CTAs per cloud, < 0 switches the grid kernel off
            if (forced > 0 && forced >= cpc_min) cpc_min = forced;
            if (coop && forced >= 0 && cpc_min <= FPS_MAX_CPC && cpc_min <= sms && N < (1 << 24) - 1 &&
                (long long)N * 4 >= 2LL * FPS_MAX_CPC * FPS_XWORDS * 4) {                       // (the exchange slots live in `scratch`)
                int per_wave = sms / (int)cpc_min;
                if (per_wave > B) per_wave = B;
                const int waves = (B + per_wave - 1) / per_wave;
                per_wave = (B + waves - 1) / waves;
                int cpc = forced > 0 ? (int)cpc_min : sms / per_wave;
                if (cpc > FPS_MAX_CPC) cpc = FPS_MAX_CPC;
                const long long chunk = (N + cpc - 1) / cpc;
                grid_cpc = cpc, grid_per_wave = per_wave;
                grid_t = forced > 0 ? 0.0 : waves * (2.6 + 0.063e-3 * (double)chunk);          // (us per iteration; exchange latency measured on a B200)
            }
        }
        for (int cs = 2; cs <= 16; cs *= 2) {
            const long long chunk = (N + cs - 1) / cs;
            const size_t smem = (size_t)chunk * 12;
            if (chunk > (long long)FPS_PPT * FPS_THREADS || smem > 225 * 1024) continue;
            GNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(B * cs)), cfg.blockDim = dim3(FPS_THREADS), cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr, cfg.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, fps_cluster_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
                (void)cudaGetLastError();
                continue;
            }
            const int waves = (B + max_clusters - 1) / max_clusters;
            const double t = waves * (2.1 + 0.063e-3 * (double)chunk);
            if (t < best_t * 0.97) best_t = t, best_cs = cs;           // ties go to the smaller cluster
        }
        if (const int e = opt(OPT_FPS_CLUSTER)) best_cs = e, grid_cpc = 0;    // tuning aid
        if (grid_cpc > 1 && N > 2048 && grid_t < best_t * 0.97) {
            const int cpc = grid_cpc;
            int chunk = (int)((N + cpc - 1) / cpc);
            const size_t smem = (size_t)chunk * 12;
            GNB_CUDA(cudaFuncSetAttribute(fps_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            GNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fps_grid_kernel, FPS_THREADS, smem));
            if (per_sm >= 1) {
                unsigned* xbuf = reinterpret_cast<unsigned*>(scratch);
                GNB_CUDA(cudaMemsetAsync(xbuf, 0, (size_t)B * 2 * cpc * FPS_XWORDS * 4, (cudaStream_t)stream));
                long long NN = N;
                const long long* st = (const long long*)start;
                long long* oi = (long long*)out_idx;
                bool ok = true;
                for (int b0 = 0; b0 < B && ok; b0 += grid_per_wave) {
                    const int nb = B - b0 < grid_per_wave ? B - b0 : grid_per_wave;
                    void* args[] = {(void*)&xyz, (void*)&NN, (void*)&npoint, (void*)&st, (void*)&cpc, (void*)&chunk, (void*)&b0, (void*)&xbuf,
                                    (void*)&oi, (void*)&out_xyz};
                    ok = cudaLaunchCooperativeKernel((const void*)fps_grid_kernel, dim3((unsigned)(nb * cpc)), dim3(FPS_THREADS), args, smem,
                                                     (cudaStream_t)stream) == cudaSuccess;
                    if (!ok && b0 > 0) GNB_CUDA(cudaGetLastError());       // (a later wave cannot fail when the first one fitted)
                }
                if (ok) return 0;
                (void)cudaGetLastError();                     // (e.g. not all CTAs co-resident right now: take the cluster kernel)
            }
        }
        if (best_cs > 1 && N > 2048) {
            const int cs = best_cs;
            const long long chunk = (N + cs - 1) / cs;
            const size_t smem = (size_t)chunk * 12;
            GNB_CHECK_ARG(chunk <= (long long)FPS_PPT * FPS_THREADS && smem <= 225 * 1024, "gnb_farthest_point_sample: cluster size %d too small for %lld points", cs, (long long)N);
            GNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(B * cs));
            cfg.blockDim = dim3(FPS_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr, cfg.numAttrs = 1;
            GNB_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel, xyz, (long long)N, npoint, (const long long*)start, cs, (int)chunk,
                                        (long long*)out_idx, out_xyz));
            return 0;
        }
    }
    fps_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(xyz, N, npoint, (const long long*)start, scratch, (long long*)out_idx, out_xyz);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_sample_points_on_rays(const int64_t* h_idxs, const int64_t* w_idxs, const float* depths, const float* intrinsics,
                                         const float* poses, const float* gaussian_depths, int B, int S, int N, int M, float delta,
                                         float min_dist, float* xyz_world, float* z, void* stream) {
    GNB_CHECK_ARG(B >= 0 && S >= 0 && N >= 2 && M >= 0, "gnb_sample_points_on_rays: bad shape (N >= 2 stratified samples)");
    if (B == 0 || S == 0) return 0;
    GNB_CHECK_ARG(h_idxs && w_idxs && depths && intrinsics && poses && xyz_world && z && (M == 0 || gaussian_depths),
                  "gnb_sample_points_on_rays: null pointer");
    RayKP kp;
    kp.h_idx = (const long long*)h_idxs, kp.w_idx = (const long long*)w_idxs, kp.depth = depths, kp.intr = intrinsics, kp.pose = poses;
    kp.gauss = gaussian_depths, kp.B = B, kp.S = S, kp.N = N, kp.M = M, kp.delta = delta, kp.min_dist = min_dist;
    kp.xyz = xyz_world, kp.z = z;
    const long long n = (long long)B * S * (1 + N + M);
    ray_points_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// sample_valid_depth_pixels (reference src/models/utils.py:340-363): per depth map, `argwhere(depth != 0)` (row-major
// list of the valid pixels, 16 bytes each), `randperm(n_valid)[:S]`, gather.  Here the list is never built: a count
// pass gives every image row's number of valid pixels and their exclusive prefix (and n_valid, which the caller needs
// on the host for the reference's randperm call -- the same synchronisation the reference has in argwhere); the select
// pass finds the k-th valid pixel of the map for each of the S drawn ranks k: binary search over the row prefix, then
// one warp walks the row 32 pixels at a time with ballot / popc.  Same (h, w) as the reference for the same ranks.
// ---------------------------------------------------------------------------------------------------------------
namespace gnb {

// grid (H, B): block = one image row; row_count[b][h]
__global__ void valid_row_count_kernel(const float* __restrict__ depth, int H, int W, int* __restrict__ row_count) {
    const int h = blockIdx.x, b = blockIdx.y;
    const float* row = depth + ((long long)b * H + h) * W;
    int n = 0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) n += (__ldg(row + w) != 0.0f) ? 1 : 0;
    n = __reduce_add_sync(FULL, n);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
        row_count[b * H + h] = t;
    }
}

// one block per map: exclusive prefix over the H rows (in place), total to n_valid[b]
__global__ void valid_row_scan_kernel(int* __restrict__ row_count, int H, int* __restrict__ n_valid) {
    const int b = blockIdx.x;
    int* rc = row_count + (long long)b * H;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int h0 = 0; h0 < H; h0 += blockDim.x) {
        const int h = h0 + threadIdx.x;
        const int v = h < H ? rc[h] : 0;
        // block-wide inclusive scan (blockDim.x = 256: warp scans + scan of the 8 warp totals)
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        __shared__ int wt[8];
        if ((threadIdx.x & 31) == 31) wt[threadIdx.x >> 5] = x;
        __syncthreads();
        int base = carry;
        for (int i = 0; i < (int)(threadIdx.x >> 5); ++i) base += wt[i];
        if (h < H) rc[h] = base + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_valid[b] = carry;
}

// one warp per (map, sample): rank k -> (h, w) of the k-th valid pixel in row-major order
__global__ void valid_select_kernel(const float* __restrict__ depth, int B, int H, int W, const int* __restrict__ row_prefix,
                                    const int* __restrict__ n_valid, const long long* __restrict__ rank, int S,
                                    long long* __restrict__ h_out, long long* __restrict__ w_out) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)B * S) return;
    const int b = (int)(wid / S);
    const long long k = rank[wid];
    const int* rp = row_prefix + (long long)b * H;
    if (k < 0 || k >= n_valid[b]) {                       // caller error (rank outside the valid list)
        if (lane == 0) h_out[wid] = -1, w_out[wid] = -1;
        return;
    }
    int lo = 0, hi = H - 1;                                // last row whose prefix <= k
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (rp[mid] <= k) lo = mid; else hi = mid - 1;
    }
    int need = (int)(k - rp[lo]);                          // the need-th valid pixel of row lo
    const float* row = depth + ((long long)b * H + lo) * W;
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        const unsigned m = __ballot_sync(FULL, w < W && __ldg(row + w) != 0.0f);
        const int c = __popc(m);
        if (need < c) {
            // position of the need-th set bit
            unsigned mm = m;
            for (int i = 0; i < need; ++i) mm &= mm - 1;
            if (lane == 0) h_out[wid] = lo, w_out[wid] = w0 + (__ffs(mm) - 1);
            return;
        }
        need -= c;
    }
}

}  // namespace gnb

extern "C" int gnb_valid_pixel_count(const float* depth, int B, int H, int W, int32_t* row_prefix, int32_t* n_valid, void* stream) {
    GNB_CHECK_ARG(B >= 0 && H > 0 && W > 0, "gnb_valid_pixel_count: bad shape");
    if (B == 0) return 0;
    GNB_CHECK_ARG(depth && row_prefix && n_valid, "gnb_valid_pixel_count: null pointer");
    gnb::valid_row_count_kernel<<<dim3((unsigned)H, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(depth, H, W, row_prefix);
    GNB_LAUNCH_CHECK();
    gnb::valid_row_scan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(row_prefix, H, n_valid);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_valid_pixel_select(const float* depth, int B, int H, int W, const int32_t* row_prefix, const int32_t* n_valid,
                                      const int64_t* rank, int S, int64_t* h_idxs, int64_t* w_idxs, void* stream) {
    GNB_CHECK_ARG(B >= 0 && H > 0 && W > 0 && S >= 0, "gnb_valid_pixel_select: bad shape");
    if (B == 0 || S == 0) return 0;
    GNB_CHECK_ARG(depth && row_prefix && n_valid && rank && h_idxs && w_idxs, "gnb_valid_pixel_select: null pointer");
    const long long threads = (long long)B * S * 32;
    gnb::valid_select_kernel<<<gnb::ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        depth, B, H, W, row_prefix, n_valid, (const long long*)rank, S, (long long*)h_idxs, (long long*)w_idxs);
    GNB_LAUNCH_CHECK();
    return 0;
}

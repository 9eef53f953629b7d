// Microbenchmark: round-trip latency of a flag ping-pong between two CTAs on different SMs through global memory (L2), for the
// store / load flavours a grid-wide FPS exchange could use.   nvcc -arch=sm_100a -o flag_pingpong flag_pingpong.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void put(unsigned* p, unsigned v) {
    if (MODE == 0) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 2) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 3) asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 4) { unsigned o; asm volatile("atom.relaxed.gpu.global.exch.b32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); }
    if (MODE == 5) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned get(unsigned* p) {
    unsigned v;
    if (MODE == 0) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 1 || MODE == 3 || MODE == 4) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 5) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int MODE>
__global__ void pingpong(unsigned* flags, int iters, long long* cycles, int nthreads_spin) {
    // block 0 writes flags[0] = i, waits flags[32] == i; block (gridDim-1) mirrors.  Other blocks exit at once.
    if (blockIdx.x != 0 && blockIdx.x != gridDim.x - 1) return;
    const bool a = blockIdx.x == 0;
    unsigned* mine = flags + (a ? 0 : 32), *theirs = flags + (a ? 32 : 0);
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int i = 1; i <= iters; ++i) {
            if (a) { put<MODE>(mine, i); while (get<MODE>(theirs) != (unsigned)i) {} }
            else { while (get<MODE>(theirs) != (unsigned)i) {} put<MODE>(mine, i); }
        }
    }
    __syncthreads();
    if (a && threadIdx.x == 0) *cycles = clock64() - t0;
}
int main() {
    unsigned* flags; long long* cyc;
    cudaMalloc(&flags, 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"st.volatile / ld.volatile", "st.relaxed.gpu / ld.relaxed.gpu", "st.release.gpu / ld.acquire.gpu", "red.max / ld.relaxed.gpu", "atom.exch / ld.relaxed.gpu", "st.wt / ld.cv"};
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int threads : {32, 1024}) {
    for (int mode = 0; mode < 6; ++mode) {
        cudaMemset(flags, 0, 1024);
        const int iters = 2000;
        void* args[] = {&flags, (void*)&iters, &cyc, (void*)&threads};
        const void* f = mode == 0 ? (const void*)pingpong<0> : mode == 1 ? (const void*)pingpong<1> : mode == 2 ? (const void*)pingpong<2> : mode == 3 ? (const void*)pingpong<3> : mode == 4 ? (const void*)pingpong<4> : (const void*)pingpong<5>;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchCooperativeKernel(f, dim3(148), dim3(threads), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%4d threads  %-34s: %7.1f ns per round trip (%.0f cycles)  [%s]\n", threads, names[mode], ms * 1e6 / iters, (double)c / iters, cudaGetErrorString(e));
    }}
    return 0;
}

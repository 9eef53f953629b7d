// tcgen05 / TMEM tensor-core decoder (fp16 or bf16 operands, fp32 accumulate) with the point-query
// sampler fused into its prologue (sm_100a).
//
// Replaces GenNerf.forward (reference src/models/model.py:207-248) in ONE kernel:
//   map_features (sampler, sample.cuh)                      model.py:163-204
//   PositionalEncoding                                      positional_encoding.py:28-40
//   ResnetFC: lin_in, n_blocks x [lin_z, fc_0, fc_1], lin_out   resnetfc.py:134-189 (trap T8)
//   TSDFHeadSimple                                          heads3d.py:36-50
//
// Tile: 128 query rows per cluster.  Hidden width d_hidden is split over NSPLIT = 1 or 2 CTAs
// (2 when d_hidden > 256): CTA r owns hidden units [r*HN, (r+1)*HN), HN <= 256, because the
// fp32 residual stream x (128 x HN) must stay resident in TMEM next to the fc_0 output
// (128 x HN): 2*HN <= 512 columns.
//
// Per CTA:
//   TMEM   cols [0,HN)      x    residual stream, fp32, accumulated in place by lin_in, lin_z, fc_1
//          cols [256,256+HN) net  fc_0 output; reused for the lin_out tile
//   SMEM   A buffer         128 x K bf16 activations, K-major, 128B-swizzled 64-column chunks
//          code tile        positional encoding (+ two constant-1 columns that carry biases)
//          W ring           NSTAGE x 16 KB weight tiles (<=128 n-rows x 64 k), pre-swizzled in HBM,
//                           streamed with 1-D bulk async copies (no tensor map needed)
//   warps  0: weight producer   1: MMA issuer   2: TMEM alloc + A-chunk exchange   3: arrival forwarder (2-CTA)
//          4-11: epilogue / prologue, two warps per TMEM lane quadrant (thread = query row = TMEM lane)
//
// Dataflow per tile: prologue samples features + encodes xyz -> bf16 operand tiles; then MMA groups
//   G0 = [lin_in, lin_z_0] -> x;  per block: E(relu(x)) -> [fc_0] -> net;  E(relu(net+b0)) ->
//   [fc_1, lin_z_next] -> x;  finally E(relu(x+b1_last)) -> [lin_out] -> epilogue (+b_out, tanh head).
// An epilogue round converts one TMEM accumulator into the next layer's A operand chunk by
// chunk; each chunk has its own mbarrier, so the MMA of a layer starts as soon as its first
// chunk is ready.  With NSPLIT=2 each CTA produces half of the chunks and pushes them to its
// peer with a bulk shared::cta -> shared::cluster copy that completes on the peer's barrier.
// lin_z biases (with lin_in's folded into lin_z_0's and fc_1's into the next lin_z) ride in two extra K columns of the
// code tile as a 16-bit hi+lo pair; fc_0 / lin_out / last fc_1 biases are added in fp32 by the epilogue.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "sample.cuh"

namespace gnb {

int check_decoder_weights(const GnbDecoderWeights* w, const char* who);

namespace tc {

constexpr int BM = 128;               // rows per tile
constexpr int CHUNK = BM * 128;       // bytes of one 128-row x 64-col bf16 chunk
constexpr int NET_COL = 256;          // TMEM column of the net / out accumulator
constexpr int MAX_CHUNKS = 16;
constexpr int MAX_STAGES = 8;
constexpr int NTHREADS = 384;
constexpr int EPI_WARP0 = 4;          // epilogue warps 4..11: two per TMEM lane quadrant (warp & 3)
constexpr int ROLE_WARP0 = 0;         // warp 0 weight producer, 1 MMA issuer, 2 TMEM alloc + chunk exchange, 3 forwarder
constexpr int EPI_GROUPS = 2;         // group g = (warp - 4) / 4 converts columns [32g, 32g+32) of every own chunk

struct Dims {
    int d_feat, d_code, Hd, nb, d_out, d_geo;
    int nsplit, HN;                   // N-halves of the hidden width (1 or 2), hidden units (TMEM columns) per half
    int two;                          // 1: cta_group::2 pairs (256 query rows per cluster, each CTA loads half of B)
    int csize;                        // CTAs per cluster = nsplit * (two ? 2 : 1)
    int WN;                           // weight rows one CTA streams per k-chunk of a hidden layer = HN / (two ? 2 : 1)
    int NOUTC;                        // lin_out weight rows per CTA
    int KF, KZ, KH;                   // 64-wide K chunks of lin_in, lin_z, hidden layers
    int NOUT;                         // lin_out rows padded to a multiple of 16
    int OWN;                          // activation chunks this CTA produces per layer = HN / 64
    int AOWN;                         // A-buffer slots for own chunks (and the lin_in operand) = max(KF, OWN)
    int RS;                           // A-buffer slots cycled through by the chunks the peer pushes (0, 1 or 2)
    int ACH;                          // A-buffer slots in total = AOWN + RS
    int nstage;                       // weight-ring stages
    int stage_bytes;                  // bytes of one stage: the weight rows of one step x 64 k (single-CTA issue: one k-chunk
                                      // of <= 256 rows; cta_group::2: two k-chunks of <= 128 rows)
    int early;                        // 1: a dedicated warp stages the NEXT tile's lin_in / lin_z operands (sampling, encoding) while
                                      // this tile's layers run; needs its own feature buffer (KFB chunks) next to the code tile
    int wide;                         // 1: lin_in has more k-chunks than shared memory can hold next to the layers' operands (KF > 8, e.g.
                                      // the reference's default latent 512 + 32): operand-image mode only, the chunks are streamed --
                                      // chunk 0 waits in the feature buffer, the others pass through the tile's own activation slots
    int KFB;                          // chunks of the feature buffer (KF, or 1 in wide mode)
    long long packed_per_rank;        // bytes
};

__host__ __device__ inline int n_tiles_of(int rows) { return (rows + 127) / 128; }

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.  try_wait parks the thread in
// hardware between polls, so waiting roles do not take issue slots from the warps that have work.
// Before trapping, the first thread to time out leaves {1, block, thread, source line, barrier, parity} in the
// host-mapped report buffer (gnb_debug_hang_report), which stays readable after the context has died.
// Suspend-time hint of mbarrier.try_wait (ns; 0 = none, the default).  Without a hint a waiting warp comes back from the hardware
// wait every ~80 cycles and re-issues its poll loop: ncu counts 21 M polls x 10 instructions of the 8 epilogue warps on acc_ready
// alone, a quarter of all instructions the kernel executes.  With CUTLASS's 0x989680 the polls all but vanish -- and the kernel
// gets SLOWER (1 Mi rows 6.52 -> 6.66 ms, the bench's query phase 114.9 -> 116.3 ms, A/B on one box): a suspended warp wakes up
// later than a polling one notices the phase flip, and every layer waits on exactly those wake-ups.  Kept as a build switch.
#ifndef GNB_TRYWAIT_HINT
#define GNB_TRYWAIT_HINT 0
#endif
__device__ int* g_hang_report = nullptr;
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int line) {
    int* r = g_hang_report;
    if (r) {
        const int i = atomicAdd(r, 1);
        if (i < 160) {
            int* e = r + 8 + i * 6;
            e[0] = (int)blockIdx.x, e[1] = (int)threadIdx.x, e[2] = line, e[3] = (int)bar, e[4] = (int)parity, e[5] = 0;
            __threadfence_system();
        }
    }
    for (int k = 0; k < 200; ++k) __nanosleep(1000000);      // let the other stuck threads report too
    __trap();
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity, int line) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
#if GNB_TRYWAIT_HINT
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(GNB_TRYWAIT_HINT)
            : "memory");
        if (!done && spin > (1u << 22)) mbar_timeout(bar, parity, line);
    }
}
#define mbar_wait(bar, parity) mbar_wait_((bar), (parity), __LINE__)
// Parking wait for a whole warp (epilogue): lane 0 waits, the others sleep at the warp barrier.
#define mbar_wait_park(bar, parity) mbar_wait_((bar), (parity), __LINE__)
// Whole-warp wait with warp-uniform control flow and ONE polling lane.  An mbarrier query is a per-thread
// operation on one shared-memory word: a full warp polling costs ~32 serialised queries (measured ~390 cycles per
// wait even on a completed barrier), and 256 epilogue threads polling slow every other waiter of the SM down.
// Lane 0 queries, the vote makes the result (and the loop) uniform, so operands of the following tcgen05
// instructions stay in uniform registers.
#define mbar_wait_warp(bar, parity) mbar_wait_((bar), (parity), __LINE__)
// Two barriers at once with both queries in flight together (one query costs ~180 cycles of latency on the issuing
// thread even when the phase is complete, tools/microbench/mma_rate.cu) -- what an MMA issuer needs per k-step:
// "A chunk ready" and "weight stage landed".
__device__ __forceinline__ void mbar_wait2_(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b, int line) {
    uint32_t da = 0, db = 0;
    for (uint32_t spin = 0; !(da & db); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
#if GNB_TRYWAIT_HINT
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3, %6;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5, %6;\n\t"
#else
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
#endif
            "selp.u32 %0, 1, 0, p;\n\t"
            "selp.u32 %1, 1, 0, q;\n\t}"
            : "=r"(da), "=r"(db)
            : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b), "r"(GNB_TRYWAIT_HINT)
            : "memory");
        if (!(da & db) && spin > (1u << 22)) mbar_timeout(da ? bar_b : bar_a, da ? par_b : par_a, line);
    }
}
#define mbar_wait2(a, pa, b, pb) mbar_wait2_((a), (pa), (b), (pb), __LINE__)
__device__ __forceinline__ void mbar_wait2_warp(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
    mbar_wait(bar_a, par_a);
    mbar_wait(bar_b, par_b);
}
// Up to three barriers (bar == 0: skip), all queries in flight together.
__device__ __forceinline__ void mbar_wait3_warp(uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, uint32_t b2, uint32_t p2) {
    if (b0) mbar_wait(b0, p0);
    if (b1) mbar_wait(b1, p1);
    mbar_wait(b2, p2);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// global -> own shared memory, completion on an own mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// own shared memory -> peer CTA's shared memory, completion on the PEER's mbarrier
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

// K-major, 128B-swizzled operand tile (rows x 64 bf16, 8-row groups 1024 B apart), sm_100 format:
// start>>4 [0,14) | LBO=1 [16,30) | SBO=1024>>4 [32,46) | version=1 [46,48) | SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A format [7,10) and B format [10,13) (0 = f16,
// 1 = bf16), K-major both, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int n, bool bf16, int m = BM) {
    const uint32_t f = bf16 ? 1u : 0u;
    return (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One 64-wide k-step = four K16 MMAs in ONE asm block: the descriptors of the 2nd..4th are derived inside (+32 bytes of
// K each, i.e. +2 in the 16-byte address field), so the issuing thread moves one set of operands to uniform registers
// per step instead of one per MMA.
__device__ __forceinline__ void umma_f16_x4(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, 1;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// Warp-converged issue.  tcgen05.mma / tcgen05.commit are uniform-datapath instructions (UTCHMMA / UTCBAR take uniform
// registers and issue once per warp).  Inside a divergent `if (lane == 0)` region ptxas cannot prove that the operands are
// warp-uniform and wraps EVERY such instruction in an "ELECT ... R2UR ... BRA.U.ANY" serialisation loop (~80 cycles per MMA
// on the issuing thread: tools/microbench/umma_pair_test.cu measures 775 cycles for 8 MMAs + commit issued by one thread
// against 512 -- the tensor pipe's own time -- from a converged warp).  So the issuing warp stays converged, computes its
// operands with all 32 lanes and elects the issuing lane inside the asm block.
__device__ __forceinline__ void umma_f16_x4_u(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, 1;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_u(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc_u(uint32_t bar, uint16_t mask) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// bulk copies from a converged warp (UBLKCP is a uniform-datapath instruction as well): expect_tx + copy by one elected lane
__device__ __forceinline__ void bulk_g2s_u(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
                 "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2peer_u(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
// 2-CTA variants: one instruction drives the tensor cores of both SMs of the pair (M = 256: 128 rows from
// each CTA's A tile, each CTA supplies half of the N rows of B); commits can signal any set of CTAs
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc_2cta(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// cta_group::2 from a converged warp (see umma_f16_x4_u)
__device__ __forceinline__ void umma_f16_x4_2cta_u(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, 1;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc_2cta_u(uint32_t bar, uint16_t mask) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (address from map_to_cta)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two fp32 -> one packed pair of 16-bit operands (fp16 saturates at +-65504 instead of overflowing)
template <bool BF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
    if constexpr (BF16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    } else {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
    }
}
// relu + (fp16: saturate) + round + pack in ONE F2FP instruction
template <bool BF16>
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi) {
    uint32_t r;
    if constexpr (BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool BF16>
__device__ __forceinline__ float round16(float a) {
    if constexpr (BF16) return __bfloat162float(__float2bfloat16_rn(a));
    else return __half2float(__float2half_rn(fminf(fmaxf(a, -65504.0f), 65504.0f)));
}
// fp16 saturation probe on packed operands: a half is +-65504 (0x7BFF / 0xFBFF) exactly when satfinite clipped it (or the
// value was that number, which is as good as clipped); |h| + 0x0401 then carries into bit 15.  2 integer ops per 4 halves.
__device__ __forceinline__ uint32_t sat_probe(uint32_t a, uint32_t b) {
    return (((a & 0x7FFF7FFFu) + 0x04010401u) | ((b & 0x7FFF7FFFu) + 0x04010401u)) & 0x80008000u;
}
// byte offset of the 16-byte unit u (8 columns) of row r inside a swizzled chunk
__device__ __forceinline__ uint32_t chunk_off(int r, int u) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)); }

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct TcKP {
    Dims d;
    GnbDecoderWeights w;          // fp32 biases, head, encoding options (matrices are read from `packed`)
    const unsigned char* packed;
    const float* xyz;             // (n_rows,3), or (n_rows,d_code) codes when w.use_code == 2
    const float4* sorted;         // brick-sorted mode (xyz == null): row r is the record (x, y, z, query index); outputs go to the
                                  //   rows the indices name.  The staging warp then walks the volume brick by brick (L1 / L2 hits)
    const float* gaxes;           // dense-grid mode (xyz == null): the gnx + gny + gnz axis coordinates; query row r of a
    int gnx, gny, gnz;            //   scene is the grid point (x[i], y[j], z[k]), r = (i*gny + j)*gnz + k  (utils.py:926-935)
    const float* feat;            // (n_rows,d_feat) when !fused
    uint16_t* save;               // optional (training forward): the 16-bit activations every layer consumed, [2 nb + 1][n_rows][Hd]:
                                  //   slab 2i = relu(x_i + alpha lin_z_i), 2i + 1 = relu(fc_0 of block i), 2 nb = relu(x) into lin_out
    const unsigned char* feat_img;   // or (!fused) the 16-bit operand image of lin_in (GnbSampleParams.image): one bulk copy per tile
    int fused;
    SampleKP s;                   // sampler (fused); s.out = optional fp32 feature output
    long long n_rows;
    float* out;                   // (n_rows,d_out) or null
    float* tsdf;                  // (n_rows) or null
    int n_tiles, n_clusters;
    long long* dbg;               // optional per-phase clock64() trace of cluster 0 (profiling aid), or null
    int nocopy;                   // GNB_DEBUG_NO_WCOPY (timing experiment, WRONG results): weight stages are signalled but not copied
};

// shared-memory carve-up (offsets from the 1024-aligned base)
struct Smem {
    uint32_t a, code, feat, ring, bias, bars;     // byte offsets
    uint32_t total;
};
__host__ __device__ inline Smem smem_layout(const Dims& d) {
    Smem s;
    s.a = 0;
    s.code = s.a + d.ACH * CHUNK;
    s.feat = d.early ? s.code + d.KZ * CHUNK : s.a;          // lin_in operand: own buffer, or the first own chunk slots
    s.ring = s.code + d.KZ * CHUNK + (d.early ? d.KFB * CHUNK : 0);
    s.bias = s.ring + d.nstage * d.stage_bytes;
    // fp32 table: bias[HN] | b_out[NOUT] | head_w[d_geo] | head_b.  bias[] is time-shared: the epilogue threads load
    // fc_0's bias of block i into it during round 2i (which adds no bias) for round 2i+1, and the last fc_1 bias
    // after round 2nb-1 -- a table of all blocks' biases (5 KB at d_hidden 512) is shared memory the operand buffers need
    uint32_t nbias = (uint32_t)(d.HN + d.NOUT + d.d_geo + 1);
    s.bars = (s.bias + nbias * 4 + 15) & ~15u;
    // barriers: w_full[8] w_empty[8] a_ready[8] rready[4] rfree[4] acc_ready in_ready (12 spare) | tmem slot
    s.total = s.bars + (2 * MAX_STAGES + MAX_CHUNKS + 2 + 12) * 8 + 16;
    return s;
}

// The per-tile program: every role (and the host-side packer) walks the same sequence of GEMM ops.
//   kind 0 lin_in (A = feature chunks), 1 lin_z (A = code chunks), 2 hidden (A = activation chunks)
// Order: lin_in, lin_z_0 | fc0_0 | lin_z_1, fc1_0 | fc0_1 | lin_z_2, fc1_1 | ... | fc0_last | fc1_last | lin_out
// ("|" = end of an accumulation group: the epilogue converts the accumulator).  lin_z_{i+1} is issued BEFORE fc1_i:
// its operand (the code tile) is always there, so its MMAs run while the epilogue is still converting fc0_i's
// output -- the tensor pipe has work during that layer-to-layer bubble.  x is free at that point: the round that
// read x (the operand of fc0_i) had written all its chunks before fc0_i's MMAs could be issued.
enum : int { M_LIN_IN = 0, M_LIN_Z = 1, M_FC0 = 2, M_FC1 = 3, M_LIN_OUT = 4 };
struct Op {
    int a_kind, kchunks, rows, d_col, first_overwrites, group_end;
    int mat, blk;                     // which matrix (M_*) of which block
};
__host__ __device__ inline int num_ops(const Dims& d) { return 2 + 3 * d.nb; }
__host__ __device__ inline Op get_op(const Dims& d, int o) {
    if (o == 0) return Op{0, d.KF, d.WN, 0, 1, 0, M_LIN_IN, 0};
    if (o == 1) return Op{1, d.KZ, d.WN, 0, 0, 1, M_LIN_Z, 0};
    if (o == 1 + 3 * d.nb) return Op{2, d.KH, d.NOUTC, NET_COL, 1, 1, M_LIN_OUT, 0};
    const int q = o - 2;
    int i, j;                         // j: 0 fc0_i, 1 lin_z_{i+1}, 2 fc1_i
    if (q < 3 * (d.nb - 1)) { i = q / 3, j = q % 3; }
    else { i = d.nb - 1, j = (q - 3 * (d.nb - 1)) == 0 ? 0 : 2; }
    if (j == 0) return Op{2, d.KH, d.WN, NET_COL, 1, 1, M_FC0, i};
    if (j == 1) return Op{1, d.KZ, d.WN, 0, 0, 0, M_LIN_Z, i + 1};
    return Op{2, d.KH, d.WN, 0, 0, 1, M_FC1, i};
}
// k-chunk visiting order of an activation op on N-half `half`: own chunk 0, the peer's chunk 0, own 1, peer 1, ...
// (the peer's r-th chunk lands about one copy latency after the own r-th is written, so the MMAs never wait
// for a whole half).  Returns the global 64-wide k-chunk index; t even = own, t odd = pushed by the peer.
__host__ __device__ inline int act_kchunk(int nsplit, int own, int half, int t, int interleave) {
    if (nsplit == 1) return t;
    if (!interleave) return (half * own + t) % (2 * own);          // own chunks first, then the peer's
    const int r = t >> 1;
    return (t & 1) ? ((half ^ 1) * own + r) : (half * own + r);
}

// Operand tiles of lin_in (features + bias columns) and lin_z (positional code + bias columns) of ONE query row:
// writes the 16-byte units u0, u0+ustep, ... of the row into the swizzled code and feature chunks.  Features are
// sampled here (fused query) or read from the feature tensor.
template <bool BF16>
__device__ __forceinline__ void stage_inputs(const TcKP& p, unsigned char* sm, const Smem& L, int half, int row, long long grow,
                                             int u0, int ustep, int what, uint32_t& ovf,       // what: 1 = code tile, 2 = feature chunks
                                             int chunk_stride = CHUNK) {                          // bytes between 64-column chunks of a tile
    const Dims& d = p.d;
    const GnbDecoderWeights& w = p.w;
    const bool live = grow < p.n_rows;
    float xyz3[3] = {0.f, 0.f, 0.f};
    long long orow = grow;                       // row of the outputs (differs from the processing row in sorted mode)
    if (live && w.use_code != 2) {
        if (p.sorted) {
            const float4 rec = __ldg(p.sorted + grow);
            xyz3[0] = rec.x, xyz3[1] = rec.y, xyz3[2] = rec.z;
            orow = (long long)__float_as_int(rec.w);
        } else if (p.gaxes) {            // the query grid is never materialised: derive the point from the row index
            const long long r = grow % ((long long)p.gnx * p.gny * p.gnz);
            const int k = (int)(r % p.gnz), j = (int)((r / p.gnz) % p.gny), i = (int)(r / ((long long)p.gnz * p.gny));
            xyz3[0] = __ldg(p.gaxes + i), xyz3[1] = __ldg(p.gaxes + p.gnx + j), xyz3[2] = __ldg(p.gaxes + p.gnx + p.gny + k);
        } else {
            xyz3[0] = __ldg(p.xyz + grow * 3), xyz3[1] = __ldg(p.xyz + grow * 3 + 1), xyz3[2] = __ldg(p.xyz + grow * 3 + 2);
        }
    }
    if (what & 1) {
    float code[64 * 4];                          // d_code + 2 <= 64*KZ (KZ <= 4)
    const int kz = d.KZ * 64;
    for (int k = 0; k < kz; ++k) code[k] = 0.0f;
    if (live) {
        if (w.use_code == 2) {
            for (int k = 0; k < d.d_code; ++k) code[k] = __ldg(p.xyz + grow * d.d_code + k);
        } else {
            if (w.use_code == 0) { code[0] = xyz3[0], code[1] = xyz3[1], code[2] = xyz3[2]; }
            else {
                int o = 0;
                if (w.include_input) { code[0] = xyz3[0], code[1] = xyz3[1], code[2] = xyz3[2]; o = 3; }
                const float half_pi = (float)(3.14159265358979323846 * 0.5);
                for (int f = 0; f < 2 * w.num_freqs; ++f) {
                    const float freq = w.freq_factor * exp2f((float)(f >> 1));
                    const float phase = (f & 1) ? half_pi : 0.0f;
                    for (int dd = 0; dd < 3; ++dd) code[o + f * 3 + dd] = sinf(__fadd_rn(phase, __fmul_rn(xyz3[dd], freq)));
                }
            }
        }
    }
    code[d.d_code] = 1.0f, code[d.d_code + 1] = 1.0f;      // bias hi / lo columns
    for (int c = 0; c < d.KZ; ++c)
        for (int u = u0; u < 8; u += ustep) {
            const float* v = code + c * 64 + u * 8;
            uint4 pk = make_uint4(pack16<BF16>(v[0], v[1]), pack16<BF16>(v[2], v[3]), pack16<BF16>(v[4], v[5]), pack16<BF16>(v[6], v[7]));
            if constexpr (!BF16) ovf |= sat_probe(pk.x, pk.y) | sat_probe(pk.z, pk.w);
            *reinterpret_cast<uint4*>(sm + L.code + c * chunk_stride + chunk_off(row, u)) = pk;
        }
    }
    if (!(what & 2)) return;
    TriCorners tcn;
    BiCorners bc[3];
    const int b = (p.fused && live) ? (int)(orow / p.s.Q) : 0;
    if (p.fused && live) {
        if (p.s.volume) trilinear_setup(p.s, xyz3[0], xyz3[1], xyz3[2], tcn);
        if (p.s.Cp > 0) planes_setup(p.s, xyz3[0], xyz3[1], xyz3[2], bc);
    }
    for (int c = 0; c < d.KF; ++c)
        for (int u = u0; u < 8; u += ustep) {
            float v[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = c * 64 + u * 8 + h * 4;         // first of 4 feature columns
                float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && k < d.d_feat) {
                    if (!p.fused) {
                        const float* src = p.feat + grow * d.d_feat + k;
                        if ((d.d_feat & 3) == 0) f4 = ldg4(src);
                        else {
                            f4.x = __ldg(src);
                            if (k + 1 < d.d_feat) f4.y = __ldg(src + 1);
                            if (k + 2 < d.d_feat) f4.z = __ldg(src + 2);
                            if (k + 3 < d.d_feat) f4.w = __ldg(src + 3);
                        }
                    } else {
                        // (host guarantees C_p % 4 == 0, C % 4 == 0 and unit channel strides here)
                        Vals<4> r = (k < p.s.Cp) ? sample_planes<4>(p.s, bc, b, k) : sample_volume<4>(p.s, tcn, b, k - p.s.Cp);
                        f4 = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
                        if (p.s.out && half == 0) *reinterpret_cast<float4*>(p.s.out + orow * p.s.out_stride + k) = f4;
                    }
                }
                v[h * 4 + 0] = f4.x, v[h * 4 + 1] = f4.y, v[h * 4 + 2] = f4.z, v[h * 4 + 3] = f4.w;
            }
            uint4 pk = make_uint4(pack16<BF16>(v[0], v[1]), pack16<BF16>(v[2], v[3]), pack16<BF16>(v[4], v[5]), pack16<BF16>(v[6], v[7]));
            if constexpr (!BF16) ovf |= sat_probe(pk.x, pk.y) | sat_probe(pk.z, pk.w);
            *reinterpret_cast<uint4*>(sm + L.feat + c * chunk_stride + chunk_off(row, u)) = pk;
        }
}

// trace slot layout: dbg[role*4096 + k]; role 0 = MMA thread, 1 = epilogue warp 4 lane 0, 2 = producer
// (compiled in only with -DGNB_TC_TRACE: even a predicated clock read per k-step costs the single-thread MMA
//  issue loop ~10 %)
#ifdef GNB_TC_TRACE
#define GNB_TRACE(role, k) do { if (p.dbg && blockIdx.x == 0 && (k) < 4096) p.dbg[(role) * 4096 + (k)] = clock64(); } while (0)
#define GNB_TRACE_ON(p) ((p).dbg != nullptr)
#else
#define GNB_TRACE(role, k) do { } while (0)
#define GNB_TRACE_ON(p) false
#endif

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// TWO = cta_group::2: ranks (2i, 2i+1) of the cluster form an MMA pair working on 256 query rows (128 each);
// rank>>1 selects the N-half of the hidden width, the even rank ("leader") issues every MMA of the pair.
// STG: four extra warps (12..15) stage the next tile's inputs, 32 query rows each, instead of warp 3 alone.  One warp is a
// chain of latency-bound gather batches (8 / 12 LDG.128 per lane and batch): enough for a volume-only prologue, but with volume
// AND plane features (20 corner reads of 128 B per query) it needs ~140 k cycles per tile against ~110 k for the tile's layers.
// 16 warps cap the kernel at 128 registers per thread (168 with 12), which costs the epilogue ~5 %: used only when both
// feature sources are sampled.
template <bool BF16, bool TWO, bool STG>
__global__ void __launch_bounds__(STG ? NTHREADS + 128 : NTHREADS, 1) decoder_tc_kernel(const __grid_constant__ TcKP p) {
    extern __shared__ unsigned char smem_raw[];
    const Dims& d = p.d;
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Smem L = smem_layout(d);
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + L.bars;
    auto w_full = [&](int s) { return bar0 + 8u * s; };
    auto w_empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    auto a_ready = [&](int t) { return bar0 + 8u * (2 * MAX_STAGES + t); };          // own chunk t written (this CTA)
    auto rready = [&](int sl) { return bar0 + 8u * (2 * MAX_STAGES + 8 + sl); };     // peer's chunk landed in remote slot sl
    auto rfree = [&](int sl) { return bar0 + 8u * (2 * MAX_STAGES + 12 + sl); };     // the PEER has consumed what I pushed into its slot sl
    const uint32_t acc_ready = bar0 + 8u * (2 * MAX_STAGES + MAX_CHUNKS);
    const uint32_t in_ready = acc_ready + 8;
    const uint32_t in_free = in_ready + 8;                   // (early staging) the code tile may be overwritten: the tile's last lin_z has retired
    const uint32_t feat_free = in_free + 8;                  // (early staging) the feature buffer may be overwritten: lin_in has retired
    // wide latent (lin_in chunks 1.. streamed through the own activation slots): the slots are free for the next tile (lin_out
    // has retired) / chunk landed in slot j / the MMAs that read slot j have retired
    const uint32_t a_free = feat_free + 8;
    auto s_full = [&](int j) { return a_free + 8u + 8u * j; };
    auto s_empty = [&](int j) { return a_free + 8u + 8u * (4 + j); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bars + (2 * MAX_STAGES + MAX_CHUNKS + 2 + 12) * 8);
    float* bias_s = reinterpret_cast<float*>(sm + L.bias);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (d.csize > 1) ? cluster_rank() : 0u;
    const int cluster_id = blockIdx.x / d.csize;
    const int own_chunks = d.OWN;
    const uint32_t half = TWO ? (rank >> 1) : rank;          // which N-half of the hidden width this CTA works on
    const uint32_t mrow = TWO ? (rank & 1u) : 0u;             // which 128-row tile of the cluster's rows
    const bool leader = !TWO || mrow == 0;                    // issues the MMAs
    const uint32_t peer = TWO ? (rank ^ 2u) : (rank ^ 1u);    // same rows, other N-half: chunk exchange partner
    const uint32_t lead_rank = rank & ~1u;                    // (TWO) leader of my pair
    const int rows_per_cluster = TWO ? 2 * BM : BM;

    // ---- one-time setup ---------------------------------------------------------------------
    if (threadIdx.x == 0) {
        // on a pair leader a weight stage / pushed chunk is "full" when its own copy AND the partner's have landed:
        // the partner forwards its local completion as a second arrival on the leader's barrier (one wait per k-step)
        const int both = (TWO && leader) ? 2 : 1;
        for (int s = 0; s < d.nstage; ++s) { mbar_init(w_full(s), both); mbar_init(w_empty(s), 1); }
        // a chunk / the input tiles are complete when every epilogue warp of the CTA -- and, on a pair leader,
        // of its partner too -- has arrived
        const int warps_in = 4 * EPI_GROUPS * ((TWO && leader) ? 2 : 1);
        for (int t = 0; t < 8; ++t) mbar_init(a_ready(t), warps_in);
        for (int sl = 0; sl < 4; ++sl) { mbar_init(rready(sl), both); mbar_init(rfree(sl), 1); }
        mbar_init(acc_ready, d.nsplit);                   // every issuing CTA of the cluster commits to every CTA
        mbar_init(in_ready, d.early ? (STG ? 4 : 1) : warps_in);
        mbar_init(in_free, 1);
        mbar_init(feat_free, 1);
        mbar_init(a_free, 1);
        for (int j = 0; j < 4; ++j) { mbar_init(s_full(j), 1); mbar_init(s_empty(j), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ROLE_WARP0 + 2) {
        if constexpr (TWO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    {   // fp32 bias table
        const GnbDecoderWeights& w = p.w;
        float* bo = bias_s + d.HN;
        for (int i = threadIdx.x; i < d.NOUT; i += (int)blockDim.x) bo[i] = i < d.d_out ? __ldg(w.lin_out_b + i) : 0.0f;
        float* hw = bo + d.NOUT;
        for (int i = threadIdx.x; i < d.d_geo; i += (int)blockDim.x) hw[i] = __ldg(w.head_w + i);
        if (threadIdx.x == 0) hw[d.d_geo] = __ldg(w.head_b);
    }
    tc_fence_before();
    __syncthreads();
    if (d.csize > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const unsigned char* wstream = p.packed + (long long)rank * d.packed_per_rank;
    const int nops = num_ops(d);
    const int my_tiles = cluster_id < p.n_tiles ? (p.n_tiles - 1 - cluster_id) / p.n_clusters + 1 : 0;

    if (warp == ROLE_WARP0) {
        // =============================== weight producer ======================================
        {
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (TWO) {
                // cta_group::2: a stage is a pair of 16 KB slots holding TWO consecutive k-chunks of the op (they are
                // contiguous in the packed stream: one bulk copy); one barrier round-trip per 128 columns of K
                const int nst = d.nstage;
                for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters) {
                    const unsigned char* src = wstream;
                    for (int o = 0; o < nops; ++o) {
                        const Op op = get_op(d, o);
                        for (int kc = 0; kc < op.kchunks; kc += 2) {
                            const uint32_t bytes = (uint32_t)min(2, op.kchunks - kc) * (uint32_t)op.rows * 128u;
                            mbar_wait(w_empty(stage), phase ^ 1);
                            bulk_g2s_u(sbase + L.ring + stage * d.stage_bytes, src, bytes, w_full(stage));
                            src += bytes;
                            if (++stage == nst) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            } else {
                // one step = one k-chunk of an op: its rows x 64 k slice (<= 256 rows = 32 KB) arrives as ONE bulk copy into
                // one ring stage.  A copy costs the issuing thread ~400 cycles whatever its size (16 KB -> 40 B/clk, 32 KB ->
                // 80 B/clk per SM, tools/microbench/l2_stream.cu), so big copies are what keeps the ingest rate up.
                for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters) {
                    const unsigned char* src = wstream;
                    for (int o = 0; o < nops; ++o) {
                        const Op op = get_op(d, o);
                        const uint32_t bytes = (uint32_t)op.rows * 128u;
                        for (int kc = 0; kc < op.kchunks; ++kc) {       // (the whole warp, converged: see umma_f16_x4_u)
                            mbar_wait(w_empty(stage), phase ^ 1);
                            if (p.nocopy) mbar_expect_tx_u(w_full(stage), 0);
                            else bulk_g2s_u(sbase + L.ring + stage * d.stage_bytes, src, bytes, w_full(stage));
                            src += bytes;
                            if (++stage == d.nstage) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == ROLE_WARP0 + 1 && !leader) {
        // ===================== (cta_group::2 partner) forward pushed-chunk arrivals to the leader ===========
        // one lane per remote slot, so that forwarding never queues behind another slot
        if (d.nsplit > 1 && lane < d.RS) {
            const uint32_t total = (uint32_t)my_tiles * (uint32_t)(2 * d.nb + 1) * (uint32_t)own_chunks;
            uint32_t use = 0;
            for (uint32_t i = lane; i < total; i += d.RS, ++use) {
                mbar_expect_tx(rready(lane), CHUNK);
                mbar_wait(rready(lane), use & 1);
                mbar_arrive_remote(map_to_cta(rready(lane), lead_rank));
            }
        }
    } else if (warp == ROLE_WARP0 + 3 && !leader) {
        // ===================== (cta_group::2 partner) forward weight-stage arrivals to the leader ============
        // one lane per ring stage (a stage = two k-chunks)
        const int nst = d.nstage;
        if (lane < nst) {
            uint32_t per_tile = 0;
            for (int o = 0; o < nops; ++o) { const Op op = get_op(d, o); per_tile += (op.kchunks + 1) / 2; }
            const uint32_t total = (uint32_t)my_tiles * per_tile;
            uint32_t use = 0;
            for (uint32_t i = lane; i < total; i += nst, ++use) {
                mbar_wait(w_full(lane), use & 1);
                mbar_arrive_remote(map_to_cta(w_full(lane), lead_rank));
            }
        }
    } else if (warp == ROLE_WARP0 + 1) {
        // =============================== MMA issuer ==========================================
        // The whole warp walks the program (warp-uniform control flow and operands); one elected
        // lane issues.  Two adjacent ring slots holding the two 128-row halves of a 256-row slice
        // are consumed by ONE N=256 instruction per k-step (half the issue work, 25% less
        // shared-memory operand traffic than two N=128 instructions).
        if constexpr (TWO) {
            // ---- cta_group::2 leader: steps of TWO k-chunks (128 columns of K): one combined wait on
            //      [first chunk ready, second chunk ready, weight stage landed], 8 MMAs, the commits ----
            const int nst = d.nstage;
            const uint16_t pairmask = (uint16_t)(3u << (half * 2)), othermask = (uint16_t)(3u << ((half ^ 1u) * 2));
            int stage = 0;
            uint32_t phase = 0, round = 0, tiles_done = 0, rtotal = 0;
            for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters, ++tiles_done) {
                mbar_wait_warp(in_ready, tiles_done & 1);
                tc_fence_after();
                int tk = tiles_done * 64;
                [[maybe_unused]] int step_no = 0;
                if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                for (int o = 0; o < nops; ++o) {
                    const Op op = get_op(d, o);
                    if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                    long long wait_all = 0;
                    const uint32_t idesc = umma_idesc(2 * op.rows, BF16, 2 * BM);
                    const uint32_t dcol = tmem + op.d_col;
                    for (int t = 0; t < op.kchunks; t += 2) {
                        const int nchunk = min(2, op.kchunks - t);
                        uint32_t a_addr[2] = {0, 0}, a_bar[2] = {0, 0}, a_par[2] = {0, 0};
                        int rslot = -1;
                        for (int c = 0; c < nchunk; ++c) {
                            const int tt = t + c;
                            if (op.a_kind == 0) a_addr[c] = sbase + L.feat + tt * CHUNK;
                            else if (op.a_kind == 1) a_addr[c] = sbase + L.code + tt * CHUNK;
                            else {
                                const bool own_chunk = d.nsplit == 1 || (tt & 1) == 0;
                                const int r = d.nsplit == 1 ? tt : (tt >> 1);
                                if (own_chunk) {
                                    a_bar[c] = a_ready(r), a_par[c] = round & 1;
                                    a_addr[c] = sbase + L.a + r * CHUNK;
                                } else {
                                    rslot = (int)(rtotal % (uint32_t)d.RS);
                                    mbar_expect_tx_u(rready(rslot), CHUNK);
                                    a_bar[c] = rready(rslot), a_par[c] = (rtotal / (uint32_t)d.RS) & 1;
                                    a_addr[c] = sbase + L.a + (d.AOWN + rslot) * CHUNK;
                                    ++rtotal;
                                }
                            }
                        }
                        const long long c1 = GNB_TRACE_ON(p) ? clock64() : 0;
#ifdef GNB_TC_TRACE
                        // finer trace (role 3): per step [reached, own chunk seen, pushed chunk seen, weight stage seen]
                        if (GNB_TRACE_ON(p)) {
                            const int sk = (tiles_done * 96 + step_no) * 4;
                            if (lane == 0) GNB_TRACE(3, sk);
                            if (a_bar[0]) mbar_wait(a_bar[0], a_par[0]);
                            if (lane == 0) GNB_TRACE(3, sk + 1);
                            if (a_bar[1]) mbar_wait(a_bar[1], a_par[1]);
                            if (lane == 0) GNB_TRACE(3, sk + 2);
                            mbar_wait(w_full(stage), phase);
                            if (lane == 0) GNB_TRACE(3, sk + 3);
                            ++step_no;
                        } else
#endif
                        mbar_wait3_warp(a_bar[0], a_par[0], a_bar[1], a_par[1], w_full(stage), phase);
                        if (GNB_TRACE_ON(p)) wait_all += clock64() - c1;
                        tc_fence_after();
                        const uint32_t b_base = sbase + L.ring + stage * d.stage_bytes;
                        const uint32_t first = (op.first_overwrites && t == 0) ? 0u : 1u;
                        umma_f16_x4_2cta_u(dcol, umma_desc(a_addr[0]), umma_desc(b_base), idesc, first);
                        if (nchunk == 2) umma_f16_x4_2cta_u(dcol, umma_desc(a_addr[1]), umma_desc(b_base + op.rows * 128), idesc, 1u);
                        umma_commit_mc_2cta_u(w_empty(stage), pairmask);     // both CTAs of the pair refill this stage
                        if (rslot >= 0) umma_commit_mc_2cta_u(rfree(rslot), othermask);
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                    }
                    if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                    if (GNB_TRACE_ON(p) && lane == 0 && blockIdx.x == 0 && tiles_done < 8) {
                        p.dbg[2 * 4096 + (tiles_done * 32 + o) * 2] = wait_all;
                        p.dbg[2 * 4096 + (tiles_done * 32 + o) * 2 + 1] = 0;
                    }
                    if (op.a_kind == 2) ++round;
                    if (op.group_end) {
                        umma_commit_mc_2cta_u(acc_ready, (uint16_t)((1u << d.csize) - 1));
                    }
                }
            }
        } else {
        // ---- single-CTA issue: the WHOLE warp walks the program, converged (see umma_f16_x4_u); every lane polls the
        //      barriers (uniform addresses: ~110 cycles per wait on a complete barrier, against ~230 for one polling lane plus a
        //      broadcast).  (Two issuing threads in different warps -- own chunks / pushed chunks, accumulating into the same
        //      TMEM columns -- were tried in round 1: tcgen05.commit arrivals got lost and results were corrupted.)
        {
            int stage = 0;
            uint32_t phase = 0, round = 0, tiles_done = 0, rtotal = 0, wfill = 0;
            const uint16_t allmask = (uint16_t)((1u << d.nsplit) - 1);
            for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters, ++tiles_done) {
                mbar_wait(in_ready, tiles_done & 1);
                tc_fence_after();
                int tk = tiles_done * 64;
                if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                for (int o = 0; o < nops; ++o) {
                    const Op op = get_op(d, o);
                    if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                    const uint32_t idesc = umma_idesc(op.rows, BF16);
                    const uint32_t dcol = tmem + op.d_col;
                    for (int t = 0; t < op.kchunks; ++t) {
                        uint32_t a_addr;
                        int rslot = -1;                                // >= 0: this chunk sits in a remote slot
                        int wslot = -1;                                // >= 0: (wide latent) lin_in chunk streamed through own slot wslot
                        if (op.a_kind == 0 && d.wide && t > 0) {
                            wslot = (int)(wfill % (uint32_t)d.OWN);
                            a_addr = sbase + L.a + wslot * CHUNK;
                            mbar_wait2(s_full(wslot), (wfill / (uint32_t)d.OWN) & 1, w_full(stage), phase);
                            ++wfill;
                        } else if (op.a_kind == 0) {
                            a_addr = sbase + L.feat + t * CHUNK;
                            mbar_wait(w_full(stage), phase);
                        } else if (op.a_kind == 1) {
                            a_addr = sbase + L.code + t * CHUNK;
                            mbar_wait(w_full(stage), phase);
                        } else if (t < own_chunks) {                   // own chunks first ...
                            a_addr = sbase + L.a + t * CHUNK;
                            mbar_wait2(a_ready(t), round & 1, w_full(stage), phase);
                        } else {                                       // ... then the peer's, in the order it pushes them
                            rslot = (int)(rtotal % (uint32_t)d.RS);
                            a_addr = sbase + L.a + (d.AOWN + rslot) * CHUNK;
                            mbar_expect_tx_u(rready(rslot), CHUNK);
                            mbar_wait2(rready(rslot), (rtotal / (uint32_t)d.RS) & 1, w_full(stage), phase);
                            ++rtotal;
                        }
                        tc_fence_after();
                        const uint64_t da = umma_desc(a_addr), db = umma_desc(sbase + L.ring + stage * d.stage_bytes);
                        umma_f16_x4_u(dcol, da, db, idesc, (op.first_overwrites && t == 0) ? 0u : 1u);
                        umma_commit_u(w_empty(stage));            // frees the ring stage when these MMAs retire
                        if (wslot >= 0) umma_commit_u(s_empty(wslot));     // (wide latent) the staging warp may refill the slot
                        // MMAs that read a remote slot: tell the PEER it may push into it again
                        if (rslot >= 0) umma_commit_mc_u(rfree(rslot), (uint16_t)(1u << peer));
                        if (++stage == d.nstage) { stage = 0; phase ^= 1; }
                    }
                    if (lane == 0) { GNB_TRACE(0, tk); } ++tk;
                    if (op.a_kind == 2) ++round;
                    // (early staging) the last lin_z of the tile has been issued: when it retires, the staging warp may
                    // overwrite the code tile and the feature buffer with the next tile's operands
                    if (d.early && op.mat == M_LIN_Z && op.blk == d.nb - 1) umma_commit_u(in_free);
                    if (d.early && o == 0) umma_commit_u(feat_free);
                    // (wide latent) lin_out was the last reader of the own activation slots: the next tile's lin_in chunks may land
                    if (d.wide && o == nops - 1) umma_commit_u(a_free);
                    if (op.group_end) {
                        // (early staging) this tile's first group may complete before the PEER has finished the previous
                        // tile: its arrival must not count for the previous tile's last phase of acc_ready
                        if (d.early && o == 1 && tiles_done > 0) mbar_wait(acc_ready, (tiles_done * (uint32_t)(2 * d.nb + 2) - 1) & 1);
                        if (d.nsplit > 1) umma_commit_mc_u(acc_ready, allmask);
                        else umma_commit_u(acc_ready);
                    }
                }
            }
        }
        }   // !TWO
    } else if (!TWO && (STG ? warp >= EPI_WARP0 + 4 * EPI_GROUPS : warp == ROLE_WARP0 + 3)) {
        // =============================== input staging (early mode) ============================
        // One warp prepares the NEXT tile's lin_in / lin_z operands (xyz -> positional code, feature sampling or load,
        // fp16/bf16 pack into the swizzled chunks) while the current tile's layers run, so the tensor pipe does not idle
        // for a prologue between tiles.  Lane l stages rows l, l+32, l+64, l+96.
        if (d.early) {
            const int rr0 = STG ? warp - (EPI_WARP0 + 4 * EPI_GROUPS) : 0, rr1 = STG ? rr0 + 1 : BM / 32;
            uint32_t it = 0, ovf = 0, wfill = 0;
            if (d.wide) {
                // wide latent (operand image only): chunk 0 goes into the feature buffer early, the code tile follows, and chunks
                // 1.. are streamed through the tile's own activation slots once the previous tile's lin_out has retired
                for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters, ++it) {
                    const unsigned char* img = p.feat_img + (long long)tile * d.KF * CHUNK;
                    if (it > 0) mbar_wait(feat_free, (it - 1) & 1);
                    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                                 "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(
                                     sbase + L.feat),
                                 "l"(img), "r"((uint32_t)CHUNK), "r"(in_ready)
                                 : "memory");
                    if (it > 0) mbar_wait(in_free, (it - 1) & 1);
                    for (int rr = rr0; rr < rr1; ++rr) {
                        const int row = rr * 32 + lane;
                        stage_inputs<BF16>(p, sm, L, (int)half, row, (long long)tile * BM + row, 0, 1, 1, ovf);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_expect_tx(in_ready, (uint32_t)CHUNK);
                    if (it > 0) mbar_wait(a_free, (it - 1) & 1);
                    for (int c = 1; c < d.KF; ++c, ++wfill) {
                        const int j = (int)(wfill % (uint32_t)d.OWN);
                        const uint32_t use = wfill / (uint32_t)d.OWN;
                        if (use > 0) mbar_wait(s_empty(j), (use - 1) & 1);
                        bulk_g2s_u(sbase + L.a + j * CHUNK, img + (long long)c * CHUNK, (uint32_t)CHUNK, s_full(j));
                    }
                }
            } else
            for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters, ++it) {
                // features first (the slow part: gathers): their buffer is free as soon as the previous tile's lin_in has
                // retired, almost a whole tile before they are needed; the code tile only after its last lin_z
                if (it > 0) mbar_wait(feat_free, (it - 1) & 1);
                const uint32_t img_bytes = (uint32_t)d.KF * CHUNK;
                if (p.feat_img) {
                    // the tile's feature chunks exist as an image already: one bulk copy, counted on in_ready (the expect_tx
                    // comes with this warp's arrival below; the transaction count may run negative inside the phase)
                    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                                 "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(
                                     sbase + L.feat),
                                 "l"(p.feat_img + (long long)tile * img_bytes), "r"(img_bytes), "r"(in_ready)
                                 : "memory");
                } else {
                    for (int rr = rr0; rr < rr1; ++rr) {
                        const int row = rr * 32 + lane;
                        stage_inputs<BF16>(p, sm, L, (int)half, row, (long long)tile * BM + row, 0, 1, 2, ovf);
                    }
                }
                if (it > 0) mbar_wait(in_free, (it - 1) & 1);
                for (int rr = rr0; rr < rr1; ++rr) {
                    const int row = rr * 32 + lane;
                    stage_inputs<BF16>(p, sm, L, (int)half, row, (long long)tile * BM + row, 0, 1, 1, ovf);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (p.feat_img) mbar_expect_tx(in_ready, img_bytes);
                    else mbar_arrive(in_ready);
                }
            }
            if (ovf && p.w.status) atomicOr(p.w.status, 1);
        }
    } else if (warp == ROLE_WARP0 + 2) {
        // =============================== A-chunk exchange (NSPLIT=2) ==========================
        // Pushes every own chunk into one of the peer's RS remote slots (cycled).  A slot is reused
        // only after the peer's MMAs that read it have retired (rfree, signalled by the peer's
        // multicast tcgen05.commit).
        if (d.nsplit > 1) {                                   // (the whole warp, converged)
            uint32_t round = 0, ptotal = 0;
            for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters) {
                for (int r = 0; r < 2 * d.nb + 1; ++r, ++round) {
                    for (int t = 0; t < own_chunks; ++t, ++ptotal) {
                        const uint32_t sl = ptotal % (uint32_t)d.RS, use = ptotal / (uint32_t)d.RS;
                        mbar_wait(a_ready(t), round & 1);     // every epilogue warp wrote + fenced chunk t
                        if (lane == 0) GNB_TRACE(2, 1024 + (int)ptotal * 3);
                        if (use > 0) mbar_wait(rfree(sl), (use - 1) & 1);
                        if (lane == 0) GNB_TRACE(2, 1024 + (int)ptotal * 3 + 1);
                        const uint32_t src = sbase + L.a + t * CHUNK;
                        const uint32_t dst = sbase + L.a + (d.AOWN + sl) * CHUNK;
                        bulk_s2peer_u(map_to_cta(dst, peer), src, CHUNK, map_to_cta(rready(sl), peer));
                        if (lane == 0) GNB_TRACE(2, 1024 + (int)ptotal * 3 + 2);
                    }
                }
            }
        }
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4 * EPI_GROUPS) {
        // =============================== prologue + epilogue ===================================
        const int q = warp & 3;                              // TMEM lane quadrant of this warp
        const int eg = (warp - EPI_WARP0) >> 2;              // epilogue group 0 / 1
        const int row = q * 32 + lane;                       // query row inside the tile == TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const GnbDecoderWeights& w = p.w;
        uint32_t grp = 0;                                    // acc_ready uses so far
        uint32_t ovf = 0;                                    // fp16 operands that saturated (reported through w.status)
        for (int tile = cluster_id; tile < p.n_tiles; tile += p.n_clusters) {
            const long long grow = (long long)tile * rows_per_cluster + mrow * BM + row;
            const bool live = grow < p.n_rows;
            int ek = (int)(grp / (2 * d.nb + 2)) * 64;
            if (threadIdx.x == EPI_WARP0 * 32) { GNB_TRACE(1, ek); } ++ek;
            // ---------------- prologue: operand tiles of lin_in and lin_z -------------------
            if (!d.early) {
                stage_inputs<BF16>(p, sm, L, (int)half, row, grow, eg, EPI_GROUPS, 3, ovf);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(in_ready);
                    else mbar_arrive_remote(map_to_cta(in_ready, lead_rank));      // the pair leader issues for both
                }
                if (threadIdx.x == EPI_WARP0 * 32) { GNB_TRACE(1, ek); } ++ek;
            } else { ++ek; }
            // ---------------- epilogue rounds: accumulator -> next A operand -----------------
            const int et = threadIdx.x - EPI_WARP0 * 32;                 // 0..255: this thread's slot of the shared bias[] (HN <= 256)
            for (int r = 0; r < 2 * d.nb + 1; ++r) {
                mbar_wait_park(acc_ready, grp & 1);
                ++grp;
                tc_fence_after();
                if (threadIdx.x == EPI_WARP0 * 32) { GNB_TRACE(1, ek); } ++ek;
                const bool from_net = (r & 1) == 1;                       // rounds: x, net, x, net, ..., x(last)
                const float* bias = (from_net || r == 2 * d.nb) ? bias_s : nullptr;
                // even round 2i, i < nb: nobody reads bias[] now (every thread finished round 2i-1 before this round's
                // accumulator could be complete) -- load fc_0's bias of block i for the next round
                float bnext = 0.0f;
                const bool load_b0 = !from_net && r < 2 * d.nb && et < d.HN;
                if (load_b0) bnext = __ldg(w.fc0_b[r >> 1] + half * d.HN + et);
                const uint32_t src_col = from_net ? NET_COL : 0;
                // both groups work on every chunk (group g converts its columns [32g, 32g+32)), so the
                // first chunk -- which releases the MMAs of this layer -- is ready as early as possible
                uint32_t v[2][32];
                tmem_ld32(tlane + src_col + eg * 32, v[0]);
#pragma unroll 1
                for (int t = 0; t < own_chunks; t += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (t + h >= own_chunks) break;
                        tmem_ld_wait();
                        if (t + h + 1 < own_chunks) tmem_ld32(tlane + src_col + (t + h + 1) * 64 + eg * 32, v[(h + 1) & 1]);
                        // (published to the other epilogue threads by the arrive below -> MMAs -> acc_ready chain)
                        if (load_b0 && t + h == own_chunks - 1) bias_s[et] = bnext;
                        unsigned char* dst = sm + L.a + (t + h) * CHUNK;
                        const float* bch = bias ? bias + (t + h) * 64 + eg * 32 : nullptr;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {                    // 4 x 16-byte units of 8 columns
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[h][u * 8 + e]);
                            if (bch) {
                                const float4 b0v = *reinterpret_cast<const float4*>(bch + u * 8);
                                const float4 b1v = *reinterpret_cast<const float4*>(bch + u * 8 + 4);
                                f[0] += b0v.x, f[1] += b0v.y, f[2] += b0v.z, f[3] += b0v.w;
                                f[4] += b1v.x, f[5] += b1v.y, f[6] += b1v.z, f[7] += b1v.w;
                            }
                            uint4 pk = make_uint4(pack16_relu<BF16>(f[0], f[1]), pack16_relu<BF16>(f[2], f[3]),
                                                  pack16_relu<BF16>(f[4], f[5]), pack16_relu<BF16>(f[6], f[7]));
                            if constexpr (!BF16) ovf |= sat_probe(pk.x, pk.y) | sat_probe(pk.z, pk.w);
                            *reinterpret_cast<uint4*>(dst + chunk_off(row, eg * 4 + u)) = pk;
                            if (p.save && live)          // the backward pass reads what the next layer's MMAs read
                                *reinterpret_cast<uint4*>(p.save + ((long long)r * p.n_rows + grow) * d.Hd + half * d.HN + (t + h) * 64 +
                                                          eg * 32 + u * 8) = pk;
                        }
                        tc_fence_before();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(a_ready(t + h));                             // local: exchange thread (+ MMA on a leader)
                            if (!leader) mbar_arrive_remote(map_to_cta(a_ready(t + h), lead_rank));
                        }
                    }
                }
                if (threadIdx.x == EPI_WARP0 * 32) { GNB_TRACE(1, ek); } ++ek;
                if (r == 2 * d.nb - 1) {
                    // the last round adds the last fc_1 bias: swap it into bias[] once every thread has read fc_0's
                    const float b1v = et < d.HN ? __ldg(w.fc1_b[d.nb - 1] + half * d.HN + et) : 0.0f;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et < d.HN) bias_s[et] = b1v;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
            // ---------------- final epilogue: lin_out tile -> global, TSDF head -----------------
            {
                mbar_wait_park(acc_ready, grp & 1);
                ++grp;
                tc_fence_after();
                const float* bo = bias_s + d.HN;
                const float* hw = bo + d.NOUT;
                float head = hw[d.d_geo];
                const long long orow = (p.sorted && live) ? (long long)__float_as_int(__ldg(reinterpret_cast<const float*>(p.sorted + grow) + 3)) : grow;
                for (int part = 0; part < (eg == 0 ? d.NOUT / 16 : 0); ++part) {
                    uint32_t v[16];
                    tmem_ld16(tlane + NET_COL + part * 16, v);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = part * 16 + e;
                        f[e] = __uint_as_float(v[e]) + bo[n];
                        if (n < d.d_geo) head = fmaf(f[e], hw[n], head);
                    }
                    if (live && half == 0 && p.out) {
                        float* o = p.out + orow * d.d_out + part * 16;
                        if ((d.d_out & 3) == 0) {
#pragma unroll
                            for (int e = 0; e < 16; e += 4)
                                if (part * 16 + e < d.d_out) *reinterpret_cast<float4*>(o + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (part * 16 + e < d.d_out) o[e] = f[e];
                        }
                    }
                }
                if (live && half == 0 && eg == 0 && p.tsdf) p.tsdf[orow] = tanhf(head);
                if (ovf && live && p.w.status) atomicOr(p.w.status, 1);
                ovf = 0;
                tc_fence_before();
                if (threadIdx.x == EPI_WARP0 * 32) { GNB_TRACE(1, ek); } ++ek;
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (d.csize > 1) cluster_sync_all();
    if (warp == ROLE_WARP0 + 2) {
        __syncwarp();
        tc_fence_after();
        if constexpr (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear matrices -> bf16, K-major, 128B-swizzled tiles in stream order
// ---------------------------------------------------------------------------------------------
struct PackOp {
    const float* W;       // (rows_true, K_true) row-major
    int rows_true, K_true;
    int n0;               // first row of this CTA's slice
    int rows;             // rows of the op in this CTA (HN or NOUT)
    int kchunks;
    int order_half;       // activation ops: N-half whose visiting order (act_kchunk) is used; -1 = identity
    int nsplit, own, interleave;
    float scale;          // multiplies W and biasA
    const float* scale_dev;   // optional: the scale is read from the device instead (GnbDecoderWeights.alpha_dev)
    const float* biasA;   // bias columns at k == K_true (hi) and K_true + 1 (lo): scale*biasA[n] + biasB[n]
    const float* biasB;
    int bias_cols;
    long long dst_off;    // byte offset of the op inside the rank's stream
};

// all ops of one CTA rank's weight stream in ONE launch (blockIdx.y = op): a training step re-packs the image every forward
constexpr int MAX_PACK_OPS = 3 * 8 + 2;
struct PackBatch {
    PackOp op[MAX_PACK_OPS];
};
template <bool BF16>
__global__ void pack_kernel(const __grid_constant__ PackBatch batch, unsigned char* __restrict__ dst_rank) {
    const PackOp& op = batch.op[blockIdx.y];
    // one thread per 16-byte unit: (t, nt, n_local, u)
    const int ntile = n_tiles_of(op.rows);
    const long long units_per_k = (long long)op.rows * 8;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= units_per_k * op.kchunks) return;
    const int t = (int)(idx / units_per_k);
    int rem = (int)(idx % units_per_k);
    const int n_in_op = rem / 8, u = rem % 8;
    const int nt = n_in_op / 128, nl = n_in_op % 128;
    const int kc = op.order_half < 0 ? t : act_kchunk(op.nsplit, op.own, op.order_half, t, op.interleave);
    const int n = op.n0 + n_in_op;
    float v[8];
    const float scale = op.scale_dev ? __ldg(op.scale_dev) : op.scale;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kc * 64 + u * 8 + e;
        float x = 0.0f;
        if (n < op.rows_true) {
            if (k < op.K_true) x = scale * op.W[(long long)n * op.K_true + k];
            else if (op.bias_cols && k < op.K_true + 2) {
                float b = (op.biasA ? scale * op.biasA[n] : 0.0f) + (op.biasB ? op.biasB[n] : 0.0f);
                float hi = round16<BF16>(b);
                x = (k == op.K_true) ? hi : (b - hi);
            }
        }
        v[e] = x;
    }
    // tiles of one k-chunk are consecutive: rows of earlier n-tiles precede
    long long off = op.dst_off + ((long long)t * op.rows + (long long)nt * 128) * 128 + chunk_off(nl, u);
    (void)ntile;
    *reinterpret_cast<uint4*>(dst_rank + off) = make_uint4(pack16<BF16>(v[0], v[1]), pack16<BF16>(v[2], v[3]), pack16<BF16>(v[4], v[5]), pack16<BF16>(v[6], v[7]));
}

static int make_dims(const GnbDecoderWeights* w, Dims& d, const char* who) {
    int rc = check_decoder_weights(w, who);
    if (rc) return rc;
    d.d_feat = w->d_feat, d.d_code = w->d_code, d.Hd = w->d_hidden, d.nb = w->n_blocks, d.d_out = w->d_out, d.d_geo = w->d_geo;
    if (d.Hd % 64 != 0 || d.Hd > 512 || d.Hd < 64 || (d.Hd > 256 && d.Hd % 128 != 0)) {
        set_error("%s: d_hidden %d is not supported by the tcgen05 path (multiple of 64 up to 256, or 384 / 512)", who, d.Hd);
        return GNB_E_UNSUPPORTED;
    }
    if (w->tc_dtype != GNB_TC_FP16 && w->tc_dtype != GNB_TC_BF16) { set_error("%s: unknown tc_dtype %d", who, w->tc_dtype); return GNB_E_INVALID; }
    if (d.nb < 1) { set_error("%s: n_blocks must be >= 1", who); return GNB_E_UNSUPPORTED; }
    d.nsplit = d.Hd > 256 ? 2 : 1;
    d.HN = d.Hd / d.nsplit;
    // cta_group::2 pairs (each CTA streams half of B, 256 rows per cluster) are built and parity-tested but opt-in:
    // measured this round they run at 0.91x of the single-CTA issue because a cluster of 4 leaves 16 SMs idle
    // and the per-layer epilogue -> exchange -> MMA latency chain, not the weight ingest, then bounds a layer
    d.two = opt(OPT_TC_TWO_CTA) ? 1 : 0;
    if (d.HN % 32 != 0) d.two = 0;
    d.csize = d.nsplit * (d.two ? 2 : 1);
    d.WN = d.two ? d.HN / 2 : d.HN;
    d.KF = (d.d_feat + 63) / 64;             // no bias columns: lin_in's bias rides with lin_z_0's (same accumulation group)
    d.KZ = (d.d_code + 2 + 63) / 64;
    d.KH = d.Hd / 64;
    d.NOUT = d.two ? (d.d_out + 31) / 32 * 32 : (d.d_out + 15) / 16 * 16;     // each CTA of a pair holds NOUT/2 rows
    d.NOUTC = d.two ? d.NOUT / 2 : d.NOUT;
    if (d.KZ > 4 || d.NOUT > 256 || d.KF > 64) {
        set_error("%s: d_code %d / d_out %d / d_feat %d too large for the tcgen05 path", who, d.d_code, d.d_out, d.d_feat);
        return GNB_E_UNSUPPORTED;
    }
    d.wide = d.KF > 8 ? 1 : 0;
    d.KFB = d.wide ? 1 : d.KF;
    if (d.wide && d.two) {
        set_error("%s: d_feat %d (more than 8 k-chunks) needs the default kernel (GNB_TC_TWO_CTA off)", who, d.d_feat);
        return GNB_E_UNSUPPORTED;
    }
    d.OWN = d.HN / 64;
    d.AOWN = (!d.wide && d.KF > d.OWN) ? d.KF : d.OWN;
    d.RS = d.nsplit > 1 ? (d.OWN >= 2 ? 2 : 1) : 0;
    d.ACH = d.AOWN + d.RS;
    // ring stage: cta_group::2 = two k-chunks of WN <= 128 rows; otherwise one k-chunk of the widest op
    {
        const int rows = d.WN > d.NOUTC ? d.WN : d.NOUTC;
        d.stage_bytes = d.two ? 2 * CHUNK : (rows * 128 + 1023) / 1024 * 1024;
    }
    d.early = 0;
    d.nstage = MAX_STAGES;
    while (d.nstage >= 2 && smem_layout(d).total + 1024 > 227 * 1024) --d.nstage;
    if (!d.two && !opt(OPT_TC_NO_EARLY)) {
        // early input staging costs KF more chunks of shared memory: take it when the weight ring keeps its depth
        Dims e = d;
        e.early = 1;
        e.AOWN = e.OWN, e.ACH = e.AOWN + e.RS;             // the lin_in operand no longer borrows the own chunk slots
        e.nstage = MAX_STAGES;
        while (e.nstage >= 2 && smem_layout(e).total + 1024 > 227 * 1024) --e.nstage;
        if (e.nstage >= d.nstage || e.nstage >= 4 || d.wide) d = e;
    }
    if (d.wide && (!d.early || opt(OPT_TC_NO_EARLY))) {
        set_error("%s: d_feat %d (more than 8 k-chunks) needs the early-staging kernel (GNB_TC_NO_EARLY off)", who, d.d_feat);
        return GNB_E_UNSUPPORTED;
    }
    if (d.nstage < 2) { set_error("%s: tile does not fit in shared memory", who); return GNB_E_UNSUPPORTED; }
    long long bytes = 0;
    for (int o = 0; o < num_ops(d); ++o) {
        const Op op = get_op(d, o);
        bytes += (long long)op.kchunks * op.rows * 128;
    }
    d.packed_per_rank = bytes;
    return 0;
}

}  // namespace tc
}  // namespace gnb

#include "decoder_tp.cuh"

using namespace gnb;
using namespace gnb::tc;

extern "C" int64_t gnb_decoder_packed_bytes(const GnbDecoderWeights* w) {
    Dims d;
    if (make_dims(w, d, "gnb_decoder_packed_bytes")) return 0;
    {
        TpDims td;
        if (tp_applies(w, td)) return tp_packed_bytes(td);
    }
    return d.packed_per_rank * d.csize;
}

extern "C" int gnb_decoder_pack_tc(const GnbDecoderWeights* w, void* packed, void* stream) {
    Dims d;
    int rc = make_dims(w, d, "gnb_decoder_pack_tc");
    if (rc) return rc;
    GNB_CHECK_ARG(packed, "gnb_decoder_pack_tc: null output");
    cudaStream_t st = (cudaStream_t)stream;
    {
        TpDims td;
        if (tp_applies(w, td)) return tp_pack(w, td, packed, st);
    }
    for (int rank = 0; rank < d.csize; ++rank) {
        const int half = d.two ? rank >> 1 : rank, mrow = d.two ? (rank & 1) : 0;
        unsigned char* dst = (unsigned char*)packed + (long long)rank * d.packed_per_rank;
        long long off = 0, max_units = 0;
        PackBatch batch;
        int n_batch = 0;
        if (num_ops(d) > MAX_PACK_OPS) { set_error("gnb_decoder_pack_tc: too many ops"); return GNB_E_INVALID; }
        auto launch = [&](PackOp op) -> int {
            op.dst_off = off;
            const long long units = (long long)op.rows * 8 * op.kchunks;
            max_units = units > max_units ? units : max_units;
            batch.op[n_batch++] = op;
            off += (long long)op.kchunks * op.rows * 128;
            return 0;
        };
        const int n0 = half * d.HN + mrow * d.WN;
        for (int o = 0; o < num_ops(d); ++o) {            // stream order == program order (get_op)
            const Op g = get_op(d, o);
            PackOp op;
            switch (g.mat) {
            case M_LIN_IN:
                op = PackOp{w->lin_in_w, d.Hd, d.d_feat, n0, d.WN, d.KF, -1, d.nsplit, d.OWN, d.two, 1.0f, nullptr, nullptr, nullptr, 0, 0};
                break;
            case M_LIN_Z:
                // x += alpha * (Wz code + bz)   [+ b1 of the previous block, or lin_in's bias for block 0: lin_in and lin_z_0
                // accumulate into x in the same group, so one pair of bias columns serves both and a 64-wide feature vector
                // (volume 32 + planes 32) stays ONE k-chunk]
                op = PackOp{w->lin_z_w[g.blk], d.Hd, d.d_code, n0, d.WN, d.KZ, -1, d.nsplit, d.OWN, d.two, w->alpha, w->alpha_dev, w->lin_z_b[g.blk],
                            g.blk > 0 ? w->fc1_b[g.blk - 1] : w->lin_in_b, 1, 0};
                break;
            case M_FC0:
                op = PackOp{w->fc0_w[g.blk], d.Hd, d.Hd, n0, d.WN, d.KH, half, d.nsplit, d.OWN, d.two, 1.0f, nullptr, nullptr, nullptr, 0, 0};
                break;
            case M_FC1:
                op = PackOp{w->fc1_w[g.blk], d.Hd, d.Hd, n0, d.WN, d.KH, half, d.nsplit, d.OWN, d.two, 1.0f, nullptr, nullptr, nullptr, 0, 0};
                break;
            default:
                op = PackOp{w->lin_out_w, d.d_out, d.Hd, mrow * d.NOUTC, d.NOUTC, d.KH, half, d.nsplit, d.OWN, d.two, 1.0f, nullptr, nullptr, nullptr, 0, 0};
                break;
            }
            if ((rc = launch(op))) return rc;
        }
        if (off != d.packed_per_rank) { set_error("gnb_decoder_pack_tc: internal size mismatch"); return GNB_E_INVALID; }
        const dim3 grid((unsigned)ceil_div(max_units, 256), (unsigned)n_batch);
        if (w->tc_dtype == GNB_TC_BF16) pack_kernel<true><<<grid, 256, 0, st>>>(batch, dst);
        else pack_kernel<false><<<grid, 256, 0, st>>>(batch, dst);
        GNB_LAUNCH_CHECK();
    }
    return 0;
}

// debugging aid (not part of the public header): returns a host pointer to 8 ints that a timed-out barrier wait fills
// in before they trap (int[0] = number of reports, then from int[8] on {block, thread, source line, barrier address, parity, 0}
// per report); readable after the launch failed
extern "C" int* gnb_debug_hang_report() {
    static int* host = nullptr;
    if (!host) {
        if (cudaHostAlloc((void**)&host, 4096, cudaHostAllocMapped) != cudaSuccess) return nullptr;
        for (int i = 0; i < 1024; ++i) host[i] = 0;
        int* dev = nullptr;
        if (cudaHostGetDevicePointer((void**)&dev, host, 0) != cudaSuccess) return nullptr;
        if (cudaMemcpyToSymbol(gnb::tc::g_hang_report, &dev, sizeof(dev)) != cudaSuccess) return nullptr;
    }
    return host;
}

static long long* g_trace = nullptr;
// profiling aid (not part of the public header): device buffer of 3*4096 int64 that receives
// clock64() stamps of cluster 0's roles; pass null to switch tracing off again
extern "C" void gnb_debug_set_trace(long long* dev_buf) { g_trace = dev_buf; }

static int launch_tc(const GnbDecoderWeights* w, const void* packed, TcKP& kp, void* stream) {
    int rc = make_dims(w, kp.d, "gnb_decode_tc");
    if (rc) return rc;
    GNB_CHECK_ARG(packed, "gnb_decode_tc: weights are not packed");
    {
        TpDims td;
        if (!kp.feat_img && !kp.save && tp_applies(w, td)) {
            if (kp.n_rows == 0) return 0;
            return tp_launch(w, td, packed, kp, stream, g_trace);
        }
    }
    const Dims& d = kp.d;
    if (d.wide && !kp.feat_img) {
        set_error("gnb_decode_tc: d_feat %d (more than 8 k-chunks of lin_in) is served through the operand image only: "
                  "gnb_features_to_image / the sampler's image output + gnb_decode_image_tc", d.d_feat);
        return GNB_E_UNSUPPORTED;
    }
    if (kp.feat_img && (!d.early || d.two)) {
        set_error("gnb_decode_image_tc: needs the default kernel (early staging, single-CTA issue)");
        return GNB_E_UNSUPPORTED;
    }
    kp.w = *w;
    kp.packed = (const unsigned char*)packed;
    if (kp.n_rows == 0) return 0;
    int dev = 0, sms = 0, cc = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GNB_CUDA(cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc != 10) { set_error("gnb_decode_tc: needs an sm_100 device (found sm_%d0)", cc); return GNB_E_ARCH; }
    kp.dbg = g_trace;
    kp.nocopy = opt(OPT_DEBUG_NO_WCOPY);
    const int rows_per_cluster = d.two ? 2 * BM : BM;
    kp.n_tiles = (int)((kp.n_rows + rows_per_cluster - 1) / rows_per_cluster);
    kp.n_clusters = sms / d.csize;
    if (const int m = opt(OPT_DEBUG_MAX_CLUSTERS)) {          // profiling aid: fewer resident clusters
        if (m > 0 && m < kp.n_clusters) kp.n_clusters = m;
    }
    if (kp.n_clusters > kp.n_tiles) kp.n_clusters = kp.n_tiles;
    const size_t smem = smem_layout(d).total + 1024;
    const bool bf = w->tc_dtype == GNB_TC_BF16;
    // extra staging warps: fused query that samples BOTH a volume and planes, single-CTA issue with early staging
    const bool stg = !d.two && d.early && kp.fused && kp.s.volume && kp.s.Cp > 0 && !opt(OPT_TC_NO_STG);
    auto kernel = d.two ? (bf ? decoder_tc_kernel<true, true, false> : decoder_tc_kernel<false, true, false>)
                        : stg ? (bf ? decoder_tc_kernel<true, false, true> : decoder_tc_kernel<false, false, true>)
                              : (bf ? decoder_tc_kernel<true, false, false> : decoder_tc_kernel<false, false, false>);
    GNB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kp.n_clusters * d.csize);
    cfg.blockDim = dim3(stg ? NTHREADS + 128 : NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = d.csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    {
        // persistent kernel with a static tile walk: never launch more clusters than can be co-resident
        // (clusters must fit inside one GPC, so fewer than SMs / cluster size may be available)
        int max_clusters = 0;
        cfg.gridDim = dim3(sms / d.csize * d.csize);
        GNB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
        if (opt(OPT_DEBUG_PRINT)) fprintf(stderr, "[gnb] decoder: cluster size %d, max co-resident clusters %d\n", d.csize, max_clusters);
        if (max_clusters > 0 && kp.n_clusters > max_clusters) kp.n_clusters = max_clusters;
        cfg.gridDim = dim3(kp.n_clusters * d.csize);
    }
    GNB_CUDA(cudaLaunchKernelEx(&cfg, kernel, kp));
    return 0;
}

extern "C" int gnb_decode_tc(const GnbDecoderWeights* w, const void* packed, const float* xyz, const float* feat,
                               int64_t n_rows, float* out, float* tsdf, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0, "gnb_decode_tc: bad arguments");
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(xyz && feat && (out || tsdf), "gnb_decode_tc: bad arguments");
    TcKP kp = {};
    kp.xyz = xyz, kp.feat = feat, kp.fused = 0, kp.n_rows = n_rows, kp.out = out, kp.tsdf = tsdf;
    return launch_tc(w, packed, kp, stream);
}

extern "C" int gnb_decode_tc_save(const GnbDecoderWeights* w, const void* packed, const float* xyz, const float* feat,
                                    int64_t n_rows, float* out, float* tsdf, void* activations, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0, "gnb_decode_tc_save: bad arguments");
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(xyz && feat && (out || tsdf) && activations, "gnb_decode_tc_save: bad arguments");
    GNB_CHECK_ARG((reinterpret_cast<uintptr_t>(activations) & 15) == 0 && w && w->d_hidden % 8 == 0,
                  "gnb_decode_tc_save: the activation buffer must be 16-byte aligned");
    TcKP kp = {};
    kp.xyz = xyz, kp.feat = feat, kp.fused = 0, kp.n_rows = n_rows, kp.out = out, kp.tsdf = tsdf;
    kp.save = (uint16_t*)activations;
    return launch_tc(w, packed, kp, stream);
}

extern "C" int gnb_decoder_image_kchunks(const GnbDecoderWeights* w) {
    Dims d;
    if (make_dims(w, d, "gnb_decoder_image_kchunks")) return 0;
    if (!d.early || d.two) return 0;         // the image is read by the early-staging warp of the default kernel only
    return d.KF;
}

extern "C" int gnb_decode_image_tc(const GnbDecoderWeights* w, const void* packed, const float* xyz, const void* image,
                                     int64_t n_rows, float* out, float* tsdf, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0, "gnb_decode_image_tc: bad arguments");
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(xyz && image && (out || tsdf), "gnb_decode_image_tc: bad arguments");
    GNB_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0, "gnb_decode_image_tc: the image must be 16-byte aligned");
    TcKP kp = {};
    kp.xyz = xyz, kp.feat = nullptr, kp.feat_img = (const unsigned char*)image, kp.fused = 0, kp.n_rows = n_rows, kp.out = out, kp.tsdf = tsdf;
    return launch_tc(w, packed, kp, stream);
}

static int query_fused_common(const GnbSampleParams* s, const GnbDecoderWeights* w, TcKP& kp, float* out, float* tsdf, const char* who) {
    (void)who;
    int rc = fill_sample_kp(s, kp.s);
    if (rc) return rc;
    GNB_CHECK_ARG(w && w->use_code != 2, "gnb_query_fused_tc: the fused query encodes xyz itself (use_code 0 or 1)");
    GNB_CHECK_ARG(w->d_feat == kp.s.C + kp.s.Cp, "gnb_query_fused_tc: d_feat %d != C_p + C = %d", w->d_feat, kp.s.C + kp.s.Cp);
    if (kp.s.total == 0) return 0;
    GNB_CHECK_ARG(out || tsdf, "gnb_query_fused_tc: no output requested");
    GNB_CHECK_ARG(!s->out || (s->out_stride >= w->d_feat && s->out_stride % 4 == 0), "gnb_query_fused_tc: bad feature output stride");
    bool ok = true;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    // float4 loads / stores: base pointers must be 16-byte aligned too (a sliced view may not be)
    ok = ok && al16(kp.s.volume) && al16(kp.s.plane[0]) && al16(kp.s.plane[1]) && al16(kp.s.plane[2]) && al16(s->out);
    if (kp.s.volume) ok = ok && kp.s.vsc == 1 && kp.s.C % 4 == 0 && kp.s.vsb % 4 == 0 && kp.s.vsx % 4 == 0 && kp.s.vsy % 4 == 0 && kp.s.vsz % 4 == 0;
    if (kp.s.Cp > 0) ok = ok && kp.s.psc == 1 && kp.s.Cp % 4 == 0 && kp.s.psb % 4 == 0 && kp.s.psh % 4 == 0 && kp.s.psw % 4 == 0;
    if (!ok) {
        set_error("gnb_query_fused_tc: the fused path needs 16-byte aligned channels-last volume / planes with channel counts % 4 == 0 "
                  "(use gnb_sample_features + gnb_decode_tc otherwise)");
        return GNB_E_UNSUPPORTED;
    }
    kp.xyz = s->xyz, kp.feat = nullptr, kp.fused = 1, kp.n_rows = kp.s.total, kp.out = out, kp.tsdf = tsdf;
    return 0;
}

extern "C" int gnb_query_fused_tc(const GnbSampleParams* s, const GnbDecoderWeights* w, const void* packed, float* out,
                                    float* tsdf, void* stream) {
    TcKP kp = {};
    int rc = query_fused_common(s, w, kp, out, tsdf, "gnb_query_fused_tc");
    if (rc || kp.s.total == 0) return rc;
    return launch_tc(w, packed, kp, stream);
}

extern "C" int64_t gnb_query_fused_sorted_scratch_bytes(const GnbSampleParams* s) { return bin_sort_scratch_bytes(s); }

extern "C" int gnb_query_fused_sorted_tc(const GnbSampleParams* s, const GnbDecoderWeights* w, const void* packed, float* out,
                                           float* tsdf, void* scratch, int64_t scratch_bytes, void* stream) {
    TcKP kp = {};
    int rc = query_fused_common(s, w, kp, out, tsdf, "gnb_query_fused_sorted_tc");
    if (rc || kp.s.total == 0) return rc;
    const float4* sorted = nullptr;
    GnbSampleParams ss = *s;
    ss.out = nullptr;                                  // (the sort plan only validates it)
    if ((rc = bin_sort(&ss, scratch, scratch_bytes, stream, &sorted))) return rc;
    kp.xyz = nullptr, kp.sorted = sorted;
    return launch_tc(w, packed, kp, stream);
}

extern "C" int gnb_query_grid_fused_tc(const GnbSampleParams* s, const int32_t* h_grid3, const float* axes, const GnbDecoderWeights* w,
                                         const void* packed, float* out, float* tsdf, void* stream) {
    GNB_CHECK_ARG(s && h_grid3 && axes, "gnb_query_grid_fused_tc: null argument");
    GNB_CHECK_ARG(h_grid3[0] > 0 && h_grid3[1] > 0 && h_grid3[2] > 0, "gnb_query_grid_fused_tc: bad grid");
    GnbSampleParams sg = *s;
    sg.n_query = (int64_t)h_grid3[0] * h_grid3[1] * h_grid3[2];
    sg.xyz = axes;                                     // (only so that the shared argument checks see a non-null pointer)
    TcKP kp = {};
    int rc = query_fused_common(&sg, w, kp, out, tsdf, "gnb_query_grid_fused_tc");
    if (rc || kp.s.total == 0) return rc;
    kp.xyz = nullptr, kp.gaxes = axes, kp.gnx = h_grid3[0], kp.gny = h_grid3[1], kp.gnz = h_grid3[2];
    return launch_tc(w, packed, kp, stream);
}

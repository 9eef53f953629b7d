// "Transposed pair" variant of the tcgen05 decoder (included by decoder_tc.cu; same entry points, same TcKP).
//
// Same network as decoder_tc_kernel (reference resnetfc.py:134-189, positional_encoding.py:28-40, heads3d.py:36-50,
// model.py:207-248), computed as  D[hidden unit, query] = W[hidden, k] x Act[k, query]  by a cluster of TWO CTAs issuing
// tcgen05.mma.cta_group::2:
//   M = 256 hidden units per instruction, 128 per CTA (A = weights, K-major tiles streamed from L2, each CTA its own rows),
//   N = 128 queries per tile; B = activations, MN-major (row = k, 64 queries = 128 bytes per row), each CTA holds the B
//       half of ITS 64 queries and the tensor cores share it across the pair.
// Against the query-major layout of decoder_tc_kernel (M = 128 queries per CTA, N = hidden half, activation chunks pushed
// to the peer: every CTA needs all K for its rows) this
//   * halves the activation bytes a CTA holds (64 KB for K = 512) -> a FOUR-stage 32 KB weight ring fits (3 stages bounded
//     the k-step cadence at ~850 cycles against 630 of MMA time),
//   * halves the bytes exchanged per layer (a CTA produces 256 hidden units x 128 queries and needs 512 x its 64 queries),
//   * runs the MMAs at the tensor pipe's full rate (512 cycles per 32 KB stage measured, tools/microbench/umma_pair_test.cu:
//     the B half of the partner does not cross this SM's shared-memory port), and
//   * splits a layer into two accumulation groups (M-tiles of 256 hidden units), so the epilogue of the first overlaps the
//     MMAs of the second.
// Per CTA: TMEM x_j at columns [128 j, 128 j + 128), net_j at [256 + 128 j, ...), j = M-tile; biases are per hidden unit = per
// TMEM lane, i.e. one scalar per epilogue thread (no bias operand columns); lin_out's rows live on CTA 0 (lanes = output
// features), the TSDF head is a warp reduction over the d_geo <= 32 feature lanes.
// Warps: 0 weight producer | 1 MMA issue (CTA 0) / weight-stage forwarder (CTA 1) | 2 TMEM alloc + push of the partner's half
// | 3 input staging (own 64 queries of the NEXT tile) | 4-11 epilogue: quadrant q = warp & 3 <-> TMEM lanes 32q.., group
// eg = (warp - 4) >> 2 <-> queries [64 eg, 64 eg + 64) of the tile (eg == rank: kept, else staged and pushed to the partner).
#pragma once

namespace gnb {
namespace tc {

constexpr int TP_BN = 128;                 // queries per tile
constexpr int TP_ACT_CHUNK = 8192;         // one k-chunk of the B operand: 64 k-rows x 64 own queries x 2 B
constexpr int TP_TILE = 16384;             // one A tile: 128 hidden rows x 64 k x 2 B
constexpr int TP_STAGE = 2 * TP_TILE;      // ring stage: two k-chunks of one M-tile
constexpr int TP_NSTAGE = 4;
constexpr int TP_THREADS = 384;

struct TpSmem {
    uint32_t act, xs, code, feat, ring, bars, total;
};
__host__ __device__ inline TpSmem tp_smem_layout(int KF) {
    TpSmem s;
    s.act = 0;
    s.xs = s.act + 8 * TP_ACT_CHUNK;                 // bounce buffer of the half that goes to the partner: 2 chunks
    s.code = s.xs + 2 * TP_ACT_CHUNK;
    s.feat = s.code + TP_ACT_CHUNK;
    s.ring = s.feat + KF * TP_ACT_CHUNK;
    s.bars = s.ring + TP_NSTAGE * TP_STAGE;
    s.total = s.bars + 40 * 8 + 16;
    return s;
}

// ---- the per-tile program: a sequence of ring stages -------------------------------------------------------------
// kind 0: lin_in (B = feature tile, K-major)   1: lin_z (B = code tile, K-major)   2: hidden / lin_out (B = ACT chunks, MN-major)
struct TpStage {
    int kind, j, blk;            // M-tile, block
    int kc0, nkc;                // k-chunks of the stage (hidden: 2; lin_in / lin_z: 1)
    int n16;                     // K16 MMAs per k-chunk
    int d_col;                   // TMEM column of the accumulator
    int overwrite;               // first MMA of the stage overwrites the accumulator
    int acc_end;                 // last stage of an accumulation group: commit acc(j) (3: the lin_out group)
    int kh0_end;                 // commit kh0_free after this stage (M-tile 1 has read K-half 0 of the layer's input)
    int feat_end, code_end;      // commit feat_free / in_free after this stage
    int bytes, rows;             // bytes this CTA copies (rows x 128 B per k-chunk tile)
    int is_out;
};
struct TpDims {
    int J, nb, KF, d_feat, d_code, d_out, d_geo, Hd, rows_out;
    int n_stages;
    long long stream_bytes;      // weight stream of one CTA per tile
    long long table_off;         // byte offset of the fp32 bias table in the packed buffer
};
__host__ __device__ inline int tp_num_stages(const TpDims& d) {
    return d.J * (d.KF + 1) + d.nb * d.J * 8 + (d.nb - 1) * d.J + 4;
}
// stage s of the program (rank = CTA rank: only the byte counts depend on it)
__host__ __device__ inline TpStage tp_stage(const TpDims& d, int s, int rank) {
    TpStage st = {};
    st.rows = 128, st.nkc = 1;
    const int n_in = d.J * (d.KF + 1);
    if (s < n_in) {                                              // lin_in(j) x KF, lin_z_0(j)
        const int j = s / (d.KF + 1), t = s % (d.KF + 1);
        st.j = j, st.blk = 0, st.d_col = 128 * j;
        if (t < d.KF) {
            st.kind = 0, st.kc0 = t;
            const int kk = d.d_feat - 64 * t;
            st.n16 = ((kk < 64 ? kk : 64) + 15) / 16;
            st.overwrite = t == 0;
            st.feat_end = (j == d.J - 1 && t == d.KF - 1);
        } else {
            st.kind = 1, st.kc0 = 0, st.n16 = (d.d_code + 15) / 16;
            st.acc_end = 1;
            st.code_end = (d.nb == 1 && j == d.J - 1);
        }
        st.bytes = TP_TILE;
        return st;
    }
    s -= n_in;
    const int per_blk_full = d.J * 4 + d.J * 5;                 // fc0: J x 4 stages; then J x (lin_z_{i+1} + 4 fc1 stages)
    int i = 0;
    // blocks 0 .. nb-2 have the lin_z of the next block, the last block has not
    while (i < d.nb - 1 && s >= per_blk_full) s -= per_blk_full, ++i;
    const bool last = i == d.nb - 1;
    if (last && s >= d.J * 8) {                                  // lin_out
        s -= d.J * 8;
        st.kind = 2, st.j = 0, st.blk = d.nb, st.kc0 = 2 * s, st.nkc = 2, st.n16 = 4, st.d_col = 256;
        st.overwrite = s == 0, st.acc_end = s == 3 ? 3 : 0, st.is_out = 1;
        st.rows = rank == 0 ? d.rows_out : 0;
        st.bytes = 2 * st.rows * 128;
        return st;
    }
    st.blk = i;
    if (s < d.J * 4) {                                           // fc0_i(j), 4 stages
        const int j = s / 4, p = s % 4;
        st.kind = 2, st.j = j, st.kc0 = 2 * p, st.nkc = 2, st.n16 = 4, st.d_col = 256 + 128 * j;
        st.overwrite = p == 0, st.acc_end = p == 3;
        st.kh0_end = (d.J == 2 && j == 1 && p == 1);
        st.bytes = TP_STAGE;
        return st;
    }
    s -= d.J * 4;
    const int per_j = last ? 4 : 5;
    const int j = s / per_j;
    int p = s % per_j;
    st.j = j, st.d_col = 128 * j;
    if (!last && p == 0) {                                       // lin_z_{i+1}(j)
        st.kind = 1, st.blk = i + 1, st.kc0 = 0, st.n16 = (d.d_code + 15) / 16;
        st.code_end = (i + 1 == d.nb - 1 && j == d.J - 1);
        st.bytes = TP_TILE;
        return st;
    }
    if (!last) --p;
    st.kind = 2, st.kc0 = 2 * p, st.nkc = 2, st.n16 = 4;       // fc1_i(j)
    st.acc_end = p == 3;
    st.kh0_end = (d.J == 2 && j == 1 && p == 1);
    st.bytes = TP_STAGE;
    return st;
}

// kind::f16 instruction descriptor, M = 256 over the pair, N = 128; B MN-major for the activation chunks
__device__ __forceinline__ uint32_t tp_idesc(bool bf16, bool b_mn_major) {
    const uint32_t f = bf16 ? 1u : 0u;
    return (1u << 4) | (f << 7) | (f << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(TP_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void tp_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tp_commit(uint32_t bar, uint16_t mask) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_u(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_u(uint32_t bar_cluster) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n\t}" ::"r"(bar_cluster)
                 : "memory");
}
// offset of 16-byte unit u of 128-byte row r in a 128B-swizzled tile (8-row atoms of 1024 B): A tiles (row = hidden unit, 64 k),
// K-major B tiles (row = query, 64 k) and MN-major B chunks (row = k, 64 queries) all use it
__device__ __host__ inline uint32_t tp_off(int r, int u) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)); }

#ifdef GNB_TC_TRACE
#define TP_TRACE(role, k) do { if (p.dbg && blockIdx.x == 0 && lane == 0 && (k) < 4096) p.dbg[(role) * 4096 + (k)] = clock64(); } while (0)
#else
#define TP_TRACE(role, k) do { } while (0)
#endif

// The program, decoded once on the host (tp_stage has divisions and a loop: ~1 000 cycles per stage when every lane of the
// converged issuing warp evaluates it) into two words per stage, read from the kernel parameters with a uniform index:
//   x: kind [0,2) | j [2,3) | kc0 [3,6) | nkc-1 [6,7) | n16 [7,10) | d_col/128 [10,12) | overwrite [12] | acc_end [13,15) |
//      kh0_end [15] | feat_end [16] | code_end [17] | is_out [18] | layer_end [19]
//   y: bytes CTA 0 copies | bytes CTA 1 copies << 16   (in units of 128 B)
constexpr int TP_MAX_STAGES = 136;
struct TpKP {
    TcKP k;                      // queries, sampler, outputs, weights (biases / head / encoding options) as in decoder_tc_kernel
    TpDims d;
    const float* table;          // [Hd][2 nb + 1] fp32: cumulative x-bias before block i (i = 0..nb), then fc_0 biases
    uint2 prog[TP_MAX_STAGES];
};
__host__ inline uint2 tp_encode(const TpDims& d, int s) {
    const TpStage a = tp_stage(d, s, 0), b = tp_stage(d, s, 1);
    const int layer_end = (a.acc_end && a.kind == 2 && !a.is_out && a.j == d.J - 1) ? 1 : 0;
    uint2 r;
    r.x = (uint32_t)a.kind | ((uint32_t)a.j << 2) | ((uint32_t)a.kc0 << 3) | ((uint32_t)(a.nkc - 1) << 6) | ((uint32_t)a.n16 << 7) |
          ((uint32_t)(a.d_col / 128) << 10) | ((uint32_t)a.overwrite << 12) | ((uint32_t)a.acc_end << 13) | ((uint32_t)a.kh0_end << 15) |
          ((uint32_t)a.feat_end << 16) | ((uint32_t)a.code_end << 17) | ((uint32_t)a.is_out << 18) | ((uint32_t)layer_end << 19);
    r.y = (uint32_t)(a.bytes / 128) | ((uint32_t)(b.bytes / 128) << 16);
    return r;
}
// four K16 MMAs of one 64-wide k-chunk in ONE asm block (descriptor steps: A +32 B, B +32 B K-major / +2 048 B MN-major)
__device__ __forceinline__ void tp_mma_x4(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate, uint32_t bstep) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, b1, a2, b2, a3, b3, bs;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "cvt.u64.u32 bs, %5;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, bs;\n\t"
        "add.s64 a2, %1, 4;\n\tadd.s64 b2, b1, bs;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, b2, bs;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, 1;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(bstep)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(TP_THREADS, 1) decoder_tp_kernel(const __grid_constant__ TpKP P) {
    extern __shared__ unsigned char smem_raw[];
    const TcKP& p = P.k;
    const TpDims& d = P.d;
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TpSmem L = tp_smem_layout(d.KF);
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + L.bars;
    auto w_full = [&](int s) { return bar0 + 8u * s; };                 // leader: own copy landed + partner's forward; partner: own copy
    auto w_empty = [&](int s) { return bar0 + 8u * (4 + s); };
    auto act_ready = [&](int j) { return bar0 + 8u * (8 + j); };        // leader: K-half j of the next B operand written (8 warps) + landed at the partner
    auto land = [&](int j) { return bar0 + 8u * (10 + j); };            // the partner's half of K-half j landed HERE
    auto acc = [&](int j) { return bar0 + 8u * (12 + j); };             // accumulation group of M-tile j complete; 2: lin_out
    const uint32_t kh0_free = bar0 + 8u * 15, in_ready = bar0 + 8u * 16, in_free = bar0 + 8u * 17, feat_free = bar0 + 8u * 18;
    const uint32_t xs_full = bar0 + 8u * 19, xs_free = bar0 + 8u * 20;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + L.bars + 40 * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int J = d.J, nb = d.nb;
    const int n_rounds = 2 * nb + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TP_NSTAGE; ++s) { mbar_init(w_full(s), leader ? 2 : 1); mbar_init(w_empty(s), 1); }
        for (int j = 0; j < 2; ++j) { mbar_init(act_ready(j), 8 + 1); mbar_init(land(j), 1); }
        for (int j = 0; j < 3; ++j) mbar_init(acc(j), 1);
        mbar_init(kh0_free, 1);
        mbar_init(in_ready, 2);
        mbar_init(in_free, 1);
        mbar_init(feat_free, 1);
        mbar_init(xs_full, 4);
        mbar_init(xs_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = cluster_id < p.n_tiles ? (p.n_tiles - 1 - cluster_id) / p.n_clusters + 1 : 0;
    const unsigned char* wstream = p.packed + (long long)rank * d.stream_bytes;
    const int nst = d.n_stages;

    if (warp == 0) {
        // =============================== weight producer (converged) ============================
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const unsigned char* src = wstream;
            for (int s = 0; s < nst; ++s) {
                const uint2 pg = P.prog[s];
                const uint32_t bytes = ((pg.y >> (16 * rank)) & 0xffffu) * 128u;
                const bool is_out = (pg.x >> 18) & 1u;
                mbar_wait(w_empty(slot), phase ^ 1);
                const uint32_t dst = sbase + L.ring + slot * TP_STAGE;
                if (is_out) {
                    // lin_out: only the real output rows are copied (the rest of the tile is stale, finite data whose result
                    // rows nobody reads); CTA 1 has none and just signals
                    if (bytes > 0) {
                        const uint32_t half = bytes / 2;
                        mbar_expect_tx_u(w_full(slot), 2 * half);
                        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                                     "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
                                     "l"(src), "r"(half), "r"(w_full(slot))
                                     : "memory");
                        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                                     "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst + TP_TILE),
                                     "l"(src + half), "r"(half), "r"(w_full(slot))
                                     : "memory");
                    } else {
                        mbar_arrive_u(w_full(slot));
                    }
                } else {
                    if (P.k.nocopy) mbar_expect_tx_u(w_full(slot), 0);
                    else bulk_g2s_u(dst, src, bytes, w_full(slot));
                }
                src += bytes;
                if (++slot == TP_NSTAGE) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && !leader) {
        // =============================== partner: forward "my weight stage landed" to the leader ==========
        int slot = 0;
        uint32_t phase = 0;
        const long long total = (long long)my_tiles * nst;
        for (long long i = 0; i < total; ++i) {
            mbar_wait(w_full(slot), phase);
            mbar_arrive_remote_u(map_to_cta(w_full(slot), 0));
            if (++slot == TP_NSTAGE) { slot = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // =============================== MMA issue (leader, converged warp) =====================
        const uint32_t idesc_k = tp_idesc(BF16, false), idesc_mn = tp_idesc(BF16, true);
        int slot = 0;
        uint32_t phase = 0, layer = 0;               // layer: hidden layers whose B operand is ACT (rounds consumed so far)
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(in_ready, it & 1);
            tc_fence_after();
            int kh_ready = 0;                        // K-halves of the current layer's operand already waited for
            for (int s = 0; s < nst; ++s) {
                const uint32_t pg = P.prog[s].x;
                const uint32_t kind = pg & 3u, mj = (pg >> 2) & 1u, kc0 = (pg >> 3) & 7u, nkc = ((pg >> 6) & 1u) + 1u, n16 = (pg >> 7) & 7u;
                const uint32_t dcol = tmem + ((pg >> 10) & 3u) * 128u, overwrite = (pg >> 12) & 1u, acc_end = (pg >> 13) & 3u;
                TP_TRACE(3, it * 128 + s);                       // stage reached
                if (kind == 2) {
                    const int kh = (int)(kc0 >> 2);              // K-half of this stage's chunks
                    if (mj == 0 && (kc0 & 3u) == 0 && !(kh_ready & (1 << kh))) {
                        mbar_wait(act_ready(kh), layer & 1);
                        mbar_wait(land(kh), layer & 1);
                        kh_ready |= 1 << kh;
                    }
                }
                TP_TRACE(0, it * 128 + s);                       // before the weight wait
                mbar_wait(w_full(slot), phase);
                tc_fence_after();
                TP_TRACE(2, it * 128 + s);                       // issue
                const uint32_t a_base = sbase + L.ring + slot * TP_STAGE;
                if (kind == 2) {                                  // hidden layer / lin_out: two k-chunks, 8 MMAs, B = ACT chunks (MN-major)
                    const uint32_t b_base = sbase + L.act + kc0 * TP_ACT_CHUNK;
                    tp_mma_x4(dcol, umma_desc(a_base), umma_desc(b_base), idesc_mn, overwrite ? 0u : 1u, 128u);
                    tp_mma_x4(dcol, umma_desc(a_base + TP_TILE), umma_desc(b_base + TP_ACT_CHUNK), idesc_mn, 1u, 128u);
                } else {                                          // lin_in / lin_z: one k-chunk, n16 MMAs, B = feature / code tile (K-major)
                    const uint64_t da = umma_desc(a_base);
                    const uint64_t db = umma_desc(kind == 0 ? sbase + L.feat + kc0 * TP_ACT_CHUNK : sbase + L.code);
                    for (uint32_t k = 0; k < n16; ++k) tp_mma(dcol, da + 2 * k, db + 2 * k, idesc_k, (overwrite && k == 0) ? 0u : 1u);
                }
                (void)nkc;
                tp_commit(w_empty(slot), 3);
                if ((pg >> 16) & 1u) tp_commit(feat_free, 3);
                if ((pg >> 17) & 1u) tp_commit(in_free, 3);
                if ((pg >> 15) & 1u) tp_commit(kh0_free, 3);
                if (acc_end) {
                    tp_commit(acc(acc_end == 3 ? 2 : (int)mj), 3);
                    // the layer is complete when its last M-tile is: the next stages read the NEXT round's operand
                    if ((pg >> 19) & 1u) { ++layer; kh_ready = 0; }
                }
                if (++slot == TP_NSTAGE) { slot = 0; phase ^= 1; }
            }
            ++layer;                                 // lin_out consumed round 2 nb
        }
    } else if (warp == 3) {
        // =============================== input staging: own 64 queries of the next tile ===========
        Smem SL = {};
        SL.code = L.code, SL.feat = L.feat;
        uint32_t ovf = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = cluster_id + it * p.n_clusters;
            const long long row0 = (long long)tile * TP_BN + rank * 64;
            if (it > 0) mbar_wait(feat_free, (it - 1) & 1);
            for (int rr = 0; rr < 2; ++rr) stage_inputs<BF16>(p, sm, SL, 0, rr * 32 + lane, row0 + rr * 32 + lane, 0, 1, 2, ovf, TP_ACT_CHUNK);
            if (it > 0) mbar_wait(in_free, (it - 1) & 1);
            for (int rr = 0; rr < 2; ++rr) stage_inputs<BF16>(p, sm, SL, 0, rr * 32 + lane, row0 + rr * 32 + lane, 0, 1, 1, ovf, TP_ACT_CHUNK);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(in_ready); else mbar_arrive_remote(map_to_cta(in_ready, 0)); }
        }
        if (ovf && p.w.status) atomicOr(p.w.status, 1);
    } else if (warp == 2) {
        // =============================== push of the partner's half (converged) ==================
        // per round and M-tile: the 4 warps of the remote group have staged 128 k-rows x 64 queries (16 KB = 2 chunks) in XS
        uint32_t n = 0;
        const long long total = (long long)my_tiles * n_rounds * J;
        for (long long i = 0; i < total; ++i, ++n) {
            const int j = (int)(i % J);
            const uint32_t rnd = (uint32_t)(i / J);                   // global round index: phase of land(j)
            mbar_wait(xs_full, n & 1);
            mbar_expect_tx_u(land(j), 2 * TP_ACT_CHUNK);              // MY landing of the partner's push of the same round / M-tile
            bulk_s2peer_u(map_to_cta(sbase + L.act + (4 * j + 2 * rank) * TP_ACT_CHUNK, peer), sbase + L.xs, 2 * TP_ACT_CHUNK,
                          map_to_cta(land(j), peer));
            // A copy that completes on an mbarrier is not part of a bulk async-group, so the sender cannot wait for its own
            // reads: the RECEIVER acknowledges.  The partner's push of the same round / M-tile has landed here -> its bounce
            // buffer is free again (and, on CTA 1, the leader learns that K-half j is complete here as well).
            mbar_wait(land(j), rnd & 1);
            mbar_arrive_remote_u(map_to_cta(xs_free, peer));
            if (!leader) mbar_arrive_remote_u(map_to_cta(act_ready(j), 0));
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===============================================
        const int q = warp & 3, eg = (warp - 4) >> 2;
        const int hl = q * 32 + lane;                              // hidden unit within this CTA's 128 of an M-tile == TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const bool keep = (uint32_t)eg == rank;                    // this group's queries belong to this CTA
        const int tstride = 2 * nb + 1;
        uint32_t ovf = 0, xs_use = 0;
        uint32_t rnd = 0, fcl = 0;                                 // global round / hidden-layer counters (barrier phases)
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = cluster_id + it * p.n_clusters;
            for (int r = 0; r < n_rounds; ++r, ++rnd) {
                const bool from_net = (r & 1) == 1;
                uint32_t pk[2][32];                                // M-tile 0's converted half is held until K-half 0 may be overwritten
                for (int j = 0; j < J; ++j) {
                    mbar_wait(acc(j), rnd & 1);
                    tc_fence_after();
                    if (warp == 4 || warp == 8) { TP_TRACE(1, (warp == 8 ? 2048 : 0) + it * 128 + r * 8 + j * 4 + 0); }
                    const int h = j * 256 + (int)rank * 128 + hl;
                    const float bias = __ldg(P.table + (long long)h * tstride + (from_net ? nb + 1 + (r >> 1) : (r >> 1)));
                    const uint32_t col = (from_net ? 256u : 0u) + 128u * j + 64u * eg;
                    uint32_t v[32];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        tmem_ld32(tlane + col + hh * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            pk[j & 1][hh * 16 + e] = pack16_relu<BF16>(__uint_as_float(v[2 * e]) + bias, __uint_as_float(v[2 * e + 1]) + bias);
                    }
                    if constexpr (!BF16) {
#pragma unroll
                        for (int e = 0; e < 32; e += 2) ovf |= sat_probe(pk[j & 1][e], pk[j & 1][e + 1]);
                    }
                    // K-half 0 of the operand is still read by M-tile 1's first stages of the layer that produced this accumulator
                    if (warp == 4 || warp == 8) { TP_TRACE(1, (warp == 8 ? 2048 : 0) + it * 128 + r * 8 + j * 4 + 1); }
                    if (J == 2 && j == 0 && r > 0) mbar_wait(kh0_free, fcl & 1);
                    if (warp == 4 || warp == 8) { TP_TRACE(1, (warp == 8 ? 2048 : 0) + it * 128 + r * 8 + j * 4 + 2); }
                    unsigned char* dst;
                    if (keep) dst = sm + L.act + (4 * j + 2 * (int)rank + (q >> 1)) * TP_ACT_CHUNK;
                    else {
                        if (xs_use > 0) mbar_wait(xs_free, (xs_use - 1) & 1);
                        ++xs_use;
                        dst = sm + L.xs + (q >> 1) * TP_ACT_CHUNK;
                    }
                    const int row = (q & 1) * 32 + lane;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<uint4*>(dst + tp_off(row, u)) =
                            make_uint4(pk[j & 1][4 * u], pk[j & 1][4 * u + 1], pk[j & 1][4 * u + 2], pk[j & 1][4 * u + 3]);
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (keep) { if (leader) mbar_arrive(act_ready(j)); else mbar_arrive_remote(map_to_cta(act_ready(j), 0)); }
                        else mbar_arrive(xs_full);
                    }
                    if (warp == 4 || warp == 8) { TP_TRACE(1, (warp == 8 ? 2048 : 0) + it * 128 + r * 8 + j * 4 + 3); }
                }
                if (r > 0) ++fcl;
            }
            // ---------------- final: lin_out accumulator (CTA 0 lanes = output features) -> out, TSDF ----------------
            mbar_wait(acc(2), it & 1);
            tc_fence_after();
            if (leader) {
                const int o = hl;                                    // output feature of this lane
                const float bo = o < d.d_out ? __ldg(p.w.lin_out_b + o) : 0.0f;
                const float hwv = o < d.d_geo ? __ldg(p.w.head_w + o) : 0.0f;
                const float hb = __ldg(p.w.head_b);
                const long long qrow0 = (long long)tile * TP_BN + 64 * eg;
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld32(tlane + 256u + 64u * eg + hh * 32, v);
                    tmem_ld_wait();
                    float part[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float val = __uint_as_float(v[e]) + bo;
                        part[e] = val * hwv;
                        const long long grow = qrow0 + hh * 32 + e;
                        if (o < d.d_out && grow < p.n_rows && p.out) {
                            const long long orow = p.sorted ? (long long)__float_as_int(__ldg(reinterpret_cast<const float*>(p.sorted + grow) + 3)) : grow;
                            p.out[orow * d.d_out + o] = val;
                        }
                    }
                    if (q == 0 && p.tsdf) {                          // d_geo <= 32: the geometric features are lanes 0..d_geo-1 of warp quadrant 0
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                            for (int i = 0; i < off; ++i) {
                                const float lo = part[i], hi = part[i + off];
                                const float send = (lane & off) ? lo : hi, kept = (lane & off) ? hi : lo;
                                part[i] = kept + __shfl_xor_sync(FULL, send, off);
                            }
                        }
                        const long long grow = qrow0 + hh * 32 + lane;
                        if (grow < p.n_rows) {
                            const long long orow = p.sorted ? (long long)__float_as_int(__ldg(reinterpret_cast<const float*>(p.sorted + grow) + 3)) : grow;
                            p.tsdf[orow] = tanhf(part[0] + hb);
                        }
                    }
                }
            }
            tc_fence_before();
        }
        if (ovf && p.w.status) atomicOr(p.w.status, 1);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// packing: fp32 nn.Linear matrices -> the two CTAs' weight streams (A tiles in program order) + the fp32 bias table
// ---------------------------------------------------------------------------------------------------------------------
struct TpPackOp {
    const float* W;          // (rows_true, K_true) row-major
    int rows_true, K_true;
    int n0;                  // first matrix row of this tile (hidden unit / output feature of tile row 0)
    int rows;                // rows written (<= 128)
    int kc;                  // 64-wide k-chunk
    float scale;
    const float* scale_dev;  // optional device copy of the scale (GnbDecoderWeights.alpha_dev)
    long long dst_off;
};
template <bool BF16>
__global__ void tp_pack_kernel(TpPackOp op, unsigned char* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 16-byte unit
    if (idx >= op.rows * 8) return;
    const int r = idx >> 3, u = idx & 7, n = op.n0 + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = op.kc * 64 + u * 8 + e;
        v[e] = (n < op.rows_true && k < op.K_true) ? (op.scale_dev ? __ldg(op.scale_dev) : op.scale) * op.W[(long long)n * op.K_true + k] : 0.0f;
    }
    *reinterpret_cast<uint4*>(dst + op.dst_off + tp_off(r, u)) =
        make_uint4(pack16<BF16>(v[0], v[1]), pack16<BF16>(v[2], v[3]), pack16<BF16>(v[4], v[5]), pack16<BF16>(v[6], v[7]));
}
// table[h][i] i <= nb: bias of x before block i's fc_0 (i = nb: before lin_out) = b_in + sum_{t<=min(i,nb-1)} alpha bz_t + sum_{t<i} b1_t;
// table[h][nb + 1 + i]: fc_0 bias of block i
__global__ void tp_table_kernel(GnbDecoderWeights w, float* __restrict__ table) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= w.d_hidden) return;
    const int nb = w.n_blocks, stride = 2 * nb + 1;
    float cb = w.lin_in_b[h];
    const float alpha = w.alpha_dev ? __ldg(w.alpha_dev) : w.alpha;
    for (int i = 0; i <= nb; ++i) {
        if (i < nb) cb += alpha * w.lin_z_b[i][h];
        if (i > 0) cb += w.fc1_b[i - 1][h];
        table[(long long)h * stride + i] = cb;
    }
    for (int i = 0; i < nb; ++i) table[(long long)h * stride + nb + 1 + i] = w.fc0_b[i][h];
}

// applies to: d_hidden 512, one feature chunk (d_feat <= 64), d_code <= 64, d_out <= 128, d_geo <= 32 (head reduced inside one warp)
static bool tp_applies(const GnbDecoderWeights* w, TpDims& d) {
    if (!opt(OPT_TC_PAIR) || opt(OPT_TC_TWO_CTA)) return false;     // opt-in (GNB_TC_PAIR=1): slower than the query-major kernel, tools/experiments/README.md
    if (w->d_hidden != 512) return false;                        // (two M-tiles of 256 hidden units; K = 512 = 4 stages of 2 k-chunks)
    if (w->d_feat > 64 || w->d_feat < 1 || w->d_code > 64 || w->d_out > 128 || w->d_geo > 32 || w->d_geo > w->d_out || w->n_blocks < 1) return false;
    d.Hd = w->d_hidden, d.J = d.Hd / 256, d.nb = w->n_blocks, d.KF = (w->d_feat + 63) / 64, d.d_feat = w->d_feat, d.d_code = w->d_code;
    d.d_out = w->d_out, d.d_geo = w->d_geo, d.rows_out = (w->d_out + 7) / 8 * 8;
    d.n_stages = tp_num_stages(d);
    if (d.n_stages > TP_MAX_STAGES) return false;
    long long b = 0;
    for (int s = 0; s < d.n_stages; ++s) {
        const TpStage st = tp_stage(d, s, 0);
        b += st.bytes;
    }
    d.stream_bytes = (b + 1023) / 1024 * 1024;
    d.table_off = 2 * d.stream_bytes;
    return tp_smem_layout(d.KF).total + 1024 <= 227 * 1024;
}
static long long tp_packed_bytes(const TpDims& d) { return d.table_off + (long long)d.Hd * (2 * d.nb + 1) * 4; }

static int tp_pack(const GnbDecoderWeights* w, const TpDims& d, void* packed, cudaStream_t st) {
    const bool bf = w->tc_dtype == GNB_TC_BF16;
    for (int rank = 0; rank < 2; ++rank) {
        unsigned char* dst = (unsigned char*)packed + (long long)rank * d.stream_bytes;
        long long off = 0;
        for (int s = 0; s < d.n_stages; ++s) {
            const TpStage sg = tp_stage(d, s, rank);
            for (int c = 0; c < sg.nkc; ++c) {
                TpPackOp op = {};
                op.kc = sg.kc0 + c, op.scale = 1.0f, op.rows = sg.rows;
                op.n0 = sg.j * 256 + rank * 128;
                op.rows_true = d.Hd;
                if (sg.kind == 0) op.W = w->lin_in_w, op.K_true = d.d_feat;
                else if (sg.kind == 1) op.W = w->lin_z_w[sg.blk], op.K_true = d.d_code, op.scale = w->alpha, op.scale_dev = w->alpha_dev;
                else if (sg.is_out) op.W = w->lin_out_w, op.K_true = d.Hd, op.rows_true = d.d_out, op.n0 = 0;
                else {
                    // stage index inside the block tells fc_0 from fc_1: d_col >= 256 <-> net accumulator <-> fc_0
                    op.W = sg.d_col >= 256 ? w->fc0_w[sg.blk] : w->fc1_w[sg.blk], op.K_true = d.Hd;
                }
                op.dst_off = off;
                if (op.rows > 0) {
                    if (bf) tp_pack_kernel<true><<<ceil_div(op.rows * 8, 256), 256, 0, st>>>(op, dst);
                    else tp_pack_kernel<false><<<ceil_div(op.rows * 8, 256), 256, 0, st>>>(op, dst);
                    GNB_LAUNCH_CHECK();
                }
                off += (long long)sg.rows * 128;
            }
        }
        if (off > d.stream_bytes) { set_error("gnb_decoder_pack_tc: internal size mismatch (pair layout)"); return GNB_E_INVALID; }
    }
    tp_table_kernel<<<ceil_div(d.Hd, 128), 128, 0, st>>>(*w, reinterpret_cast<float*>((unsigned char*)packed + d.table_off));
    GNB_LAUNCH_CHECK();
    return 0;
}

static int tp_launch(const GnbDecoderWeights* w, const TpDims& d, const void* packed, TcKP& kp, void* stream, long long* trace) {
    TpKP P = {};
    kp.w = *w;
    kp.packed = (const unsigned char*)packed;
    kp.d = Dims{};
    kp.d.d_feat = d.d_feat, kp.d.d_code = d.d_code, kp.d.Hd = d.Hd, kp.d.nb = d.nb, kp.d.d_out = d.d_out, kp.d.d_geo = d.d_geo;
    kp.d.KF = d.KF, kp.d.KZ = 1;
    P.d = d;
    if (d.n_stages > TP_MAX_STAGES) { set_error("gnb_decode_tc: program too long for the pair kernel"); return GNB_E_UNSUPPORTED; }
    for (int s = 0; s < d.n_stages; ++s) P.prog[s] = tp_encode(d, s);
    P.table = reinterpret_cast<const float*>((const unsigned char*)packed + d.table_off);
    int dev = 0, sms = 0, cc = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GNB_CUDA(cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc != 10) { set_error("gnb_decode_tc: needs an sm_100 device (found sm_%d0)", cc); return GNB_E_ARCH; }
    kp.dbg = trace;
    kp.nocopy = opt(OPT_DEBUG_NO_WCOPY);
    kp.n_tiles = (int)((kp.n_rows + TP_BN - 1) / TP_BN);
    kp.n_clusters = sms / 2;
    if (const int m = opt(OPT_DEBUG_MAX_CLUSTERS)) {
        if (m > 0 && m < kp.n_clusters) kp.n_clusters = m;
    }
    if (kp.n_clusters > kp.n_tiles) kp.n_clusters = kp.n_tiles;
    const size_t smem = tp_smem_layout(d.KF).total + 1024;
    auto kernel = (w->tc_dtype == GNB_TC_BF16) ? decoder_tp_kernel<true> : decoder_tp_kernel<false>;
    GNB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(TP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    int max_clusters = 0;
    cfg.gridDim = dim3(sms / 2 * 2);
    GNB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    if (max_clusters > 0 && kp.n_clusters > max_clusters) kp.n_clusters = max_clusters;
    cfg.gridDim = dim3(kp.n_clusters * 2);
    P.k = kp;
    GNB_CUDA(cudaLaunchKernelEx(&cfg, kernel, P));
    return 0;
}

}  // namespace tc
}  // namespace gnb

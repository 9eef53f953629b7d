// Backward links of the ResNet-MLP decoder.
//
// Under autograd the reference's ResnetFC (src/models/components/resnetfc.py:54-63, 134-189) turns every Linear + ReLU link
// of the backward pass into a GEMM followed by three element-wise / reduction kernels: threshold_backward (the ReLU mask),
// the residual add of the skip connection, and a column sum for the bias gradient.  With the 16-bit activations the tcgen05
// forward kernel saved (gnb_decode_tc_save) those three are ONE pass over the (n, d) gradient here:
//
//     out[r, c]  = (act[r, c] > 0 ? pre[r, c] : 0) + (res ? res[r, c] : 0)
//     colsum[c] += sum_r out[r, c]                                          (bias gradient of the layer that produced out)
//     act32[r, c] = (float)act[r, c]                                        (optional: left operand of the wgrad GEMM)
//
// Every element is read once and written once by the same thread, so `out` may alias `pre` or `res`.
#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace gnb {

struct LinkKP {
    const float* pre;
    const uint16_t* act;
    const float* res;
    float* out;
    float* act32;
    float* colsum;
    int64_t n, ld_pre, ld_act, ld_res, ld_out, ld_act32;
    int d, rows_per_block;
};

template <bool BF16>
__device__ __forceinline__ float act_to_float(uint16_t b) {
    if (BF16) return __uint_as_float((unsigned)b << 16);
    return __half2float(__ushort_as_half(b));
}

constexpr int LINK_COLS = 128;   // float4 column groups per block (512 columns)
constexpr int LINK_UNROLL = 4;   // rows in flight per thread

// block = 256 threads = 2 row lanes x 128 column groups of 4 floats; grid = (row blocks, column slabs of 512)
template <bool BF16, bool HAS_RES, bool WRITE_ACT>
__global__ void __launch_bounds__(256) mlp_grad_link_kernel(const LinkKP p) {
    __shared__ float4 part[LINK_COLS];
    const int cg = blockIdx.y * LINK_COLS + (threadIdx.x & (LINK_COLS - 1));
    const int rl = threadIdx.x >> 7;
    const bool live = cg * 4 < p.d;
    const int64_t r0 = (int64_t)blockIdx.x * p.rows_per_block;
    const int64_t r1 = min(r0 + (int64_t)p.rows_per_block, p.n);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        const int c = cg * 4;
        for (int64_t r = r0 + rl; r < r1; r += 2 * LINK_UNROLL) {
            float4 g[LINK_UNROLL], q[LINK_UNROLL];
            uint2 a[LINK_UNROLL];
#pragma unroll
            for (int u = 0; u < LINK_UNROLL; ++u) {
                const int64_t ru = r + 2 * u;
                if (ru < r1) {
                    g[u] = *reinterpret_cast<const float4*>(p.pre + ru * p.ld_pre + c);
                    a[u] = *reinterpret_cast<const uint2*>(p.act + ru * p.ld_act + c);
                    if (HAS_RES) q[u] = *reinterpret_cast<const float4*>(p.res + ru * p.ld_res + c);
                }
            }
#pragma unroll
            for (int u = 0; u < LINK_UNROLL; ++u) {
                const int64_t ru = r + 2 * u;
                if (ru < r1) {
                    float4 f;
                    f.x = act_to_float<BF16>((uint16_t)(a[u].x & 0xffffu));
                    f.y = act_to_float<BF16>((uint16_t)(a[u].x >> 16));
                    f.z = act_to_float<BF16>((uint16_t)(a[u].y & 0xffffu));
                    f.w = act_to_float<BF16>((uint16_t)(a[u].y >> 16));
                    float4 o;
                    o.x = f.x > 0.f ? g[u].x : 0.f;
                    o.y = f.y > 0.f ? g[u].y : 0.f;
                    o.z = f.z > 0.f ? g[u].z : 0.f;
                    o.w = f.w > 0.f ? g[u].w : 0.f;
                    if (HAS_RES) o.x += q[u].x, o.y += q[u].y, o.z += q[u].z, o.w += q[u].w;
                    *reinterpret_cast<float4*>(p.out + ru * p.ld_out + c) = o;
                    if (WRITE_ACT) *reinterpret_cast<float4*>(p.act32 + ru * p.ld_act32 + c) = f;
                    s.x += o.x, s.y += o.y, s.z += o.z, s.w += o.w;
                }
            }
        }
    }
    if (p.colsum == nullptr) return;
    if (rl == 1) part[threadIdx.x & (LINK_COLS - 1)] = s;
    __syncthreads();
    if (rl == 0 && live) {
        const float4 t = part[threadIdx.x];
        float* cs = p.colsum + cg * 4;
        atomicAdd(cs + 0, s.x + t.x);
        atomicAdd(cs + 1, s.y + t.y);
        atomicAdd(cs + 2, s.z + t.z);
        atomicAdd(cs + 3, s.w + t.w);
    }
}

// Gradient entering the decoder from its two outputs (reference heads3d.py:36-50 + model.py:226-246 under autograd):
//   s[r]   = g_tsdf[r] * (1 - tsdf[r]^2)                              (through tanh)
//   G[r,c] = g_out[r,c] + (c < d_geo ? s[r] * head_w[c] : 0)          (written to `G`)
//   d_head_w[c] += sum_r s[r] * out[r,c] (c < d_geo),  d_head_b += sum_r s[r],  d_lin_out_b[c] += sum_r G[r,c]
// one warp per 32 rows x d_out columns (d_out <= 256), lanes = columns.
__global__ void __launch_bounds__(256) mlp_grad_head_kernel(const float* g_out, const float* g_tsdf, const float* out, const float* tsdf,
                                                            const float* head_w, int64_t n, int d_out, int d_geo, float* G,
                                                            float* d_head_w, float* d_head_b, float* d_lin_out_b, int rows_per_block) {
    extern __shared__ float sm[];            // [3][d_out] block sums: bias, head weight; [1] head bias
    float* sb = sm;
    float* sw = sm + d_out;
    float* shb = sm + 2 * d_out;
    for (int i = threadIdx.x; i < 2 * d_out + 1; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + (int64_t)rows_per_block, n);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int c0 = 0; c0 < d_out; c0 += 32) {
        const int c = c0 + lane;
        const bool cl = c < d_out;
        const float hw = (cl && c < d_geo) ? head_w[c] : 0.f;
        float ab = 0.f, aw = 0.f, ah = 0.f;
        for (int64_t r = r0 + warp; r < r1; r += nwarp) {
            float s = 0.f;
            if (g_tsdf) {
                const float t = tsdf[r];
                s = g_tsdf[r] * (1.f - t * t);
            }
            if (cl) {
                float g = (g_out ? g_out[r * d_out + c] : 0.f) + s * hw;
                G[r * d_out + c] = g;
                ab += g;
                if (c < d_geo) aw += s * out[r * d_out + c];
            }
            ah += s;
        }
        if (cl) {
            atomicAdd(sb + c, ab);
            if (c < d_geo) atomicAdd(sw + c, aw);
        }
        if (c0 == 0 && lane == 0) atomicAdd(shb, ah);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d_out; i += blockDim.x) {
        if (d_lin_out_b) atomicAdd(d_lin_out_b + i, sb[i]);
        if (d_head_w && i < d_geo) atomicAdd(d_head_w + i, sw[i]);
    }
    if (threadIdx.x == 0 && d_head_b) atomicAdd(d_head_b, *shb);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_mlp_grad_link(const float* pre, int64_t ld_pre, const void* act, int64_t ld_act, int act_dtype,
                                 const float* res, int64_t ld_res, float* out, int64_t ld_out, float* act32, int64_t ld_act32,
                                 float* colsum, int64_t n_rows, int d, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0 && d >= 4 && d % 4 == 0, "gnb_mlp_grad_link: n_rows %lld, d %d (a multiple of 4)", (long long)n_rows, d);
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(pre && act && out, "gnb_mlp_grad_link: null pointer");
    GNB_CHECK_ARG(act_dtype == GNB_TC_FP16 || act_dtype == GNB_TC_BF16, "gnb_mlp_grad_link: act_dtype %d", act_dtype);
    GNB_CHECK_ARG(ld_pre >= d && ld_act >= d && ld_out >= d && (!res || ld_res >= d) && (!act32 || ld_act32 >= d),
                  "gnb_mlp_grad_link: a row stride is smaller than d");
    GNB_CHECK_ARG(ld_pre % 4 == 0 && ld_act % 4 == 0 && ld_out % 4 == 0 && (!res || ld_res % 4 == 0) && (!act32 || ld_act32 % 4 == 0),
                  "gnb_mlp_grad_link: row strides must be multiples of 4 elements");
    GNB_CHECK_ARG(al16(pre) && al16(out) && (!res || al16(res)) && (!act32 || al16(act32)) &&
                      (reinterpret_cast<uintptr_t>(act) & 7) == 0,
                  "gnb_mlp_grad_link: pointers must be 16-byte aligned (activations: 8-byte)");
    LinkKP p;
    p.pre = pre, p.act = (const uint16_t*)act, p.res = res, p.out = out, p.act32 = act32, p.colsum = colsum;
    p.n = n_rows, p.ld_pre = ld_pre, p.ld_act = ld_act, p.ld_res = ld_res, p.ld_out = ld_out, p.ld_act32 = ld_act32, p.d = d;
    // ONE balanced wave: two blocks are resident per SM (91-93 registers), so the rows are cut into at most 2 x SMs equal
    // blocks per column slab (ncu at n = 23 200 with fixed 64-row blocks: 363 blocks = 1.23 waves, SMs idle 36 % of the time);
    // the column sums then leave the kernel as <= 2 x SMs x d atomics whatever n is
    int dev = 0, sms = 148;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int slabs = (d / 4 + LINK_COLS - 1) / LINK_COLS;
    int64_t blocks = 2 * (int64_t)sms / slabs;
    if (blocks < 1) blocks = 1;
    int64_t rpb = (n_rows + blocks - 1) / blocks;
    rpb = rpb < 16 ? 16 : ((rpb + 7) & ~(int64_t)7);
    p.rows_per_block = (int)rpb;
    dim3 grid((unsigned)((n_rows + rpb - 1) / rpb), (unsigned)slabs);
    cudaStream_t st = (cudaStream_t)stream;
    const int sel = (act_dtype == GNB_TC_BF16 ? 4 : 0) | (res ? 2 : 0) | (act32 ? 1 : 0);
    switch (sel) {
        case 0: mlp_grad_link_kernel<false, false, false><<<grid, 256, 0, st>>>(p); break;
        case 1: mlp_grad_link_kernel<false, false, true><<<grid, 256, 0, st>>>(p); break;
        case 2: mlp_grad_link_kernel<false, true, false><<<grid, 256, 0, st>>>(p); break;
        case 3: mlp_grad_link_kernel<false, true, true><<<grid, 256, 0, st>>>(p); break;
        case 4: mlp_grad_link_kernel<true, false, false><<<grid, 256, 0, st>>>(p); break;
        case 5: mlp_grad_link_kernel<true, false, true><<<grid, 256, 0, st>>>(p); break;
        case 6: mlp_grad_link_kernel<true, true, false><<<grid, 256, 0, st>>>(p); break;
        default: mlp_grad_link_kernel<true, true, true><<<grid, 256, 0, st>>>(p); break;
    }
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_mlp_grad_head(const float* g_out, const float* g_tsdf, const float* out, const float* tsdf, const float* head_w,
                                 int64_t n_rows, int d_out, int d_geo, float* G, float* d_head_w, float* d_head_b,
                                 float* d_lin_out_b, void* stream) {
    GNB_CHECK_ARG(n_rows >= 0 && d_out >= 1 && d_out <= 1024 && d_geo >= 0 && d_geo <= d_out,
                  "gnb_mlp_grad_head: n_rows %lld, d_out %d, d_geo %d", (long long)n_rows, d_out, d_geo);
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(G && (g_out || g_tsdf), "gnb_mlp_grad_head: null pointer");
    GNB_CHECK_ARG(!g_tsdf || (tsdf && out && head_w), "gnb_mlp_grad_head: g_tsdf needs tsdf, out and head_w");
    int64_t rpb = (n_rows + 591) / 592;
    rpb = rpb < 64 ? 64 : rpb;
    const unsigned grid = (unsigned)((n_rows + rpb - 1) / rpb);
    mlp_grad_head_kernel<<<grid, 256, sizeof(float) * (2 * d_out + 1), (cudaStream_t)stream>>>(
        g_out, g_tsdf, out, tsdf, head_w, n_rows, d_out, d_geo, G, d_head_w, d_head_b, d_lin_out_b, (int)rpb);
    GNB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole backward pass of the decoder as ONE C call (row a15): the chain of links above around plain library GEMMs.
// cuBLAS is a library dependency of the *training* entry point only and is bound at first use with dlopen (the inference
// library has no link-time dependency on it; inside a PyTorch process the loader hands back the copy torch already mapped).
// ---------------------------------------------------------------------------------------------------------------------
namespace gnb {

struct Blas {
    void* so = nullptr;
    cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
    cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
    cublasStatus_t (*set_math)(cublasHandle_t, cublasMath_t) = nullptr;
    cublasStatus_t (*sgemm)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const float*, const float*, int,
                            const float*, int, const float*, float*, int) = nullptr;
    cublasHandle_t handle[64] = {};
    bool ok = false;
};
static Blas g_blas;
static std::mutex g_blas_mu;

static int blas_handle(cublasHandle_t* h) {
    std::lock_guard<std::mutex> lock(g_blas_mu);
    Blas& b = g_blas;
    if (!b.ok) {
        const char* names[] = {"libcublas.so.12", "libcublas.so"};
        for (const char* nm : names)
            if (!b.so) b.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (!b.so) { set_error("gnb_decode_train_bwd: cannot load libcublas.so.12 (%s)", dlerror()); return GNB_E_UNSUPPORTED; }
        b.create = (decltype(b.create))dlsym(b.so, "cublasCreate_v2");
        b.set_stream = (decltype(b.set_stream))dlsym(b.so, "cublasSetStream_v2");
        b.set_math = (decltype(b.set_math))dlsym(b.so, "cublasSetMathMode");
        b.sgemm = (decltype(b.sgemm))dlsym(b.so, "cublasSgemm_v2");
        if (!b.create || !b.set_stream || !b.set_math || !b.sgemm) { set_error("gnb_decode_train_bwd: cuBLAS symbols missing"); return GNB_E_UNSUPPORTED; }
        b.ok = true;
    }
    int dev = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("gnb_decode_train_bwd: device %d", dev); return GNB_E_INVALID; }
    if (!b.handle[dev]) {
        cublasStatus_t st = b.create(&b.handle[dev]);
        if (st != CUBLAS_STATUS_SUCCESS) { b.handle[dev] = nullptr; set_error("gnb_decode_train_bwd: cublasCreate failed (%d)", (int)st); return GNB_E_UNSUPPORTED; }
    }
    *h = b.handle[dev];
    return 0;
}

// row-major C (M x N, ldc) = op(A) op(B) [+ C]; A is stored (ta ? K x M : M x K), B (tb ? N x K : K x N), both row-major
static int gemm_rm(cublasHandle_t h, bool ta, bool tb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float* Cm, int64_t ldc, float beta = 0.f) {
    const float one = 1.f;
    cublasStatus_t st = g_blas.sgemm(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)N, (int)M, (int)K, &one, B,
                                     (int)ldb, A, (int)lda, &beta, Cm, (int)ldc);
    if (st != CUBLAS_STATUS_SUCCESS) { set_error("gnb_decode_train_bwd: cublasSgemm failed (%d)", (int)st); return GNB_E_UNSUPPORTED; }
    return 0;
}

struct FinishKP {
    const float* gz;        // (n, 16k): sum_i S_i Wz_i, padded columns are zero
    const float* code;      // (n, dcp) zero-padded codes
    const float* dwz;       // (nb*H, dcp): S_i^T code
    const float* bz[8];     // lin_z biases
    const float* sum_s[8];  // column sums of S_i (= the bias gradient buffers of lin_in / fc_1 of block i-1)
    float* d_wz[8];         // (H, dc) each
    float* d_bz[8];         // (H) each, overwritten
    float* d_alpha;         // accumulated
    float* g_code;          // (n, dc) or null, overwritten
    const float* alpha_dev;
    float alpha;
    int64_t n;
    int nb, H, dc, dcp;
};

__global__ void __launch_bounds__(256) mlp_grad_finish_kernel(const FinishKP p) {
    const float alpha = p.alpha_dev ? __ldg(p.alpha_dev) : p.alpha;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int64_t i = tid; i < p.n * p.dcp; i += nth) {                  // <gz, code> and g_code = alpha gz
        const int64_t r = i / p.dcp;
        const int c = (int)(i % p.dcp);
        const float g = p.gz[i];
        acc = fmaf(g, p.code[i], acc);
        if (p.g_code && c < p.dc) p.g_code[r * p.dc + c] = alpha * g;
    }
    for (int64_t i = tid; i < (int64_t)p.nb * p.H; i += nth) {          // lin_z biases: alpha * colsum(S_i); alpha: <colsum, bz>
        const int b = (int)(i / p.H), h = (int)(i % p.H);
        const float s = p.sum_s[b][h];
        p.d_bz[b][h] = alpha * s;
        acc = fmaf(s, p.bz[b][h], acc);
    }
    for (int64_t i = tid; i < (int64_t)p.nb * p.H * p.dc; i += nth) {   // lin_z weights: alpha * S_i^T code
        const int64_t row = i / p.dc;
        const int c = (int)(i % p.dc);
        p.d_wz[row / p.H][(row % p.H) * p.dc + c] = alpha * p.dwz[row * p.dcp + c];
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(p.d_alpha, t);
    }
}

struct BwdPlan {
    int64_t n, G, GX, gn, pre, a32, code, gz, wz, dwz, total;   // offsets in floats
    int dcp;
};
static BwdPlan bwd_plan(const GnbDecoderWeights* w, int64_t n) {
    BwdPlan p;
    const int64_t H = w->d_hidden, nb = w->n_blocks;
    auto up = [](int64_t v) { return (v + 63) & ~(int64_t)63; };
    p.n = n;
    p.dcp = (w->d_code + 3) & ~3;
    int64_t o = 0;
    p.G = o, o += up(n * w->d_out);
    p.GX = o, o += up(n * (nb + 1) * H);
    p.gn = o, o += up(n * H);
    p.pre = o, o += up(n * H);
    p.a32 = o, o += up(n * H);
    p.code = o, o += up(n * p.dcp);
    p.gz = o, o += up(n * p.dcp);
    p.wz = o, o += up(nb * H * p.dcp);
    p.dwz = o, o += up(nb * H * p.dcp);
    p.total = o;
    return p;
}

}  // namespace gnb

extern "C" int64_t gnb_decode_train_bwd_workspace_bytes(const GnbDecoderWeights* w, int64_t n_rows) {
    if (!w || n_rows < 0) return -1;
    return bwd_plan(w, n_rows).total * (int64_t)sizeof(float);
}

extern "C" int gnb_decode_train_bwd(const GnbDecoderWeights* w, const float* code, const float* feat, const float* out, const float* tsdf,
                                    const void* activations, const float* g_out, const float* g_tsdf, int64_t n_rows,
                                    const GnbDecoderGrads* g, void* workspace, int64_t workspace_bytes, int tf32, void* stream) {
    GNB_CHECK_ARG(w && g, "gnb_decode_train_bwd: null weights / gradients");
    GNB_CHECK_ARG(w->use_code == 2, "gnb_decode_train_bwd: the weights must describe given codes (use_code = 2: the positional encoding "
                                    "stays with the caller, whose autograd carries d code / d xyz)");
    GNB_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31) / 8, "gnb_decode_train_bwd: n_rows %lld", (long long)n_rows);
    const int H = w->d_hidden, nb = w->n_blocks, dc = w->d_code, df = w->d_feat, dout = w->d_out, dgeo = w->d_geo;
    GNB_CHECK_ARG(H >= 4 && H % 4 == 0 && nb >= 0 && nb <= 8 && dc >= 1 && df >= 1 && dout >= 1 && dgeo >= 0 && dgeo <= dout,
                  "gnb_decode_train_bwd: bad dimensions");
    GNB_CHECK_ARG(w->tc_dtype == GNB_TC_FP16 || w->tc_dtype == GNB_TC_BF16, "gnb_decode_train_bwd: tc_dtype %d", w->tc_dtype);
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(code && feat && out && tsdf && activations && (g_out || g_tsdf) && workspace, "gnb_decode_train_bwd: null pointer");
    GNB_CHECK_ARG(g->lin_in_w && g->lin_in_b && g->lin_out_w && g->lin_out_b && g->alpha && (!g_tsdf || (g->head_w && g->head_b)),
                  "gnb_decode_train_bwd: null gradient buffer");
    for (int i = 0; i < nb; ++i)
        GNB_CHECK_ARG(g->lin_z_w[i] && g->lin_z_b[i] && g->fc0_w[i] && g->fc0_b[i] && g->fc1_w[i] && g->fc1_b[i],
                      "gnb_decode_train_bwd: null gradient buffer in block %d", i);
    const BwdPlan pl = bwd_plan(w, n_rows);
    GNB_CHECK_ARG(workspace_bytes >= pl.total * (int64_t)sizeof(float) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                  "gnb_decode_train_bwd: workspace of %lld bytes (256-byte aligned) needed", (long long)(pl.total * sizeof(float)));
    cudaStream_t st = (cudaStream_t)stream;
    cublasHandle_t h;
    int rc = blas_handle(&h);
    if (rc) return rc;
    if (g_blas.set_stream(h, st) != CUBLAS_STATUS_SUCCESS || g_blas.set_math(h, tf32 ? CUBLAS_TF32_TENSOR_OP_MATH : CUBLAS_DEFAULT_MATH) != CUBLAS_STATUS_SUCCESS) {
        set_error("gnb_decode_train_bwd: cublasSetStream / cublasSetMathMode failed");
        return GNB_E_UNSUPPORTED;
    }
    float* ws = (float*)workspace;
    float *G = ws + pl.G, *GX = ws + pl.GX, *gn = ws + pl.gn, *pre = ws + pl.pre, *a32 = ws + pl.a32;
    const int64_t n = n_rows, ldx = (int64_t)(nb + 1) * H;
    const size_t act_slab = (size_t)n * H * 2;                          // bytes of one activation slab
    const unsigned char* acts = (const unsigned char*)activations;
    auto link = [&](const float* pre_, const void* act, const float* res, int64_t ld_res, float* out_, int64_t ld_out, float* colsum) {
        return gnb_mlp_grad_link(pre_, H, act, H, w->tc_dtype, res, ld_res, out_, ld_out, a32, H, colsum, n, H, stream);
    };
    // outputs -> G, head and lin_out.bias gradients
    if ((rc = gnb_mlp_grad_head(g_out, g_tsdf, out, tsdf, w->head_w, n, dout, dgeo, G, g_tsdf ? g->head_w : nullptr,
                                g_tsdf ? g->head_b : nullptr, g->lin_out_b, stream)))
        return rc;
    // lin_out: S_nb = mask(relu input of lin_out) * (G W_out); its column sum is fc_1.bias of the last block (lin_in.bias if nb = 0)
    float* S = GX + (int64_t)nb * H;
    if ((rc = gemm_rm(h, false, false, n, H, dout, G, dout, w->lin_out_w, H, pre, H))) return rc;
    if ((rc = link(pre, acts + (size_t)(2 * nb) * act_slab, nullptr, 0, S, ldx, nb ? g->fc1_b[nb - 1] : g->lin_in_b))) return rc;
    if ((rc = gemm_rm(h, true, false, dout, H, n, G, dout, a32, H, g->lin_out_w, H))) return rc;
    for (int i = nb - 1; i >= 0; --i) {
        // fc_1: gn = mask(h_i) * (S W1), d W1 = S^T h_i
        if ((rc = gemm_rm(h, false, false, n, H, H, S, ldx, w->fc1_w[i], H, pre, H))) return rc;
        if ((rc = link(pre, acts + (size_t)(2 * i + 1) * act_slab, nullptr, 0, gn, H, g->fc0_b[i]))) return rc;
        if ((rc = gemm_rm(h, true, false, H, H, n, S, ldx, a32, H, g->fc1_w[i], H))) return rc;
        // fc_0: S_i = S_{i+1} + mask(a_i) * (gn W0), d W0 = gn^T a_i; column sum of S_i = fc_1.bias of block i-1 / lin_in.bias
        if ((rc = gemm_rm(h, false, false, n, H, H, gn, H, w->fc0_w[i], H, pre, H))) return rc;
        float* Sn = GX + (int64_t)i * H;
        if ((rc = link(pre, acts + (size_t)(2 * i) * act_slab, S, ldx, Sn, ldx, i ? g->fc1_b[i - 1] : g->lin_in_b))) return rc;
        if ((rc = gemm_rm(h, true, false, H, H, n, gn, H, a32, H, g->fc0_w[i], H))) return rc;
        S = Sn;
    }
    // lin_in
    if ((rc = gemm_rm(h, true, false, H, df, n, S, ldx, feat, df, g->lin_in_w, df))) return rc;
    if (g->g_feat && (rc = gemm_rm(h, false, false, n, df, H, S, ldx, w->lin_in_w, df, g->g_feat, df))) return rc;
    // lin_z of all blocks as two GEMMs over the concatenation [S_0 | ... | S_{nb-1}] (codes / weights zero-padded to dcp columns)
    if (nb) {
        const int dcp = pl.dcp;
        float *code_p = ws + pl.code, *gz = ws + pl.gz, *wz = ws + pl.wz, *dwz = ws + pl.dwz;
        if (dcp != dc) {
            GNB_CUDA(cudaMemsetAsync(code_p, 0, sizeof(float) * n * dcp, st));
            GNB_CUDA(cudaMemsetAsync(wz, 0, sizeof(float) * (size_t)nb * H * dcp, st));
        }
        GNB_CUDA(cudaMemcpy2DAsync(code_p, sizeof(float) * dcp, code, sizeof(float) * dc, sizeof(float) * dc, n, cudaMemcpyDeviceToDevice, st));
        for (int i = 0; i < nb; ++i)
            GNB_CUDA(cudaMemcpy2DAsync(wz + (size_t)i * H * dcp, sizeof(float) * dcp, w->lin_z_w[i], sizeof(float) * dc, sizeof(float) * dc, H,
                                       cudaMemcpyDeviceToDevice, st));
        if ((rc = gemm_rm(h, false, false, n, dcp, (int64_t)nb * H, GX, ldx, wz, dcp, gz, dcp))) return rc;
        if ((rc = gemm_rm(h, true, false, (int64_t)nb * H, dcp, n, GX, ldx, code_p, dcp, dwz, dcp))) return rc;
        FinishKP f = {};
        f.gz = gz, f.code = code_p, f.dwz = dwz, f.d_alpha = g->alpha, f.g_code = g->g_code, f.alpha_dev = w->alpha_dev, f.alpha = w->alpha;
        f.n = n, f.nb = nb, f.H = H, f.dc = dc, f.dcp = dcp;
        for (int i = 0; i < nb; ++i) {
            f.bz[i] = w->lin_z_b[i], f.sum_s[i] = i ? g->fc1_b[i - 1] : g->lin_in_b;
            f.d_wz[i] = g->lin_z_w[i], f.d_bz[i] = g->lin_z_b[i];
        }
        mlp_grad_finish_kernel<<<296, 256, 0, st>>>(f);
        GNB_LAUNCH_CHECK();
    } else if (g->g_code) {
        GNB_CUDA(cudaMemsetAsync(g->g_code, 0, sizeof(float) * n * dc, st));
    }
    return 0;
}

// fp32 CUDA-core decoder ("exact mode") for sm_100a.
//
// Replaces, fused into one kernel per tile of 32 query rows,
//   PositionalEncoding.forward   reference src/models/components/positional_encoding.py:28-40
//   ResnetFC.forward             reference src/models/components/resnetfc.py:134-189
//   ResnetBlockFC.forward        reference src/models/components/resnetfc.py:54-63
//   TSDFHeadSimple.forward       reference src/models/components/heads3d.py:36-50
// All arithmetic is fp32 FMA, so results agree with the CPU oracle to ~1e-6 relative; this is
// the parity anchor for the bf16 tcgen05 decoder (decoder_tc.cu), not the fast path.
//
// Mapping: one thread per hidden unit n (blockDim = d_hidden), 32 rows per CTA.  The residual
// stream x[32] of column n lives in registers for the whole network; the ReLU'd activations
// that feed the next layer are exchanged through one shared [32][d_hidden] buffer.  Weights
// are read straight from their nn.Linear layout (each thread streams its own row, L1-cached).
#include "common.cuh"

namespace gnb {

constexpr int DM = 32;   // rows per CTA

struct DecKP {
    GnbDecoderWeights w;
    const float* xyz;
    const float* feat;
    long long n_rows;
    float* out;
    float* tsdf;
};

// acc[m] += sum_k act[m*lda + k] * wrow[k]
template <bool VEC4>
__device__ __forceinline__ void row_dot(const float* __restrict__ wrow, const float* __restrict__ act, int lda, int K,
                                        float (&acc)[DM]) {
    if constexpr (VEC4) {
        for (int k = 0; k < K; k += 4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
            for (int m = 0; m < DM; ++m) {
                const float4 a = *reinterpret_cast<const float4*>(act + m * lda + k);
                acc[m] = fmaf(a.x, w4.x, acc[m]);
                acc[m] = fmaf(a.y, w4.y, acc[m]);
                acc[m] = fmaf(a.z, w4.z, acc[m]);
                acc[m] = fmaf(a.w, w4.w, acc[m]);
            }
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float wk = __ldg(wrow + k);
#pragma unroll
            for (int m = 0; m < DM; ++m) acc[m] = fmaf(act[m * lda + k], wk, acc[m]);
        }
    }
}

__device__ __forceinline__ void linear_col(const float* __restrict__ W, const float* __restrict__ bias, int n, int K,
                                           const float* __restrict__ act, int lda, float (&acc)[DM]) {
    const float b = bias ? __ldg(bias + n) : 0.0f;
#pragma unroll
    for (int m = 0; m < DM; ++m) acc[m] = b;
    const float* wrow = W + (long long)n * K;
    if ((K & 3) == 0 && (lda & 3) == 0 && ((reinterpret_cast<uintptr_t>(wrow) & 15) == 0))
        row_dot<true>(wrow, act, lda, K, acc);
    else
        row_dot<false>(wrow, act, lda, K, acc);
}

// code[m][:] for one row (trap T12): [x, sin(f0 x), sin(f0 x + pi/2), sin(f1 x), ...]
__device__ __forceinline__ void posenc_row(const GnbDecoderWeights& w, const float* __restrict__ xyz3, float* __restrict__ code) {
    if (!w.use_code) {
        code[0] = xyz3[0], code[1] = xyz3[1], code[2] = xyz3[2];
        return;
    }
    int o = 0;
    if (w.include_input) {
        code[0] = xyz3[0], code[1] = xyz3[1], code[2] = xyz3[2];
        o = 3;
    }
    const float half_pi = (float)(3.14159265358979323846 * 0.5);
    for (int f = 0; f < 2 * w.num_freqs; ++f) {
        const float freq = w.freq_factor * exp2f((float)(f >> 1));           // freq_factor * 2^k (exact scaling)
        const float phase = (f & 1) ? half_pi : 0.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) code[o + f * 3 + d] = sinf(__fadd_rn(phase, __fmul_rn(xyz3[d], freq)));
    }
}

__global__ void __launch_bounds__(512) decode_fp32_kernel(const __grid_constant__ DecKP p) {
    extern __shared__ __align__(16) float smem[];
    const GnbDecoderWeights& w = p.w;
    const int Hd = w.d_hidden;
    const int lda_f = (w.d_feat + 3) & ~3, lda_c = (w.d_code + 3) & ~3, lda_o = (w.d_out + 3) & ~3;
    float* act = smem;                         // [DM][Hd]
    float* featb = act + DM * Hd;              // [DM][lda_f]
    float* codeb = featb + DM * lda_f;         // [DM][lda_c]
    float* outb = codeb + DM * lda_c;          // [DM][lda_o]
    const int n = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * DM;

    // ---- stage inputs -----------------------------------------------------------------
    for (int i = threadIdx.x; i < DM * lda_f; i += blockDim.x) {
        int m = i / lda_f, k = i % lda_f;
        long long r = row0 + m;
        featb[i] = (r < p.n_rows && k < w.d_feat) ? __ldg(p.feat + r * w.d_feat + k) : 0.0f;
    }
    for (int m = threadIdx.x; m < DM; m += blockDim.x) {
        long long r = row0 + m;
        float xyz3[3] = {0.f, 0.f, 0.f};
        if (r < p.n_rows && w.use_code != 2) { xyz3[0] = p.xyz[r * 3], xyz3[1] = p.xyz[r * 3 + 1], xyz3[2] = p.xyz[r * 3 + 2]; }
        for (int k = w.d_code; k < lda_c; ++k) codeb[m * lda_c + k] = 0.0f;
        if (w.use_code == 2) {
            for (int k = 0; k < w.d_code; ++k) codeb[m * lda_c + k] = (r < p.n_rows) ? p.xyz[r * w.d_code + k] : 0.0f;
            continue;
        }
        posenc_row(w, xyz3, codeb + m * lda_c);
    }
    __syncthreads();

    float x[DM], t[DM];
    const float alpha = w.alpha_dev ? __ldg(w.alpha_dev) : w.alpha;
    linear_col(w.lin_in_w, w.lin_in_b, n, w.d_feat, featb, lda_f, x);          // resnetfc.py:149
    for (int blk = 0; blk < w.n_blocks; ++blk) {
        linear_col(w.lin_z_w[blk], w.lin_z_b[blk], n, w.d_code, codeb, lda_c, t);   // resnetfc.py:175
#pragma unroll
        for (int m = 0; m < DM; ++m) x[m] = __fadd_rn(x[m], __fmul_rn(alpha, t[m]));   // x + alpha * tz (:180)
#pragma unroll
        for (int m = 0; m < DM; ++m) act[m * Hd + n] = fmaxf(x[m], 0.0f);
        __syncthreads();
        linear_col(w.fc0_w[blk], w.fc0_b[blk], n, Hd, act, Hd, t);                   // net = fc_0(relu(x))  (:56)
        __syncthreads();
#pragma unroll
        for (int m = 0; m < DM; ++m) act[m * Hd + n] = fmaxf(t[m], 0.0f);
        __syncthreads();
        linear_col(w.fc1_w[blk], w.fc1_b[blk], n, Hd, act, Hd, t);                   // dx = fc_1(relu(net)) (:57)
#pragma unroll
        for (int m = 0; m < DM; ++m) x[m] = __fadd_rn(x[m], t[m]);                   // x_s + dx (:63)
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < DM; ++m) act[m * Hd + n] = fmaxf(x[m], 0.0f);
    __syncthreads();
    for (int no = n; no < w.d_out; no += blockDim.x) {                               // lin_out(relu(x)) (:185)
        linear_col(w.lin_out_w, w.lin_out_b, no, Hd, act, Hd, t);
#pragma unroll
        for (int m = 0; m < DM; ++m) {
            outb[m * lda_o + no] = t[m];
            if (row0 + m < p.n_rows && p.out) p.out[(row0 + m) * w.d_out + no] = t[m];
        }
    }
    __syncthreads();
    if (p.tsdf) {
        for (int m = threadIdx.x; m < DM; m += blockDim.x) {                         // tanh(fc(feat_geo)) heads3d.py:44-45
            if (row0 + m >= p.n_rows) continue;
            float a = __ldg(w.head_b);
            for (int j = 0; j < w.d_geo; ++j) a = fmaf(outb[m * lda_o + j], __ldg(w.head_w + j), a);
            p.tsdf[row0 + m] = tanhf(a);
        }
    }
}

__global__ void posenc_kernel(const float* __restrict__ x, long long n, GnbDecoderWeights w, float* __restrict__ out) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float xyz3[3] = {x[r * 3], x[r * 3 + 1], x[r * 3 + 2]};
    float code[3 + 6 * 42];
    posenc_row(w, xyz3, code);
    for (int k = 0; k < w.d_code; ++k) out[r * w.d_code + k] = code[k];
}

__global__ void tsdf_head_kernel(const float* __restrict__ g, long long n, int d_geo, long long stride,
                                 const float* __restrict__ hw, const float* __restrict__ hb, float* __restrict__ tsdf) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float a = __ldg(hb);
    for (int j = 0; j < d_geo; ++j) a = fmaf(g[r * stride + j], __ldg(hw + j), a);
    tsdf[r] = tanhf(a);
}

int check_decoder_weights(const GnbDecoderWeights* w, const char* who) {
    GNB_CHECK_ARG(w, "%s: null weights", who);
    GNB_CHECK_ARG(w->d_feat >= 1 && w->d_hidden >= 1 && w->d_out >= 1 && w->d_geo >= 1 && w->d_geo <= w->d_out,
                  "%s: bad dimensions", who);
    GNB_CHECK_ARG(w->n_blocks >= 0 && w->n_blocks <= 8, "%s: n_blocks %d not in [0,8]", who, w->n_blocks);
    int d_code = w->use_code == 1 ? (w->include_input ? 3 : 0) + 6 * w->num_freqs : (w->use_code == 2 ? w->d_code : 3);
    GNB_CHECK_ARG(w->d_code == d_code && d_code >= 1 && d_code <= 256, "%s: d_code %d does not match the encoding (%d)", who,
                  w->d_code, d_code);
    GNB_CHECK_ARG(w->lin_in_w && w->lin_in_b && w->lin_out_w && w->lin_out_b && w->head_w && w->head_b, "%s: null parameter", who);
    for (int i = 0; i < w->n_blocks; ++i)
        GNB_CHECK_ARG(w->lin_z_w[i] && w->lin_z_b[i] && w->fc0_w[i] && w->fc0_b[i] && w->fc1_w[i] && w->fc1_b[i],
                      "%s: null parameter in block %d", who, i);
    return 0;
}

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_positional_encoding(const float* x, int64_t n_rows, int num_freqs, float freq_factor, int include_input,
                                       float* out, void* stream) {
    GNB_CHECK_ARG(x && out && n_rows >= 0, "gnb_positional_encoding: bad arguments");
    GNB_CHECK_ARG(num_freqs >= 0 && num_freqs <= 42 && (include_input || num_freqs > 0), "gnb_positional_encoding: num_freqs %d",
                  num_freqs);
    if (n_rows == 0) return 0;
    GnbDecoderWeights w = {};
    w.use_code = 1, w.num_freqs = num_freqs, w.freq_factor = freq_factor, w.include_input = include_input;
    w.d_code = (include_input ? 3 : 0) + 6 * num_freqs;
    posenc_kernel<<<ceil_div(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(x, n_rows, w, out);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_tsdf_head(const float* feat_geo, int64_t n_rows, int d_geo, int64_t row_stride, const float* head_w,
                             const float* head_b, float* tsdf, void* stream) {
    GNB_CHECK_ARG(feat_geo && head_w && head_b && tsdf && n_rows >= 0 && d_geo >= 1 && row_stride >= d_geo,
                  "gnb_tsdf_head: bad arguments");
    if (n_rows == 0) return 0;
    tsdf_head_kernel<<<ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(feat_geo, n_rows, d_geo, row_stride, head_w, head_b, tsdf);
    GNB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnb_decode_fp32(const GnbDecoderWeights* w, const float* xyz, const float* feat, int64_t n_rows, float* out,
                               float* tsdf, void* stream) {
    int rc = check_decoder_weights(w, "gnb_decode_fp32");
    if (rc) return rc;
    GNB_CHECK_ARG(n_rows >= 0, "gnb_decode_fp32: bad arguments");
    if (n_rows == 0) return 0;
    GNB_CHECK_ARG(xyz && feat && (out || tsdf), "gnb_decode_fp32: bad arguments");
    if (w->d_hidden % 32 != 0 || w->d_hidden > 512) {
        set_error("gnb_decode_fp32: d_hidden %d must be a multiple of 32 and <= 512", w->d_hidden);
        return GNB_E_UNSUPPORTED;
    }
    if (n_rows == 0) return 0;
    DecKP kp;
    kp.w = *w, kp.xyz = xyz, kp.feat = feat, kp.n_rows = n_rows, kp.out = out, kp.tsdf = tsdf;
    const int lda_f = (w->d_feat + 3) & ~3, lda_c = (w->d_code + 3) & ~3, lda_o = (w->d_out + 3) & ~3;
    size_t smem = sizeof(float) * DM * (size_t)(w->d_hidden + lda_f + lda_c + lda_o);
    if (smem > 227 * 1024) {
        set_error("gnb_decode_fp32: tile needs %zu bytes of shared memory (> 227 KB)", smem);
        return GNB_E_UNSUPPORTED;
    }
    GNB_CUDA(cudaFuncSetAttribute(decode_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_fp32_kernel<<<ceil_div(n_rows, DM), w->d_hidden, smem, (cudaStream_t)stream>>>(kp);
    GNB_LAUNCH_CHECK();
    return 0;
}

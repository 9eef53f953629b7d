// Stand-alone point-query sampler: xyz -> [plane features | volume features]   (sm_100a)
// Replaces GenNerf.map_features (reference src/models/model.py:163-204).  HBM/L2-bound gather:
// G lanes share a query, each lane one float4 of channels, so every corner fetch of a
// channels-last volume/plane is one contiguous C*4-byte run.
#include "sample.cuh"

namespace gnb {

template <int VEC>
__global__ void __launch_bounds__(256) sample_kernel(const __grid_constant__ SampleKP p, int G) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = tid / G;
    const int sub = (int)(tid % G);
    if (q >= p.total) return;
    const int b = (int)(q / p.Q);
    const float x = __ldg(p.xyz + q * 3 + 0), y = __ldg(p.xyz + q * 3 + 1), z = __ldg(p.xyz + q * 3 + 2);
    float* __restrict__ out = p.out + q * p.out_stride;
    int c_off = 0;
    if (p.Cp > 0) {
        BiCorners bc[3];
        planes_setup(p, x, y, z, bc);
        for (int c = sub * VEC; c < p.Cp; c += G * VEC) {
            Vals<VEC> r = sample_planes<VEC>(p, bc, b, c);
            if constexpr (VEC == 4) *reinterpret_cast<float4*>(out + c) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
            else out[c] = r.v[0];
        }
        c_off = p.Cp;
    }
    if (p.volume) {
        TriCorners tc;
        trilinear_setup(p, x, y, z, tc);
        for (int c = sub * VEC; c < p.C; c += G * VEC) {
            Vals<VEC> r = sample_volume<VEC>(p, tc, b, c);
            if constexpr (VEC == 4) *reinterpret_cast<float4*>(out + c_off + c) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
            else out[c_off + c] = r.v[0];
        }
    }
}

int fill_sample_kp(const GnbSampleParams* s, SampleKP& kp) {
    GNB_CHECK_ARG(s, "sample: null params");
    GNB_CHECK_ARG(s->batch >= 1 && s->n_query >= 0 && s->xyz, "sample: bad batch / n_query / xyz");
    bool has_planes = s->plane[0] || s->plane[1] || s->plane[2];
    GNB_CHECK_ARG(s->volume || has_planes, "sample: neither a volume nor planes given");
    kp.xyz = s->xyz;
    kp.Q = s->n_query;
    kp.total = (long long)s->batch * s->n_query;
    kp.volume = s->volume;
    kp.nx = kp.ny = kp.nz = 1;
    kp.C = 0;
    if (s->volume) {
        GNB_CHECK_ARG(s->nx > 0 && s->ny > 0 && s->nz > 0 && s->C > 0, "sample: bad volume shape");
        kp.nx = s->nx, kp.ny = s->ny, kp.nz = s->nz, kp.C = s->C;
        kp.vsb = s->vol_stride_b, kp.vsx = s->vol_stride_x, kp.vsy = s->vol_stride_y, kp.vsz = s->vol_stride_z;
        kp.vsc = s->vol_stride_c;
        // torch.tensor([nx,ny,nz]) * voxel_size: int64 tensor x python float -> fp32 product
        kp.ext_x = (float)s->nx * s->voxel_size;
        kp.ext_y = (float)s->ny * s->voxel_size;
        kp.ext_z = (float)s->nz * s->voxel_size;
        kp.ox = s->origin[0], kp.oy = s->origin[1], kp.oz = s->origin[2];
    }
    kp.R = 1, kp.Cp = 0;
    for (int k = 0; k < 3; ++k) kp.plane[k] = s->plane[k];
    if (has_planes) {
        GNB_CHECK_ARG(s->R > 0 && s->Cp > 0, "sample: bad plane shape");
        kp.R = s->R, kp.Cp = s->Cp;
        kp.psb = s->pl_stride_b, kp.psh = s->pl_stride_h, kp.psw = s->pl_stride_w, kp.psc = s->pl_stride_c;
        kp.den = (float)(1.0 + s->padding + 10e-6);
    }
    kp.out = s->out;
    kp.out_stride = s->out_stride;
    return 0;
}

}  // namespace gnb

using namespace gnb;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int gnb_sample_features(const GnbSampleParams* s, void* stream) {
    SampleKP kp;
    int rc = fill_sample_kp(s, kp);
    if (rc) return rc;
    GNB_CHECK_ARG(s->out && s->out_stride >= kp.C + kp.Cp, "sample: bad output");
    if (kp.total == 0) return 0;
    // float4 path: unit channel stride, channel counts and every base/stride a multiple of 4
    bool vec = (s->out_stride % 4 == 0) && aligned16(s->out);
    if (kp.volume)
        vec = vec && kp.vsc == 1 && kp.C % 4 == 0 && aligned16(kp.volume) && kp.vsb % 4 == 0 && kp.vsx % 4 == 0 &&
              kp.vsy % 4 == 0 && kp.vsz % 4 == 0;
    if (kp.Cp > 0) {
        vec = vec && kp.psc == 1 && kp.Cp % 4 == 0 && kp.psb % 4 == 0 && kp.psh % 4 == 0 && kp.psw % 4 == 0;
        for (int k = 0; k < 3; ++k) vec = vec && aligned16(kp.plane[k]);
    }
    int cmax = kp.C > kp.Cp ? kp.C : kp.Cp;
    int lanes = vec ? cmax / 4 : cmax;
    int G = 1;
    while (G < lanes && G < 32) G <<= 1;
    long long threads = kp.total * G;
    unsigned blocks = (unsigned)((threads + 255) / 256);
    if (vec)
        sample_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(kp, G);
    else
        sample_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(kp, G);
    GNB_LAUNCH_CHECK();
    return 0;
}

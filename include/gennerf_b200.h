/*
 * gennerf_b200 -- C ABI of the B200-native lift-and-query path of gen-nerf.
 *
 * The reference (mrchris7/gen-nerf) is pure Python and has no FFI layer; its boundary for
 * this path is a set of Python call sites.  Each entry point below
 * names the reference function it replaces (file:line relative to the reference root).
 * gennerf_b200/ops.py binds these through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only
 *     enqueue work on it: they never synchronise, allocate or free (the caller owns all
 *     buffers, including scratch) and keep no mutable global state, so they are re-entrant
 *     across streams and threads;
 *   - return value 0 = success, otherwise a negative GNB_E_* code or a positive
 *     cudaError_t; gnb_last_error() gives a thread-local message;
 *   - sizes are element counts, strides are in elements, fp32 unless stated.
 */
#ifndef GENNERF_B200_H
#define GENNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNB_VERSION 100
#define GNB_MAX_FRAMES 64          /* frames per gnb_backproject_frames call               */

#define GNB_E_INVALID (-1)         /* bad argument                                          */
#define GNB_E_UNSUPPORTED (-2)     /* shape not supported by the sm_100a kernels            */
#define GNB_E_ARCH (-3)            /* device is not sm_100                                  */

/* feature-map layouts */
#define GNB_LAYOUT_NCHW 0          /* reference layout (B,C,H,W) contiguous                 */
#define GNB_LAYOUT_NHWC 1          /* channels-last: (B,H,W,C) contiguous                   */

int gnb_version(void);
const char* gnb_last_error(void);
/* sizeof() of the parameter structs as compiled: 0 GnbLiftParams, 1 GnbSampleParams,
 * 2 GnbDecoderWeights, 3 GnbFusionParams.  Lets a foreign-language binding verify its struct layout. */
int gnb_struct_size(int which);
/* Tuning / debugging switches (GNB_TC_TWO_CTA, GNB_TC_NO_EARLY, GNB_DEBUG_MAX_CLUSTERS, GNB_DEBUG_PRINT, GNB_LIFT_NVW,
 * GNB_SCATTER_SCALAR, GNB_FPS_SINGLE_CTA, GNB_FPS_CLUSTER, GNB_SAMPLE_GENERIC, GNB_BIN_UNIT, GNB_BIN_ROWCOPY,
 * GNB_SCATTER_TILED, GNB_BIN_PRESORTED, GNB_TC_NO_STG, GNB_TC_PAIR, GNB_DEBUG_NO_WCOPY, GNB_FPS_GRID,
 * GNB_QUERY_FUSED [read by the Python host: always the single fused query kernel]).  Every option takes its default from the environment variable of the same name,
 * read ONCE per process (never on a hot-path call); these two calls read / change it afterwards.  None of them changes
 * results beyond floating-point summation order. */
int gnb_set_option(const char* name, int value);
int gnb_get_option(const char* name, int* value);

/* -------------------------------------------------------------------------------------
 * Layout helper: (T frames) NCHW -> NHWC, one launch.  src[t] is (B,C,H,W); dst is
 * (T,B,H,W,C).  Not in the reference: it is the price of the reference's NCHW contract and
 * is skipped when the CNN already runs channels-last.
 * ----------------------------------------------------------------------------------- */
int gnb_nchw_to_nhwc(const float* const* h_src, int n_frames, float* dst,
                     int B, int C, int H, int W, void* stream);

/* -------------------------------------------------------------------------------------
 * Fused lift.  Replaces, for all T frames of one scene batch at once,
 *   coordinates()                   src/data/tsdf.py:25-40        (never materialised)
 *   backproject()                   src/models/utils.py:948-996
 *   the accumulation in encode()    src/models/model.py:121-127, voxel_net.py:120-126
 *   the normalisation               src/models/model.py:195-199, voxel_net.py:163-168
 * volume[b,v,:] = sum over frames t (in frame order) of features_t[b,:,py,px] where frame t
 * sees voxel v;  count[b,v] = number of such frames;  valid[b,v] = count > 0.
 * With mean != 0 the sum is divided by count (the north-star's masked mean; the reference
 * computes the SUM, SURVEY trap T2).
 * ----------------------------------------------------------------------------------- */
typedef struct GnbLiftParams {
    int32_t nx, ny, nz;            /* voxel grid; linear id v = (x*ny + y)*nz + z           */
    float voxel_size;
    float origin[3];
    int32_t batch;                 /* B scenes                                              */
    int32_t n_frames;              /* T <= GNB_MAX_FRAMES                                   */
    int32_t C, H, W;               /* feature maps                                          */
    int32_t feat_layout;           /* GNB_LAYOUT_NHWC: features[t] is (B,H,W,C)             */
                                   /* GNB_LAYOUT_NCHW: features[t] is (B,C,H,W); `scratch`  */
                                   /*   must hold T*B*H*W*C floats                          */
    const float* features[GNB_MAX_FRAMES];
    const float* h_projection;     /* HOST (B,T,3,4) world->pixel, row-major                */
    float* scratch;
    /* outputs */
    float* volume;                 /* element (b,v,c) at b*vol_stride_b + v*vol_stride_v +  */
    int64_t vol_stride_b;          /*   c*vol_stride_c.  channels-last: (V*C, C, 1);        */
    int64_t vol_stride_v;          /*   reference (B,C,nx,ny,nz): (C*V, 1, V)               */
    int64_t vol_stride_c;
    int32_t* count;                /* (B,V) or NULL                                         */
    uint8_t* valid;                /* (B,V) 0/1 or NULL                                     */
    int32_t accumulate;            /* != 0: volume/count already hold earlier frames        */
    int32_t mean;                  /* != 0: divide by count at the end                      */
    int32_t x_begin, x_end;        /* only voxels with x_begin <= x < x_end are computed and  */
                                   /* written (slab sharding across GPUs); 0,0 = whole grid  */
} GnbLiftParams;

int gnb_backproject_frames(const GnbLiftParams* p, void* stream);

/* Backward of gnb_backproject_frames (autograd of utils.py:991, an index_put_ with accumulate):
 * grad_features_t[b,:,py,px] += grad_volume[b,v,:] for every frame t that sees voxel v.
 * `p` as in the forward call (volume strides now describe grad_volume; `count` is read in mean
 * mode); h_grad_features: HOST array of n_frames device pointers in p->feat_layout, fully
 * overwritten.  Atomic, order-nondeterministic sums like the reference's CUDA index_put_. */
int gnb_backproject_frames_bwd(const GnbLiftParams* p, const float* grad_volume,
                               float* const* h_grad_features, void* stream);

/* Parity probe for the integer part of backproject (utils.py:979-985): pixel indices of
 * every voxel for ONE projection.  px,py are the int64 values the reference computes where
 * they are finite and fit int32, else INT32_MIN.  valid as in the reference. */
int gnb_project_indices(int nx, int ny, int nz, float voxel_size, const float* h_origin3,
                        const float* h_projection12, int H, int W,
                        int32_t* px, int32_t* py, uint8_t* valid, void* stream);

/* -------------------------------------------------------------------------------------
 * Point-query sampler.  Replaces
 *   trilinear_interpolation()            src/models/utils.py:999-1042  (F.grid_sample 3-D)
 *   GenNerf.sample_plane_feature() x3    src/models/model.py:153-161   (F.grid_sample 2-D)
 *   normalize_coordinate()               src/models/utils.py:75-98
 *   GenNerf.map_features()               src/models/model.py:163-204   (concat, planes first)
 * out[b,q,:] = [ sum over planes of bilinear(plane, q) (C_p) | trilinear(volume, q) (C) ].
 * Either part may be absent (pointer NULL).
 * ----------------------------------------------------------------------------------- */
typedef struct GnbSampleParams {
    int32_t batch;
    int64_t n_query;               /* Q per scene                                           */
    const float* xyz;              /* (B,Q,3)                                               */
    /* volume part */
    const float* volume;           /* NULL = no volume part                                 */
    int32_t nx, ny, nz, C;
    int64_t vol_stride_b, vol_stride_x, vol_stride_y, vol_stride_z, vol_stride_c;
    float voxel_size;
    float origin[3];
    /* plane part: planes in the reference order xz, xy, yz; NULL entries are skipped */
    const float* plane[3];
    int32_t R, Cp;
    int64_t pl_stride_b, pl_stride_h, pl_stride_w, pl_stride_c;  /* (B,C_p,H=R,W=R) logical */
    double padding;                /* python float, e.g. 0.1 (kept double: 1+padding+1e-5 is rounded once) */
    /* output */
    float* out;                    /* (B,Q,out_stride), plane part first                    */
    int64_t out_stride;            /* >= Cp + C                                             */
    /* optional second output (may replace `out`): the features as the tcgen05 decoder's 16-bit lin_in OPERAND IMAGE, which
     * gnb_decode_image_tc copies into shared memory with one bulk copy per tile instead of converting fp32 rows.
     * Tile t = q / 128 of the flat query index q is image_kchunks blocks of 16 KB; block k holds columns [64k, 64k+64) of
     * the tile's 128 rows, 128 bytes per row, 128B-swizzled in 8-row atoms:
     *   byte offset of the 8 columns [8u, 8u+8) of row r  =  (r/8)*1024 + (r%8)*128 + ((u ^ (r%8)) * 16).
     * The caller zero-fills the image first when Cp + C < 64*image_kchunks or B*Q is not a multiple of 128 (operand columns /
     * rows the sampler does not write).  Needs the float4 layout (unit channel strides, Cp % 4 == 0, C % 4 == 0). */
    void* image;                   /* NULL = none; ceil(B*Q/128) * image_kchunks * 16384 bytes */
    int32_t image_kchunks;         /* 64*image_kchunks >= Cp + C                            */
    int32_t image_dtype;           /* GNB_TC_FP16 / GNB_TC_BF16                             */
    int32_t* image_status;         /* optional: bit 0 is set when a feature saturated fp16   */
} GnbSampleParams;

int gnb_sample_features(const GnbSampleParams* p, void* stream);

/* Brick-binned variant of gnb_sample_features for MANY random queries per voxel (dense extraction of
 * trilinear_interpolation() / GenNerf.map_features over ~V or more points): the queries are counting-sorted by the
 * brick of voxels holding their base cell, each brick's corner voxels are staged in shared memory with bulk copies and
 * gathered from there.  Same results, bit for bit, as gnb_sample_features.  Needs a channels-last fp32 volume
 * (vol_stride_c == 1, vol_stride_z == C, C % 4 == 0) and caller-provided scratch:
 * gnb_sample_binned_scratch_bytes() returns the size for these parameters (out may still be NULL), or 0 when the
 * binned path does not apply to them -- call gnb_sample_features then. */
int64_t gnb_sample_binned_scratch_bytes(const GnbSampleParams* p);
int gnb_sample_features_binned(const GnbSampleParams* p, void* scratch, int64_t scratch_bytes, void* stream);

/* Backward of gnb_sample_features (ATen grid_sampler_{3d,2d}_backward chained through the
 * reference's coordinate normalisations).  grad_out: (B,Q,grad_out_stride) = [planes | volume].
 * grad_volume / grad_planes (HOST array of 3 device pointers) use the forward strides and must be
 * zero-initialised by the caller (they are accumulated into); grad_xyz (B,Q,3) is overwritten.
 * Any of the three may be null. */
int gnb_sample_features_bwd(const GnbSampleParams* s, const float* grad_out, int64_t grad_out_stride,
                            float* grad_volume, float* const* h_grad_planes3, float* grad_xyz,
                            void* stream);

/* Double backward of the sampler: the backward of gnb_sample_features_bwd's grad_xyz output, i.e. what
 * torch.autograd.grad(tsdf, xyz, create_graph=True) (calculate_grad, src/models/utils.py:636-649) followed by
 * loss.backward() needs for the eikonal / gradient losses (src/models/model.py:385-400).  The reference gets it for the
 * planes from its pure-PyTorch grid_sample_2d (src/models/utils.py:1117-1174, chosen in model.py:157-158); ATen has no
 * double backward for the 3-D grid sampler, so for the volume this goes beyond the reference.
 * gg_xyz (B,Q,3) = gradient w.r.t. grad_xyz.  Outputs (any may be null): g_grad_out (B,Q,g_grad_out_stride)
 * overwritten; g_volume / g_planes (forward strides) accumulated into; g_xyz (B,Q,3) overwritten (the second-order
 * cross terms of the multilinear weights). */
int gnb_sample_features_bwd2(const GnbSampleParams* s, const float* grad_out, int64_t grad_out_stride,
                             const float* gg_xyz, float* g_grad_out, int64_t g_grad_out_stride,
                             float* g_volume, float* const* h_g_planes3, float* g_xyz, void* stream);

/* -------------------------------------------------------------------------------------
 * Plane coordinates and cell indices.  Replaces normalize_coordinate() + coordinate2index()
 * src/models/utils.py:57-98 for the three planes at once.
 * coord: (3,B,N,2) fp32 in [0, 1-1e-5];  index: (3,B,N) int64 = x0 + R*x1.
 * ----------------------------------------------------------------------------------- */
int gnb_plane_coords(const float* p, int64_t n_points_total, double padding, int R,
                     float* coord, int64_t* index, void* stream);

/* -------------------------------------------------------------------------------------
 * Triplane scatter-mean.  Replaces LocalPoolPointnet.generate_plane_features() x3
 * (src/models/components/pointnet.py:72-89) = torch_scatter.scatter_mean onto R*R cells.
 * planes: (3,B,R,R,C_p) channels-last (logical (B,C_p,R,R) per plane);
 * count: (3,B,R*R) int32 (exact in both modes).
 * mode 0: warp-aggregated atomics (sums order-nondeterministic);
 * mode 1: deterministic -- every cell sums its points in ascending point index, which is
 *         bit-identical to the CPU scatter_add_ the oracle uses.  scratch >= result of
 *         gnb_scatter_scratch_bytes().
 * ----------------------------------------------------------------------------------- */
#define GNB_SCATTER_ATOMIC 0
#define GNB_SCATTER_DETERMINISTIC 1
#define GNB_SCATTER_ATOMIC_SUM 2   /* like ATOMIC but leaves SUMS in `planes` (no division): the  */
                                   /* partial result a rank all-reduces before gnb_scatter_finalize */
int64_t gnb_scatter_scratch_bytes(int B, int64_t N, int R, int mode);
int gnb_scatter_mean_planes(const float* p, const float* c, int B, int64_t N, int Cp, int R,
                            double padding, int mode, float* planes, int32_t* count,
                            void* scratch, int64_t scratch_bytes, void* stream);

/* planes[cell,:] /= max(count[cell],1) over n_cells = 3*B*R*R cells (after an all-reduce of
 * GNB_SCATTER_ATOMIC_SUM partial sums and counts). */
int gnb_scatter_finalize(float* planes, const int32_t* count, int64_t n_cells, int Cp, void* stream);

/* -------------------------------------------------------------------------------------
 * Local pooling.  Replaces LocalPoolPointnet.pool_local()
 * (src/models/components/pointnet.py:105-121): for each plane scatter-(max|mean) the point
 * features into cells, gather the cell value back to every point, sum over the 3 planes.
 * c,out: (B,N,Hd);  scratch >= gnb_pool_scratch_bytes().
 * ----------------------------------------------------------------------------------- */
#define GNB_POOL_MAX 0
#define GNB_POOL_MEAN 1
int64_t gnb_pool_scratch_bytes(int B, int64_t N, int Hd, int R);
int gnb_pool_local(const float* p, const float* c, int B, int64_t N, int Hd, int R,
                   double padding, int pool_type, float* out,
                   void* scratch, int64_t scratch_bytes, void* stream);

/* Backward passes of the triplane projection.
 * scatter-mean: grad_c[b,n,:] = sum over planes of grad_planes[cell(n),:] / max(count,1);
 *   grad_planes (3,B,R,R,C_p) channels-last, count from the forward.
 * pool_local: grad_c from grad_out (B,N,Hd); `fwd_scratch` is the scratch buffer the forward
 *   gnb_pool_local call filled (it holds the cell maxima / sums and counts) and `c` its input. */
int gnb_scatter_mean_planes_bwd(const float* p, const float* grad_planes, const int32_t* count, int B,
                                int64_t N, int Cp, int R, double padding, float* grad_c, void* stream);
int64_t gnb_pool_bwd_scratch_bytes(int B, int64_t N, int Hd, int R);
int gnb_pool_local_bwd(const float* p, const float* c, const float* grad_out, int B, int64_t N, int Hd,
                       int R, double padding, int pool_type, const void* fwd_scratch, float* grad_c,
                       void* scratch, int64_t scratch_bytes, void* stream);

/* -------------------------------------------------------------------------------------
 * Front end of the triplane branch (SURVEY 8f "next" row 1).
 * gnb_get_3d_points replaces get_3d_points() (src/models/utils.py:120-175): depth (B,H,W) and HOST
 *   projections (B,3,4) -> world points (B,H,W,3).
 * gnb_farthest_point_sample replaces farthest_point_sample() (src/models/utils.py:178-202): xyz (B,N,3),
 *   start (B) int64 = the first index (the reference draws it with torch.randint) -> out_idx (B,npoint)
 *   int64, out_xyz (B,npoint,3); scratch >= B*N floats.  Bit-identical to the reference's iteration.
 * ----------------------------------------------------------------------------------- */
int gnb_get_3d_points(const float* depth, const float* h_projection, int B, int H, int W, float* out,
                      void* stream);
int gnb_farthest_point_sample(const float* xyz, int B, int64_t N, int npoint, const int64_t* start,
                              float* scratch, int64_t* out_idx, float* out_xyz, void* stream);

/* -------------------------------------------------------------------------------------
 * Training-time ray sampler (SURVEY 8f "next" row 4).  Replaces sample_points_on_rays()
 * (src/models/utils.py:458-540): per camera b and sampled pixel s, 1+N+M depths along the ray
 * [depth | linspace(min_dist, depth+delta, N) | the M gaussian depths the caller drew with
 * normal(depth, sigma)], unprojected with intrinsics (B,3,3) and transformed by poses (B,4,4).
 * h_idxs, w_idxs int64 (B,S); depths (B,S); gaussian_depths (B,S,M);
 * xyz_world (B,S,1+N+M,3); z (B,S,1+N+M).  Everything on the device.
 * ----------------------------------------------------------------------------------- */
int gnb_sample_points_on_rays(const int64_t* h_idxs, const int64_t* w_idxs, const float* depths,
                              const float* intrinsics, const float* poses, const float* gaussian_depths,
                              int B, int S, int N, int M, float delta, float min_dist,
                              float* xyz_world, float* z, void* stream);

/* sample_valid_depth_pixels (src/models/utils.py:340-363) without materialising argwhere(depth != 0): gnb_valid_pixel_count
 * writes, per depth map (B,H,W), the exclusive prefix of the valid-pixel count of every image row (row_prefix (B,H)) and
 * the total n_valid (B); gnb_valid_pixel_select maps S ranks per map (rank (B,S) int64, each in [0, n_valid[b]) -- the
 * caller draws them exactly like the reference: torch.randperm(n_valid[b])[:S]) to the (h, w) of the rank-th valid pixel
 * in row-major order, i.e. argwhere(depth[b] != 0)[rank].  h_idxs / w_idxs (B,S) int64; -1 for a rank out of range. */
int gnb_valid_pixel_count(const float* depth, int B, int H, int W, int32_t* row_prefix, int32_t* n_valid, void* stream);
int gnb_valid_pixel_select(const float* depth, int B, int H, int W, const int32_t* row_prefix, const int32_t* n_valid,
                           const int64_t* rank, int S, int64_t* h_idxs, int64_t* w_idxs, void* stream);

/* -------------------------------------------------------------------------------------
 * TSDF fusion of posed depth maps (SURVEY 8f "next" row 3): GT generation and evaluation re-fusion.
 * gnb_tsdf_fusion_integrate replaces TSDFFusion.integrate() (src/data/tsdf.py:369-418) for n_frames frames
 *   at once; frames are applied in order, so the volumes equal n_frames sequential integrate() calls bit
 *   for bit.  The running volumes (the reference's tsdf_vol, weight_vol, color_vol, label_vol; flat voxel id
 *   v = (x*ny + y)*nz + z) are read and updated in place; initialise them as TSDFFusion.reset() does
 *   (tsdf 1, weight 0, colour 0, label -1).  trunc_margin = voxel_size * trunc_ratio (tsdf.py:341).
 * gnb_tsdf_fusion_finalize replaces the normalisation of TSDFFusion.get_tsdf() (tsdf.py:426-434):
 *   out = vol / weight where weight > 0, else vol.  tsdf_out / color_out may alias their inputs.
 * ----------------------------------------------------------------------------------- */
typedef struct GnbFusionParams {
    int32_t nx, ny, nz;
    float voxel_size;
    float origin[3];
    float trunc_margin;
    int32_t n_frames;              /* <= GNB_MAX_FRAMES per call; call again for more        */
    int32_t H, W;
    const float* h_projection;     /* HOST (n_frames,3,4) world->pixel, row-major           */
    const float* depth;            /* (n_frames,H,W)                                        */
    const float* color;            /* (n_frames,3,H,W) or NULL                              */
    const int32_t* label;          /* (n_frames,H,W) or NULL                                */
    float* tsdf_vol;               /* (V)                                                   */
    float* weight_vol;             /* (V)                                                   */
    float* color_vol;              /* (3,V) or NULL (iff color == NULL)                     */
    int32_t* label_vol;            /* (V) or NULL (iff label == NULL)                       */
    void* scratch;                 /* optional, >= gnb_tsdf_fusion_scratch_bytes(): enables  */
    int64_t scratch_bytes;         /*   the depth-band culling (same results, fewer projections) */
} GnbFusionParams;

int64_t gnb_tsdf_fusion_scratch_bytes(int n_frames, int H, int W);
int gnb_tsdf_fusion_integrate(const GnbFusionParams* p, void* stream);
int gnb_tsdf_fusion_finalize(const float* tsdf_vol, const float* weight_vol, const float* color_vol,
                             int64_t n_voxels, float* tsdf_out, float* color_out, void* stream);

/* -------------------------------------------------------------------------------------
 * Decoder.  Replaces
 *   PositionalEncoding.forward()   src/models/components/positional_encoding.py:28-40
 *   ResnetFC.forward()             src/models/components/resnetfc.py:134-189 (default options:
 *                                   ReLU, no spade / layer-norm, combine_layer > n_blocks)
 *   TSDFHeadSimple.forward()       src/models/components/heads3d.py:36-50
 *   the glue in GenNerf.forward()  src/models/model.py:226-246
 * Weight layout is the reference's nn.Linear layout (out_features, in_features) row-major.
 * ----------------------------------------------------------------------------------- */
#define GNB_TC_FP16 0              /* 11-bit significand, |x| <= 65504 (saturated)          */
#define GNB_TC_BF16 1              /* 8-bit significand, fp32 range                         */
typedef struct GnbDecoderWeights {
    int32_t d_feat;                /* lin_in in_features  (= encoder_latent)                */
    int32_t d_code;                /* lin_z in_features   (= 3 + 6*num_freqs, or 3)         */
    int32_t d_hidden;
    int32_t n_blocks;              /* <= 8                                                  */
    int32_t d_out;                 /* d_out_geo + d_out_sem                                 */
    int32_t d_geo;                 /* head input = first d_geo outputs                      */
    float alpha;                   /* ResnetFC.alpha (value of the learnable scalar)        */
    /* positional encoding */
    int32_t use_code;              /* 0: code = xyz; 1: positional encoding of xyz;         */
                                   /* 2: the `xyz` argument already holds (n_rows, d_code)  */
                                   /*    codes (stand-alone ResnetFC.forward(zx))           */
    int32_t num_freqs;
    float freq_factor;
    int32_t include_input;
    /* fp32 parameters */
    const float* lin_in_w;  const float* lin_in_b;
    const float* lin_z_w[8]; const float* lin_z_b[8];
    const float* fc0_w[8];   const float* fc0_b[8];
    const float* fc1_w[8];   const float* fc1_b[8];
    const float* lin_out_w; const float* lin_out_b;
    const float* head_w;    const float* head_b;
    int32_t tc_dtype;              /* tensor-core operand type: GNB_TC_FP16 (default) or GNB_TC_BF16 */
    int32_t* status;               /* optional device int32 (NULL = off): the tensor-core kernels OR bit 0 into it when an  */
                                   /* fp16 operand (input feature, code or activation) saturated at +-65504, i.e. the      */
                                   /* result is outside the 1e-2 contract and the fp32 decoder should be used instead      */
    const float* alpha_dev;        /* optional device float (NULL = use `alpha`): ResnetFC.alpha read on the device by the  */
                                   /* pack / fp32 kernels -- a training step (the scalar is a learnable parameter that      */
                                   /* changes every step) then needs no device-to-host read of it, i.e. no stream sync      */
} GnbDecoderWeights;

/* Stand-alone pieces (the reference's modules called on their own). */
int gnb_positional_encoding(const float* x, int64_t n_rows, int num_freqs, float freq_factor,
                            int include_input, float* out, void* stream);      /* (n,3)->(n,d_code) */
int gnb_tsdf_head(const float* feat_geo, int64_t n_rows, int d_geo, int64_t row_stride,
                  const float* head_w, const float* head_b, float* tsdf, void* stream);

/* fp32 CUDA-core decoder (exact mode, tolerance 1e-5 relative to the oracle). */
int gnb_decode_fp32(const GnbDecoderWeights* w, const float* xyz, const float* feat,
                    int64_t n_rows, float* out, float* tsdf, void* stream);

/* tcgen05/TMEM tensor-core decoder (fast mode): 16-bit operands (w->tc_dtype), fp32
 * accumulation and fp32 residual stream; |tsdf - oracle| <= 1e-2 with fp16 operands.
 * `packed` is the device buffer written by gnb_decoder_pack_tc (gnb_decoder_packed_bytes
 * bytes); it must be re-packed after the fp32 parameters change. */
int64_t gnb_decoder_packed_bytes(const GnbDecoderWeights* w);
int gnb_decoder_pack_tc(const GnbDecoderWeights* w, void* packed, void* stream);
/* gnb_decode_tc with the features given as the 16-bit operand image a sampler wrote (GnbSampleParams.image, with
 * image_kchunks = gnb_decoder_image_kchunks(w) and image_dtype = w->tc_dtype): the kernel brings a tile's features into
 * shared memory with one bulk copy instead of converting 128 fp32 rows.  Same results as gnb_query_fused_tc. */
int gnb_decoder_image_kchunks(const GnbDecoderWeights* w);   /* 0: these weights / options do not take an image */
int gnb_decode_image_tc(const GnbDecoderWeights* w, const void* packed, const float* xyz, const void* image,
                        int64_t n_rows, float* out, float* tsdf, void* stream);
/* Wide latent codes (d_feat > 512, i.e. more than 8 k-chunks of lin_in -- the reference's default Hydra config has
 * encoder_latent = 512 + 32, configs/model/gen_nerf.yaml:43,56 and src/models/model.py:35-44): the tensor-core decoder takes
 * the features through the operand image ONLY and streams its chunks through shared memory; gnb_decode_tc / gnb_query_fused_tc
 * return GNB_E_UNSUPPORTED for such weights.  gnb_features_to_image converts fp32 feature rows (n_rows, feat_stride) into that
 * image (all image_kchunks * 64 columns of every 128-row tile are written, padding as zeros; image bytes =
 * ceil(n_rows / 128) * image_kchunks * 16384); status as GnbSampleParams.image_status. */
int gnb_features_to_image(const float* feat, int64_t n_rows, int d_feat, int64_t feat_stride, int image_kchunks,
                          int image_dtype, void* image, int32_t* status, void* stream);
/* gnb_decode_tc that also stores the 16-bit activations every layer consumed -- what the backward pass of a training step
 * needs (ReLU masks for dgrad, left operands for wgrad): activations is [2*n_blocks + 1][n_rows][d_hidden] in w->tc_dtype;
 * slab 2i = relu(x_i + alpha*lin_z_i(code)) (input of blocks.i.fc_0), slab 2i+1 = relu(fc_0 output) (input of fc_1),
 * slab 2*n_blocks = relu(x) (input of lin_out).  resnetfc.py:134-189 under autograd. */
int gnb_decode_tc_save(const GnbDecoderWeights* w, const void* packed, const float* xyz, const float* feat,
                       int64_t n_rows, float* out, float* tsdf, void* activations, void* stream);
int gnb_decode_tc(const GnbDecoderWeights* w, const void* packed, const float* xyz,
                  const float* feat, int64_t n_rows, float* out, float* tsdf, void* stream);

/* Backward links of the ResNet-MLP in a training step (row a15: resnetfc.py:54-63, 134-189 under autograd, where every
 * Linear + ReLU link costs threshold_backward + the skip connection's add + a column sum for the bias gradient).  With the
 * 16-bit activations gnb_decode_tc_save stored, one pass over the (n_rows, d) gradient:
 *   out[r,c] = (act[r,c] > 0 ? pre[r,c] : 0) + (res ? res[r,c] : 0);  colsum[c] += sum_r out[r,c]  (atomic order; NULL = off);
 *   act32[r,c] = (float)act[r,c]  (NULL = off: the fp32 left operand of the layer's weight-gradient GEMM).
 * `pre` is the dgrad GEMM's result (grad @ W).  Row strides (ld_*) in elements, multiples of 4; pointers 16-byte aligned
 * (act: 8); d % 4 == 0; act_dtype GNB_TC_FP16 / GNB_TC_BF16; out may alias pre or res.  colsum must be zeroed by the caller. */
int gnb_mlp_grad_link(const float* pre, int64_t ld_pre, const void* act, int64_t ld_act, int act_dtype,
                      const float* res, int64_t ld_res, float* out, int64_t ld_out, float* act32, int64_t ld_act32,
                      float* colsum, int64_t n_rows, int d, void* stream);
/* Gradient entering the decoder from its two outputs (heads3d.py:36-50 and model.py:226-246 under autograd):
 *   s[r] = g_tsdf[r] * (1 - tsdf[r]^2);  G[r,c] = g_out[r,c] + (c < d_geo ? s[r] * head_w[c] : 0);
 *   d_head_w[c] += sum_r s[r] * out[r,c];  d_head_b += sum_r s[r];  d_lin_out_b[c] += sum_r G[r,c]   (atomic order).
 * g_out or g_tsdf may be NULL (that output received no gradient); the three sums must be zeroed by the caller. */
int gnb_mlp_grad_head(const float* g_out, const float* g_tsdf, const float* out, const float* tsdf, const float* head_w,
                      int64_t n_rows, int d_out, int d_geo, float* G, float* d_head_w, float* d_head_b,
                      float* d_lin_out_b, void* stream);

/* The decoder's whole backward pass in ONE call (row a15; loss.backward() through ResnetFC + TSDFHeadSimple in the reference,
 * resnetfc.py:134-189, heads3d.py:36-50): the two kernels above around plain library GEMMs (cuBLAS SGEMM, TF32 tensor cores
 * when `tf32` != 0 -- the reference trains with torch.set_float32_matmul_precision("high"), src/utils/utils.py:48; no gradient
 * is ever stored in 16 bits).  cuBLAS is bound with dlopen at the first call (GNB_E_UNSUPPORTED if it cannot be loaded); the
 * rest of the library does not depend on it.
 * Inputs: `w` as given to gnb_decode_tc_save (use_code = 2: `code` holds the (n_rows, d_code) codes; w->alpha or w->alpha_dev);
 * feat (n_rows, d_feat); out / tsdf / activations as gnb_decode_tc_save wrote them; g_out (n_rows, d_out) and g_tsdf (n_rows)
 * = d loss / d outputs, either may be NULL.
 * Outputs (fp32, shapes of the parameters): the weight gradients, lin_z biases, g_code (n_rows, d_code) and g_feat
 * (n_rows, d_feat) [either may be NULL] are overwritten; the other bias gradients, head_w / head_b and alpha are ACCUMULATED
 * into (atomic order) and must be zeroed by the caller.  head_w / head_b are only touched when g_tsdf is given. */
typedef struct GnbDecoderGrads {
    float* lin_in_w;  float* lin_in_b;
    float* lin_z_w[8]; float* lin_z_b[8];
    float* fc0_w[8];   float* fc0_b[8];
    float* fc1_w[8];   float* fc1_b[8];
    float* lin_out_w; float* lin_out_b;
    float* head_w;    float* head_b;
    float* alpha;
    float* g_code;    float* g_feat;
} GnbDecoderGrads;
int64_t gnb_decode_train_bwd_workspace_bytes(const GnbDecoderWeights* w, int64_t n_rows);
int gnb_decode_train_bwd(const GnbDecoderWeights* w, const float* code, const float* feat, const float* out, const float* tsdf,
                         const void* activations, const float* g_out, const float* g_tsdf, int64_t n_rows,
                         const GnbDecoderGrads* grads, void* workspace, int64_t workspace_bytes, int tf32, void* stream);

/* Fused sampler + tensor-core decoder (GenNerf.forward, model.py:207-248): xyz -> feat
 * (optional output), out (feat_geo|feat_sem), tsdf, in one kernel. */
int gnb_query_fused_tc(const GnbSampleParams* s, const GnbDecoderWeights* w, const void* packed,
                       float* out, float* tsdf, void* stream);

/* The same with the queries first counting-sorted by the 8x8x8-voxel brick of the volume they fall into (the brick sort of
 * gnb_sample_features_binned): the kernel's sampling prologue then walks the volume brick by brick, so its 8 corner reads
 * per query hit L1 / L2 instead of DRAM when the volume is larger than L2 (config 4: 805 MB), and results are written to
 * the rows the original query order names -- bit-identical outputs.  `scratch`: gnb_query_fused_sorted_scratch_bytes(s)
 * bytes (0: these parameters cannot be sorted -- no channels-last volume; use gnb_query_fused_tc). */
int64_t gnb_query_fused_sorted_scratch_bytes(const GnbSampleParams* s);
int gnb_query_fused_sorted_tc(const GnbSampleParams* s, const GnbDecoderWeights* w, const void* packed,
                              float* out, float* tsdf, void* scratch, int64_t scratch_bytes, void* stream);

/* Dense-grid extraction (GenNerf.predict_tsdf, model.py:752-790 with get_grid_coordinates, utils.py:926-935): the same
 * fused kernel, but the (nx*ny*nz, 3) query grid is never materialised -- query row r of a scene is the grid point
 * (axes[i], axes[nx + j], axes[nx + ny + k]) with r = (i*ny + j)*nz + k.  `axes` (device, nx + ny + nz floats) holds the
 * three torch.linspace axes; s->xyz and s->n_query are ignored (n_query = nx*ny*nz), h_grid3 = {nx, ny, nz} on the host.
 * No 10 000-point chunk loop, no per-chunk re-normalisation or D2H copy, and 12 B/query less HBM traffic. */
int gnb_query_grid_fused_tc(const GnbSampleParams* s, const int32_t* h_grid3, const float* axes,
                            const GnbDecoderWeights* w, const void* packed, float* out, float* tsdf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENNERF_B200_H */

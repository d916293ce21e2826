/*
 * pe_b200.h — C ABI of the B200-native volumetric renderer for PlayableEnvironments.
 *
 * The reference has no FFI of its own (it is pure Python/PyTorch); this
 * header is the boundary its Python modules bind through ctypes.  Each
 * entry point names the reference interface it replaces (paths relative to the upstream tree).
 *
 * Conventions
 *  - plain-old-data structs, raw DEVICE pointers (tensor.data_ptr()), sizes in elements;
 *  - no allocation and no ownership transfer: every buffer, including the workspace, belongs to
 *    the caller; calls are asynchronous on the supplied CUDA stream, never synchronise the
 *    device and never read device memory from the host;
 *  - all floating point tensors are contiguous fp32 unless stated; leading dims (B,O,C) of the
 *    reference are flattened into `images`;
 *  - return value 0 = success, negative = error (pe_last_error() gives a thread-local message);
 *  - re-entrant: no global mutable state, safe from several host threads on different streams
 *    (nn.DataParallel replicas, train.py:61).
 */
#ifndef PE_B200_H
#define PE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_ABI_VERSION 9
#define PE_MAX_OBJECTS 8      /* object instances composed in one call                     */
#define PE_MAX_LAYERS 12      /* backbone layers of a field / ray bender                   */
#define PE_MAX_OCTAVES 16

#define PE_MAX_PEERS 8
typedef void* pe_stream_t;    /* cudaStream_t */

enum PeNerfKind   { PE_NERF_ADAIN = 0,      /* model/nerf_models/adain_style_nerf_model.py          */
                    PE_NERF_SKYBOX_V3 = 1   /* model/nerf_models/skybox_adain_style_nerf_model_v3.py */ };
enum PeBenderKind { PE_BENDER_ZEROED = 0,   /* model/nerf_models/zeroed_ray_bender_model.py          */
                    PE_BENDER_POSITIONAL = 1/* model/nerf_models/positional_ray_bender_model.py      */ };
enum PePrecision  { PE_PRECISION_FP32 = 0,  /* CUDA-core fp32 FMA path, any shape                    */
                    PE_PRECISION_FP16 = 1,  /* tcgen05 kind::f16, fp16 operands, fp32 accumulate     */
                    PE_PRECISION_FP16X2 = 2,/* tcgen05, weights split hi+lo fp16 (2 MMA passes)      */
                    PE_PRECISION_FP16X3 = 3,/* tcgen05, weights AND activations split hi+lo (3 MMAs per k-step): fp32-class accuracy */
                    PE_PRECISION_MIXED = 4  /* tcgen05, hi+lo weights on the late trunk layers (L3-L7) only; objects with fewer
                                               than 64 samples per ray (large sample spacing amplifies raw-alpha error) run as FP16X3 */ };
enum PeError      { PE_OK = 0, PE_ERR_INVALID = -1, PE_ERR_UNSUPPORTED = -2, PE_ERR_CUDA = -3, PE_ERR_WORKSPACE = -4 };

/* Architecture + geometry of one object model
 * (model/nerf_models/ray_bending_style_nerf_model.py:17-50 and its two sub-model configs). */
typedef struct PeObjectDesc {
    int32_t nerf_kind, bender_kind;
    int32_t width, layers, skip, octaves, features;      /* nerf_model: layers_width, backbone_layers_count, skip_layer_idx, octaves, output_features */
    int32_t style_features, deformation_features;
    int32_t b_width, b_layers, b_skip, b_octaves;        /* ray_bender_model (positional)         */
    int32_t positions;                                   /* positions_count_coarse                */
    int32_t is_static;                                   /* object_ids_helper.py:47-55            */
    int32_t canonical_pose;                              /* zero the displacements (:108-110)     */
    float bbox[6];                                       /* x_lo,x_hi,y_lo,y_hi,z_lo,z_hi          */
    float z_near_min, z_far_max, empty_space_alpha;
    float b_anneal[PE_MAX_OCTAVES];                      /* annealable_positional_encoder.py:54-58, evaluated on the host from current_step */
    const void* packed;                                  /* blob written by pe_pack_object (device)*/
    int32_t aware_rounding;                              /* the blob's fp16 weight stream was rounded against activation statistics
                                                            (PeObjectParams.backbone_in_moments): the `mixed` mode then runs objects
                                                            with >= 96 samples per ray in ONE weight pass                         */
} PeObjectDesc;

/* fp32 parameter tensors of one object model, nn.Linear layout [out][in]
 * (state_dict names in SURVEY.md section 3.3). NULL where the architecture has none. */
typedef struct PeObjectParams {
    const float* backbone_w[PE_MAX_LAYERS];  const float* backbone_b[PE_MAX_LAYERS];   /* nerf_model.backbone_layers.{i} */
    const float* alpha_w;  const float* alpha_b;                                        /* nerf_model.alpha_head          */
    const float* head0_w;                                                               /* features_head.0 (no bias)      */
    const float* affine1_w; const float* affine1_b; const float* bn1_mean; const float* bn1_var;  /* features_head.1     */
    const float* head3_w;                                                               /* features_head.3 (no bias)      */
    const float* affine2_w; const float* affine2_b; const float* bn2_mean; const float* bn2_var;  /* features_head.4     */
    const float* head6_w;  const float* head6_b;                                        /* features_head.6                */
    const float* bender_w[PE_MAX_LAYERS];    const float* bender_b[PE_MAX_LAYERS];     /* ray_bender.backbone_layers.{i} */
    const float* bender_out_w;                                                          /* ray_bender.output_head (no bias) */
    /* pe_pack_object only, optional (shipped field shape): second-moment matrices E[a a^T] ([in][in], fp32) of the fp16-rounded INPUTS of
     * backbone layer l / of features_head.0, measured by the caller on sample positions inside the bounding box.  Where given, the
     * fp16 weight stream is rounded "activation-aware": per output row the fp16 neighbour (down / up) of every weight is chosen so
     * that the expected squared error e^T E[a a^T] e of the single-pass product is minimal (coordinate descent), instead of the
     * data-free zero-sum rounding.  Removes most of the systematic part of the single-pass error.          */
    const float* backbone_in_moments[PE_MAX_LAYERS];
    const float* head0_in_moments;
} PeObjectParams;

/* One composer call: model/object_composer.py:786-812 (ObjectComposer.forward). */
typedef struct PeScene {
    int32_t images;            /* B*O*C                                                      */
    int32_t rays;              /* samples_per_image                                          */
    int32_t objects;           /* object instances (transformation_matrix_w2o.size(-1))      */
    int32_t static_objects;    /* leading instances that are static (fix_object_overlaps)    */
    int32_t perturb;           /* stratified jitter + raw-alpha noise                        */
    int32_t training;          /* BatchNorm uses batch statistics (adain.py:47)              */
    int32_t fix_object_overlaps; /* utils/configuration.py:191-192                           */
    int32_t apply_activation;  /* sigmoid on features (object_composer.py:548-549)           */
    int32_t precision;         /* PePrecision                                                */
    int32_t explicit_positions;/* 0: sample along rays; 1: field evaluation on given points  */
    int32_t keep_samples;      /* 1: the forward leaves every per-sample tensor, AdaIn constant and BatchNorm sum in its workspace,
                                  which the caller keeps and hands to pe_render_backward_saved (no forward recompute there) */
    int32_t explicit_t;        /* 1: hierarchical ("fine") pass -- every object is sampled at PeInputs.sample_t[k] (all non-NULL).  The
                                  backward of such a call runs the exact fp32 kernels: on the resampled (clustered) ray parameters the
                                  tensor-core backward deviates by 5-10 % on the fine model's trunk gradients (measured on the
                                  static_fine golden, tests/gpu_fine_diag.py; cause not isolated), the fp32 one by 1e-2              */
    int32_t divergence;        /* 1: the forward also evaluates the Hutchinson divergence of every positional ray bender's displacement
                                  field (object_composer.py:582-601) from PeInputs.divergence_noise / divergence_params               */
    int32_t bent_gradients;    /* 1: the backward also takes dL/d (sample position + displacement) from PeOutGrads.bent_positions -- the
                                  backward of forward_expected_positions (object_composer.py:603-722)                                 */
    PeObjectDesc object[PE_MAX_OBJECTS];
    /* backward only, optional: 1 + the exact number of 128-sample tiles the tensor-core backward of object k will walk, as counted by
     * pe_forward_tile_counts on the kept forward of this very call; 0 = unknown (the backward then sizes its activation stash and its
     * batch count for the worst case, every slot inside the box).  A training step whose worst case exceeds the stash would otherwise
     * run in batches and repeat the recompute in each of the three BatchNorm phases.                                                  */
    int64_t bwd_tiles[PE_MAX_OBJECTS];
} PeScene;

typedef struct PeInputs {
    const float* ray_origins;        /* [images][3]                world space               */
    const float* ray_directions;     /* [images][rays][3]                                    */
    const float* w2o;                /* [images][objects][3][4]    rows of the 4x4 (:828)     */
    const float* style[PE_MAX_OBJECTS];        /* [images][style_features]                   */
    const float* deformation[PE_MAX_OBJECTS];  /* [images][deformation_features]             */
    const uint8_t* object_in_scene;  /* [images][objects]                                    */
    const float* rand[PE_MAX_OBJECTS];         /* [images][rays][P_k] uniform, replaces torch.rand (ray_helper.py:1275); NULL unless perturb */
    const float* noise[PE_MAX_OBJECTS];        /* [images][rays][P_k] normal, replaces torch.randn (object_composer.py:194) in the per-object integrate */
    const float* noise_global;       /* [images][rays][sum P]     same, for the composed scene */
    const float* positions;          /* explicit_positions: [images][rays][3] object space   */
    /* hierarchical ("fine") pass, model/object_composer.py:563-578: explicit ray parameters [images][rays][P_k] of object k (sorted
     * along P: RayHelper.create_ray_positions_weighted, utils/lib_3d/ray_helper.py:1320-1347) instead of the stratified samples of
     * create_ray_positions; NULL = stratified.  rand[k] is not read for such an object.                                          */
    const float* sample_t[PE_MAX_OBJECTS];
    /* Hutchinson divergence (scene.divergence): [images][rays][P_k][3] normal e, replaces torch.randn_like (object_composer.py:597); the
     * per-sample value e . (d displacement / d position) e enters integrated_divergence = mean(alpha |div|) (:776-778).  NULL for an
     * object: zeros.  divergence_params[k]: the fp32 parameter tensors of instance k (the vector-Jacobian product reads the nn.Linear
     * layout, like the backward).                                                                                                   */
    const float* divergence_noise[PE_MAX_OBJECTS];
    const PeObjectParams* divergence_params;
} PeInputs;

/* Result of ObjectComposer.integrate (model/object_composer.py:724-784). Any pointer may be NULL. */
typedef struct PeIntegrated {
    float* integrated_features;      /* [images][rays][features]                             */
    float* opacity;                  /* [images][rays]                                       */
    float* weights;                  /* [images][rays][P]                                    */
    float* depth;                    /* [images][rays]                                       */
    float* disparity;                /* [images][rays]                                       */
    float* integrated_displacements_magnitude;  /* [images][rays]                            */
    float* integrated_divergence;    /* [images][rays]  (zeros: Hutchinson term, see DESIGN)  */
} PeIntegrated;

/* Decoder hand-off written by the kernel that composes the scene: the rays of a frame are the concatenation of
 * the decoder's strided grids (RayHelper.sample_all_rays_strided_grid, utils/lib_3d/ray_helper.py:433-482), every grid keeps its own
 * channel range of the composed features and wants them channels-first (fold_strided_tensors + split_features_by_layer + permute,
 * model/environment_model_backpropagated_autoencoder.py:129-168, ..._multiresolution_backpropagated_autoencoder.py:29-99).  Segment q =
 * rays [ray_begin, ray_begin + ray_count) of every image, channels [channel_begin, channel_begin + channel_count) (multiples of 32),
 * grid[q] = [images][channel_count][ray_count].  With global.integrated_features == NULL only those channels are accumulated (the rest
 * of the per-sample features is never read) and only the grids are written.                                                         */
#define PE_MAX_HANDOFF 4
typedef struct PeHandoff {
    int32_t segments;                          /* 0: no hand-off                                */
    int32_t ray_begin[PE_MAX_HANDOFF];
    int32_t ray_count[PE_MAX_HANDOFF];
    int32_t channel_begin[PE_MAX_HANDOFF];
    int32_t channel_count[PE_MAX_HANDOFF];
    float*  grid[PE_MAX_HANDOFF];
} PeHandoff;

typedef struct PeOutputs {
    PeIntegrated object[PE_MAX_OBJECTS];        /* results["coarse"]["object_k"]             */
    PeIntegrated global;                        /* results["coarse"]["global"]               */
    /* per-sample tensors (optional; [images][rays][P_k] (+[features]|[3])): the return value of
     * RayBendingStyleNerfModel.forward (ray_bending_style_nerf_model.py:137-219)              */
    float* raw_features[PE_MAX_OBJECTS];
    float* raw_alphas[PE_MAX_OBJECTS];
    float* displacements[PE_MAX_OBJECTS];
    float* positions_t[PE_MAX_OBJECTS];
    /* training: BatchNorm batch statistics [2][C] (mean, unbiased variance) per AdaIn layer; the
     * caller applies the momentum update to its running buffers                                  */
    float* bn1_running[PE_MAX_OBJECTS];
    float* bn2_running[PE_MAX_OBJECTS];
    /* fused all-gather of the rendered feature grid (the one collective of the path): `peers` extra destinations of the
     * composed scene's integrated_features, [images][rays][features] each -- buffers of the OTHER GPUs of the box mapped into this
     * process (CUDA virtual-memory handles: cuMemImportFromShareableHandle + cuMemMap, e.g. torch symmetric memory) or local ones.  The fused field kernel stores every ray's features to all of them as it produces
     * them (P2P stores over NVLink, no separate collective and no SM taken from the render); other paths copy after their last kernel.
     * The caller orders the peers' reads after this call with its own cross-rank synchronisation.                                   */
    int32_t peers;
    float* peer_features[PE_MAX_PEERS];
    /* multi-object scenes (the compositor produces the scene's grid), inference: see PeHandoff */
    PeHandoff handoff;
} PeOutputs;

/* -- library ------------------------------------------------------------------------------- */
int         pe_abi_version(void);
const char* pe_last_error(void);
/* number of kernels launched by this thread since the last call of this function */
int64_t     pe_take_launch_count(void);

/* -- parameters: replaces nn.Module parameter storage of model/nerf_models/*.py --------------- */
size_t pe_packed_bytes(const PeObjectDesc* desc);
int    pe_pack_object(const PeObjectDesc* desc, const PeObjectParams* params, void* packed, pe_stream_t stream);

/* -- the hot path: replaces ObjectComposer.forward (model/object_composer.py:786-892) and, through
 *    it, forward_object :486-580, RayHelper.transform_rays / create_ray_positions
 *    (utils/lib_3d/ray_helper.py:1203-1282), RayBendingStyleNerfModel.forward, compose :399-447 and
 *    integrate :724-784.  TensorBatchifier chunking (utils/tensor_batchifier.py) is unnecessary:
 *    the workspace is O(rays), so the whole frame is one call.                                  */
size_t pe_workspace_bytes(const PeScene* scene);
int    pe_render_forward(const PeScene* scene, const PeInputs* in, const PeOutputs* out,
                         void* workspace, size_t workspace_bytes, pe_stream_t stream);

/* -- backward of the hot path: replaces the autograd graph the reference builds under ObjectComposer.forward
 *    (loss.backward() in training/trainer.py:643 replays every op of model/object_composer.py:786-892).
 *    The forward is recomputed from the inputs (nothing is saved between the two calls except what the caller
 *    passes again), gradients are ACCUMULATED (+=) into caller-zeroed fp32 buffers; a NULL pointer skips that
 *    gradient.  Instances that share one object model may point at the same parameter-gradient buffers.     */
typedef struct PeIntegratedGrads {             /* dL/d(outputs of ObjectComposer.integrate); NULL = zero     */
    const float* integrated_features;          /* [images][rays][features]                                   */
    const float* opacity;                      /* [images][rays]                                             */
    const float* weights;                      /* [images][rays][P]                                          */
    const float* depth;                        /* [images][rays]                                             */
    const float* disparity;                    /* [images][rays]                                             */
    const float* integrated_displacements_magnitude;   /* [images][rays]                                     */
} PeIntegratedGrads;

typedef struct PeOutGrads {
    PeIntegratedGrads object[PE_MAX_OBJECTS];
    PeIntegratedGrads global;
    /* scene.bent_gradients: [images][rays][P_k][3] dL/d (sample position + displacement) of object k, object space, or NULL */
    const float* bent_positions[PE_MAX_OBJECTS];
} PeOutGrads;

typedef struct PeObjectParamGrads {            /* same tensors and layouts as PeObjectParams (nn.Linear [out][in]) */
    float* backbone_w[PE_MAX_LAYERS];  float* backbone_b[PE_MAX_LAYERS];
    float* alpha_w;  float* alpha_b;
    float* head0_w;
    float* affine1_w; float* affine1_b;
    float* head3_w;
    float* affine2_w; float* affine2_b;
    float* head6_w;  float* head6_b;
    float* bender_w[PE_MAX_LAYERS];    float* bender_b[PE_MAX_LAYERS];
    float* bender_out_w;
} PeObjectParamGrads;

typedef struct PeInGrads {
    float* ray_origins;                        /* [images][3]                                                */
    float* ray_directions;                     /* [images][rays][3]                                          */
    float* w2o;                                /* [images][objects][3][4]                                    */
    float* style[PE_MAX_OBJECTS];              /* [images][style_features]                                   */
    float* deformation[PE_MAX_OBJECTS];        /* [images][deformation_features]                             */
    PeObjectParamGrads params[PE_MAX_OBJECTS]; /* per object instance                                        */
    float* sample_t[PE_MAX_OBJECTS];           /* [images][rays][P_k] dL/dt of PeInputs.sample_t[k] (written, not accumulated) or NULL */
} PeInGrads;

size_t pe_backward_workspace_bytes(const PeScene* scene);
/* `params[k]`: the fp32 parameter tensors of instance k (the transposed products W^T g read the nn.Linear layout
 * directly).  `scene`/`in` must be the ones of the forward call (including rand/noise when perturb is set).  */
int    pe_render_backward(const PeScene* scene, const PeInputs* in, const PeObjectParams* params,
                          const PeOutGrads* grad_out, const PeInGrads* grad_in,
                          void* workspace, size_t workspace_bytes, pe_stream_t stream);
/* Same, on top of the workspace a pe_render_forward call with scene->keep_samples = 1 filled (`saved_forward`, untouched since;
 * the scene must be that call's scene): what autograd's saved tensors are to the reference (training/trainer.py:643).         */
int    pe_render_backward_saved(const PeScene* scene, const PeInputs* in, const PeObjectParams* params,
                                const PeOutGrads* grad_out, const PeInGrads* grad_in,
                                const void* saved_forward, size_t saved_forward_bytes,
                                void* workspace, size_t workspace_bytes, pe_stream_t stream);

/* How many 128-sample tiles the tensor-core backward of each object will walk, from the in-box masks of a kept forward
 * (scene->keep_samples = 1): counts[k] (device memory, PE_MAX_OBJECTS entries) = sum over images of ceil(in-box samples / 128), 0 for
 * objects whose backward does not run on the tensor cores.  The caller copies the counts to the host while the loss is computed and
 * hands them back as scene->bwd_tiles[k] = 1 + counts[k] (what autograd knows from its saved tensors' shapes: the number of rows of
 * the gathered in-box samples, model/nerf_models/ray_bending_style_nerf_model.py:170-176).                                          */
int    pe_forward_tile_counts(const PeScene* scene, const void* saved_forward, size_t saved_forward_bytes, int64_t* counts,
                              pe_stream_t stream);

/* -- stand-alone operators (module-level API of the reference) ---------------------------------- */
/* PositionalEncoder.forward / AnnealablePositionalEncoder.forward
 * (model/positional_encoder.py:41-65, model/annealable_positional_encoder.py:46-76).
 * x [n][dims] -> out [n][dims*(append_original + 2*octaves)]; weights NULL or [octaves].        */
int pe_positional_encoding(const float* x, int64_t n, int32_t dims, int32_t octaves, int32_t append_original,
                           const float* weights, float* out, pe_stream_t stream);
/* RayHelper.create_camera_rays + sample_all_rays_strided_grid + transform_rays(c2w)
 * (utils/lib_3d/ray_helper.py:15-52, 433-482, 1203-1227): focal [images], c2w [images][3][4],
 * strides[n_strides]; writes directions [images][R][3], origins [images][3], positions [images][R][2]
 * with R = sum_s (H/s)*(W/s).                                                                   */
int pe_generate_rays(const float* focal, const float* c2w, int32_t images, int32_t height, int32_t width,
                     const int32_t* strides, int32_t n_strides, float* directions, float* origins,
                     float* positions, pe_stream_t stream);
/* fold_strided_tensors + run_decoder_on_results channel split
 * (model/environment_model_backpropagated_autoencoder.py:129-168,
 *  model/environment_model_multiresolution_backpropagated_autoencoder.py:59-99):
 * features [images][R][F] -> per stride s a CHW grid [images][channels_s][H/s][W/s] taking the
 * channel range starting at sum of previous channels.                                            */
int pe_fold_feature_grids(const float* features, int32_t images, int32_t height, int32_t width, int32_t n_features,
                          const int32_t* strides, const int32_t* channels, int32_t n_strides,
                          float* const* grids, pe_stream_t stream);

/* -- debug / validation ---------------------------------------------------------------------- */
/* D[128][n] = A[128][k] * B[n][k]^T (+ bias[n], NULL for none, added by the rank-1 "ones" MMA) through
 * the same tcgen05 building blocks as the fused kernel (fp16 operands, fp32 accumulate).  Used by
 * tests to validate descriptors on the device.                                                  */
int pe_debug_umma_gemm(const float* a, const float* b, const float* bias, float* d, int32_t n, int32_t k, pe_stream_t stream);
/* Validation of two operand forms on one CTA.  mode 1: both operands MN-major, read from the activation layout of the field kernel with
 * K = its 128 rows: d[m][n] = sum_r a[r][m] * b[r][n] (a: [k][128], b: [k][n], rows >= k zero; lbo / sbo = descriptor byte strides).
 * mode 2: A operand from TMEM (TS form, written with tcgen05.st): d[m][n] = sum_k a[m][k] * b[n][k] (a: [128][k], b: [n][k]).     */
int pe_debug_umma_gemm2(int32_t mode, const float* a, const float* b, float* d, int32_t n, int32_t k, int32_t lbo, int32_t sbo,
                        pe_stream_t stream);
/* One layer ([N][K_src] fp32, nn.Linear layout) through the fp16 weight packing of pe_pack_object: zero-sum rounding, or -- with
 * `moments` ([K_src][K_src]) -- the activation-aware rounding of PeObjectParams.backbone_in_moments.  hi / lo: N * K_pad fp16 each in
 * the slab layout of the tensor-core stream (element (n, k) of K = 32 slab s at s*N*64 + (k%32/8)*16N + (n/8)*128 + (n%8)*16 + (k%8)*2). */
int pe_debug_pack_layer(const float* w, const float* moments, int32_t N, int32_t K_src, int32_t K_pad, int32_t sweeps, void* hi, void* lo,
                        pe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PE_B200_H */

// Kernel argument blocks and launcher declarations shared by the translation units of libpe_b200.so.
#pragma once

#include "pe_common.cuh"

// Arguments of the per-object field kernels (fp32 CUDA-core and tcgen05).
struct PeFieldArgs {
    PeObjectDesc ob;
    PeLayout L;
    int32_t images, rays, objects, k;
    int32_t perturb, explicit_positions, training, phase, apply_activation, precision;
    int32_t pass2_mask;          // mixed mode: 0 = the default two-pass layers, else 0x10000 | mask (bit l = layer l runs hi + lo)
    const float* origins;        // [images][3]
    const float* dirs;           // [images][rays][3]
    const float* w2o;            // [images][objects][12]
    const float* deformation;    // [images][D]
    const float* rand;           // [images][rays][P]
    const float* t_in;           // [images][rays][P] explicit ray parameters (fine pass) or NULL: stratified sampling
    const float* positions;      // explicit positions [images][rays][3]
    const uint8_t* ois;          // [images][objects]
    const float* aff1;           // [images][2][W]    effective AdaIn scale/shift (BatchNorm folded)
    const float* aff2;           // [images][2][W/2]
    float* t_out;                // [images][rays][P]
    float* raw_out;              // [images][rays][P]
    float* feat_out;             // [images][rays][P][F]
    float* disp_out;             // [images][rays][P][3] or NULL
    float* dispmag_out;          // [images][rays][P]
    uint8_t* inbox_out;          // [images][rays][P]
    double* stats;               // train-mode BatchNorm sums: [2][W] then [2][W/2]
    // fused per-object integration (tcgen05 path): outputs of ObjectComposer.integrate for this object
    PeIntegrated integ;
    const float* noise;          // [images][rays][P] raw-alpha noise or NULL
    // folded-head mode of the tcgen05 kernel (both or neither): per-ray weighted sums of the last hidden layer
    float* fold_v;               // [images][rays][128]
    float* fold_s;               // [images][rays]
    // pre-pass hand-off (objects with a ray bender on the tcgen05 path): written by the fp32 kernel in phase PE_PHASE_PREPASS,
    // read by the tcgen05 kernel, which then evaluates only the listed (non-empty) tiles
    float* bent;                 // [images][rays][P][3] bent sample positions (object space)
    uint8_t* flags;              // [images][rays][P]    bit 0: in the box, bit 1: bent position in the box (field evaluated)
    const int32_t* tile_list;    // [*tile_count] tiles (of floor(128/P) rays) that hold at least one sample with bit 1
    const int32_t* tile_count;
    float* h7_out;               // [images][rays][P][W] trunk output of the evaluated samples (train-mode recompute for the backward) or NULL
    // fused all-gather (folded-head path, the object is the scene): extra destinations of integrated_features, peer-mapped or local
    int32_t peers;
    float* peer_features[PE_MAX_PEERS];
};

// Arguments of the compositing kernel (model/object_composer.py:399-447, 724-784).
struct PeCompositeArgs {
    int32_t images, rays, objects, static_objects, features, total_positions;
    int32_t fix_overlaps, perturb, do_objects, do_global;
    int32_t positions[PE_MAX_OBJECTS];
    const float* dirs;                            // world-space directions [images][rays][3]
    const float* t[PE_MAX_OBJECTS];
    const float* raw[PE_MAX_OBJECTS];
    const float* feat[PE_MAX_OBJECTS];
    const float* dispmag[PE_MAX_OBJECTS];
    const float* div[PE_MAX_OBJECTS];             // per-sample divergence of the ray bender's displacement field or NULL (zeros)
    const uint8_t* inbox[PE_MAX_OBJECTS];
    const float* noise[PE_MAX_OBJECTS];
    const float* noise_global;
    PeIntegrated object[PE_MAX_OBJECTS];
    PeIntegrated global;
    PeHandoff handoff;                            // decoder hand-off of the composed scene's features (segments == 0: none)
};

// Arguments of the style prologue: [scale|bias] = Linear(style) (adain.py:30-32) with the BatchNorm
// statistics folded in: y = x*sc + sh, sc = scale/sqrt(var+eps), sh = bias - mean*sc.
struct PeStyleArgs {
    int32_t images, style_features, channels, training;
    const float* style;          // [images][S]
    const float* aff_w;          // [2C][S]
    const float* aff_b;          // [2C]
    const float* run_mean;       // [C]
    const float* run_var;        // [C]
    const double* stats;         // training: sum[C], sum of squares[C], sample count[1]
    float* out;                  // [images][2][C]
    float* running_out;          // training: [2][C] batch mean / unbiased batch variance (host applies momentum 0.1)
};

// Arguments of the compositing backward (gradients of integrate / compose w.r.t. the per-sample tensors).
struct PeCompositeBwdArgs {
    PeCompositeArgs f;                            // forward view: per-sample t / raw / feat / dispmag / inbox of every object
    PeIntegratedGrads g_object[PE_MAX_OBJECTS];   // upstream gradients
    PeIntegratedGrads g_global;
    float* cw_obj[PE_MAX_OBJECTS];                // out [images][rays][P]: d integrated_features(object) / d feature = cw_obj * I
    float* cw_glob[PE_MAX_OBJECTS];               // out: same for the composed scene
    float* g_raw[PE_MAX_OBJECTS];                 // out: dL/d raw alpha
    float* g_t[PE_MAX_OBJECTS];                   // out: dL/d t
    float* g_dm[PE_MAX_OBJECTS];                  // out: dL/d |displacement|
    float* g_dirs;                                // accumulated [images][rays][3]: through |d| of compute_position_distances
};

// Arguments of the field backward kernel (pe_field_bwd.cu).
struct PeFieldBwdArgs {
    PeFieldArgs f;                 // the forward arguments (the tile is recomputed)
    PeObjectParams w;              // fp32 parameters, nn.Linear layout [out][in]
    PeObjectParamGrads gw;         // accumulated parameter gradients
    int32_t bwd_phase;             // 0: full backward; 1 / 2: train-mode BatchNorm reductions of the second / first AdaIn layer
    const float* cw_obj;           // [images][rays][P]
    const float* cw_glob;
    const float* g_feat_obj;       // [images][rays][F] dL/d integrated_features of this object (or NULL)
    const float* g_feat_glob;      // [images][rays][F] dL/d integrated_features of the composed scene (or NULL)
    const float* g_raw;            // [images][rays][P]
    const float* g_dm;             // [images][rays][P]
    float* g_pos;                  // out [images][rays][P][3]: dL/d sample position, object space
    float* g_od;                   // out [images][rays][P][6] (skybox field): dL/d object-space origin and direction
    float* adain_sums;             // accumulated [images][2W + W]: per image sum g (W), sum g*x (W) of AdaIn 1, then the same (W/2 each) of AdaIn 2
    double* bn_sums;               // training, accumulated: [2W] of BatchNorm 1 (sum g*sc, sum g*sc*x), then [W] of BatchNorm 2
    const float* bn_fix;           // [2W + W]: k1, k2 of BatchNorm 1 then 2 (g_x = g*sc - k1 - x*k2); zeros in eval mode
    float* g_deformation;          // accumulated [images][D]
    float* stash;                  // per-block activation stash
    int64_t stash_floats;          // floats per block
    const float* h7_cache;         // [images][rays][P][W] trunk output written by the forward recompute (tensor-core path) or NULL:
    const uint8_t* inbox_in;       // ... with its evaluated-sample mask; the BatchNorm-reduction passes (bwd_phase 1, 2) start from them
    // compacted tiles (or NULL: tiles of 32 consecutive slots): per image the slots inside the object's box, so that a tile holds
    // 32 samples that need work instead of one ray's mostly empty samples
    const int32_t* slot_list;      // [images][rays * P] slot index inside the image; the first slot_count[img] entries are valid
    const int32_t* slot_count;     // [images]
    const int32_t* tile_begin;     // [images + 1] first tile of every image in the compacted numbering; [images] = number of tiles
    // ray-bender-only mode (objects whose field backward ran on the tensor cores, pe_bwd_tc.cu): dL/d bent position of every listed
    // slot; the kernel recomputes the bender of a tile, back-propagates through it and writes g_pos
    const float* g_bent_in;        // [images][rays][P][3] or NULL
    int32_t g_bent_flag;           // which samples read g_bent_in: flag bit 2 (field evaluated; 0 means this) or 1 (inside the box)
    // Hutchinson divergence (object_composer.py:582-601): with g_bent_in = e, div_out = e . (J e) where J = d displacement / d position
    float* div_out;                // [images][rays][P] or NULL
};

// Style / BatchNorm backward (pe_backward.cu)
struct PeStyleBwdArgs {
    int32_t images, style_features, channels, training;
    const float* style;            // [images][S]
    const float* aff_w;            // [2C][S]
    const float* run_mean; const float* run_var;
    const double* stats;           // training: forward sums (sum x, sum x^2, count)
    const float* adain_sums;       // + img * stride: A[C], B[C]
    int64_t adain_stride;
    float* g_aff_w; float* g_aff_b;   // accumulated
    float* g_style;                // accumulated [images][S]
};

struct PeGeometryBwdArgs {
    PeObjectDesc ob;
    int32_t images, rays, objects, k, perturb;
    const float* origins; const float* dirs; const float* w2o; const uint8_t* ois; const float* rand;
    const float* t_in;             // explicit ray parameters of the forward (fine pass) or NULL
    float* g_t_in;                 // [images][rays][P] dL/dt of the explicit ray parameters or NULL
    const float* g_pos;            // [images][rays][P][3]
    const float* g_t;              // [images][rays][P]
    const float* g_od;             // skybox: [images][rays][P][6] or NULL
    float* g_origins;              // accumulated [images][3]
    float* g_dirs;                 // accumulated [images][rays][3]
    float* g_w2o;                  // accumulated [images][objects][12]
};

// Arguments of the tensor-core field backward (pe_bwd_tc.cu).  The samples inside the object's box are compacted per image into tiles of
// 128 rows (slot_list / slot_count / tile_begin, TB = 128); every tile owns one block of the activation / gradient STASH in the
// canonical K-major no-swizzle operand layout (chunk = 8 columns x 128 rows x fp16 = 2048 B), so that a whole operand of a tile is ONE
// contiguous bulk copy, usable K-major (dX = G W: K = features) and MN-major (dW = G^T A: K = the tile's samples).
struct PeBwdTcArgs {
    PeFieldArgs f;                 // forward view: ob, L, images, rays, k, aff1/aff2, ois, deformation, bent, flags, raw_out, feat_out, inbox_out, ...
    const int32_t* slot_list;      // [images][rays * P]
    const int32_t* slot_count;     // [images]
    const int32_t* tile_begin;     // [images + 1]
    int32_t tile_capacity;         // tiles the stash holds (tiles beyond it are dropped and flagged in *overflow)
    int32_t* overflow;             // device flag or NULL
    unsigned char* stash;          // [tile][PE_BWD_FS_CHUNKS][2048] field stash
    unsigned char* bstash;         // [tile][PE_BWD_BS_CHUNKS][2048] ray-bender stash or NULL
    const float* cw_obj;           // [images][rays][P] from the compositing backward (pe_backward.cu)
    const float* cw_glob;
    const float* g_feat_obj;       // [images][rays][F] or NULL
    const float* g_feat_glob;
    const float* g_raw;            // [images][rays][P]
    const float* g_dm;             // [images][rays][P]
    const float* scale;            // device [2]: S (power of two applied to the fp16 gradient operands), 1 / S
    const float* bn_fix;           // [2W + W]: k1, k2 of BatchNorm 1 then 2 (zeros in eval mode)
    float* g_bent;                 // out [images][rays][P][3]: dL/d (bent) sample position, object space, unscaled
    float* g_pos;                  // out [images][rays][P][3]: dL/d sample position (== g_bent without a ray bender)
    PeObjectParamGrads gw;         // accumulated parameter gradients
    float* adain_sums;             // accumulated [images][3W] (layout of PeFieldBwdArgs::adain_sums)
    double* bn_sums;               // [3W] (layout of PeFieldBwdArgs::bn_sums)
    float* g_deformation;          // accumulated [images][D] or NULL
};
#define PE_BWD_TILE 128
#define PE_BWD_FS_CHUNKS 709       // field stash chunks per tile: activations (312) + AdaIn inputs (48) + gradients (330) + ReLU-mask words (19)
#define PE_BWD_BS_CHUNKS 215       // ray-bender stash chunks per tile: activations (108), gradients (100), masks (6), clamp state (1)
bool pe_bwd_tc_object_ok(const PeObjectDesc& ob);
int pe_tcT_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream);
// every launch covers the tiles [tile0, tile0 + args.tile_capacity) of the compacted numbering (one stash batch)
int pe_launch_bwd_fwd(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream);
int pe_launch_bwd_scale(const PeBwdTcArgs& args, float* scale, unsigned int* scratch, cudaStream_t stream);
int pe_launch_bwd_chain(const PeBwdTcArgs& args, int64_t tile0, int phase, int sm_count, cudaStream_t stream);
int pe_launch_bwd_dw(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream);
size_t pe_bwd_tc_stash_bytes(int64_t tiles);
bool pe_bwd_tc_bender_ok(const PeObjectDesc& ob, const PeLayout& L);
int pe_launch_bwd_scale_bender(const PeBwdTcArgs& args, float* scale, unsigned int* scratch, cudaStream_t stream);
int pe_launch_bwd_bender(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream);   // recompute + chain + dW of the ray bender

size_t pe_field_bwd_smem_bytes();
int64_t pe_field_bwd_stash_floats(const PeObjectDesc& ob, const PeLayout& L);
int pe_field_bwd_grid(int sm_count);
int pe_launch_field_bwd(const PeFieldBwdArgs& args, int sm_count, cudaStream_t stream);
int pe_launch_count_tiles(const uint8_t* flags, int flag_mask, int images, int64_t slots_per_image, int tile_rows, int64_t* out,
                          cudaStream_t stream);
int pe_launch_compact_slots(const uint8_t* flags, int flag_mask, int images, int64_t slots_per_image, int32_t* slot_list, int32_t* slot_count,
                            int32_t* tile_begin, cudaStream_t stream, int tile_rows = 32);
int pe_launch_composite_bwd(const PeCompositeBwdArgs& args, cudaStream_t stream);
int pe_launch_style_bwd(const PeStyleBwdArgs& args, cudaStream_t stream);
int pe_launch_bn_fix(const double* fwd_stats, const double* bn_sums, int channels, float* bn_fix, cudaStream_t stream);
int pe_launch_geometry_bwd(const PeGeometryBwdArgs& args, cudaStream_t stream);

size_t pe_field_fp32_smem_bytes(const PeObjectDesc& ob, const PeLayout& L);
int pe_launch_field_fp32(const PeFieldArgs& args, int sm_count, cudaStream_t stream);
int pe_launch_field_tc(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream);
int pe_launch_tile_list(const PeFieldArgs& args, int flag_mask, int32_t* tile_list, int32_t* tile_count, cudaStream_t stream);
int pe_launch_bender_tc(const PeFieldArgs& args, int sm_count, cudaStream_t stream);
int pe_launch_sample(const PeFieldArgs& args, int sm_count, cudaStream_t stream);
int pe_launch_composite(const PeCompositeArgs& args, cudaStream_t stream);
int pe_launch_style(const PeStyleArgs& args, cudaStream_t stream);
int pe_launch_pack(const PeObjectDesc& desc, const PeLayout& L, const PeObjectParams& params, void* packed, cudaStream_t stream);
int pe_device_sm_count(int* out);

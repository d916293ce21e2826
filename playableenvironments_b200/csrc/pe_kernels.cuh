// Kernel argument blocks and launcher declarations shared by the translation units of libpe_b200.so.
#pragma once

#include "pe_common.cuh"

// Arguments of the per-object field kernels (fp32 CUDA-core and tcgen05).
struct PeFieldArgs {
    PeObjectDesc ob;
    PeLayout L;
    int32_t images, rays, objects, k;
    int32_t perturb, explicit_positions, training, phase, apply_activation, precision;
    const float* origins;        // [images][3]
    const float* dirs;           // [images][rays][3]
    const float* w2o;            // [images][objects][12]
    const float* deformation;    // [images][D]
    const float* rand;           // [images][rays][P]
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
    const uint8_t* inbox[PE_MAX_OBJECTS];
    const float* noise[PE_MAX_OBJECTS];
    const float* noise_global;
    PeIntegrated object[PE_MAX_OBJECTS];
    PeIntegrated global;
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

size_t pe_field_fp32_smem_bytes(const PeObjectDesc& ob, const PeLayout& L);
int pe_launch_field_fp32(const PeFieldArgs& args, int sm_count, cudaStream_t stream);
int pe_launch_field_tc(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream);
int pe_launch_field_tc2(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream);
int pe_launch_composite(const PeCompositeArgs& args, cudaStream_t stream);
int pe_launch_style(const PeStyleArgs& args, cudaStream_t stream);
int pe_launch_pack(const PeObjectDesc& desc, const PeLayout& L, const PeObjectParams& params, void* packed, cudaStream_t stream);
int pe_device_sm_count(int* out);

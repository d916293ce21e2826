// Shared definitions of the B200 renderer kernels: packed-parameter layout, launch helpers and the
// ray-sampling geometry (one device function per reference helper it replaces).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pe_b200.h"

#define PE_MAX_TOTAL_POSITIONS 512     // sum of positions_count over objects supported by the compositor

// ------------------------------------------------------------------------------------------------
// error plumbing (thread local; the Python wrapper raises on non-zero)
// ------------------------------------------------------------------------------------------------
void pe_set_error(const char* fmt, ...);
void pe_count_launch(int n = 1);

#define PE_CUDA_CHECK(expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            pe_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PE_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

#define PE_LAUNCH_CHECK(name)                                                                   \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            pe_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
            return PE_ERR_CUDA;                                                                 \
        }                                                                                       \
        pe_count_launch();                                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------------
// packed parameter blob
// ------------------------------------------------------------------------------------------------
// fp32 section (CUDA-core path): every Linear stored TRANSPOSED, [in][out] row-major, so that a chunk of
// input rows is one coalesced read; fp16 section (tcgen05 path): UMMA K-major no-swizzle core-matrix
// slabs, see pe_field_tc.cu.  Offsets are in BYTES from the start of the blob, 256-byte aligned.
struct PeLayout {
    // nerf field
    int32_t enc;                    // encoding size E = dims*(1+2*octaves)
    int32_t in_dims;                // 3, or 6 for the skybox field
    int32_t k_in[PE_MAX_LAYERS];    // input width of backbone layer l (E, W, or W+E at the skip)
    int64_t bb_w[PE_MAX_LAYERS], bb_b[PE_MAX_LAYERS];
    int64_t alpha_w, alpha_b;
    int64_t head0_w, head3_w, head6_w, head6_b;
    int64_t aff1_w, aff1_b, bn1_mean, bn1_var;     // affine: nn.Linear layout [2C][S] (tiny, used per image)
    int64_t aff2_w, aff2_b, bn2_mean, bn2_var;
    // ray bender
    int32_t b_enc;                  // 3*(1+2*b_octaves) + deformation_features
    int32_t b_k_in[PE_MAX_LAYERS];
    int64_t bd_w[PE_MAX_LAYERS], bd_b[PE_MAX_LAYERS];
    int64_t bd_out_w;               // [Wb][3] transposed
    // tcgen05 section
    int64_t tc_base;                // start of the fp16 slab stream (0 if the shape is unsupported)
    int64_t tc_bytes_per_pass;      // bytes of one weight pass (hi); lo pass follows at +tc_bytes_per_pass
    int64_t tcb_base;               // ray-bender slab stream [hi pass | lo pass] (0 if the bender shape is unsupported)
    int64_t tcb_bytes_per_pass;
    int64_t tcT_base;               // TRANSPOSED slab stream of the field (dX = G W of the tensor-core backward, pe_bwd_tc.cu) [hi | lo]
    int64_t tcT_bytes_per_pass;
    int64_t tcbT_base;              // transposed slab stream of the ray bender [hi | lo]
    int64_t tcbT_bytes_per_pass;
    int32_t tc_supported;
    int64_t total;
};

__host__ __device__ inline int64_t pe_min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t pe_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Field shapes handled by the tcgen05 kernel (pe_field_tc.cu): the shipped field
// (configs/tennis/193_*.yaml:141-163, configs/minecraft/013_*.yaml:138-160).
__host__ __device__ inline bool pe_tc_field_ok(const PeObjectDesc& d) {
    return d.nerf_kind == PE_NERF_ADAIN && d.width == 256 && d.layers == 8 && d.skip == 4 && d.octaves == 10 && d.features == 192 &&
           d.positions >= 1 && d.positions <= 128;
}
// ... fully fused (sampling inside the kernel): objects without a ray bender
__host__ __device__ inline bool pe_tc_shape_ok(const PeObjectDesc& d) { return pe_tc_field_ok(d) && d.bender_kind == PE_BENDER_ZEROED; }
// ... behind the sampling / ray-bender pre-pass (pe_field_fp32.cu, phase PE_PHASE_PREPASS): objects with a positional ray bender
__host__ __device__ inline bool pe_tc_prepass_ok(const PeObjectDesc& d) { return pe_tc_field_ok(d) && d.bender_kind == PE_BENDER_POSITIONAL; }
#define PE_PHASE_PREPASS 3                    // fp32 field kernel: sampling + ray bender only (bent positions, flags, displacements)
// Ray-bender shape handled by the tcgen05 bender kernel: the shipped one (configs/tennis/193_*.yaml:165-178: 6 x 128, skip at 3,
// 6 octaves, 32 deformation features)
__host__ __device__ inline bool pe_tc_bender_ok(const PeObjectDesc& d) {
    return d.bender_kind == PE_BENDER_POSITIONAL && d.b_width == 128 && d.b_layers == 6 && d.b_skip == 3 && d.b_octaves == 6 &&
           d.deformation_features == 32;
}
// bytes of one weight pass of the bender stream: L0 (K=96): 3 slabs; L1,2,4,5: 4; L3 (K=224): 7; out (N=16): 4; 6 bias slabs
__host__ __device__ inline int64_t pe_tcb_pass_bytes() { return (3LL + 16 + 7) * 128 * 32 * 2 + 4LL * 16 * 32 * 2 + 6LL * 128 * 32; }

// Transposed streams (backward): per chain step an [N' = layer inputs][K' = layer outputs] operand in K'=32 slabs.
// field:  H6T (N'128,K'192) H3T (256,128) H0T L7T L6T L5T (256,256) L4encT (64,256) L4T L3T L2T L1T (256,256) L0T (64,256)
__host__ __device__ inline int64_t pe_tcT_pass_bytes() { return 6LL * 128 * 64 + 4LL * 256 * 64 + 8LL * 8 * 256 * 64 + 2LL * 8 * 64 * 64; }
// bender: OUTT (N'128,K'32) L5T L4T (128,128) L3encT (96,128) L3T L2T L1T (128,128) L0T (96,128)
__host__ __device__ inline int64_t pe_tcbT_pass_bytes() { return 1LL * 128 * 64 + 5LL * 4 * 128 * 64 + 2LL * 4 * 96 * 64; }

#define PE_TC_AWARE_MASK 0x0C0                // ... with an activation-aware weight stream and >= 96 samples per ray
#define PE_TC_MIXED_MASK 0x0F8                // two-pass layers of the mixed mode: trunk layers L3-L7 (the head gains nothing, profiles/r2_mixed_mode.md)
#define PE_TC_SLAB_K 32                       // K elements per streamed weight slab
// number of K=32 slabs of one weight pass of the shipped field:
// L0: 64/32=2; L1-3: 8 each; L4: 320/32=10; L5-7: 8 each; H0: 8; H3 (N=128): 8; H6 (K=128,N=192): 4
__host__ __device__ inline int64_t pe_tc_pass_bytes() {
    int64_t b = 0;
    b += 2LL * 256 * 32 * 2;                  // L0
    b += 6LL * 8 * 256 * 32 * 2;              // L1-3, L5-7
    b += 10LL * 256 * 32 * 2;                 // L4
    b += 8LL * 256 * 32 * 2;                  // H0
    b += 8LL * 128 * 32 * 2;                  // H3
    b += 4LL * 192 * 32 * 2;                  // H6
    b += 8LL * 256 * 32 + 192 * 32;           // bias slabs (N x 16 fp16) of L0-L7 and H6
    return b;
}

__host__ __device__ inline PeLayout pe_layout(const PeObjectDesc& d) {
    PeLayout L = {};
    int64_t off = 0;
    auto take = [&](int64_t floats) { int64_t o = off; off = pe_align_up(off + floats * 4, 256); return o; };
    const int W = d.width, F = d.features, S = d.style_features;
    L.in_dims = d.nerf_kind == PE_NERF_SKYBOX_V3 ? 6 : 3;
    L.enc = L.in_dims * (1 + 2 * d.octaves);
    int cur = L.enc;
    for (int l = 0; l < d.layers; ++l) {
        if (l == d.skip) cur += L.enc;
        L.k_in[l] = cur;
        L.bb_w[l] = take((int64_t)cur * W);
        L.bb_b[l] = take(W);
        cur = W;
    }
    L.alpha_w = take(W); L.alpha_b = take(1);
    L.head0_w = take((int64_t)W * W);
    L.head3_w = take((int64_t)W * (W / 2));
    L.head6_w = take((int64_t)(W / 2) * F); L.head6_b = take(F);
    L.aff1_w = take((int64_t)2 * W * S); L.aff1_b = take(2 * W); L.bn1_mean = take(W); L.bn1_var = take(W);
    L.aff2_w = take((int64_t)W * S); L.aff2_b = take(W); L.bn2_mean = take(W / 2); L.bn2_var = take(W / 2);
    if (d.bender_kind == PE_BENDER_POSITIONAL) {
        L.b_enc = 3 * (1 + 2 * d.b_octaves) + d.deformation_features;
        cur = L.b_enc;
        for (int l = 0; l < d.b_layers; ++l) {
            if (l == d.b_skip) cur += L.b_enc;
            L.b_k_in[l] = cur;
            L.bd_w[l] = take((int64_t)cur * d.b_width);
            L.bd_b[l] = take(d.b_width);
            cur = d.b_width;
        }
        L.bd_out_w = take((int64_t)d.b_width * 3);
    }
    L.tc_supported = pe_tc_field_ok(d) ? 1 : 0;
    if (L.tc_supported) {
        L.tc_bytes_per_pass = pe_tc_pass_bytes();
        L.tc_base = off;
        off = pe_align_up(off + 2 * L.tc_bytes_per_pass, 256);
    }
    if (L.tc_supported && pe_tc_bender_ok(d)) {
        L.tcb_bytes_per_pass = pe_tcb_pass_bytes();
        L.tcb_base = off;
        off = pe_align_up(off + 2 * L.tcb_bytes_per_pass, 256);
    }
    if (L.tc_supported) {
        L.tcT_bytes_per_pass = pe_tcT_pass_bytes();
        L.tcT_base = off;
        off = pe_align_up(off + 2 * L.tcT_bytes_per_pass, 256);
        if (pe_tc_bender_ok(d)) {
            L.tcbT_bytes_per_pass = pe_tcbT_pass_bytes();
            L.tcbT_base = off;
            off = pe_align_up(off + 2 * L.tcbT_bytes_per_pass, 256);
        }
    }
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------------------
// geometry: IEEE-exact (no FMA contraction) so that in-box decisions agree with the fp32 CPU path
// ------------------------------------------------------------------------------------------------
struct PeRay {           // one ray in the coordinate system of one object
    float o[3], d[3];
    float z_near, z_far; // after the slab test and the [z_near_min, z_far_max] clamp
};

// RayHelper.transform_points (utils/lib_3d/ray_helper.py:1180-1201): sum_b M[a][b]*p[b] (+ M[a][3]).
__device__ __forceinline__ void pe_transform(const float* __restrict__ m34, const float p[3], bool translate, float out[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float s = __fmul_rn(p[0], m34[a * 4 + 0]);
        s = __fadd_rn(s, __fmul_rn(p[1], m34[a * 4 + 1]));
        s = __fadd_rn(s, __fmul_rn(p[2], m34[a * 4 + 2]));
        out[a] = translate ? __fadd_rn(s, m34[a * 4 + 3]) : s;
    }
}

// ObjectComposer.compute_raywise_object_z_bounds + clamps (model/object_composer.py:104-151, 522-523).
__device__ __forceinline__ void pe_z_bounds(const PeObjectDesc& ob, bool in_scene, PeRay& ray) {
    const float eps = 1e-6f;
    float z_near = -INFINITY, z_far = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float den = __fadd_rn(ray.d[a], eps);
        const float z0 = __fdiv_rn(__fsub_rn(ob.bbox[2 * a + 0], ray.o[a]), den);
        const float z1 = __fdiv_rn(__fsub_rn(ob.bbox[2 * a + 1], ray.o[a]), den);
        z_near = fmaxf(z_near, fminf(z0, z1));
        z_far = fminf(z_far, fmaxf(z0, z1));
    }
    if (z_far <= z_near || !in_scene) { z_near = 0.f; z_far = 0.f; }
    ray.z_near = fminf(fmaxf(z_near, ob.z_near_min), ob.z_far_max);
    ray.z_far = fminf(fmaxf(z_far, ob.z_near_min), ob.z_far_max);
}

// torch.linspace(0, 1, P)[p] (ATen's symmetric formula), utils/lib_3d/ray_helper.py:1253.
__device__ __forceinline__ float pe_linspace01(int p, int P) {
    if (P == 1) return 0.f;
    const float step = __fdiv_rn(1.f, (float)(P - 1));
    return p < P / 2 ? __fmul_rn(step, (float)p) : __fsub_rn(1.f, __fmul_rn(step, (float)(P - p - 1)));
}

__device__ __forceinline__ float pe_uniform_t(const PeRay& ray, int p, int P) {
    const float s = pe_linspace01(p, P);
    return __fadd_rn(__fmul_rn(ray.z_near, __fsub_rn(1.f, s)), __fmul_rn(ray.z_far, s));
}

// RayHelper.create_ray_positions (utils/lib_3d/ray_helper.py:1229-1282); `u` replaces torch.rand.
__device__ __forceinline__ float pe_sample_t(const PeRay& ray, int p, int P, bool perturb, float u) {
    const float t = pe_uniform_t(ray, p, P);
    if (!perturb) return t;
    const float lower = p == 0 ? t : __fmul_rn(__fadd_rn(t, pe_uniform_t(ray, p - 1, P)), 0.5f);
    const float upper = p == P - 1 ? t : __fmul_rn(__fadd_rn(pe_uniform_t(ray, p + 1, P), t), 0.5f);
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
}

// Hierarchical ("fine") pass: the caller supplies the ray parameters (RayHelper.create_ray_positions_weighted, ray_helper.py:1320-1347:
// coarse samples merged with inverse-CDF samples of the coarse weights, sorted) -- `t_in` [images][rays][P] replaces the stratified samples.
__device__ __forceinline__ float pe_sample_t_or(const float* __restrict__ t_in, int64_t gs, const PeRay& ray, int p, int P, bool perturb, float u) {
    return t_in ? t_in[gs] : pe_sample_t(ray, p, P, perturb, u);
}

__device__ __forceinline__ void pe_position(const PeRay& ray, float t, float x[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(ray.o[a], __fmul_rn(ray.d[a], t));
}

__device__ __forceinline__ float pe_sincos_feature(float x, int fn) { return fn ? cosf(x) : sinf(x); }

// Fourier features, layout of model/positional_encoder.py:41-65: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...],
// each block `dims` wide; optional per-octave weight (annealable_positional_encoder.py:54-76).
__device__ __forceinline__ float pe_encoding_value(const float* x, int dims, int e, const float* anneal) {
    if (e < dims) return x[e];
    const int q = e - dims;
    const int oct = q / (2 * dims);
    const int rem = q - oct * 2 * dims;
    const int fn = rem / dims;
    const int dim = rem - fn * dims;
    float v = pe_sincos_feature(__fmul_rn(exp2f((float)oct), x[dim]), fn);
    if (anneal) v = __fmul_rn(v, anneal[oct]);
    return v;
}

// compute_bounding_box_filtering_mask (ray_bending_style_nerf_model.py:62-85): inclusive bounds.
__device__ __forceinline__ bool pe_in_box(const PeObjectDesc& ob, const float x[3]) {
    return x[0] >= ob.bbox[0] && x[0] <= ob.bbox[1] && x[1] >= ob.bbox[2] && x[1] <= ob.bbox[3] &&
           x[2] >= ob.bbox[4] && x[2] <= ob.bbox[5];
}

// Build the object-space ray (transform_rays, utils/lib_3d/ray_helper.py:1203-1227).
__device__ __forceinline__ PeRay pe_make_ray(const PeObjectDesc& ob, const float* __restrict__ m34,
                                             const float* __restrict__ origin_w, const float* __restrict__ dir_w, bool in_scene) {
    PeRay ray;
    const float ow[3] = {origin_w[0], origin_w[1], origin_w[2]};
    const float dw[3] = {dir_w[0], dir_w[1], dir_w[2]};
    pe_transform(m34, ow, true, ray.o);
    pe_transform(m34, dw, false, ray.d);
    pe_z_bounds(ob, in_scene, ray);
    return ray;
}

// ------------------------------------------------------------------------------------------------------
// compose (model/object_composer.py:399-447) sorts the concatenation of the objects' sample lists by t; ties keep the concatenation
// order.  Position of every entry in that order, for the warp that owns the ray: `ut` = the concatenated t values (shared memory, n
// entries, object k at [start_k, start_k + positions[k])).  Each object's list is non-decreasing by construction (stratified or merged
// samples), so an entry's position is its own index in its list plus a binary search in every other list -- upper bound in the lists
// before it (ties go first there), lower bound in the lists after it.  A ray whose lists are not ordered (samples masked by
// fix_object_overlaps get t = 0; a NaN) takes the all-pairs count, which is the definition.  Both give the same permutation.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pe_lists_ordered(const float* __restrict__ ut, int n, const int32_t* __restrict__ positions, int objects, int lane) {
    bool ok = true;
    int start = 0;
    for (int k = 0; k < objects; ++k) {
        const int P = positions[k];
        for (int p = 1 + lane; p < P; p += 32) ok = ok && (ut[start + p] >= ut[start + p - 1]);
        if (lane == 0 && P > 0) ok = ok && (ut[start] == ut[start]);           // a single NaN entry
        start += P;
    }
    return __all_sync(0xffffffffu, ok);
}

__device__ __forceinline__ int pe_compose_rank(const float* __restrict__ ut, int n, int j, const int32_t* __restrict__ positions, int objects,
                                               bool ordered) {
    const float tj = ut[j];
    int rank = 0;
    if (!ordered) {
        for (int m = 0; m < n; ++m) {
            const float tm = ut[m];
            rank += (tm < tj || (tm == tj && m < j)) ? 1 : 0;
        }
        return rank;
    }
    int start = 0;
    for (int k = 0; k < objects; ++k) {
        const int P = positions[k];
        if (j >= start && j < start + P) {
            rank += j - start;
        } else {
            const bool before = start < j;                 // the whole list precedes j in the concatenation: ties count
            const float* __restrict__ base = ut + start;
            int lo = 0, len = P;
            while (len > 0) {
                const int half = len >> 1;
                const float v = base[lo + half];
                const bool right = before ? (v <= tj) : (v < tj);
                if (right) { lo += half + 1; len -= half + 1; } else len = half;
            }
            rank += lo;
        }
        start += P;
    }
    return rank;
}

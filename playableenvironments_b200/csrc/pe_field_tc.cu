// Fused ray-march kernel for the shipped field (8x256 trunk, skip at 4, 10 octaves, 192 features, AdaIn head)
// on the 5th-generation tensor cores: persistent, warp-specialised, one CTA per SM.
//
//   per CTA iteration: two tiles (X, Y) of 128 sample slots each
//   epilogue warps  : sample positions -> Fourier encoding -> fp16 A operand in shared memory;
//                     per layer: TMEM accumulators -> +bias/AdaIn -> ReLU -> fp16 -> next A operand (in place);
//                     last layer: alpha compositing of the tile's rays, only per-ray results reach HBM
//   MMA warp        : one elected thread issues tcgen05.mma (M=128, N<=256, K=16), both tiles share every
//                     weight slab, accumulators live in TMEM (2 x 256 columns)
//   producer warp   : streams the pre-packed fp16 weight slabs (UMMA canonical layout) with bulk async copies
//                     into a 4-stage ring; weights stay L2 resident, each slab feeds 256 rows
//
// Replaces, for one object: RayHelper.transform_rays / create_ray_positions (utils/lib_3d/ray_helper.py:1203-1282),
// compute_raywise_object_z_bounds (model/object_composer.py:104-151), RayBendingStyleNerfModel.forward with a zeroed
// bender (model/nerf_models/ray_bending_style_nerf_model.py:137-219), AdaInStyleNerfModel.compute_network_pass
// (model/nerf_models/adain_style_nerf_model.py:106-145), PositionalEncoder.forward (model/positional_encoder.py:41-65)
// and ObjectComposer.integrate (model/object_composer.py:724-784).
#include "pe_kernels.cuh"
#include "pe_umma.cuh"

namespace {
using namespace pe;

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 2048;                 // 8 K-columns of a 128-row operand: 16 row groups x 128 B
constexpr int A_CHUNKS = 40;                      // K columns 0..255: activations, 256..319: positional encoding
constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES;   // 80 KB per tile
constexpr int PE_CHUNK0 = 32;
constexpr int STAGE_BYTES = 16384;                // largest slab: 256 rows x 32 k x 2 B
constexpr int NUM_STAGES = 4;
constexpr int NUM_LAYERS = 11;                    // L0..L7, H0, H3, H6
constexpr int THREADS = 384;                      // producer, MMA, TMEM-alloc, spare + 2 x 4 epilogue warps
constexpr int SCRATCH_STRIDE = 97;                // floats per row of the compositing scratch (bank-conflict free)
constexpr int SCR_T = 50 * 1024, SCR_SH = SCR_T + 512, SCR_W = SCR_SH + 512;   // byte offsets inside the A buffer
constexpr int SMEM_BAR = 2 * A_BYTES + NUM_STAGES * STAGE_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 128;

__device__ __forceinline__ void layer_spec(int l, int& n, int& slabs, int& chunk0) {
    n = 256; slabs = 8; chunk0 = 0;
    if (l == 0) { slabs = 2; chunk0 = PE_CHUNK0; }
    else if (l == 4) { slabs = 10; }
    else if (l == 9) { n = 128; }
    else if (l == 10) { n = 192; slabs = 4; }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// store 8 consecutive K values of row `m` into the K-major no-swizzle operand (chunk = K/8)
__device__ __forceinline__ void store_a8(unsigned char* a_base, int chunk, int m, const float* v) {
    uint4 q;
    q.x = pack_half2(v[0], v[1]); q.y = pack_half2(v[2], v[3]); q.z = pack_half2(v[4], v[5]); q.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(a_base + chunk * CHUNK_BYTES + m * 16) = q;
}

struct RowState {        // what an epilogue thread remembers about its sample between layers
    float t, raw_alpha, dnorm;
    int64_t ray;         // global ray index (image * rays + r), -1 for padding rows
    int p;
    bool valid, inbox, in_scene;
};

__global__ void __launch_bounds__(THREADS, 1) pe_field_tc_kernel(const PeFieldArgs A, const PeIntegrated G2, const int num_passes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_buf[2] = {smem, smem + A_BYTES};
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;     // [2]
    uint64_t* a_ready = acc_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);

    const int P = ob.positions;
    const int rpt = TILE_M / P;                                   // rays per tile
    const int rows_used = rpt * P;
    const int tiles_per_image = (A.rays + rpt - 1) / rpt;
    const int64_t total_tiles = (int64_t)tiles_per_image * A.images;
    const int64_t total_pairs = (total_tiles + 1) / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(acc_full + g, 1); mbar_init(a_ready + g, TILE_M); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ weight producer ================================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
                const unsigned char* src = blob + L.tc_base;
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0;
                    layer_spec(l, n, slabs, chunk0);
                    const uint32_t bytes = (uint32_t)n * PE_TC_SLAB_K * 2;
                    for (int s = 0; s < slabs; ++s) {
                        for (int pass = 0; pass < num_passes; ++pass) {
                            mbar_wait(empty_bar + stage, phase ^ 1);
                            mbar_arrive_expect_tx(full_bar + stage, bytes);
                            bulk_copy_g2s(ring + stage * STAGE_BYTES, src + (int64_t)pass * L.tc_bytes_per_pass, bytes, full_bar + stage);
                            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        }
                        src += bytes;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0, ready_phase = 0;
            const uint32_t a_addr[2] = {smem_u32(a_buf[0]), smem_u32(a_buf[1])};
            const uint32_t ring_addr = smem_u32(ring);
            for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0;
                    layer_spec(l, n, slabs, chunk0);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, n);
                    const uint32_t lbo_b = (uint32_t)n * 16;          // bytes between K chunks of a slab: (n/8) core matrices
                    mbar_wait(a_ready + 0, ready_phase);
                    mbar_wait(a_ready + 1, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    for (int s = 0; s < slabs; ++s) {
                        for (int pass = 0; pass < num_passes; ++pass) {
                            mbar_wait(full_bar + stage, phase);
                            tc_fence_after();
                            const bool last = (s == slabs - 1) && (pass == num_passes - 1);
                            const uint32_t b_addr = ring_addr + stage * STAGE_BYTES;
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint32_t a_chunk = chunk0 + 4 * s + 2 * j;
                                    const uint64_t da = umma_smem_desc(a_addr[g] + a_chunk * CHUNK_BYTES, CHUNK_BYTES, 128);
                                    const uint64_t db = umma_smem_desc(b_addr + 2 * j * lbo_b, lbo_b, 128);
                                    umma_f16_ss(tmem_base + g * 256, da, db, idesc, (s | pass | j) != 0 ? 1u : 0u);
                                }
                                if (last) umma_commit(acc_full + g);
                            }
                            umma_commit(empty_bar + stage);
                            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue groups ================================
        const int g = (warp - 4) >> 2;                 // 0: tile X, 1: tile Y
        const int m = ((warp & 3) << 5) | lane;        // row of the tile == TMEM lane
        const uint32_t bar_id = 1 + g;
        unsigned char* abuf = a_buf[g];
        const uint32_t taddr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + g * 256;
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        const bool single = G2.integrated_features != nullptr || G2.opacity != nullptr || G2.weights != nullptr;   // this object IS the scene
        uint32_t acc_phase = 0;
        float* scr = reinterpret_cast<float*>(abuf);
        float* t_s = reinterpret_cast<float*>(abuf + SCR_T);
        float* sh_s = reinterpret_cast<float*>(abuf + SCR_SH);
        float* w_s = reinterpret_cast<float*>(abuf + SCR_W);

        for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
            const int64_t tile = pair * 2 + g;
            const bool tile_valid = tile < total_tiles;
            const int img = tile_valid ? (int)(tile / tiles_per_image) : 0;
            const int ray0 = tile_valid ? (int)(tile - (int64_t)img * tiles_per_image) * rpt : 0;

            // ---- sampling + positional encoding -> A columns 256..319 ----
            RowState st;
            st.valid = false; st.inbox = false; st.t = 0.f; st.raw_alpha = ob.empty_space_alpha; st.dnorm = 0.f; st.ray = -1; st.p = 0;
            st.in_scene = A.ois ? A.ois[(int64_t)img * A.objects + A.k] != 0 : true;
            float x[3] = {0.f, 0.f, 0.f};
            if (tile_valid && m < rows_used) {
                const int rl = m / P;
                const int r = ray0 + rl;
                if (r < A.rays) {
                    st.valid = true;
                    st.p = m - rl * P;
                    st.ray = (int64_t)img * A.rays + r;
                    const float* dw = A.dirs + st.ray * 3;
                    const PeRay ray = pe_make_ray(ob, A.w2o + ((int64_t)img * A.objects + A.k) * 12, A.origins + (int64_t)img * 3, dw, st.in_scene);
                    const float u = A.perturb ? A.rand[st.ray * P + st.p] : 0.f;
                    st.t = pe_sample_t(ray, st.p, P, A.perturb != 0, u);
                    pe_position(ray, st.t, x);
                    st.inbox = pe_in_box(ob, x);
                    st.dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dw[0], dw[0]), __fmul_rn(dw[1], dw[1])), __fmul_rn(dw[2], dw[2])));
                }
            }
            {
                const float xn[3] = {__fdiv_rn(x[0], size[0]), __fdiv_rn(x[1], size[1]), __fdiv_rn(x[2], size[2])};
                float enc[64];
                enc[0] = xn[0]; enc[1] = xn[1]; enc[2] = xn[2];
#pragma unroll
                for (int o = 0; o < 10; ++o) {
                    const float f = (float)(1 << o);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        float sn, cs;
                        sincosf(__fmul_rn(f, xn[a]), &sn, &cs);
                        enc[3 + 6 * o + a] = sn;
                        enc[3 + 6 * o + 3 + a] = cs;
                    }
                }
                enc[63] = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) store_a8(abuf, PE_CHUNK0 + c, m, enc + 8 * c);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(a_ready + g);

            // ---- the 11 tensor-core layers ----
            for (int l = 0; l < NUM_LAYERS; ++l) {
                mbar_wait(acc_full + g, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                if (l < 10) {
                    const int n = (l == 9) ? 128 : 256;
                    const float* bias = nullptr; const float* sc = nullptr; const float* sh = nullptr;
                    if (l < 8) bias = reinterpret_cast<const float*>(blob + L.bb_b[l]);
                    else if (l == 8) { sc = A.aff1 + (int64_t)img * 512; sh = sc + 256; }
                    else { sc = A.aff2 + (int64_t)img * 256; sh = sc + 128; }
                    const float* aw = reinterpret_cast<const float*>(blob + L.alpha_w);
                    float alpha_acc = 0.f;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c0, v);
                        tmem_wait_ld();
                        float y[32];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 b4, s4;
                            if (l < 8) {
                                b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + q);
                                y[4 * q + 0] = fmaxf(__uint_as_float(v[4 * q + 0]) + b4.x, 0.f);
                                y[4 * q + 1] = fmaxf(__uint_as_float(v[4 * q + 1]) + b4.y, 0.f);
                                y[4 * q + 2] = fmaxf(__uint_as_float(v[4 * q + 2]) + b4.z, 0.f);
                                y[4 * q + 3] = fmaxf(__uint_as_float(v[4 * q + 3]) + b4.w, 0.f);
                            } else {
                                s4 = __ldg(reinterpret_cast<const float4*>(sc + c0) + q);
                                b4 = __ldg(reinterpret_cast<const float4*>(sh + c0) + q);
                                y[4 * q + 0] = fmaxf(fmaf(__uint_as_float(v[4 * q + 0]), s4.x, b4.x), 0.f);
                                y[4 * q + 1] = fmaxf(fmaf(__uint_as_float(v[4 * q + 1]), s4.y, b4.y), 0.f);
                                y[4 * q + 2] = fmaxf(fmaf(__uint_as_float(v[4 * q + 2]), s4.z, b4.z), 0.f);
                                y[4 * q + 3] = fmaxf(fmaf(__uint_as_float(v[4 * q + 3]), s4.w, b4.w), 0.f);
                            }
                        }
                        if (l == 7) {           // alpha head (adain_style_nerf_model.py:138) in fp32 on the un-rounded trunk output
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 w4 = __ldg(reinterpret_cast<const float4*>(aw + c0) + q);
                                alpha_acc = fmaf(y[4 * q + 0], w4.x, alpha_acc);
                                alpha_acc = fmaf(y[4 * q + 1], w4.y, alpha_acc);
                                alpha_acc = fmaf(y[4 * q + 2], w4.z, alpha_acc);
                                alpha_acc = fmaf(y[4 * q + 3], w4.w, alpha_acc);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 32; ++q) y[q] = fminf(y[q], 65504.f);     // fp16 range
#pragma unroll
                        for (int c = 0; c < 4; ++c) store_a8(abuf, (c0 >> 3) + c, m, y + 8 * c);
                    }
                    if (l == 7) st.raw_alpha = alpha_acc + __ldg(reinterpret_cast<const float*>(blob + L.alpha_b));
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(a_ready + g);
                } else {
                    // ---- last layer: features in TMEM -> volume rendering of the tile's rays ----
                    const int64_t gs = st.valid ? st.ray * P + st.p : 0;
                    float raw = (st.inbox && st.in_scene) ? st.raw_alpha : ob.empty_space_alpha;
                    if (st.valid) {
                        if (A.raw_out) A.raw_out[gs] = raw;
                        if (A.t_out) A.t_out[gs] = st.t;
                        if (A.inbox_out) A.inbox_out[gs] = st.inbox ? 1 : 0;
                        if (A.dispmag_out) A.dispmag_out[gs] = 0.f;
                    }
                    t_s[m] = st.t;
                    named_bar_sync(bar_id, TILE_M);
                    float alpha = 0.f;
                    if (st.valid) {
                        const float delta = __fmul_rn(st.p == P - 1 ? 1e10f : __fsub_rn(t_s[m + 1], st.t), st.dnorm);
                        if (A.noise) raw = __fadd_rn(raw, A.noise[gs]);
                        alpha = __fsub_rn(1.f, expf(__fmul_rn(-fmaxf(raw, 0.f), delta)));
                    }
                    sh_s[m] = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
                    named_bar_sync(bar_id, TILE_M);
                    float T = 1.f;
                    for (int j = m - st.p; j < m; ++j) T *= sh_s[j];                  // exclusive cumprod (:207-212)
                    const float w = st.valid ? alpha * T : 0.f;
                    w_s[m] = w;
                    if (st.valid) {
                        if (A.integ.weights) A.integ.weights[gs] = w;
                        if (single && G2.weights) G2.weights[gs] = w;
                    }
                    const float wf = st.inbox ? w : 0.f;
                    const float* b6 = reinterpret_cast<const float*>(blob + L.head6_b);
                    for (int half = 0; half < 2; ++half) {
                        for (int c0 = 0; c0 < 96; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld32(taddr + half * 96 + c0, v);
                            tmem_wait_ld();
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                float f = __uint_as_float(v[q]) + __ldg(b6 + half * 96 + c0 + q);
                                if (A.apply_activation) f = 1.f / (1.f + expf(-f));
                                if (A.feat_out && st.valid) A.feat_out[gs * 192 + half * 96 + c0 + q] = st.inbox ? f : 0.f;
                                scr[m * SCRATCH_STRIDE + c0 + q] = wf * f;
                            }
                        }
                        named_bar_sync(bar_id, TILE_M);
                        for (int item = m; item < rpt * 96; item += TILE_M) {
                            const int rl = item / 96, c = item - rl * 96;
                            const int r = ray0 + rl;
                            if (tile_valid && r < A.rays) {
                                float s = 0.f;
                                for (int j = rl * P; j < rl * P + P; ++j) s += scr[j * SCRATCH_STRIDE + c];
                                const int64_t o = ((int64_t)img * A.rays + r) * 192 + half * 96 + c;
                                if (A.integ.integrated_features) A.integ.integrated_features[o] = s;
                                if (single && G2.integrated_features) G2.integrated_features[o] = s;
                            }
                        }
                        named_bar_sync(bar_id, TILE_M);
                    }
                    if (st.valid && st.p == 0) {          // per-ray scalars (:758-772)
                        float opacity = 0.f, depth = 0.f;
                        for (int j = m; j < m + P; ++j) { opacity += w_s[j]; depth += w_s[j] * t_s[j]; }
                        const float qd = depth / opacity;
                        const float disparity = 1.f / (qd != qd ? qd : fmaxf(qd, 1e-10f));
                        const PeIntegrated* outs[2] = {&A.integ, &G2};
                        for (int oi = 0; oi < (single ? 2 : 1); ++oi) {
                            const PeIntegrated& O = *outs[oi];
                            if (O.opacity) O.opacity[st.ray] = opacity;
                            if (O.depth) O.depth[st.ray] = depth;
                            if (O.disparity) O.disparity[st.ray] = disparity;
                            if (O.integrated_displacements_magnitude) O.integrated_displacements_magnitude[st.ray] = 0.f;
                            if (O.integrated_divergence) O.integrated_divergence[st.ray] = 0.f;
                        }
                    }
                    named_bar_sync(bar_id, TILE_M);       // scratch is dead before the next tile's encoding overwrites it
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// weight packing into the slab stream
// ------------------------------------------------------------------------------------------------------
// slab element (n, kk) lives at (kk/8)*(N*16) + (n/8)*128 + (n%8)*16 + (kk%8)*2  (K-major, no swizzle)
__global__ void pe_tc_pack_layer_kernel(const float* __restrict__ w, int N, int K_src, int K_pad, unsigned char* __restrict__ hi,
                                        unsigned char* __restrict__ lo) {
    const int64_t total = (int64_t)N * K_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / K_pad), k = (int)(i - (int64_t)n * K_pad);
        const float v = k < K_src ? w[(int64_t)n * K_src + k] : 0.f;
        const __half h = __float2half_rn(v);
        const __half r = __float2half_rn(v - __half2float(h));
        const int slab = k / PE_TC_SLAB_K, kk = k - slab * PE_TC_SLAB_K;
        const int64_t off = (int64_t)slab * N * PE_TC_SLAB_K * 2 + (int64_t)(kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = r;
    }
}

// D = A * B^T on one CTA through the same building blocks (validation of descriptors / TMEM addressing)
__global__ void __launch_bounds__(128, 1) pe_debug_umma_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int n, int k) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sa = smem;                         // 128 x k
    unsigned char* sb = smem + 128 * 256 * 2;         // n x k
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 256 * 2 + 256 * 256 * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = threadIdx.x;
    for (int c = 0; c < k / 8; ++c) {
        float v[8];
        for (int q = 0; q < 8; ++q) v[q] = a[(int64_t)m * k + c * 8 + q];
        store_a8(sa, c, m, v);
    }
    for (int64_t i = threadIdx.x; i < (int64_t)n * k; i += blockDim.x) {
        const int row = (int)(i / k), kk = (int)(i - (int64_t)row * k);
        const int64_t off = (int64_t)(kk >> 3) * (n * 16) + (row >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(sb + off) = __float2half_rn(b[i]);
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, n);
        const uint32_t lbo_b = (uint32_t)n * 16;
        for (int j = 0; j < k / 16; ++j) {
            const uint64_t da = umma_smem_desc(smem_u32(sa) + 2 * j * CHUNK_BYTES, CHUNK_BYTES, 128);
            const uint64_t db = umma_smem_desc(smem_u32(sb) + 2 * j * lbo_b, lbo_b, 128);
            umma_f16_ss(tmem_base, da, db, idesc, j != 0 ? 1u : 0u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < n; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (((uint32_t)warp * 32u) << 16) + c0, v);
        tmem_wait_ld();
        for (int q = 0; q < 32; ++q)
            if (c0 + q < n) d[(int64_t)m * n + c0 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
    (void)lane;
}

}  // namespace

int pe_tc_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream) {
    unsigned char* hi = (unsigned char*)packed + L.tc_base;
    unsigned char* lo = hi + L.tc_bytes_per_pass;
    struct Item { const float* w; int N, K_src, K_pad; };
    const Item items[NUM_LAYERS] = {
        {p.backbone_w[0], 256, 63, 64},   {p.backbone_w[1], 256, 256, 256}, {p.backbone_w[2], 256, 256, 256}, {p.backbone_w[3], 256, 256, 256},
        {p.backbone_w[4], 256, 319, 320}, {p.backbone_w[5], 256, 256, 256}, {p.backbone_w[6], 256, 256, 256}, {p.backbone_w[7], 256, 256, 256},
        {p.head0_w, 256, 256, 256},       {p.head3_w, 128, 256, 256},       {p.head6_w, 192, 128, 128}};
    int64_t off = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const Item& it = items[l];
        if (!it.w) { pe_set_error("missing parameter tensor for tensor-core layer %d", l); return PE_ERR_INVALID; }
        const int64_t total = (int64_t)it.N * it.K_pad;
        pe_tc_pack_layer_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(it.w, it.N, it.K_src, it.K_pad, hi + off, lo + off);
        PE_LAUNCH_CHECK("pe_tc_pack_layer_kernel");
        off += total * 2;
    }
    if (off != L.tc_bytes_per_pass) { pe_set_error("internal: tensor-core weight stream size mismatch"); return PE_ERR_INVALID; }
    return PE_OK;
}

int pe_launch_field_tc(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream) {
    if (!pe_tc_shape_ok(args.ob) || args.training || args.explicit_positions || args.phase != 0) {
        pe_set_error("tensor-core field kernel: unsupported configuration");
        return PE_ERR_UNSUPPORTED;
    }
    const int num_passes = args.precision == PE_PRECISION_FP16X2 ? 2 : 1;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    const int rpt = TILE_M / args.ob.positions;
    const int64_t tiles = (int64_t)((args.rays + rpt - 1) / rpt) * args.images;
    const int64_t pairs = (tiles + 1) / 2;
    if (pairs == 0) return PE_OK;
    const int grid = (int)pe_min64(pairs, sm_count);
    pe_field_tc_kernel<<<grid, THREADS, SMEM_TOTAL, stream>>>(args, global_out, num_passes);
    PE_LAUNCH_CHECK("pe_field_tc_kernel");
    return PE_OK;
}

extern "C" int pe_debug_umma_gemm(const float* a, const float* b, float* d, int32_t n, int32_t k, pe_stream_t stream) {
    if (n < 16 || n > 256 || n % 16 || k < 16 || k > 256 || k % 16) { pe_set_error("debug gemm: n,k multiples of 16 up to 256"); return PE_ERR_INVALID; }
    const int smem = 128 * 256 * 2 + 256 * 256 * 2 + 64;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_debug_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pe_debug_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, n, k);
    PE_LAUNCH_CHECK("pe_debug_umma_kernel");
    return PE_OK;
}

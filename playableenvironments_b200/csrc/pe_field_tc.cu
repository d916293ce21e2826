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
#include <stdlib.h>

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
constexpr int SMEM_ONES = SMEM_BAR + 128;         // 256-byte "ones" operand of the rank-1 bias update
constexpr int SMEM_TOTAL = SMEM_ONES + 256;
// per-tile constants staged in the (dead after L4) positional-encoding columns of the A buffer
constexpr int CST_BASE = PE_CHUNK0 * CHUNK_BYTES; // byte offset inside the A buffer
constexpr int CST_SC1 = 0, CST_SH1 = 256, CST_SC2 = 512, CST_SH2 = 640, CST_AW = 768, CST_FLOATS = 1024;

__device__ __forceinline__ void layer_spec(int l, int& n, int& slabs, int& chunk0, bool& has_bias) {
    n = 256; slabs = 8; chunk0 = 0; has_bias = true;
    if (l == 0) { slabs = 2; chunk0 = PE_CHUNK0; }
    else if (l == 4) { slabs = 10; }
    else if (l == 8) { has_bias = false; }
    else if (l == 9) { n = 128; has_bias = false; }
    else if (l == 10) { n = 192; slabs = 4; }
}

// relu + saturating conversion of two fp32 to packed fp16 (low half = a, high half = b)
__device__ __forceinline__ uint32_t relu_pack_half2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

__device__ __forceinline__ void store_a8_relu(unsigned char* a_base, int chunk, int m, const float* v) {
    uint4 q;
    q.x = relu_pack_half2(v[0], v[1]); q.y = relu_pack_half2(v[2], v[3]); q.z = relu_pack_half2(v[4], v[5]); q.w = relu_pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(a_base + chunk * CHUNK_BYTES + m * 16) = q;
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

#define PE_TS(i) do { if (ts_on) ts[(i)] = clock64(); } while (0)

struct RowState {        // what an epilogue thread remembers about its sample between layers
    float t, raw_alpha, dnorm;
    int64_t ray;         // global ray index (image * rays + r), -1 for padding rows
    int p;
    bool valid, inbox, in_scene;
};

// wait for outstanding tcgen05.ld; the registers are threaded through so no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld_regs(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// Epilogue of one hidden layer for row m: TMEM accumulators (bias already added by the rank-1 MMA) ->
// [AdaIn affine] -> ReLU -> fp16 -> A operand of the next layer (in place).
//   MODE 0: trunk layer            y = relu(acc)
//   MODE 1: trunk output (L7)      y = relu(acc), also accumulates the alpha head dot product in fp32
//   MODE 2: AdaIn layer            y = relu(acc * sc[c] + sh[c])   (BatchNorm folded into sc/sh, adain.py:58-59)
template <int MODE, int N>
__device__ __forceinline__ float hidden_epilogue(uint32_t taddr, unsigned char* abuf, int m, const float* __restrict__ c0s,
                                                 const float* __restrict__ c1s) {
    uint32_t v[2][32];
    float alpha = 0.f;
    tmem_ld32(taddr, v[0]);
#pragma unroll
    for (int c = 0; c < N / 32; ++c) {
        tmem_wait_ld_regs(v[c & 1]);
        if (c + 1 < N / 32) tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        float y[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = __uint_as_float(v[c & 1][q]);
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(c0s + c * 32 + 4 * q);      // smem broadcast
                alpha = fmaf(fmaxf(y[4 * q + 0], 0.f), w4.x, alpha);
                alpha = fmaf(fmaxf(y[4 * q + 1], 0.f), w4.y, alpha);
                alpha = fmaf(fmaxf(y[4 * q + 2], 0.f), w4.z, alpha);
                alpha = fmaf(fmaxf(y[4 * q + 3], 0.f), w4.w, alpha);
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 s4 = *reinterpret_cast<const float4*>(c0s + c * 32 + 4 * q);
                const float4 b4 = *reinterpret_cast<const float4*>(c1s + c * 32 + 4 * q);
                y[4 * q + 0] = fmaf(y[4 * q + 0], s4.x, b4.x);
                y[4 * q + 1] = fmaf(y[4 * q + 1], s4.y, b4.y);
                y[4 * q + 2] = fmaf(y[4 * q + 2], s4.z, b4.z);
                y[4 * q + 3] = fmaf(y[4 * q + 3], s4.w, b4.w);
            }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) store_a8_relu(abuf, c * 4 + cc, m, y + 8 * cc);
    }
    return alpha;
}

__global__ void __launch_bounds__(THREADS, 1) pe_field_tc_kernel(const PeFieldArgs A, const PeIntegrated G2, const int num_passes, const int dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_buf[2] = {smem, smem + A_BYTES};
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;     // [2]
    uint64_t* a_ready = acc_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);
    unsigned char* ones = smem + SMEM_ONES;

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
    if (threadIdx.x < 128) {
        // "ones" operand: one 8x8 core matrix whose rows are [1,1,0,...] (hi and lo bias terms) + one zero core matrix;
        // the descriptor replicates it over all 16 row groups with SBO = 0
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        reinterpret_cast<__half*>(ones)[threadIdx.x] = __float2half_rn((r < 8 && c < 2) ? 1.f : 0.f);
    }
    fence_proxy_async();
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
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
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
                    if (has_bias) {          // [bias_hi | bias_lo | 0 ...] x 16 K columns
                        const uint32_t bbytes = (uint32_t)n * 32;
                        mbar_wait(empty_bar + stage, phase ^ 1);
                        mbar_arrive_expect_tx(full_bar + stage, bbytes);
                        bulk_copy_g2s(ring + stage * STAGE_BYTES, src, bbytes, full_bar + stage);
                        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        src += bbytes;
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
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
                for (int l = 0; l < NUM_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
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
                            const bool last = !has_bias && (s == slabs - 1) && (pass == num_passes - 1);
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
                    if (has_bias) {          // D += ones(128x16) * [bias_hi | bias_lo | 0..]^T : the bias, at fp32-class accuracy
                        mbar_wait(full_bar + stage, phase);
                        tc_fence_after();
                        const uint64_t db = umma_smem_desc(ring_addr + stage * STAGE_BYTES, lbo_b, 128);
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            umma_f16_ss(tmem_base + g * 256, ones_desc, db, idesc, 1u);
                            umma_commit(acc_full + g);
                        }
                        umma_commit(empty_bar + stage);
                        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue groups ================================
        const int g = (warp - 4) >> 2;                 // 0: tile X, 1: tile Y
        const int m = ((warp & 3) << 5) | lane;        // row of the tile == TMEM lane
        const uint32_t bar_id = 1 + g;
        unsigned char* abuf = smem + g * A_BYTES;      // derived from the __shared__ base so accesses compile to LDS/STS
        const uint32_t taddr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + g * 256;
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        const bool single = G2.integrated_features != nullptr || G2.opacity != nullptr || G2.weights != nullptr;   // this object IS the scene
        uint32_t acc_phase = 0;
        float* scr = reinterpret_cast<float*>(abuf);
        float* t_s = reinterpret_cast<float*>(abuf + SCR_T);
        float* sh_s = reinterpret_cast<float*>(abuf + SCR_SH);
        float* w_s = reinterpret_cast<float*>(abuf + SCR_W);
        float* cst = reinterpret_cast<float*>(abuf + CST_BASE);
        const float alpha_bias = __ldg(reinterpret_cast<const float*>(blob + L.alpha_b));
        const float* alpha_w = reinterpret_cast<const float*>(blob + L.alpha_w);

        long long* ts = reinterpret_cast<long long*>(A.stats);
        int iter = 0;
        for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x, ++iter) {
            const bool ts_on = (dbg & 8) && blockIdx.x == 1 && m == 0 && g == 0 && iter == 3;
            PE_TS(0);
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
                // Fourier features sin/cos(2^o * x): the argument is reduced EXACTLY (x/(2 pi) as a two-float value, scaled by the
                // power of two, integer part dropped), then evaluated with the SFU on [-pi, pi] (abs error < 5e-7, far below the
                // fp16 rounding of the operand).  Same values as positional_encoder.py:59-64 up to that error.
                const float xn[3] = {__fdiv_rn(x[0], size[0]), __fdiv_rn(x[1], size[1]), __fdiv_rn(x[2], size[2])};
                float enc[64];
                enc[0] = xn[0]; enc[1] = xn[1]; enc[2] = xn[2];
                float tp[3], tl[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float c_hi = 0.15915494f, c_lo = 6.4206382e-9f;      // 1/(2 pi) = c_hi + c_lo
                    tp[a] = xn[a] * c_hi;
                    tl[a] = fmaf(xn[a], c_lo, fmaf(xn[a], c_hi, -tp[a]));
                }
#pragma unroll
                for (int o = 0; o < 10; ++o) {
                    const float f = (float)(1 << o);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        float sn = 0.f, cs = 1.f;
                        if (!(dbg & 1)) {
                            const float turns = tp[a] * f;                           // exact (power of two)
                            const float fr = (turns - rintf(turns)) + tl[a] * f;     // fractional turns in [-0.5, 0.5]
                            const float ang = fr * 6.2831855f;
                            sn = __sinf(ang);
                            cs = __cosf(ang);
                        }
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
            PE_TS(1);

            // ---- the 10 hidden tensor-core layers ----
            float4 pre0 = make_float4(0.f, 0.f, 0.f, 0.f), pre1 = pre0;
            for (int l = 0; l < 10; ++l) {
                mbar_wait(acc_full + g, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                PE_TS(2 + 2 * l);
                if (l == 4) {
                    // the encoding columns are dead once L4 has run: reuse them for the constants of the later epilogues
                    // (AdaIn scale/shift of this image and the alpha-head weights); loads overlap this layer's epilogue
                    const float* a1 = A.aff1 + (int64_t)img * 512;
                    const float* a2 = A.aff2 + (int64_t)img * 256;
                    const int i0 = m * 4, i1 = 512 + m * 4;                       // 1024 floats, 8 per thread
                    pre0 = __ldg(reinterpret_cast<const float4*>(a1 + i0));      // sc1|sh1
                    pre1 = i1 < 768 ? __ldg(reinterpret_cast<const float4*>(a2 + (i1 - 512))) : __ldg(reinterpret_cast<const float4*>(alpha_w + (i1 - 768)));
                }
                if (l == 7) named_bar_sync(bar_id, TILE_M);                        // constants written by the whole group at l == 4
                if (dbg & 4) { /* timing experiment: no epilogue work */ }
                else if (l < 7) hidden_epilogue<0, 256>(taddr, abuf, m, nullptr, nullptr);
                else if (l == 7) st.raw_alpha = hidden_epilogue<1, 256>(taddr, abuf, m, cst + CST_AW, nullptr) + alpha_bias;
                else if (l == 8) hidden_epilogue<2, 256>(taddr, abuf, m, cst + CST_SC1, cst + CST_SH1);
                else hidden_epilogue<2, 128>(taddr, abuf, m, cst + CST_SC2, cst + CST_SH2);
                if (l == 4) {
                    *reinterpret_cast<float4*>(cst + m * 4) = pre0;
                    *reinterpret_cast<float4*>(cst + 512 + m * 4) = pre1;
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(a_ready + g);
                PE_TS(3 + 2 * l);
            }
            {
                // ---- last layer: features in TMEM -> volume rendering of the tile's rays ----
                mbar_wait(acc_full + g, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                PE_TS(22);
                const int64_t gs = st.valid ? st.ray * P + st.p : 0;
                float raw = (st.inbox && st.in_scene) ? st.raw_alpha : ob.empty_space_alpha;
                if (dbg & 2) { tc_fence_before(); continue; }    // timing experiment: no compositing
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
                // exclusive cumprod of (1 - alpha + 1e-10) along the samples of each ray (compute_weights :199-214):
                // segmented warp scan + carry across the warps a ray spans
                const float shifted = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
                const bool head = st.p == 0;
                float incl = shifted;
                bool closed = head;                              // a segment head lies in [first lane of the scan window, lane]
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, incl, d);
                    const bool fu = __shfl_up_sync(0xffffffffu, closed ? 1 : 0, d) != 0;
                    if (lane >= d && !closed) { incl *= up; closed = fu; }
                }
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0 || head) excl = 1.f;
                const int wq = warp & 3;
                if (lane == 31) { sh_s[wq] = incl; sh_s[4 + wq] = closed ? 1.f : 0.f; }
                named_bar_sync(bar_id, TILE_M);
                // lanes before the first head of their warp continue a ray started in an earlier warp
                const bool open = !closed;
                float T = excl;
                if (open) {
                    for (int v = wq - 1; v >= 0; --v) {
                        T *= sh_s[v];
                        if (sh_s[4 + v] != 0.f) break;
                    }
                }
                const float w = st.valid ? alpha * T : 0.f;
                w_s[m] = w;
                if (st.valid) {
                    if (A.integ.weights) A.integ.weights[gs] = w;
                    if (single && G2.weights) G2.weights[gs] = w;
                }
                const float wf = st.inbox ? w : 0.f;
                PE_TS(23);
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[2][32];
                    tmem_ld32(taddr + half * 96, v[0]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        tmem_wait_ld_regs(v[c & 1]);
                        if (c + 1 < 3) tmem_ld32(taddr + half * 96 + (c + 1) * 32, v[(c + 1) & 1]);
                        if (!A.apply_activation && !A.feat_out) {                  // common case: nothing per-sample leaves the SM
#pragma unroll
                            for (int q = 0; q < 32; ++q) scr[m * SCRATCH_STRIDE + c * 32 + q] = wf * __uint_as_float(v[c & 1][q]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                float f = __uint_as_float(v[c & 1][q]);            // head-6 bias already added by the rank-1 MMA
                                if (A.apply_activation) f = 1.f / (1.f + expf(-f));
                                if (A.feat_out && st.valid) A.feat_out[gs * 192 + half * 96 + c * 32 + q] = st.inbox ? f : 0.f;
                                scr[m * SCRATCH_STRIDE + c * 32 + q] = wf * f;
                            }
                        }
                    }
                    PE_TS(24 + 3 * half);
                    named_bar_sync(bar_id, TILE_M);
                    PE_TS(25 + 3 * half);
                    for (int item = m; item < rpt * 96; item += TILE_M) {
                        const int rl = item / 96, c = item - rl * 96;
                        const int r = ray0 + rl;
                        if (tile_valid && r < A.rays) {
                            const float* col = scr + rl * P * SCRATCH_STRIDE + c;
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                            int j = 0;
                            for (; j + 4 <= P; j += 4) {
                                s0 += col[(j + 0) * SCRATCH_STRIDE]; s1 += col[(j + 1) * SCRATCH_STRIDE];
                                s2 += col[(j + 2) * SCRATCH_STRIDE]; s3 += col[(j + 3) * SCRATCH_STRIDE];
                            }
                            for (; j < P; ++j) s0 += col[j * SCRATCH_STRIDE];
                            const float sum = (s0 + s1) + (s2 + s3);
                            const int64_t o = ((int64_t)img * A.rays + r) * 192 + half * 96 + c;
                            if (A.integ.integrated_features) A.integ.integrated_features[o] = sum;
                            if (single && G2.integrated_features) G2.integrated_features[o] = sum;
                        }
                    }
                    named_bar_sync(bar_id, TILE_M);
                    PE_TS(26 + 3 * half);
                }
                // per-ray scalars (:758-772): one warp per ray, lanes stride the samples
                for (int rl = wq; rl < rpt; rl += 4) {
                    const int r = ray0 + rl;
                    if (!tile_valid || r >= A.rays) continue;
                    float opacity = 0.f, depth = 0.f;
                    for (int j = lane; j < P; j += 32) { const float wj = w_s[rl * P + j]; opacity += wj; depth += wj * t_s[rl * P + j]; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        opacity += __shfl_xor_sync(0xffffffffu, opacity, o);
                        depth += __shfl_xor_sync(0xffffffffu, depth, o);
                    }
                    if (lane == 0) {
                        const int64_t gr = (int64_t)img * A.rays + r;
                        const float qd = depth / opacity;
                        const float disparity = 1.f / (qd != qd ? qd : fmaxf(qd, 1e-10f));
                        const PeIntegrated* outs[2] = {&A.integ, &G2};
                        for (int oi = 0; oi < (single ? 2 : 1); ++oi) {
                            const PeIntegrated& O = *outs[oi];
                            if (O.opacity) O.opacity[gr] = opacity;
                            if (O.depth) O.depth[gr] = depth;
                            if (O.disparity) O.disparity[gr] = disparity;
                            if (O.integrated_displacements_magnitude) O.integrated_displacements_magnitude[gr] = 0.f;
                            if (O.integrated_divergence) O.integrated_divergence[gr] = 0.f;
                        }
                    }
                }
                PE_TS(30);
                tc_fence_before();
                named_bar_sync(bar_id, TILE_M);       // scratch is dead before the next tile's encoding overwrites it
                PE_TS(31);
                if (ts_on) {
                    printf("PE_TS pe=%lld", ts[1] - ts[0]);
                    for (int i = 0; i < 10; ++i) printf(" | L%d wait=%lld epi=%lld", i, ts[2 + 2 * i] - ts[1 + 2 * i], ts[3 + 2 * i] - ts[2 + 2 * i]);
                    printf(" | L10 wait=%lld weights=%lld h0: ld+sts=%lld bar=%lld sum=%lld h1: ld+sts=%lld bar=%lld sum=%lld scalars=%lld endbar=%lld total=%lld\n",
                           ts[22] - ts[21], ts[23] - ts[22], ts[24] - ts[23], ts[25] - ts[24], ts[26] - ts[25], ts[27] - ts[26], ts[28] - ts[27],
                           ts[29] - ts[28], ts[30] - ts[29], ts[31] - ts[30], ts[31] - ts[0]);
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
__device__ __forceinline__ int64_t slab_offset(int N, int n, int k) {
    const int slab = k / PE_TC_SLAB_K, kk = k - slab * PE_TC_SLAB_K;
    return (int64_t)slab * N * PE_TC_SLAB_K * 2 + (int64_t)(kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
}

// One thread per output row.  hi = fp16(w) with ZERO-SUM rounding: walking along K, each weight is rounded to the fp16
// neighbour (down or up) that keeps the running sum of rounding errors of the row closest to zero.  Post-ReLU activations
// have a large common positive mean, so the systematic part sum_k dW[n][k] * mean(a) of the single-pass error cancels
// (measured: -10..-30 % error on the rendered frame); lo = fp16(w - hi) is the second pass of the fp16x2 mode.
__global__ void pe_tc_pack_layer_kernel(const float* __restrict__ w, int N, int K_src, int K_pad, unsigned char* __restrict__ hi,
                                        unsigned char* __restrict__ lo) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float run = 0.f;
    for (int k = 0; k < K_pad; ++k) {
        const float v = k < K_src ? w[(int64_t)n * K_src + k] : 0.f;
        const __half near = __float2half_rn(v);
        const float fn = __half2float(near);
        __half other = near;
        if (fn != v) other = fn < v ? __float2half_ru(v) : __float2half_rd(v);
        const float e_near = fn - v, e_other = __half2float(other) - v;
        const bool pick_other = fabsf(run + e_other) < fabsf(run + e_near);
        const __half h = pick_other ? other : near;
        run += pick_other ? e_other : e_near;
        const int64_t off = slab_offset(N, n, k);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
    }
}

// bias slab of a layer: N rows x 16 K columns, column 0 = fp16(bias), column 1 = fp16(bias - column 0), rest 0
__global__ void pe_tc_pack_bias_kernel(const float* __restrict__ bias, int N, unsigned char* __restrict__ dst) {
    const int total = N * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / 16, kk = i - n * 16;
        const float v = bias[n];
        const __half h = __float2half_rn(v);
        __half out = __float2half_rn(0.f);
        if (kk == 0) out = h;
        if (kk == 1) out = __float2half_rn(v - __half2float(h));
        const int64_t off = (int64_t)(kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(dst + off) = out;
    }
}

// D = A * B^T on one CTA through the same building blocks (validation of descriptors / TMEM addressing)
__global__ void __launch_bounds__(128, 1) pe_debug_umma_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ bias,
                                                                float* __restrict__ d, int n, int k) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sa = smem;                         // 128 x k
    unsigned char* sb = smem + 128 * 256 * 2;         // n x k
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 256 * 2 + 256 * 256 * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    unsigned char* ones = smem + 128 * 256 * 2 + 256 * 256 * 2 + 64;     // 256 B
    unsigned char* sbias = ones + 256;                                    // n x 16 halves
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = threadIdx.x;
    {
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        reinterpret_cast<__half*>(ones)[threadIdx.x] = __float2half_rn((r < 8 && c < 2) ? 1.f : 0.f);
    }
    if (bias) {
        for (int i = threadIdx.x; i < n * 16; i += blockDim.x) {
            const int row = i / 16, kk = i - row * 16;
            const float v = bias[row];
            const __half h = __float2half_rn(v);
            __half out = __float2half_rn(0.f);
            if (kk == 0) out = h;
            if (kk == 1) out = __float2half_rn(v - __half2float(h));
            const int64_t off = (int64_t)(kk >> 3) * (n * 16) + (row >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2;
            *reinterpret_cast<__half*>(sbias + off) = out;
        }
    }
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
        if (bias) umma_f16_ss(tmem_base, umma_smem_desc(smem_u32(ones), 128, 0), umma_smem_desc(smem_u32(sbias), lbo_b, 128), idesc, 1u);
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
    struct Item { const float* w; const float* b; int N, K_src, K_pad; };
    const Item items[NUM_LAYERS] = {
        {p.backbone_w[0], p.backbone_b[0], 256, 63, 64},   {p.backbone_w[1], p.backbone_b[1], 256, 256, 256},
        {p.backbone_w[2], p.backbone_b[2], 256, 256, 256}, {p.backbone_w[3], p.backbone_b[3], 256, 256, 256},
        {p.backbone_w[4], p.backbone_b[4], 256, 319, 320}, {p.backbone_w[5], p.backbone_b[5], 256, 256, 256},
        {p.backbone_w[6], p.backbone_b[6], 256, 256, 256}, {p.backbone_w[7], p.backbone_b[7], 256, 256, 256},
        {p.head0_w, nullptr, 256, 256, 256},               {p.head3_w, nullptr, 128, 256, 256},
        {p.head6_w, p.head6_b, 192, 128, 128}};
    int64_t off = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const Item& it = items[l];
        if (!it.w || (l != 8 && l != 9 && !it.b)) { pe_set_error("missing parameter tensor for tensor-core layer %d", l); return PE_ERR_INVALID; }
        const int64_t total = (int64_t)it.N * it.K_pad;
        pe_tc_pack_layer_kernel<<<(it.N + 63) / 64, 64, 0, stream>>>(it.w, it.N, it.K_src, it.K_pad, hi + off, lo + off);
        PE_LAUNCH_CHECK("pe_tc_pack_layer_kernel");
        off += total * 2;
        if (it.b) {
            pe_tc_pack_bias_kernel<<<(it.N * 16 + 255) / 256, 256, 0, stream>>>(it.b, it.N, hi + off);
            PE_LAUNCH_CHECK("pe_tc_pack_bias_kernel");
            off += (int64_t)it.N * 32;
        }
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
    const char* dbg_env = getenv("PE_TC_DEBUG");
    pe_field_tc_kernel<<<grid, THREADS, SMEM_TOTAL, stream>>>(args, global_out, num_passes, dbg_env ? atoi(dbg_env) : 0);
    PE_LAUNCH_CHECK("pe_field_tc_kernel");
    return PE_OK;
}

extern "C" int pe_debug_umma_gemm(const float* a, const float* b, const float* bias, float* d, int32_t n, int32_t k, pe_stream_t stream) {
    if (n < 16 || n > 256 || n % 16 || k < 16 || k > 256 || k % 16) { pe_set_error("debug gemm: n,k multiples of 16 up to 256"); return PE_ERR_INVALID; }
    const int smem = 128 * 256 * 2 + 256 * 256 * 2 + 64 + 256 + 256 * 32;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_debug_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pe_debug_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, bias, d, n, k);
    PE_LAUNCH_CHECK("pe_debug_umma_kernel");
    return PE_OK;
}

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
#include "pe_tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace pe;
using namespace pe_tc;

constexpr int SPLIT = 1;                          // threads per tile row in the epilogue groups
constexpr int THREADS = 128 + 2 * 128 * SPLIT;    // producer, MMA, TMEM-alloc, spare + 2 groups of 4*SPLIT epilogue warps
constexpr int SMEM_BAR = 2 * A_BYTES + NUM_STAGES * STAGE_BYTES;
constexpr int SMEM_ONES = SMEM_BAR + 256;         // 256-byte "ones" operand of the rank-1 bias update
constexpr int SMEM_TOTAL = SMEM_ONES + 256;

// kStats: the train-mode instantiation (statistics phases of BatchNorm); the eval instantiation carries none of that code
// kFoldOnly: the headline instantiation (folded head, sampling inside the kernel, eval)
// kX3: the fp16x3 instantiation (one tile per iteration, both epilogue groups on its rows); the other modes do not carry its code
// kPasses: weight passes per k-step known at compile time (1: fp16, 2: fp16x2 / fp16x3), 0: per layer from `pass2_mask`
// (bit l set: layer l runs hi + lo weight passes -- the "mixed" precision mode gives the late layers two passes)
template <bool kStats, bool kFoldOnly, bool kX3, int kPasses>
__global__ void __launch_bounds__(THREADS, 1) pe_field_tc_kernel(const PeFieldArgs A, const PeIntegrated G2, const int pass2_mask, const int x3_arg, const int fold, const int dbg) {
    constexpr int x3 = kX3 ? 1 : 0;
    auto layer_passes = [&](int l) { return kPasses ? kPasses : (((pass2_mask >> l) & 1) ? 2 : 1); };
    (void)x3_arg;
    // fold: folded-head mode (pe_tc_common.cuh): head layer 6 is applied per ray by pe_head6_fold_kernel, 10 MMA layers per tile
    // x3: fp16x3 mode — ONE tile per iteration; buffer 0 holds the high halves of the activations, buffer 1 the low halves;
    // per k-step A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32-class accuracy on the tensor cores); epilogue group Y idles
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;     // [2]
    uint64_t* a_ready = acc_full + 2;                // [2]
    uint64_t* h6_full = a_ready + 2;                 // [2] folded-head mode: per-ray sums of a tile are written
    uint64_t* h6_done = h6_full + 2;                 // [2] ... and consumed by the head-6 warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h6_done + 2);
    unsigned char* ones = smem + SMEM_ONES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);

    const int P = ob.positions;
    const int rpt = TILE_M / P;                                   // rays per tile
    const int tiles_per_image = (A.rays + rpt - 1) / rpt;
    // pre-pass mode: the kernel walks the list of non-empty tiles built on the device
    const int64_t total_tiles = A.tile_count ? (int64_t)__ldg(A.tile_count) : (int64_t)tiles_per_image * A.images;
    const int64_t total_pairs = x3 ? total_tiles : (total_tiles + 1) / 2;        // iterations of this kernel
    // train-mode BatchNorm (adain.py:47) needs the batch statistics of the two AdaIn inputs before it can go on: phase 1 stops after
    // head layer 0 and accumulates its column sums, phase 2 after head layer 3 (three launches, two global reductions)
    const int stat_phase = (kStats && A.training) ? A.phase : 0;
    const int num_layers = stat_phase == 1 ? 9 : (stat_phase == 2 ? 10 : (fold ? NUM_LAYERS - 1 : NUM_LAYERS));
    // train mode with a trunk cache (A.h7_out): phase 1 (the first launch) writes every row's trunk output, phases 2 and 0 start at
    // head layer 0 from it instead of evaluating encoding + trunk two more times
    const int first_layer = (kStats && A.training && A.h7_out != nullptr && A.phase != 1) ? 8 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(acc_full + g, 1); mbar_init(a_ready + g, x3 ? 8 : 4 * SPLIT); mbar_init(h6_full + g, x3 ? 8 : 4 * SPLIT); mbar_init(h6_done + g, 1); }
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
                for (int l = 0; l < num_layers; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t bytes = (uint32_t)n * PE_TC_SLAB_K * 2;
                    if (l < first_layer) { src += (int64_t)slabs * bytes + (has_bias ? n * 32 : 0); continue; }
                    const int num_passes = layer_passes(l);
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
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(smem); R.a_addr[1] = smem_u32(smem + A_BYTES);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = layer_passes(0); R.x3 = x3;
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
                for (int l = first_layer; l < num_layers; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    if (!kPasses) R.num_passes = layer_passes(l);
                    // transposed layers (folded head: head layer 3; statistics phases: the layer whose column sums are wanted)
                    const bool swap = (fold && l == 9) || (stat_phase != 0 && l == num_layers - 1);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, swap ? TILE_M : n);
                    const uint32_t lbo_b = (uint32_t)n * 16;          // bytes between K chunks of a slab: (n/8) core matrices
                    mbar_wait(a_ready + 0, ready_phase);
                    if (!x3) mbar_wait(a_ready + 1, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    // folded-head mode: head layer 3 is issued transposed (D^T = W3 * A^T: weights as the M operand, the tile's
                    // samples as N) so that its epilogue can sum over a ray's samples inside one thread; both operands are
                    // K-major in the same canonical layout, so the two descriptors simply swap roles
                    if (kStats && swap && n == 256) mma_layer<true, 2, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                    else if (!kPasses && kFoldOnly) {
                        // mixed mode, headline instantiation: one copy of the issue loop per pass count, so that both run with a
                        // compile-time trip count like the single-mode instantiations
                        if (R.num_passes == 2) {
                            R.num_passes = 2;
                            if (swap) mma_layer<true, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                            else mma_layer<false, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                        } else {
                            R.num_passes = 1;
                            if (swap) mma_layer<true, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                            else mma_layer<false, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                        }
                    }
                    else if (swap) mma_layer<true, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                    else mma_layer<false, 1, x3>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                    if (has_bias) {          // D += ones(128x16) * [bias_hi | bias_lo | 0..]^T : the bias, at fp32-class accuracy
                        mbar_wait(full_bar + R.stage, R.phase);
                        tc_fence_after();
                        const uint64_t db = umma_smem_desc(R.ring_addr + R.stage * STAGE_BYTES, lbo_b, 128);
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (x3 && g == 1) break;
                            umma_f16_ss(tmem_base + g * 256, ones_desc, db, idesc, 1u);
                            umma_commit(acc_full + g);
                        }
                        umma_commit(empty_bar + R.stage);
                        if (++R.stage == NUM_STAGES) { R.stage = 0; R.phase ^= 1; }
                    }
                }
            }
        }
    } else if (fold && (warp == 2 || (warp == 3 && !x3))) {
        // ================================ head layer 6, once per ray (folded-head mode) ================================
        // integrated_features[ray] = W6 * v[ray] + b6 * s[ray] in exact fp32 (adain_style_nerf_model.py:88 commutes with the
        // linear volume-rendering sum of object_composer.py:749); v, s were just written by this CTA's epilogue group
        const int g = warp - 2;
        const float* w6t = reinterpret_cast<const float*>(blob + L.head6_w);      // [128][192] (transposed nn.Linear weight)
        const float* b6 = reinterpret_cast<const float*>(blob + L.head6_b);
        float* out0 = A.integ.integrated_features;
        float* out1 = G2.integrated_features;
        uint32_t hphase = 0;
        for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
            const int64_t tile = x3 ? pair : pair * 2 + g;
            if (lane == 0) mbar_wait(h6_full + g, hphase);
            __syncwarp();
            hphase ^= 1;
            if (tile >= total_tiles) { if (lane == 0) mbar_arrive(h6_done + g); continue; }
            const int img = (int)(tile / tiles_per_image);
            const int ray0 = (int)(tile - (int64_t)img * tiles_per_image) * rpt;
            const int nr = min(rpt, A.rays - ray0);                                 // <= 4 (positions >= 32)
            const int64_t gr0 = (int64_t)img * A.rays + ray0;
            const float* v = A.fold_v + gr0 * FOLD_K;
            float acc[4][6];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float sr = r < nr ? A.fold_s[gr0 + r] : 0.f;
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[r][j] = __ldg(b6 + lane + 32 * j) * sr;
            }
#pragma unroll 2
            for (int k = 0; k < FOLD_K; k += 4) {
                float w[4][6];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int j = 0; j < 6; ++j) w[kk][j] = __ldg(w6t + (k + kk) * 192 + lane + 32 * j);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r < nr) {
                        const float4 q = *reinterpret_cast<const float4*>(v + r * FOLD_K + k);
#pragma unroll
                        for (int j = 0; j < 6; ++j)
                            acc[r][j] = fmaf(q.x, w[0][j], fmaf(q.y, w[1][j], fmaf(q.z, w[2][j], fmaf(q.w, w[3][j], acc[r][j]))));
                    }
                }
            }
            // destinations: the object's grid, the scene's grid, and -- fused all-gather -- every peer's copy of the grid (the same
            // 128-byte warp stores, P2P over NVLink).  The destination loop is the OUTER, rolled one: 24 stores of code, not 24 per target.
#pragma unroll 1
            for (int dsti = -2; dsti < A.peers; ++dsti) {
                float* dst = dsti == -2 ? out0 : (dsti == -1 ? out1 : A.peer_features[dsti]);
                if (!dst) continue;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r < nr) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) dst[(gr0 + r) * 192 + lane + 32 * j] = acc[r][j];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(h6_done + g);
        }
    } else if (warp >= 4) {
        // ================================ epilogue groups ================================
        const int g = (warp - 4) / (4 * SPLIT);        // 0: tile X, 1: tile Y
        const int gw = (warp - 4) % (4 * SPLIT);       // warp inside the group
        TileCtx X;
        X.A = &A; X.G2 = &G2;
        X.abuf = smem + g * A_BYTES;                   // derived from the __shared__ base so accesses compile to LDS/STS
        X.abuf_lo = smem + A_BYTES;
        X.wq = warp & 3;                               // TMEM lane quadrant this warp may access
        X.m = (X.wq << 5) | lane; X.lane = lane; X.gw = gw;
        X.half = gw >> 2;                              // the two warps of a quadrant split the columns of their 32 rows
        X.taddr = tmem_base + (((uint32_t)X.wq * 32u) << 16) + g * 256;
        X.bar_id = 1 + g;
        X.P = P; X.rpt = rpt; X.rows_used = rpt * P; X.tiles_per_image = tiles_per_image; X.total_tiles = total_tiles;
        X.size[0] = ob.bbox[1] - ob.bbox[0]; X.size[1] = ob.bbox[3] - ob.bbox[2]; X.size[2] = ob.bbox[5] - ob.bbox[4];
        X.alpha_bias = __ldg(reinterpret_cast<const float*>(blob + L.alpha_b));
        X.alpha_w = reinterpret_cast<const float*>(blob + L.alpha_w);
        X.dbg = dbg;
        X.fold = fold != 0;
        X.stat_phase = stat_phase;
        X.resume = first_layer != 0;
        X.single = G2.integrated_features != nullptr || G2.opacity != nullptr || G2.weights != nullptr;   // this object IS the scene
        Sync1 sync{acc_full + g, a_ready + g, 0u, lane, h6_full + g, h6_done + g, 0u};
        if constexpr (kX3) {
            // one tile per iteration: the two epilogue groups become the two column halves of the same 128 rows
            X.abuf = smem; X.half = g; X.gw = warp - 4; X.bar_id = 1;
            X.taddr = tmem_base + (((uint32_t)X.wq * 32u) << 16);
            Sync1 sync3{acc_full, a_ready, 0u, lane, h6_full, h6_done, 0u};
            TileAhead ahead;
            for (int64_t tile = blockIdx.x; tile < total_pairs; tile += gridDim.x) epilogue_tile<2, true, kStats, kFoldOnly>(X, tile, sync3, ahead, tile + gridDim.x);
        } else {
            TileAhead ahead;
            for (int64_t pair = blockIdx.x; pair < total_pairs; pair += gridDim.x)
                epilogue_tile<SPLIT, false, kStats, kFoldOnly>(X, pair * 2 + g, sync, ahead, (pair + gridDim.x) * 2 + g);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------
// Ray bender (model/nerf_models/positional_ray_bender_model.py:81-163) on the tensor cores, shipped shape:
//   [PE(x/size, 6 octaves, annealed) | deformation(32)] (71 -> K 96) -> 6 x [Linear 128 + ReLU], [h | input] before layer 3
//   -> Linear(128 -> 3, no bias) * size -> clamp into the box -> bent position, second in-box mask.
// Same building blocks as pe_field_tc_kernel in its fp16x3 form (activations and weights as hi + lo fp16 pairs, three MMAs
// per k-step: fp32-class results -- the displacement feeds ten octaves of Fourier features downstream), one 128-row tile per
// iteration over the device-built list of tiles with samples inside the box.  Reads the positions written by the fp32 sampling
// pass (PE_PHASE_SAMPLE) and overwrites them with the bent ones.
// ------------------------------------------------------------------------------------------------------
constexpr int B_LAYERS = 7;                       // 6 hidden + output
constexpr int B_ENC_CHUNK0 = 16;                  // A operand: K columns 0..127 activations, 128..223 the bender's input (71, padded to 96)
constexpr int B_A_CHUNKS = 28;
constexpr int B_A_BYTES = B_A_CHUNKS * CHUNK_BYTES;
constexpr int B_THREADS = 256;
constexpr int B_SMEM_BAR = 2 * B_A_BYTES + NUM_STAGES * STAGE_BYTES;
constexpr int B_SMEM_ONES = B_SMEM_BAR + 128;
constexpr int B_SMEM_TOTAL = B_SMEM_ONES + 256;

__host__ __device__ __forceinline__ void bender_layer_spec(int l, int& n, int& slabs, int& chunk0, bool& has_bias) {
    n = 128; slabs = 4; chunk0 = 0; has_bias = true;
    if (l == 0) { slabs = 3; chunk0 = B_ENC_CHUNK0; }
    else if (l == 3) { slabs = 7; }                // [h | input]: chunks 0..27 are contiguous
    else if (l == 6) { n = 16; has_bias = false; } // 3 outputs, padded to the smallest MMA N
}

constexpr uint32_t B_SMALL_OFF = 128;            // TMEM columns 128..255: accumulator of the small partial products (MmaRing::small_off)

// Epilogue of a hidden bender layer with the split accumulator: y = relu(large + small) -> hi + lo fp16 A operand of the next layer
__device__ __forceinline__ void bender_hidden_epilogue(uint32_t tcol, unsigned char* a_hi, unsigned char* a_lo, int m) {
    uint32_t v[32], w[32];
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        tmem_ld32(tcol + c * 32, v);
        tmem_ld32(tcol + B_SMALL_OFF + c * 32, w);
        tmem_wait_ld_regs(v);
        tmem_wait_ld_regs(w);
        float y[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = __fadd_rn(__uint_as_float(v[q]), __uint_as_float(w[q]));
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) store_a8_hilo(a_hi, a_lo, c * 4 + cc, m, y + 8 * cc, true);
    }
}

__global__ void __launch_bounds__(B_THREADS, 1) pe_bender_tc_kernel(const PeFieldArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + B_A_BYTES;
    unsigned char* ring = smem + 2 * B_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + B_SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;     // [1]
    uint64_t* a_ready = acc_full + 1;                // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    unsigned char* ones = smem + B_SMEM_ONES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    const int P = ob.positions;
    const int rpt = TILE_M / P;
    const int tiles_per_image = (A.rays + rpt - 1) / rpt;
    const int64_t total_tiles = (int64_t)__ldg(A.tile_count);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(acc_full, 1); mbar_init(a_ready, 4);
        mbar_fence_init();
    }
    if (threadIdx.x < 128) {
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        reinterpret_cast<__half*>(ones)[threadIdx.x] = __float2half_rn((r < 8 && c < 2) ? 1.f : 0.f);
    }
    fence_proxy_async();
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {                            // weight producer: [hi | lo] passes of every slab, then the bias slab
            int stage = 0; uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const unsigned char* src = blob + L.tcb_base;
                for (int l = 0; l < B_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    bender_layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t bytes = (uint32_t)n * PE_TC_SLAB_K * 2;
                    for (int s = 0; s < slabs; ++s) {
                        for (int pass = 0; pass < 2; ++pass) {
                            mbar_wait(empty_bar + stage, phase ^ 1);
                            mbar_arrive_expect_tx(full_bar + stage, bytes);
                            bulk_copy_g2s(ring + stage * STAGE_BYTES, src + (int64_t)pass * L.tcb_bytes_per_pass, bytes, full_bar + stage);
                            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                        }
                        src += bytes;
                    }
                    if (has_bias) {
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
        if (elect_one()) {                            // MMA issuer
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(a_hi); R.a_addr[1] = smem_u32(a_lo);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = 2; R.x3 = 2; R.small_off = B_SMALL_OFF;
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int l = 0; l < B_LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    bender_layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, n);
                    const uint32_t lbo_b = (uint32_t)n * 16;
                    mbar_wait(a_ready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    mma_layer<false, 1, 2>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
                    if (has_bias) {
                        mbar_wait(full_bar + R.stage, R.phase);
                        tc_fence_after();
                        const uint64_t db = umma_smem_desc(R.ring_addr + R.stage * STAGE_BYTES, lbo_b, 128);
                        umma_f16_ss(tmem_base, ones_desc, db, idesc, 1u);
                        umma_commit(acc_full);
                        umma_commit(empty_bar + R.stage);
                        if (++R.stage == NUM_STAGES) { R.stage = 0; R.phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // epilogue warps: thread = row of the tile = TMEM lane
        const int wq = warp & 3;
        const int m = (wq << 5) | lane;
        const uint32_t taddr = tmem_base + (((uint32_t)wq * 32u) << 16);
        Sync1 sync{acc_full, a_ready, 0u, lane, nullptr, nullptr, 0u};
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        const int rows_used = rpt * P;
        for (int64_t it = blockIdx.x; it < total_tiles; it += gridDim.x) {
            const int64_t tile = A.tile_list[it];
            const int img = (int)(tile / tiles_per_image);
            const int ray0 = (int)(tile - (int64_t)img * tiles_per_image) * rpt;
            const int rl = m / P;
            const int r = ray0 + rl;
            const bool valid = m < rows_used && r < A.rays;
            const int64_t gs = valid ? ((int64_t)img * A.rays + r) * P + (m - rl * P) : 0;
            const int flag0 = valid ? (A.flags[gs] & 1) : 0;
            float x[3] = {0.f, 0.f, 0.f};
            if (flag0) { x[0] = A.bent[gs * 3]; x[1] = A.bent[gs * 3 + 1]; x[2] = A.bent[gs * 3 + 2]; }
            // input of the bender: annealed Fourier features of x / size (positional_ray_bender_model.py:96-100) | deformation code
            {
                const float xn[3] = {__fdiv_rn(x[0], size[0]), __fdiv_rn(x[1], size[1]), __fdiv_rn(x[2], size[2])};
                const float* dfm = A.deformation + (int64_t)img * 32;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = 8 * c + i;
                        v[i] = e < 39 ? pe_encoding_value(xn, 3, e, ob.b_anneal) : (e < 71 ? __ldg(dfm + (e - 39)) : 0.f);
                    }
                    store_a8_hilo(a_hi, a_lo, B_ENC_CHUNK0 + c, m, v, false);
                }
            }
            sync.arrive_ready();
            for (int l = 0; l < 6; ++l) {
                sync.wait_acc();
                bender_hidden_epilogue(taddr, a_hi, a_lo, m);
                sync.arrive_ready();
            }
            sync.wait_acc();
            uint32_t v[16], vs[16];
            tmem_ld16(taddr, v);
            tmem_ld16(taddr + B_SMALL_OFF, vs);
            tmem_wait_ld_regs16(v);
            tmem_wait_ld_regs16(vs);
#pragma unroll
            for (int a = 0; a < 3; ++a) v[a] = __float_as_uint(__fadd_rn(__uint_as_float(v[a]), __uint_as_float(vs[a])));
            tc_fence_before();
            if (valid) {
                float bent[3], d2 = 0.f;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float dsp = __fmul_rn(__uint_as_float(v[a]), size[a]);
                    dsp = fmaxf(dsp, __fsub_rn(ob.bbox[2 * a], x[a]));           // clamp_output :116-140
                    dsp = fminf(dsp, __fsub_rn(ob.bbox[2 * a + 1], x[a]));
                    if (ob.canonical_pose) dsp = __fmul_rn(dsp, 0.f);
                    if (!flag0) dsp = 0.f;
                    bent[a] = __fadd_rn(x[a], dsp);
                    d2 += dsp * dsp;
                    if (A.disp_out) A.disp_out[gs * 3 + a] = dsp;
                }
                if (flag0) {
                    A.bent[gs * 3] = bent[0]; A.bent[gs * 3 + 1] = bent[1]; A.bent[gs * 3 + 2] = bent[2];
                    A.flags[gs] = (uint8_t)(1 | (pe_in_box(ob, bent) ? 2 : 0));  // inner mask of the field, adain_style_nerf_model.py:171-184
                }
                if (A.dispmag_out) A.dispmag_out[gs] = sqrtf(d2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------------
// Sampling pass of the tensor-core ray-bender path, one thread per sample slot: transform_rays, z bounds, create_ray_positions,
// outer in-box mask (the same device functions as the field kernels, so the values are identical), and the empty-space values
// of every per-sample output; the bender and field kernels then only touch the tiles that hold samples inside the box.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pe_sample_kernel(const PeFieldArgs A) {
    const PeObjectDesc& ob = A.ob;
    const int P = ob.positions;
    const int64_t per_image = (int64_t)A.rays * P;
    const int64_t total = per_image * A.images;
    for (int64_t gs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gs < total; gs += (int64_t)gridDim.x * blockDim.x) {
        const int img = (int)(gs / per_image);
        const int64_t rem = gs - (int64_t)img * per_image;
        const int r = (int)(rem / P), p = (int)(rem - (int64_t)r * P);
        const bool in_scene = A.ois ? A.ois[(int64_t)img * A.objects + A.k] != 0 : true;
        const PeRay ray = pe_make_ray(ob, A.w2o + ((int64_t)img * A.objects + A.k) * 12, A.origins + (int64_t)img * 3,
                                      A.dirs + ((int64_t)img * A.rays + r) * 3, in_scene);
        const float u = (A.perturb && !A.t_in) ? A.rand[gs] : 0.f;
        const float t = pe_sample_t_or(A.t_in, gs, ray, p, P, A.perturb != 0, u);
        float x[3];
        pe_position(ray, t, x);
        A.t_out[gs] = t;
        A.raw_out[gs] = ob.empty_space_alpha;
        A.inbox_out[gs] = 0;
        // bit 0: inside the box before bending, bit 1: to be evaluated by the field (set by the ray bender for bent samples; a zeroed
        // bender bends nothing, so the two masks coincide)
        A.flags[gs] = pe_in_box(ob, x) ? (ob.bender_kind == PE_BENDER_ZEROED ? 3 : 1) : 0;
        A.bent[gs * 3] = x[0]; A.bent[gs * 3 + 1] = x[1]; A.bent[gs * 3 + 2] = x[2];
        if (A.disp_out) { A.disp_out[gs * 3] = 0.f; A.disp_out[gs * 3 + 1] = 0.f; A.disp_out[gs * 3 + 2] = 0.f; }
        if (A.dispmag_out) A.dispmag_out[gs] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------
// pre-pass mode: list of the tiles (floor(128/P) rays each) that hold at least one sample to evaluate
// ------------------------------------------------------------------------------------------------------
__global__ void pe_tile_list_kernel(const uint8_t* __restrict__ flags, int flag_mask, int rays, int P, int rpt, int tiles_per_image,
                                    int64_t total_tiles, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool any = false;
    if (tile < total_tiles) {
        const int img = (int)(tile / tiles_per_image);
        const int ray0 = (int)(tile - (int64_t)img * tiles_per_image) * rpt;
        const int n = min(rpt, rays - ray0) * P;
        const uint8_t* f = flags + ((int64_t)img * rays + ray0) * P;
        for (int i = 0; i < n && !any; ++i) any = (f[i] & flag_mask) != 0;
    }
    // one atomic per warp; the order of the list does not matter (tiles are independent)
    const unsigned mask = __ballot_sync(0xffffffffu, any);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && mask) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (any) list[base + __popc(mask & ((1u << lane) - 1))] = (int32_t)tile;
}

// ------------------------------------------------------------------------------------------------------
// weight packing into the slab stream
// ------------------------------------------------------------------------------------------------------
// slab element (n, kk) lives at (kk/8)*(N*16) + (n/8)*128 + (n%8)*16 + (kk%8)*2  (K-major, no swizzle)
__host__ __device__ __forceinline__ int64_t slab_offset(int N, int n, int k) {
    const int slab = k / PE_TC_SLAB_K, kk = k - slab * PE_TC_SLAB_K;
    return (int64_t)slab * N * PE_TC_SLAB_K * 2 + (int64_t)(kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
}

// One thread per output row.  hi = fp16(w) with ZERO-SUM rounding: walking along K, each weight is rounded to the fp16
// neighbour (down or up) that keeps the running sum of rounding errors of the row closest to zero.  Post-ReLU activations
// have a large common positive mean, so the systematic part sum_k dW[n][k] * mean(a) of the single-pass error cancels
// (measured: -10..-30 % error on the rendered frame); lo = fp16(w - hi) is the second pass of the fp16x2 mode.
constexpr int PACK_ROWS = 8;           // rows of a layer per block of the zero-sum packing
__host__ __device__ inline size_t pack_layer_smem(int K_pad) { return (size_t)3 * PACK_ROWS * (K_pad + 1) * 4 + (size_t)PACK_ROWS * 16 * 4; }

// table of layers for one launch (blockIdx.y = item)
constexpr int TC_PACK_TABLE = 24;
struct TcPackTable {
    const float* w[TC_PACK_TABLE];           // layer weights [N_real][K_src] (zero-sum items) / bias [N] (bias items)
    unsigned char* hi[TC_PACK_TABLE];
    unsigned char* lo[TC_PACK_TABLE];
    int32_t N[TC_PACK_TABLE], K_src[TC_PACK_TABLE], K_pad[TC_PACK_TABLE], N_real[TC_PACK_TABLE];
};

__device__ __forceinline__ void pack_layer_rows(const float* __restrict__ w, int N, int K_src, int K_pad, unsigned char* __restrict__ hi,
                                                unsigned char* __restrict__ lo, int N_real, int block) {
    // N: rows of the slab layout; rows >= N_real (padding up to the smallest MMA N) are zero.  This kernel runs for every layer of
    // every model whenever a parameter changes, i.e. every training step, so only the decision itself is sequential: (a) all threads
    // load the block's PACK_ROWS rows (coalesced) and compute both candidates' rounding errors; (b) one thread per row walks along K
    // with nothing but `run` in its dependency chain and records the picks as bits; (c) all threads build the hi / lo halves and store
    // the slab layout in 16-byte pieces.
    extern __shared__ __align__(16) unsigned char pack_smem[];
    const int ld = K_pad + 1;
    float* ws = reinterpret_cast<float*>(pack_smem);                   // [PACK_ROWS][ld] weights
    float* en = ws + PACK_ROWS * ld;                                   // error of the nearest fp16
    float* eo = en + PACK_ROWS * ld;                                   // error of the other neighbour
    uint32_t* bits = reinterpret_cast<uint32_t*>(eo + PACK_ROWS * ld); // [PACK_ROWS][16] picks (K_pad <= 512)
    const int n0 = block * PACK_ROWS, tid = threadIdx.x;
    if (n0 >= N) return;
    auto candidates = [](float v, __half& near, __half& other) {
        near = __float2half_rn(v);
        const float fn = __half2float(near);
        other = near;
        if (fn != v) other = fn < v ? __float2half_ru(v) : __float2half_rd(v);
    };
    for (int idx = tid; idx < PACK_ROWS * K_pad; idx += blockDim.x) {
        const int r = idx / K_pad, k = idx - r * K_pad, n = n0 + r;
        const float v = (k < K_src && n < N_real && n < N) ? w[(int64_t)n * K_src + k] : 0.f;
        __half near, other;
        candidates(v, near, other);
        ws[r * ld + k] = v;
        en[r * ld + k] = __half2float(near) - v;
        eo[r * ld + k] = __half2float(other) - v;
    }
    __syncthreads();
    if (tid < PACK_ROWS) {
        const float* a = en + tid * ld;
        const float* b = eo + tid * ld;
        float run = 0.f;
        for (int k0 = 0; k0 < K_pad; k0 += 32) {
            uint32_t word = 0;
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) {
                const float e_near = a[k0 + kk], e_other = b[k0 + kk];
                const bool pick_other = fabsf(run + e_other) < fabsf(run + e_near);
                run += pick_other ? e_other : e_near;
                word |= (pick_other ? 1u : 0u) << kk;
            }
            bits[tid * 16 + (k0 >> 5)] = word;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < PACK_ROWS * (K_pad / 8); idx += blockDim.x) {
        const int r = idx % PACK_ROWS, k8 = idx / PACK_ROWS, n = n0 + r;
        if (n >= N) continue;
        const uint32_t word = bits[r * 16 + (k8 >> 2)] >> ((k8 & 3) * 8);
        __align__(16) __half h8[8];
        __align__(16) __half l8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float v = ws[r * ld + k8 * 8 + q];
            __half near, other;
            candidates(v, near, other);
            const __half h = ((word >> q) & 1u) ? other : near;
            h8[q] = h;
            l8[q] = __float2half_rn(v - __half2float(h));
        }
        const int64_t off = slab_offset(N, n, k8 * 8);
        *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h8);
        *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l8);
    }
}

__global__ void __launch_bounds__(128) pe_tc_pack_layer_kernel(const float* __restrict__ w, int N, int K_src, int K_pad, unsigned char* __restrict__ hi,
                                                               unsigned char* __restrict__ lo, int N_real) {
    pack_layer_rows(w, N, K_src, K_pad, hi, lo, N_real, blockIdx.x);
}

// every zero-sum layer of a model in ONE launch (grid.x = row blocks of the tallest layer, grid.y = layers)
__global__ void __launch_bounds__(128) pe_tc_pack_layers_kernel(const __grid_constant__ TcPackTable T) {
    const int it = blockIdx.y;
    pack_layer_rows(T.w[it], T.N[it], T.K_src[it], T.K_pad[it], T.hi[it], T.lo[it], T.N_real[it], blockIdx.x);
}

// Activation-aware rounding of one layer's hi stream (one block per output row, thread j = input j).  Every weight may go to either
// fp16 neighbour; with e the row's vector of rounding errors and C = E[a a^T] the second moments of the layer's (fp16-rounded) inputs,
// the expected squared error of the single-pass product is e^T C e.  Coordinate descent: visit the inputs in order, pick for input k
// the neighbour that minimises the quadratic form given all other choices (thread j keeps r_j = sum_i e_i C[i][j]); `sweeps` further
// passes refine the greedy solution.  Emulated on the CPU oracle this takes the single-pass error of the rendered frame from 4.8e-4
// (zero-sum rounding) to 2.1e-4 relative L2 (exact weights: 1.5e-4) with C measured on uniform positions in the box.
__global__ void pe_tc_pack_layer_aware_kernel(const float* __restrict__ w, const float* __restrict__ C, int N, int K_src, int K_pad,
                                              unsigned char* __restrict__ hi, unsigned char* __restrict__ lo, int sweeps) {
    __shared__ float delta_s[2];
    const int n = blockIdx.x, j = threadIdx.x;
    const float v = j < K_src ? w[(int64_t)n * K_src + j] : 0.f;
    const __half near = __float2half_rn(v);
    const float fn = __half2float(near);
    __half other = near;
    if (fn != v) other = fn < v ? __float2half_ru(v) : __float2half_rd(v);
    const float e0 = fn - v, e1 = __half2float(other) - v;
    const float cjj = j < K_src ? C[(int64_t)j * K_src + j] : 0.f;
    float r = 0.f, e = 0.f;
    bool pick = false;
    int it = 0;
    float c_next = j < K_src ? C[j] : 0.f;                    // row 0 of C
    for (int sweep = 0; sweep <= sweeps; ++sweep) {
        for (int k = 0; k < K_src; ++k, ++it) {
            const float c_row = c_next;
            const int kn = k + 1 < K_src ? k + 1 : 0;
            c_next = j < K_src ? C[(int64_t)kn * K_src + j] : 0.f;          // prefetch the next row behind the barrier
            if (j == k) {
                const float rr = r - e * cjj;                               // r_k without this input's own contribution
                const float c0 = fmaf(2.f * e0, rr, e0 * e0 * cjj), c1 = fmaf(2.f * e1, rr, e1 * e1 * cjj);
                const bool p = c1 < c0;
                const float en = p ? e1 : e0;
                delta_s[it & 1] = en - e;
                e = en; pick = p;
            }
            __syncthreads();
            const float d = delta_s[it & 1];
            if (d != 0.f) r = fmaf(d, c_row, r);
        }
    }
    if (j < K_pad) {
        const __half h = pick ? other : near;
        const int64_t off = slab_offset(N, n, j);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
    }
}

// bias slab of a layer: N rows x 16 K columns, column 0 = fp16(bias), column 1 = fp16(bias - column 0), rest 0
__global__ void pe_tc_pack_biases_kernel(const __grid_constant__ TcPackTable T) {      // every bias slab of a model in one launch
    const int it = blockIdx.y;
    const float* __restrict__ bias = T.w[it];
    unsigned char* __restrict__ dst = T.hi[it];
    const int N = T.N[it], total = N * 16;
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

// Validation of two operand forms the next kernels build on (tests/test_gpu_parity.py::test_umma_operand_forms):
//   mode 1: both operands MN-major, read from the SAME shared-memory layout the field kernel uses for activations
//           (element (row r, column c) at (c/8)*2048 + r*16 + (c%8)*2): with K = the 128 rows, D[m][n] = sum_r X[r][m] * Y[r][n]
//           (the dW = G^T * A product of the backward); lbo / sbo: byte strides passed by the caller
//   mode 2: A operand from TMEM (TS form), written by tcgen05.st: D[m][n] = sum_k A[m][k] * B[n][k]
__global__ void __launch_bounds__(128, 1) pe_debug_umma2_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b,
                                                                 float* __restrict__ d, int n, int k, int lbo, int sbo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sa = smem;                         // 128 rows x 256 columns, activation layout (64 KB)
    unsigned char* sb = smem + 65536;                 // same
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    const int r = threadIdx.x;                        // row of the stored matrices / TMEM lane
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (mode == 1) {
        // X = a: [128 rows (K)][128 columns (M)], Y = b: [128 rows (K)][n columns (N)], rows >= k are zero
        for (int c = 0; c < 128 / 8; ++c) {
            float v[8];
            for (int q = 0; q < 8; ++q) v[q] = r < k ? a[(int64_t)r * 128 + c * 8 + q] : 0.f;
            store_a8(sa, c, r, v);
        }
        for (int c = 0; c < n / 8; ++c) {
            float v[8];
            for (int q = 0; q < 8; ++q) v[q] = r < k ? b[(int64_t)r * n + c * 8 + q] : 0.f;
            store_a8(sb, c, r, v);
        }
    } else {
        // B = b: [n][k] K-major slab layout; A = a: [128][k] -> TMEM columns 256 .. 256 + k/2
        for (int64_t i = threadIdx.x; i < (int64_t)n * k; i += blockDim.x) {
            const int row = (int)(i / k), kk = (int)(i - (int64_t)row * k);
            const int64_t off = (int64_t)(kk >> 3) * (n * 16) + (row >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2;
            *reinterpret_cast<__half*>(sb + off) = __float2half_rn(b[i]);
        }
        for (int j = 0; j < k / 16; ++j) {
            uint32_t v[8];
            for (int q = 0; q < 8; ++q) v[q] = pack_half2(a[(int64_t)r * k + j * 16 + 2 * q], a[(int64_t)r * k + j * 16 + 2 * q + 1]);
            tmem_st8(tmem_base + (((uint32_t)warp * 32u) << 16) + 256 + j * 8, v);
        }
        tmem_wait_st();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && elect_one()) {
        if (mode == 1) {
            const uint32_t idesc = umma_idesc_f16_major(128, n, 1, 1);
            for (int j = 0; j < 128 / 16; ++j) {      // K = the 128 stored rows: 16 rows (2 groups of 8, 128 B each) per MMA
                const uint64_t da = umma_smem_desc(smem_u32(sa) + j * 256, (uint32_t)lbo, (uint32_t)sbo);
                const uint64_t db = umma_smem_desc(smem_u32(sb) + j * 256, (uint32_t)lbo, (uint32_t)sbo);
                umma_f16_ss(tmem_base, da, db, idesc, j != 0 ? 1u : 0u);
            }
        } else {
            const uint32_t idesc = umma_idesc_f16(128, n);
            const uint32_t lbo_b = (uint32_t)n * 16;
            for (int j = 0; j < k / 16; ++j) {
                const uint64_t db = umma_smem_desc(smem_u32(sb) + 2 * j * lbo_b, lbo_b, 128);
                umma_f16_ts(tmem_base, tmem_base + 256 + j * 8, db, idesc, j != 0 ? 1u : 0u);
            }
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
            if (c0 + q < n) d[(int64_t)r * n + c0 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int pe_tc_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream) {
    unsigned char* hi = (unsigned char*)packed + L.tc_base;
    unsigned char* lo = hi + L.tc_bytes_per_pass;
    struct Item { const float* w; const float* b; int N, K_src, K_pad; const float* moments; };
    const Item items[NUM_LAYERS] = {
        {p.backbone_w[0], p.backbone_b[0], 256, 63, 64, p.backbone_in_moments[0]},   {p.backbone_w[1], p.backbone_b[1], 256, 256, 256, p.backbone_in_moments[1]},
        {p.backbone_w[2], p.backbone_b[2], 256, 256, 256, p.backbone_in_moments[2]}, {p.backbone_w[3], p.backbone_b[3], 256, 256, 256, p.backbone_in_moments[3]},
        {p.backbone_w[4], p.backbone_b[4], 256, 319, 320, p.backbone_in_moments[4]}, {p.backbone_w[5], p.backbone_b[5], 256, 256, 256, p.backbone_in_moments[5]},
        {p.backbone_w[6], p.backbone_b[6], 256, 256, 256, p.backbone_in_moments[6]}, {p.backbone_w[7], p.backbone_b[7], 256, 256, 256, p.backbone_in_moments[7]},
        {p.head0_w, nullptr, 256, 256, 256, p.head0_in_moments},                     {p.head3_w, nullptr, 128, 256, 256, nullptr},
        {p.head6_w, p.head6_b, 192, 128, 128, nullptr}};
    // zero-sum layers and bias slabs are collected into tables and packed by ONE launch each (this runs every training step);
    // activation-aware layers (inference packs) keep their own launches
    TcPackTable layers = {}, biases = {};
    int n_layers = 0, n_biases = 0, max_rows = 0, max_kpad = 0, max_bias_n = 0;
    auto add_layer = [&](const float* w, int N, int K_src, int K_pad, unsigned char* h, unsigned char* l_, int N_real) {
        layers.w[n_layers] = w; layers.hi[n_layers] = h; layers.lo[n_layers] = l_;
        layers.N[n_layers] = N; layers.K_src[n_layers] = K_src; layers.K_pad[n_layers] = K_pad; layers.N_real[n_layers] = N_real;
        ++n_layers;
        max_rows = N > max_rows ? N : max_rows; max_kpad = K_pad > max_kpad ? K_pad : max_kpad;
    };
    auto add_bias = [&](const float* b, int N, unsigned char* dst) {
        biases.w[n_biases] = b; biases.hi[n_biases] = dst; biases.N[n_biases] = N;
        ++n_biases;
        max_bias_n = N > max_bias_n ? N : max_bias_n;
    };
    int64_t off = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const Item& it = items[l];
        if (!it.w || (l != 8 && l != 9 && !it.b)) { pe_set_error("missing parameter tensor for tensor-core layer %d", l); return PE_ERR_INVALID; }
        const int64_t total = (int64_t)it.N * it.K_pad;
        if (it.moments) {
            pe_tc_pack_layer_aware_kernel<<<it.N, (it.K_pad + 31) / 32 * 32, 0, stream>>>(it.w, it.moments, it.N, it.K_src, it.K_pad, hi + off, lo + off, 1);
            PE_LAUNCH_CHECK("pe_tc_pack_layer_aware_kernel");
        } else {
            add_layer(it.w, it.N, it.K_src, it.K_pad, hi + off, lo + off, it.N);
        }
        off += total * 2;
        if (it.b) {
            add_bias(it.b, it.N, hi + off);
            off += (int64_t)it.N * 32;
        }
    }
    if (off != L.tc_bytes_per_pass) { pe_set_error("internal: tensor-core weight stream size mismatch"); return PE_ERR_INVALID; }
    if (L.tcb_base) {
        // ray-bender stream: the 6 hidden layers (K padded 71 -> 96, 199 -> 224) with their bias slabs, then the 3-row output layer
        unsigned char* bhi = (unsigned char*)packed + L.tcb_base;
        unsigned char* blo = bhi + L.tcb_bytes_per_pass;
        int64_t boff = 0;
        for (int l = 0; l < 6; ++l) {
            if (!p.bender_w[l] || !p.bender_b[l]) { pe_set_error("missing ray-bender parameter tensor of layer %d", l); return PE_ERR_INVALID; }
            const int K_src = l == 0 ? 71 : (l == 3 ? 199 : 128), K_pad = l == 0 ? 96 : (l == 3 ? 224 : 128);
            add_layer(p.bender_w[l], 128, K_src, K_pad, bhi + boff, blo + boff, 128);
            boff += (int64_t)128 * K_pad * 2;
            add_bias(p.bender_b[l], 128, bhi + boff);
            boff += 128 * 32;
        }
        if (!p.bender_out_w) { pe_set_error("missing ray-bender output layer"); return PE_ERR_INVALID; }
        add_layer(p.bender_out_w, 16, 128, 128, bhi + boff, blo + boff, 3);
        boff += (int64_t)16 * 128 * 2;
        if (boff != L.tcb_bytes_per_pass) { pe_set_error("internal: ray-bender weight stream size mismatch"); return PE_ERR_INVALID; }
    }
    static_assert(NUM_LAYERS + 7 <= TC_PACK_TABLE, "pack table too small");
    if (n_layers) {
        pe_tc_pack_layers_kernel<<<dim3((max_rows + PACK_ROWS - 1) / PACK_ROWS, n_layers), 128, pack_layer_smem(max_kpad), stream>>>(layers);
        PE_LAUNCH_CHECK("pe_tc_pack_layers_kernel");
    }
    if (n_biases) {
        pe_tc_pack_biases_kernel<<<dim3((max_bias_n * 16 + 255) / 256, n_biases), 256, 0, stream>>>(biases);
        PE_LAUNCH_CHECK("pe_tc_pack_biases_kernel");
    }
    return PE_OK;
}

int pe_launch_field_tc(const PeFieldArgs& args, const PeIntegrated& global_out, int sm_count, cudaStream_t stream) {
    const bool prepass = args.bent != nullptr;
    if (!(prepass ? pe_tc_field_ok(args.ob) : pe_tc_shape_ok(args.ob)) || args.explicit_positions || args.phase < 0 || args.phase > 2 ||
        (args.phase != 0 && (!args.training || !args.stats))) {
        pe_set_error("tensor-core field kernel: unsupported configuration");
        return PE_ERR_UNSUPPORTED;
    }
    if (prepass && (!args.flags || !args.tile_list || !args.tile_count || !args.t_out || !args.feat_out || !args.raw_out || args.fold_v)) {
        pe_set_error("tensor-core field kernel: incomplete pre-pass hand-off");
        return PE_ERR_INVALID;
    }
    const int x3 = args.precision == PE_PRECISION_FP16X3 ? 1 : 0;
    int num_passes = args.precision == PE_PRECISION_FP16 ? 1 : 2;
    // per-layer weight passes: bit l of the mask = layer l (0-7 trunk, 8 head 0, 9 head 3, 10 head 6) runs hi + lo.
    // "mixed" (PE_PRECISION_MIXED): two passes where the systematic fp16 rounding of the weights matters most (trunk layers L3-L7,
    // measured: profiles/r2_mixed_mode.md), one pass for the early trunk and the head; PE_TC_PASS2_MASK overrides the choice.
    int pass2_mask = num_passes == 2 ? 0x7FF : 0;
    bool mixed = args.precision == PE_PRECISION_MIXED;
    if (mixed) {
        const char* menv = getenv("PE_TC_PASS2_MASK");
        pass2_mask = menv ? (int)strtol(menv, nullptr, 0) : ((args.pass2_mask & 0x10000) ? (args.pass2_mask & 0xFFFF) : PE_TC_MIXED_MASK);
        if (pass2_mask == 0) { mixed = false; num_passes = 1; }          // no two-pass layer left: the single-pass instantiation
    }
    const int fold = (args.fold_v != nullptr && args.phase == 0) ? 1 : 0;
    if (fold && (!args.fold_s || args.feat_out || args.apply_activation || args.ob.positions % 32)) {
        pe_set_error("tensor-core field kernel: folded head needs positions %% 32 == 0, no per-sample features, no output activation");
        return PE_ERR_UNSUPPORTED;
    }
    const int dbg = getenv("PE_TC_TIMELINE") ? atoi(getenv("PE_TC_TIMELINE")) : 0;
    const int rpt = TILE_M / args.ob.positions;
    const int64_t tiles = (int64_t)((args.rays + rpt - 1) / rpt) * args.images;
    const int64_t pairs = x3 ? tiles : (tiles + 1) / 2;
    if (pairs == 0) return PE_OK;
    const int grid = (int)pe_min64(pairs, sm_count);
    // one instantiation per (train-mode statistics, headline specialisation, fp16x3) combination that occurs
    using Kernel = void (*)(const PeFieldArgs, const PeIntegrated, const int, const int, const int, const int);
    Kernel kernel;
    if (args.training) kernel = x3 ? pe_field_tc_kernel<true, false, true, 2> : pe_field_tc_kernel<true, false, false, 0>;
    else if (fold && !prepass) kernel = x3 ? pe_field_tc_kernel<false, true, true, 2> : (mixed ? pe_field_tc_kernel<false, true, false, 0> : (num_passes == 1 ? pe_field_tc_kernel<false, true, false, 1> : pe_field_tc_kernel<false, true, false, 2>));
    else kernel = x3 ? pe_field_tc_kernel<false, false, true, 2> : pe_field_tc_kernel<false, false, false, 0>;
    PE_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    kernel<<<grid, THREADS, SMEM_TOTAL, stream>>>(args, global_out, pass2_mask, x3, fold, dbg);
    PE_LAUNCH_CHECK("pe_field_tc_kernel");
    return PE_OK;
}

int pe_launch_tile_list(const PeFieldArgs& args, int flag_mask, int32_t* tile_list, int32_t* tile_count, cudaStream_t stream) {
    const int rpt = TILE_M / args.ob.positions;
    const int tiles_per_image = (args.rays + rpt - 1) / rpt;
    const int64_t tiles = (int64_t)tiles_per_image * args.images;
    PE_CUDA_CHECK(cudaMemsetAsync(tile_count, 0, sizeof(int32_t), stream));
    if (tiles == 0) return PE_OK;
    pe_tile_list_kernel<<<(unsigned)((tiles + 127) / 128), 128, 0, stream>>>(args.flags, flag_mask, args.rays, args.ob.positions, rpt, tiles_per_image,
                                                                             tiles, tile_list, tile_count);
    PE_LAUNCH_CHECK("pe_tile_list_kernel");
    return PE_OK;
}

int pe_launch_sample(const PeFieldArgs& args, int sm_count, cudaStream_t stream) {
    if (!args.t_out || !args.raw_out || !args.inbox_out || !args.flags || !args.bent || args.explicit_positions) {
        pe_set_error("sampling pass: incomplete arguments");
        return PE_ERR_INVALID;
    }
    const int64_t total = (int64_t)args.images * args.rays * args.ob.positions;
    if (total == 0) return PE_OK;
    pe_sample_kernel<<<(int)pe_min64((total + 255) / 256, (int64_t)sm_count * 8), 256, 0, stream>>>(args);
    PE_LAUNCH_CHECK("pe_sample_kernel");
    return PE_OK;
}

int pe_launch_bender_tc(const PeFieldArgs& args, int sm_count, cudaStream_t stream) {
    if (!pe_tc_prepass_ok(args.ob) || !pe_tc_bender_ok(args.ob) || !args.L.tcb_base || !args.bent || !args.flags || !args.tile_list ||
        !args.tile_count || !args.deformation) {
        pe_set_error("tensor-core ray bender: unsupported configuration");
        return PE_ERR_UNSUPPORTED;
    }
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bender_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_TOTAL));
    const int rpt = TILE_M / args.ob.positions;
    const int64_t tiles = (int64_t)((args.rays + rpt - 1) / rpt) * args.images;
    if (tiles == 0) return PE_OK;
    pe_bender_tc_kernel<<<(int)pe_min64(tiles, sm_count), B_THREADS, B_SMEM_TOTAL, stream>>>(args);
    PE_LAUNCH_CHECK("pe_bender_tc_kernel");
    return PE_OK;
}

// One layer through the weight packing (zero-sum rounding, or activation-aware when `moments` is given) into plain slab buffers of
// N * K_pad fp16 each -- tests/test_gpu_parity.py::test_activation_aware_rounding unpacks them and checks the rounding choices.
extern "C" int pe_debug_pack_layer(const float* w, const float* moments, int32_t N, int32_t K_src, int32_t K_pad, int32_t sweeps, void* hi,
                                   void* lo, pe_stream_t stream) {
    if (N < 8 || N % 8 || K_pad % 32 || K_src > K_pad || K_pad > (moments ? 1024 : 448) || !w || !hi || !lo) { pe_set_error("debug pack: bad shape"); return PE_ERR_INVALID; }
    if (moments) pe_tc_pack_layer_aware_kernel<<<N, K_pad, 0, (cudaStream_t)stream>>>(w, moments, N, K_src, K_pad, (unsigned char*)hi, (unsigned char*)lo, sweeps);
    else pe_tc_pack_layer_kernel<<<(N + PACK_ROWS - 1) / PACK_ROWS, 128, pack_layer_smem(K_pad), (cudaStream_t)stream>>>(w, N, K_src, K_pad, (unsigned char*)hi, (unsigned char*)lo, N);
    PE_LAUNCH_CHECK("pe_debug_pack_layer");
    return PE_OK;
}

extern "C" int pe_debug_umma_gemm2(int32_t mode, const float* a, const float* b, float* d, int32_t n, int32_t k, int32_t lbo, int32_t sbo,
                                   pe_stream_t stream) {
    if ((mode != 1 && mode != 2) || n < 32 || n > 256 || n % 32 || k < 16 || k > 128 || k % 16) {
        pe_set_error("debug gemm2: mode 1|2, n multiple of 32 up to 256, k multiple of 16 up to 128");
        return PE_ERR_INVALID;
    }
    const int smem = 131072 + 64;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_debug_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pe_debug_umma2_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, a, b, d, n, k, lbo, sbo);
    PE_LAUNCH_CHECK("pe_debug_umma2_kernel");
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

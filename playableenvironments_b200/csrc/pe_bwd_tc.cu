// Backward of the shipped field (8x256 trunk, skip at 4, 10 octaves, AdaIn head, 192 features) on the 5th-generation tensor cores.
// What autograd does for the reference by replaying model/nerf_models/adain_style_nerf_model.py:106-145, model/layers/adain.py:21-61 and
// model/positional_encoder.py:41-65 op by op (training/trainer.py:643).
//
// The samples inside the object's box are compacted per image into tiles of 128 rows (pe_launch_compact_slots).  Three kernels walk them:
//   pe_bwd_fwd_kernel    recomputes the field forward of a tile pair on tcgen05 (hi + lo weight passes) and writes every layer's
//                        activations to the tile's STASH block in fp16, in the very K-major no-swizzle operand layout the MMAs read
//                        (chunk = 8 columns x 128 rows = 2048 B), plus one ReLU-mask bit per activation;
//   pe_bwd_chain_kernel  dX chain: G_{l-1} = (G_l W_l) * relu'(h_{l-1}) with the TRANSPOSED weight stream (PeLayout::tcT_base), AdaIn /
//                        train-mode BatchNorm backward in the epilogues, down to the gradient of the sample position; every G_l goes to
//                        the stash (fp16, scaled by a power of two S chosen per call);
//   pe_bwd_dw_kernel     dW_l = G_l^T A_l, db_l = G_l^T 1: both operands MN-major straight from the stash (K = the tile's 128 samples,
//                        form validated by tests/test_gpu_parity.py::test_umma_operand_forms), accumulated in TMEM over a range of
//                        tiles, added once to the fp32 parameter gradients.
#include "pe_tc_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {
using namespace pe;
using namespace pe_tc;

// ---- stash map: chunk offsets inside a tile's block (PE_BWD_FS_CHUNKS chunks of 2048 B) ----
// The stash feeds dW = G^T A, a sum over ~10^5..10^6 samples: plain fp16 copies are enough there (measured: adding the lo halves of G and A
// as two more products changes no parameter gradient by more than 1e-4 of its scale) -- unlike the operands of the recompute and of the dX
// chain, which stay hi + lo pairs in shared memory (see "Numerics" below).
__host__ __device__ constexpr int FS_H(int l) { return l < 4 ? 32 * l : 136 + 32 * (l - 4); }   // h0..h3 | enc | h4..h7: [h3 | enc] is contiguous
constexpr int FS_ENC = 128, FS_Y1 = 264, FS_Y2 = 296;          // activations: [0, 312)
constexpr int FS_X1 = 312, FS_X2 = 344;                        // AdaIn inputs: [312, 360)
constexpr int FS_GF = 360;                                     // 24 chunks (192 columns); a 128-row block starting at column 128 over-reads into GP(0)
__host__ __device__ constexpr int FS_GP(int l) { return FS_GF + 24 + 32 * l; }
constexpr int FS_GX1 = FS_GF + 24 + 256, FS_GX2 = FS_GX1 + 32;  // gradients: [360, 688)
constexpr int FS_GRAW = 688, FS_MASK = 690;                    // 16-column operand [graw hi | graw lo | 0 ...]; ReLU-mask words
// mask words (uint32, [word][row]): h_l -> words 8l .. 8l+7, y1 -> 64..71, y2 -> 72..75
constexpr int MASK_Y1 = 64, MASK_Y2 = 72, MASK_WORDS = 76;
static_assert(FS_GX2 + 16 == FS_GRAW, "stash map");
static_assert(FS_MASK * 2048 + MASK_WORDS * 512 <= PE_BWD_FS_CHUNKS * 2048, "stash block too small");
constexpr int64_t FS_BYTES = (int64_t)PE_BWD_FS_CHUNKS * CHUNK_BYTES;

constexpr int TCT_WEXP = 10;                      // the transposed weight stream holds 2^10 * W (hi + lo)
constexpr int THREADS = 384;                      // producer, MMA, TMEM-alloc, spare + 2 epilogue groups of 4 warps
constexpr int SMEM_BAR = 2 * A_BYTES + NUM_STAGES * STAGE_BYTES;
constexpr int SMEM_ONES = SMEM_BAR + 256;
constexpr int SMEM_TOTAL = SMEM_ONES + 256;
// chain kernel: constants of the AdaIn / alpha steps live in the encoding columns of the A buffer (float offsets from CST_BASE) until
// step 6 overwrites them with the encoding gradient of the skip layer
constexpr int CC_SC1 = 0, CC_SC2 = 256, CC_AW = 384, CC_K11 = 640, CC_K21 = 896, CC_K12 = 1152, CC_K22 = 1280, CC_SUMA = 1408, CC_SUMB = 1664;

// ---- per-row bookkeeping shared by the forward-recompute and the chain kernel ----
struct BRow {
    bool store;        // the tile exists (its stash block may be written)
    bool listed;       // the row holds a sample of the compacted list
    bool active;       // ... that the field evaluated (inner in-box mask)
    bool in_scene;
    int img, slot;
    int64_t gs;        // global sample index img * rays * P + slot
    float x[3];        // (bent) sample position, object space
};

__device__ __forceinline__ void load_row(const PeBwdTcArgs& B, int64_t tile, int64_t tile_end, int m, int& img_cursor, BRow& r) {
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    r.store = tile < tile_end;
    r.listed = false; r.active = false; r.in_scene = true; r.img = 0; r.slot = -1; r.gs = 0;
    r.x[0] = r.x[1] = r.x[2] = 0.f;
    if (!r.store) return;
    while (tile >= B.tile_begin[img_cursor + 1]) ++img_cursor;
    r.img = img_cursor;
    const int P = ob.positions;
    const int64_t spi = (int64_t)A.rays * P;
    const int64_t e = (tile - B.tile_begin[r.img]) * PE_BWD_TILE + m;
    r.in_scene = A.ois ? A.ois[(int64_t)r.img * A.objects + A.k] != 0 : true;
    if (e >= B.slot_count[r.img]) return;
    r.slot = B.slot_list[(int64_t)r.img * spi + e];
    r.gs = (int64_t)r.img * spi + r.slot;
    r.listed = true;
    if (A.bent) {                 // sampled and bent by the forward recompute's pre-pass
        r.active = (A.flags[r.gs] & 2) != 0;
        if (r.active) { r.x[0] = A.bent[r.gs * 3]; r.x[1] = A.bent[r.gs * 3 + 1]; r.x[2] = A.bent[r.gs * 3 + 2]; }
    } else {
        const int ray = r.slot / P, p = r.slot - ray * P;
        const PeRay pr = pe_make_ray(ob, A.w2o + ((int64_t)r.img * A.objects + A.k) * 12, A.origins + (int64_t)r.img * 3,
                                     A.dirs + ((int64_t)r.img * A.rays + ray) * 3, r.in_scene);
        const float u = (A.perturb && !A.t_in) ? A.rand[r.gs] : 0.f;
        const float t = pe_sample_t_or(A.t_in, r.gs, pr, p, P, A.perturb != 0, u);
        float x[3];
        pe_position(pr, t, x);
        r.active = pe_in_box(ob, x);
        if (r.active) { r.x[0] = x[0]; r.x[1] = x[1]; r.x[2] = x[2]; }
    }
}

__device__ __forceinline__ void unpack8(const uint4 q, float* v) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint4 pack8_sat(const float* v) {
    uint4 q;
    q.x = pack_half2_sat(v[0], v[1]); q.y = pack_half2_sat(v[2], v[3]); q.z = pack_half2_sat(v[4], v[5]); q.w = pack_half2_sat(v[6], v[7]);
    return q;
}

// lane j ends up with the sum over the warp's 32 lanes of v[j] (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = upper ? v[i] : v[i + s];
            const float recv = __shfl_xor_sync(0xffffffffu, send, s);
            v[i] = (upper ? v[i + s] : v[i]) + recv;
        }
    }
    return v[0];
}

// weight producer of both kernels: streams `steps` operands ([n rows][32 k] slabs, `passes` weight passes each) per tile pair
struct StepSpec { int n, slabs; };
__device__ __forceinline__ StepSpec chain_step(int s) {
    // H6T (N'128,K'192) H3T (256,128) H0T L7T L6T L5T (256,256) L4encT (64,256) L4T L3T L2T L1T (256,256) L0T (64,256)
    StepSpec st; st.n = 256; st.slabs = 8;
    if (s == 0) { st.n = 128; st.slabs = 6; }
    else if (s == 1) { st.slabs = 4; }
    else if (s == 6 || s == 11) { st.n = 64; }
    return st;
}

// =====================================================================================================================
// 1. forward recompute with stash
// =====================================================================================================================
// fp16x3 form (activations and weights as hi + lo pairs, three MMAs per k-step), ONE tile per iteration: ReLU masks and activations
// agree with an fp32 evaluation to ~1e-6 -- the gradients' sums cancel so heavily that fp16-level differences of the recompute alone
// show up at the 1e-2 .. 1e-1 level on the position gradients (measured, profiles/r2_bwd_tc.md).
__device__ __forceinline__ void split_store8(unsigned char* hi_dst, unsigned char* lo_dst, const float* v, bool relu) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = v[2 * i], b = v[2 * i + 1];
        h[i] = relu ? relu_pack_half2(a, b) : pack_half2_sat(a, b);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        l[i] = pack_half2_sat(a - hf.x, b - hf.y);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo_dst) *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// MODE 0: trunk layer  y = relu(acc);  MODE 2: AdaIn layer  x = acc (stashed), y = relu(x * sc + sh)
// smem: hi / lo operand buffers (or NULL: last layer); st_y: stash chunk base of the layer (plain fp16 copy)
template <int MODE, int N>
__device__ __forceinline__ void fwd_epilogue(uint32_t tcol, unsigned char* a_hi, unsigned char* a_lo, int m, const float* __restrict__ c0s,
                                             const float* __restrict__ c1s, unsigned char* st_y, unsigned char* st_x, uint32_t* mask_words, bool store) {
    uint32_t v[2][32];
    tmem_ld32(tcol, v[0]);
#pragma unroll
    for (int c = 0; c < N / 32; ++c) {
        tmem_wait_ld_regs(v[c & 1]);
        if (c + 1 < N / 32) tmem_ld32(tcol + (c + 1) * 32, v[(c + 1) & 1]);
        float y[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = __uint_as_float(v[c & 1][q]);
        if (MODE == 2) {
            if (store) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(st_x + (c * 4 + cc) * CHUNK_BYTES + m * 16) = pack8_sat(y + 8 * cc);
            }
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
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 32; ++q) bits |= (y[q] > 0.f) ? (1u << q) : 0u;
        if (store) mask_words[c * PE_BWD_TILE + m] = bits;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int off = (c * 4 + cc) * CHUNK_BYTES + m * 16;
            if (a_hi) split_store8(a_hi + off, a_lo + off, y + 8 * cc, true);
            if (store) split_store8(st_y + off, nullptr, y + 8 * cc, true);
        }
    }
}

constexpr int CHAIN_THREADS = 256;       // producer, MMA, TMEM-alloc, spare + 4 epilogue warps
constexpr int FCHAIN_THREADS = 384;      // field chain: two epilogue groups (column halves of the same rows)
constexpr int FCHAIN_XCH = SMEM_BAR + 256, FCHAIN_SMEM = FCHAIN_XCH + 2048;     // + the halves' exchange area
constexpr uint32_t CHAIN_BAR = 1;

__global__ void __launch_bounds__(CHAIN_THREADS, 1) pe_bwd_fwd_kernel(const PeBwdTcArgs B, const int64_t tile0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;
    uint64_t* a_ready = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);
    unsigned char* ones = smem + SMEM_ONES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    const int64_t total = B.tile_begin[A.images];
    const int64_t tile_end = pe_min64(total, tile0 + B.tile_capacity);
    const int64_t tiles = tile_end > tile0 ? tile_end - tile0 : 0;
    constexpr int LAYERS = 10;                       // L0..L7, head 0, head 3

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
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const unsigned char* src = blob + L.tc_base;
                for (int l = 0; l < LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t bytes = (uint32_t)n * PE_TC_SLAB_K * 2;
                    for (int s = 0; s < slabs; ++s) {
                        for (int pass = 0; pass < 2; ++pass) {
                            mbar_wait(empty_bar + stage, phase ^ 1);
                            mbar_arrive_expect_tx(full_bar + stage, bytes);
                            bulk_copy_g2s(ring + stage * STAGE_BYTES, src + (int64_t)pass * L.tc_bytes_per_pass, bytes, full_bar + stage);
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
        if (elect_one()) {
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(smem); R.a_addr[1] = smem_u32(smem + A_BYTES);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = 2; R.x3 = 1;
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int l = 0; l < LAYERS; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    layer_spec(l, n, slabs, chunk0, has_bias);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, n);
                    const uint32_t lbo_b = (uint32_t)n * 16;
                    mbar_wait(a_ready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    mma_layer<false, 1, 1>(R, l, n, slabs, chunk0, has_bias, idesc, lbo_b);
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
        const int wq = warp & 3;
        const int m = (wq << 5) | lane;
        unsigned char* a_hi = smem;
        unsigned char* a_lo = smem + A_BYTES;
        float* cst = reinterpret_cast<float*>(a_hi + CST_BASE);
        const uint32_t taddr = tmem_base + (((uint32_t)wq * 32u) << 16);
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        Sync1 sync{acc_full, a_ready, 0u, lane, nullptr, nullptr, 0u};
        int img_cursor = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int64_t tile = tile0 + t;
            BRow r;
            load_row(B, tile, tile_end, m, img_cursor, r);
            unsigned char* st = B.stash + t * FS_BYTES;
            uint32_t* mask = reinterpret_cast<uint32_t*>(st + (int64_t)FS_MASK * CHUNK_BYTES);
            {
                // Fourier features (positional_encoder.py:59-64): exact argument reduction + SFU, like the forward kernels; hi + lo operand
                const float xn[3] = {__fdiv_rn(r.x[0], size[0]), __fdiv_rn(r.x[1], size[1]), __fdiv_rn(r.x[2], size[2])};
                float tp[3], tl[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float c_hi = 0.15915494f, c_lo = 6.4206382e-9f;
                    tp[a] = xn[a] * c_hi;
                    tl[a] = fmaf(xn[a], c_lo, fmaf(xn[a], c_hi, -tp[a]));
                }
                float enc[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 0) encode_half<0>(xn, tp, tl, enc); else encode_half<1>(xn, tp, tl, enc);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int off = (4 * h + c) * CHUNK_BYTES + m * 16;
                        split_store8(a_hi + PE_CHUNK0 * CHUNK_BYTES + off, a_lo + PE_CHUNK0 * CHUNK_BYTES + off, enc + 8 * c, false);
                        if (r.store) split_store8(st + FS_ENC * CHUNK_BYTES + off, nullptr, enc + 8 * c, false);
                    }
                }
            }
            sync.arrive_ready();
            float4 pre[2];
#pragma unroll 1
            for (int l = 0; l < LAYERS; ++l) {
                sync.wait_acc();
                if (l == 4) {
                    // the encoding columns are dead once L4 has run: they take this image's AdaIn scale / shift (sc1|sh1 512, sc2|sh2 256)
                    const float* a1 = A.aff1 + (int64_t)r.img * 512;
                    const float* a2 = A.aff2 + (int64_t)r.img * 256;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int i0 = (j * 128 + m) * 4;
                        pre[j] = i0 < 512 ? __ldg(reinterpret_cast<const float4*>(a1 + i0))
                                          : (i0 < 768 ? __ldg(reinterpret_cast<const float4*>(a2 + (i0 - 512))) : make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
                if (l == 7) named_bar_sync(CHAIN_BAR, 128);
                if (l < 8) fwd_epilogue<0, 256>(taddr, a_hi, a_lo, m, nullptr, nullptr, st + FS_H(l) * CHUNK_BYTES, nullptr, mask + 8 * l * PE_BWD_TILE, r.store);
                else if (l == 8) fwd_epilogue<2, 256>(taddr, a_hi, a_lo, m, cst + CST_SC1, cst + CST_SH1, st + FS_Y1 * CHUNK_BYTES, st + FS_X1 * CHUNK_BYTES,
                                                      mask + MASK_Y1 * PE_BWD_TILE, r.store);
                else fwd_epilogue<2, 128>(taddr, nullptr, nullptr, m, cst + CST_SC2, cst + CST_SH2, st + FS_Y2 * CHUNK_BYTES, st + FS_X2 * CHUNK_BYTES,
                                          mask + MASK_Y2 * PE_BWD_TILE, r.store);
                if (l == 4) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) *reinterpret_cast<float4*>(cst + (j * 128 + m) * 4) = pre[j];
                }
                if (l < LAYERS - 1) sync.arrive_ready();
            }
            tc_fence_before();
            named_bar_sync(CHAIN_BAR, 128);      // the constants are dead before the next tile's encoding overwrites them
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// =====================================================================================================================
// 2. dX chain
// =====================================================================================================================
// Numerics.  Gradients are badly scaled (compositing weights span many decades) and their sums over samples cancel heavily (d sin(512 x)/dx),
// so the chain runs in the fp32-class form of the forward's fp16x3 mode: every gradient operand is kept as a hi + lo fp16 pair (two operand
// buffers, ONE tile per iteration), weights as hi (+ lo: PE_BWD_CHAIN_WLO=1) slabs, two MMAs per k-step (G_hi W_hi + G_lo W_hi [+ G_hi W_lo]), and every ROW is
// normalised by its own power of two s_m (the chain is linear per row) so that its values sit at ~2^8 whatever the sample's weight.
// The stash copy of G (the M operand of dW, summed over rows) carries the call-wide scale S instead: G_norm * S / s_m, fp16 hi only.
struct ChainCtx {
    unsigned char *a_hi, *a_lo;
    float* cst;
    uint32_t taddr;
    int m, lane;
    int kS;            // S = 2^kS
    int lo_on;         // 0: diagnostic -- the lo halves of the gradient operands are zeroed (plain fp16 chain)
    int h;             // column half handled by this thread (field chain: two epilogue groups share every row), 0 otherwise
    int par;           // parity of the row-maximum exchange slots
    float* xch;        // [2][2][128] exchange area of the two halves of a row
    int km;            // exponent of the row's scale: the operand currently stored in the A buffers holds (true gradient) * 2^km
    float mop;         // largest magnitude of that stored row
};
// power of two that brings a row whose largest magnitude is `mx` to [128, 256), keeping the row exponent within +-100
__device__ __forceinline__ int renorm_exp(float mx, int km) {
    if (!(mx > 0.f) || !(mx < INFINITY)) return 0;
    int e;
    frexpf(mx, &e);
    return max(-100 - km, min(100 - km, 8 - e));
}

// the two threads of a row agree on the row's largest magnitude (NH == 1: one thread per row, nothing to do)
template <int NH>
__device__ __forceinline__ float row_max_exchange(ChainCtx& C, float mx) {
    if (NH == 1) return mx;
    float* slot = C.xch + C.par * 256;
    slot[C.h * 128 + C.m] = mx;
    named_bar_sync(CHAIN_BAR, 256);
    C.par ^= 1;
    return fmaxf(slot[C.m], slot[128 + C.m]);
}

// gradient operand (hi + lo fp16 pair) of 8 consecutive columns of row m, saturating
__device__ __forceinline__ void store_g8_hilo(unsigned char* a_hi, unsigned char* a_lo, int chunk, int m, const float* v, int lo_on = 1) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_half2_sat(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        l[i] = lo_on ? pack_half2_sat(v[2 * i] - hf.x, v[2 * i + 1] - hf.y) : 0u;
    }
    *reinterpret_cast<uint4*>(a_hi + chunk * CHUNK_BYTES + m * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_lo + chunk * CHUNK_BYTES + m * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

// G = mask ? acc : 0;  with_alpha: G = mask ? acc + graw * aw[c] : 0 (the alpha head joins at the trunk output)
template <int NB = 8, int NH = 1>
__device__ __forceinline__ void chain_epilogue_plain(ChainCtx& C, const uint32_t* __restrict__ mask_words, unsigned char* st_g, bool with_alpha,
                                                     float graw, bool store) {
    // NH == 2: this thread handles the NB / 2 column blocks of half C.h; C.mop leaves as this half's maximum (the caller exchanges it)
    constexpr int MB = NB / NH;
    const int c0 = NH == 1 ? 0 : C.h * MB;
    const int fexp = renorm_exp(C.mop, C.km);
    const int km = C.km + fexp;
    const float f = ldexpf(1.f, fexp - TCT_WEXP), r_stash = ldexpf(1.f, C.kS - km), graw_n = with_alpha ? ldexpf(graw, km) : 0.f;
    uint32_t bits[MB];
#pragma unroll
    for (int w = 0; w < MB; ++w) bits[w] = mask_words[(c0 + w) * PE_BWD_TILE + C.m];
    uint32_t v[2][32];
    float mx = 0.f;
    tmem_ld32(C.taddr + c0 * 32, v[0]);
#pragma unroll
    for (int cb = 0; cb < MB; ++cb) {
        const int c = c0 + cb;
        tmem_wait_ld_regs(v[cb & 1]);
        if (cb + 1 < MB) tmem_ld32(C.taddr + (c + 1) * 32, v[(cb + 1) & 1]);
        float y[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            float a = __uint_as_float(v[cb & 1][q]) * f;
            if (with_alpha) a = fmaf(graw_n, C.cst[CC_AW + c * 32 + q], a);
            y[q] = ((bits[cb] >> q) & 1u) ? a : 0.f;
            mx = fmaxf(mx, fabsf(y[q]));
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            store_g8_hilo(C.a_hi, C.a_lo, c * 4 + cc, C.m, y + 8 * cc, C.lo_on);
            if (store) {
                float z[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = y[8 * cc + i] * r_stash;
                const int off = (c * 4 + cc) * CHUNK_BYTES + C.m * 16;
                split_store8(st_g + off, nullptr, z, false);
            }
        }
    }
    C.km = km; C.mop = mx;
}

// AdaIn step (adain.py:51-61 backward): g = mask ? acc : 0;  per image A[c] += sum g, Bx[c] += sum g x;  G = g sc - k1 - x k2 on the rows the
// field evaluated (k1 = k2 = 0 in eval mode).  Two passes over the accumulators: the first finds the row's largest output (the BatchNorm
// terms are the same for every row, so a row with a tiny own gradient can grow by many decades here), the second stores it re-normalised.
// sums: accumulate A / Bx (true units) into the tile's shared sums.
template <int N, int NH = 1>
__device__ __forceinline__ void chain_epilogue_adain(ChainCtx& C, const uint32_t* __restrict__ mask_words, const unsigned char* st_x,
                                                     unsigned char* st_g, const float* sc, const float* k1, const float* k2, bool active,
                                                     bool sums, bool store) {
    constexpr int MB = N / 32 / NH;
    const int c0 = NH == 1 ? 0 : C.h * MB;
    uint32_t bits[MB];
#pragma unroll
    for (int w = 0; w < MB; ++w) bits[w] = mask_words[(c0 + w) * PE_BWD_TILE + C.m];
    float* sumA = C.cst + CC_SUMA;
    float* sumB = C.cst + CC_SUMB;
    int km = C.km + renorm_exp(C.mop, C.km);
    float mx = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) { km += renorm_exp(row_max_exchange<NH>(C, mx), km); mx = 0.f; }
        const float f = ldexpf(1.f, km - C.km - TCT_WEXP), s_row = ldexpf(1.f, km), inv_s_row = ldexpf(1.f, -km), r_stash = ldexpf(1.f, C.kS - km);
        uint32_t v[32];
#pragma unroll 1
        for (int cb = 0; cb < MB; ++cb) {
            const int c = c0 + cb;
            uint4 xq[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) xq[cc] = *reinterpret_cast<const uint4*>(st_x + (c * 4 + cc) * CHUNK_BYTES + C.m * 16);
            tmem_ld32(C.taddr + c * 32, v);
            tmem_wait_ld_regs(v);
            float x[32], g[32];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) unpack8(xq[cc], x + 8 * cc);
#pragma unroll
            for (int q = 0; q < 32; ++q) g[q] = ((bits[cb] >> q) & 1u) ? __uint_as_float(v[q]) * f : 0.f;
            if (sums && pass == 1) {
                float a[32], b[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) { a[q] = g[q] * inv_s_row; b[q] = a[q] * x[q]; }
                const float sa = warp_transpose_sum(a, C.lane);
                const float sb = warp_transpose_sum(b, C.lane);
                atomicAdd(sumA + c * 32 + C.lane, sa);
                atomicAdd(sumB + c * 32 + C.lane, sb);
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int col = c * 32 + q;
                g[q] = active ? fmaf(g[q], sc[col], -s_row * fmaf(x[q], k2[col], k1[col])) : 0.f;
                mx = fmaxf(mx, fabsf(g[q]));
            }
            if (pass == 1) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    store_g8_hilo(C.a_hi, C.a_lo, c * 4 + cc, C.m, g + 8 * cc, C.lo_on);
                    if (store) {
                        float z[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) z[i] = g[8 * cc + i] * r_stash;
                        const int off = (c * 4 + cc) * CHUNK_BYTES + C.m * 16;
                        split_store8(st_g + off, nullptr, z, false);
                    }
                }
            }
        }
    }
    C.km = km; C.mop = mx;
}

// phase 0: full chain (+ per-image AdaIn sums for the style backward); phase 1 / 2 (train mode): stop after the second / first AdaIn
// layer of the head (walking backwards) and accumulate the cross-sample sums of its BatchNorm backward
__global__ void __launch_bounds__(FCHAIN_THREADS, 1) pe_bwd_chain_kernel(const PeBwdTcArgs B, const int64_t tile0, const int phase, const int lo_on) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem + 2 * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;
    uint64_t* a_ready = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    const int64_t total = B.tile_begin[A.images];
    const int64_t tile_end = pe_min64(total, tile0 + B.tile_capacity);
    const int64_t tiles = tile_end > tile0 ? tile_end - tile0 : 0;
    const int steps = phase == 1 ? 1 : (phase == 2 ? 2 : 12);
    constexpr int W = 256;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(acc_full, 1); mbar_init(a_ready, 8);
        mbar_fence_init();
    }
    fence_proxy_async();
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const unsigned char* src = blob + L.tcT_base;
                for (int s = 0; s < steps; ++s) {
                    const StepSpec st = chain_step(s);
                    const uint32_t bytes = (uint32_t)st.n * PE_TC_SLAB_K * 2;
                    for (int k = 0; k < st.slabs; ++k) {
                        for (int pass = 0; pass < ((lo_on & 2) ? 2 : 1); ++pass) {
                            mbar_wait(empty_bar + stage, ph ^ 1);
                            mbar_arrive_expect_tx(full_bar + stage, bytes);
                            bulk_copy_g2s(ring + stage * STAGE_BYTES, src + (int64_t)pass * L.tcT_bytes_per_pass, bytes, full_bar + stage);
                            if (++stage == NUM_STAGES) { stage = 0; ph ^= 1; }
                        }
                        src += bytes;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(smem); R.a_addr[1] = smem_u32(smem + A_BYTES);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = (lo_on & 2) ? 2 : 1; R.x3 = 1;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int s = 0; s < steps; ++s) {
                    const StepSpec st = chain_step(s);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, st.n);
                    const uint32_t lbo_b = (uint32_t)st.n * 16;
                    mbar_wait(a_ready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    mma_layer<false, 1, 1>(R, s, st.n, st.slabs, 0, false, idesc, lbo_b);
                }
            }
        }
    } else if (warp >= 4) {
        // two epilogue groups of 4 warps share every row: group h handles half of the columns of each step
        ChainCtx C;
        C.lane = lane; C.lo_on = lo_on & 1;
        C.h = (warp - 4) >> 2; C.par = 0;
        C.xch = reinterpret_cast<float*>(smem + FCHAIN_XCH);
        C.m = ((warp & 3) << 5) | lane;
        C.a_hi = smem; C.a_lo = smem + A_BYTES;
        C.cst = reinterpret_cast<float*>(smem + CST_BASE);
        C.taddr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
        const float S = B.scale[0];
        const float max_aw = B.scale[2];
        {
            int e;
            frexpf(S, &e);               // S = 2^(e - 1)
            C.kS = e - 1;
        }
        const int m = C.m, h = C.h, tid = h * 128 + m;
        float* cst = C.cst;
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        const float* alpha_w = reinterpret_cast<const float*>(blob + L.alpha_w);
        const int P = ob.positions, F = ob.features;
        Sync1 sync{acc_full, a_ready, 0u, lane, nullptr, nullptr, 0u};
        int img_cursor = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int64_t tile = tile0 + t;
            BRow r;
            load_row(B, tile, tile_end, m, img_cursor, r);
            unsigned char* st = B.stash + t * FS_BYTES;
            const uint32_t* mask = reinterpret_cast<const uint32_t*>(st + (int64_t)FS_MASK * CHUNK_BYTES);
            const bool store = phase == 0;
            // ---- constants of this image: AdaIn scales, BatchNorm fix terms, alpha-head weights; zeroed sums ----
            {
                const float* sc1 = A.aff1 + (int64_t)r.img * 2 * W;
                const float* sc2 = A.aff2 + (int64_t)r.img * W;
                {
                    const int i = tid;          // 256 threads, W = 256
                    cst[CC_SC1 + i] = sc1[i];
                    cst[CC_AW + i] = __ldg(alpha_w + i);
                    cst[CC_K11 + i] = B.bn_fix[i];
                    cst[CC_K21 + i] = B.bn_fix[W + i];
                    cst[CC_SUMA + i] = 0.f; cst[CC_SUMB + i] = 0.f;
                }
                if (h == 0) {
                    cst[CC_SC2 + m] = sc2[m];
                    cst[CC_K12 + m] = B.bn_fix[2 * W + m];
                    cst[CC_K22 + m] = B.bn_fix[2 * W + W / 2 + m];
                }
            }
            // ---- upstream gradient of the per-sample features: cw_obj dL/dF_obj[ray] + cw_glob dL/dF_glob[ray], row-normalised -> operand,
            //      call-scaled -> stash ----
            float graw = 0.f;
            {
                float cwo = 0.f, cwg = 0.f;
                const float* gfo = nullptr;
                const float* gfg = nullptr;
                if (r.active) {
                    const int64_t ray = (int64_t)r.img * A.rays + r.slot / P;
                    if (B.g_feat_obj) { cwo = B.cw_obj[r.gs]; gfo = B.g_feat_obj + ray * F; }
                    if (B.g_feat_glob) { cwg = B.cw_glob[r.gs]; gfg = B.g_feat_glob + ray * F; }
                    if (r.in_scene) graw = B.g_raw[r.gs];
                }
                auto feature_grad8 = [&](int c, float* v) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    if (gfo) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(gfo + c * 8)), b = __ldg(reinterpret_cast<const float4*>(gfo + c * 8 + 4));
                        v[0] = cwo * a.x; v[1] = cwo * a.y; v[2] = cwo * a.z; v[3] = cwo * a.w; v[4] = cwo * b.x; v[5] = cwo * b.y; v[6] = cwo * b.z; v[7] = cwo * b.w;
                    }
                    if (gfg) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(gfg + c * 8)), b = __ldg(reinterpret_cast<const float4*>(gfg + c * 8 + 4));
                        v[0] = fmaf(cwg, a.x, v[0]); v[1] = fmaf(cwg, a.y, v[1]); v[2] = fmaf(cwg, a.z, v[2]); v[3] = fmaf(cwg, a.w, v[3]);
                        v[4] = fmaf(cwg, b.x, v[4]); v[5] = fmaf(cwg, b.y, v[5]); v[6] = fmaf(cwg, b.z, v[6]); v[7] = fmaf(cwg, b.w, v[7]);
                    }
                };
                float mx = fabsf(graw) * max_aw;
#pragma unroll 2
                for (int c = 12 * h; c < 12 * h + 12; ++c) {
                    float v[8];
                    feature_grad8(c, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const float av = fabsf(v[i]); if (av < INFINITY) mx = fmaxf(mx, av); }
                }
                mx = row_max_exchange<2>(C, mx);
                C.km = renorm_exp(mx, 0);                  // mx * 2^km in [128, 256)
                C.mop = ldexpf(mx, C.km);
                const float s_row = ldexpf(1.f, C.km);
#pragma unroll 2
                for (int c = 12 * h; c < 12 * h + 12; ++c) {
                    float v[8], z[8];
                    feature_grad8(c, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { z[i] = v[i] * S; v[i] *= s_row; }
                    store_g8_hilo(C.a_hi, C.a_lo, c, m, v, C.lo_on);
                    if (store && r.store) {
                        const int off = (FS_GF + c) * CHUNK_BYTES + m * 16;
                        split_store8(st + off, nullptr, z, false);
                    }
                }
                if (store && r.store && h == 0) {          // columns 0, 1 of a 16-column operand: S * dL/d raw alpha as hi, lo (d alpha_head.weight = h7^T graw)
                    const float gs = graw * S;
                    const float hi = __half2float(__float2half_rn(fminf(fmaxf(gs, -65504.f), 65504.f)));
                    float v[8] = {hi, gs - hi, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    *reinterpret_cast<uint4*>(st + FS_GRAW * CHUNK_BYTES + m * 16) = pack8_sat(v);
                    *reinterpret_cast<uint4*>(st + (FS_GRAW + 1) * CHUNK_BYTES + m * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            if (phase == 0 && B.gw.alpha_b && h == 0) {            // d alpha_head.bias = sum graw
                float s = graw;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && s != 0.f) atomicAdd(B.gw.alpha_b, s);
            }
            sync.arrive_ready();
            named_bar_sync(CHAIN_BAR, 256);                // constants visible to the group
            float* asum = B.adain_sums + (int64_t)r.img * 3 * W;
            const bool st_ok = store && r.store;

            auto flush_sums = [&](int N, int off_a, int off_b, const float* sc, double* bn, bool to_bn) {
                named_bar_sync(CHAIN_BAR, 256);
                for (int c = tid; c < N; c += 256) {
                    const float a = cst[CC_SUMA + c], b = cst[CC_SUMB + c];
                    if (r.store) {
                        if (to_bn) {            // cross-sample terms of the train-mode BatchNorm backward: S1 = sum g sc, S2 = sum g sc x
                            atomicAdd(bn + c, (double)(a * sc[c]));
                            atomicAdd(bn + N + c, (double)(b * sc[c]));
                        } else {
                            if (a != 0.f) atomicAdd(asum + off_a + c, a);
                            if (b != 0.f) atomicAdd(asum + off_b + c, b);
                        }
                    }
                    cst[CC_SUMA + c] = 0.f; cst[CC_SUMB + c] = 0.f;
                }
                named_bar_sync(CHAIN_BAR, 256);
            };

            // ---- step 0: gy2 = gF H6 -> AdaIn 2 ----
            sync.wait_acc();
            chain_epilogue_adain<128, 2>(C, mask + MASK_Y2 * PE_BWD_TILE, st + FS_X2 * CHUNK_BYTES, st + FS_GX2 * CHUNK_BYTES, cst + CC_SC2, cst + CC_K12,
                                      cst + CC_K22, r.active, phase != 2, st_ok);
            if (phase != 2) flush_sums(W / 2, 2 * W, 2 * W + W / 2, cst + CC_SC2, B.bn_sums + 2 * W, phase == 1);
            if (phase == 1) { tc_fence_before(); named_bar_sync(CHAIN_BAR, 256); continue; }
            sync.arrive_ready();
            C.mop = row_max_exchange<2>(C, C.mop);
            // ---- step 1: gy1 = gx2 H3 -> AdaIn 1 ----
            sync.wait_acc();
            chain_epilogue_adain<256, 2>(C, mask + MASK_Y1 * PE_BWD_TILE, st + FS_X1 * CHUNK_BYTES, st + FS_GX1 * CHUNK_BYTES, cst + CC_SC1, cst + CC_K11,
                                      cst + CC_K21, r.active, true, st_ok);
            flush_sums(W, 0, W, cst + CC_SC1, B.bn_sums, phase == 2);
            if (phase == 2) { tc_fence_before(); named_bar_sync(CHAIN_BAR, 256); continue; }
            sync.arrive_ready();
            C.mop = row_max_exchange<2>(C, C.mop);
            // ---- steps 2-10 share ONE copy of the plain epilogue (instruction fetch: profiles/r2_bwd_tc.md) ----
            //   step 2: gh7 = gx1 H0 + graw alpha_w -> relu'(h7) (the alpha head joins at the trunk output)
            //   steps 3-5: trunk layers 7, 6, 5 -> gradients of the pre-activations of layers 6, 5, 4
            //   step 6: the encoding half of the skip layer's input gradient, parked (hi + lo) in the encoding columns
            //   steps 7-10: trunk layers 4 (hidden half), 3, 2, 1 -> gradients of the pre-activations of layers 3, 2, 1, 0
            int km_parked = 0;
#pragma unroll 1
            for (int step = 2; step <= 10; ++step) {
                sync.wait_acc();
                if (step == 6) {
                    uint32_t v[32];
                    tmem_ld32(C.taddr + 32 * h, v);
                    tmem_wait_ld_regs(v);
                    named_bar_sync(CHAIN_BAR, 256);            // every reader of the constants is done
                    const int fexp = renorm_exp(C.mop, C.km);  // (the operand stays: the next step reads it again with the same factor)
                    km_parked = C.km + fexp;
                    const float f = ldexpf(1.f, fexp - TCT_WEXP);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        float y[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v[8 * cc + i]) * f;
                        store_g8_hilo(C.a_hi, C.a_lo, PE_CHUNK0 + h * 4 + cc, m, y);
                    }
                    sync.arrive_ready();
                    continue;
                }
                const int l = step < 6 ? 9 - step : 10 - step;         // layer whose pre-activation gradient this step produces
                chain_epilogue_plain<8, 2>(C, mask + 8 * l * PE_BWD_TILE, st + FS_GP(l) * CHUNK_BYTES, step == 2, graw, st_ok);
                sync.arrive_ready();
                C.mop = row_max_exchange<2>(C, C.mop);
            }
            // ---- step 11: encoding gradient = layer 0's input gradient + the parked half; positional_encoder.py:59-64 backward ----
            sync.wait_acc();
            {
                uint32_t v[32];
                tmem_ld32(C.taddr + 32 * h, v);
                tmem_wait_ld_regs(v);
                // true gradient = accumulator / 2^km + parked / 2^km_parked; this thread holds encoding columns 32 h .. 32 h + 31
                const float fa = ldexpf(1.f, -C.km - TCT_WEXP), fp = ldexpf(1.f, -km_parked);
                float ge[32];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float ph[8], pl[8];
                    unpack8(*reinterpret_cast<const uint4*>(C.a_hi + (PE_CHUNK0 + h * 4 + cc) * CHUNK_BYTES + m * 16), ph);
                    unpack8(*reinterpret_cast<const uint4*>(C.a_lo + (PE_CHUNK0 + h * 4 + cc) * CHUNK_BYTES + m * 16), pl);
#pragma unroll
                    for (int i = 0; i < 8; ++i) ge[cc * 8 + i] = fmaf(__uint_as_float(v[8 * cc + i]), fa, (ph[i] + pl[i]) * fp);
                }
                // positional_encoder.py:59-64 backward over this thread's columns: [x | sin, cos per octave], column e = 3 + 6 o + 3 fn + a
                float gx[3] = {0.f, 0.f, 0.f};
                if (r.active) {
                    const float xn[3] = {__fdiv_rn(r.x[0], size[0]), __fdiv_rn(r.x[1], size[1]), __fdiv_rn(r.x[2], size[2])};
                    auto columns = [&](auto half) {
                        constexpr int H = decltype(half)::value;
                        // rolled: 32 inlined sincosf bodies per half made this epilogue the largest piece of the kernel's code, and the
                        // kernel's instruction fetch stalls (profiles/r2_bwd_tc.md) are what this step can give back
#pragma unroll 1
                        for (int i = 0; i < 32; ++i) {
                            const int e = 32 * H + i;
                            if (e < 3) gx[e] += ge[i];
                            else if (e < 63) {
                                const int o = (e - 3) / 6, rr = (e - 3) % 6, fn = rr / 3, a = rr % 3;
                                const float f = (float)(1 << o);
                                float sn, cs;
                                sincosf(__fmul_rn(f, xn[a]), &sn, &cs);
                                gx[a] = fmaf(f * (fn ? -sn : cs), ge[i], gx[a]);
                            }
                        }
                    };
                    if (h == 0) columns(std::integral_constant<int, 0>{}); else columns(std::integral_constant<int, 1>{});
#pragma unroll
                    for (int a = 0; a < 3; ++a) gx[a] /= size[a];
                }
                // the two halves of the row meet in the exchange area (barrier: a slow thread may still be reading the last row maximum)
                named_bar_sync(CHAIN_BAR, 256);
                if (h == 1) { C.xch[m * 3] = gx[0]; C.xch[m * 3 + 1] = gx[1]; C.xch[m * 3 + 2] = gx[2]; }
                named_bar_sync(CHAIN_BAR, 256);
                if (h == 0 && r.listed) {
                    float* dst = (B.g_bent ? B.g_bent : B.g_pos) + r.gs * 3;
                    dst[0] = gx[0] + C.xch[m * 3]; dst[1] = gx[1] + C.xch[m * 3 + 1]; dst[2] = gx[2] + C.xch[m * 3 + 2];
                }
            }
            tc_fence_before();
            named_bar_sync(CHAIN_BAR, 256);        // the parked gradient is dead before the next tile's constants overwrite it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// =====================================================================================================================
// 2b. ray bender (model/nerf_models/positional_ray_bender_model.py:81-163), shipped shape: 6 x 128, skip at 3, 6 annealed octaves + 32
//     deformation features.  Same three-kernel scheme on its own stash block (PE_BWD_BS_CHUNKS chunks per tile):
//     recompute (fp16x3 with all four partial products, like pe_bender_tc_kernel) -> dX chain (from dL/d bent position, left by the
//     field's chain, and dL/d |displacement|) -> dW through pe_bwd_dw_kernel with the bender's item table.
// =====================================================================================================================
__host__ __device__ constexpr int BS_H(int l) { return l < 3 ? 16 * l : 60 + 16 * (l - 3); }     // bh0..bh2 | input (12) | bh3..bh5: [bh2 | input] contiguous
constexpr int BS_ENC = 48;                                     // activations [0, 108)
constexpr int BS_GOUT = 108;                                   // 4 chunks (3 real columns); a 128-row block over-reads into GP(0)
__host__ __device__ constexpr int BS_GP(int l) { return 112 + 16 * l; }   // gradients [108, 208)
constexpr int BS_MASK = 208, BS_AUX = 214;                     // mask words [layer][4][row]; per row: clamp bits, displacement xyz (fp32)
static_assert(BS_AUX + 1 == PE_BWD_BS_CHUNKS, "bender stash map");
constexpr int64_t BS_BYTES = (int64_t)PE_BWD_BS_CHUNKS * CHUNK_BYTES;
constexpr int BB_A_CHUNKS = 28, BB_ENC_CHUNK0 = 16;            // operand buffers: K columns 0..127 activations, 128..223 the bender's input
constexpr int BB_A_BYTES = BB_A_CHUNKS * CHUNK_BYTES;
constexpr int BB_SMEM_BAR = 2 * BB_A_BYTES + NUM_STAGES * STAGE_BYTES;
constexpr int BB_SMEM_ONES = BB_SMEM_BAR + 128;
constexpr int BB_SMEM_TOTAL = BB_SMEM_ONES + 256;

__host__ __device__ __forceinline__ void bb_layer_spec(int l, int& n, int& slabs, int& chunk0, bool& has_bias) {
    n = 128; slabs = 4; chunk0 = 0; has_bias = true;
    if (l == 0) { slabs = 3; chunk0 = BB_ENC_CHUNK0; }
    else if (l == 3) { slabs = 7; }
    else if (l == 6) { n = 16; has_bias = false; }
}
// chain steps: OUTT (N'128,K'32) L5T L4T (128,128) L3encT (96,128) L3T L2T L1T (128,128) L0T (96,128)
__device__ __forceinline__ StepSpec bb_chain_step(int s) {
    StepSpec st; st.n = 128; st.slabs = 4;
    if (s == 0) st.slabs = 1;
    else if (s == 3 || s == 7) st.n = 96;
    return st;
}

// row of a bender tile: the sample BEFORE bending (the list holds exactly the samples inside the box)
__device__ __forceinline__ void load_row_prebend(const PeBwdTcArgs& B, int64_t tile, int64_t tile_end, int m, int& img_cursor, BRow& r) {
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    r.store = tile < tile_end;
    r.listed = false; r.active = false; r.in_scene = true; r.img = 0; r.slot = -1; r.gs = 0;
    r.x[0] = r.x[1] = r.x[2] = 0.f;
    if (!r.store) return;
    while (tile >= B.tile_begin[img_cursor + 1]) ++img_cursor;
    r.img = img_cursor;
    const int P = ob.positions;
    const int64_t spi = (int64_t)A.rays * P;
    const int64_t e = (tile - B.tile_begin[r.img]) * PE_BWD_TILE + m;
    r.in_scene = A.ois ? A.ois[(int64_t)r.img * A.objects + A.k] != 0 : true;
    if (e >= B.slot_count[r.img]) return;
    r.slot = B.slot_list[(int64_t)r.img * spi + e];
    r.gs = (int64_t)r.img * spi + r.slot;
    r.listed = true;
    const int ray = r.slot / P, p = r.slot - ray * P;
    const PeRay pr = pe_make_ray(ob, A.w2o + ((int64_t)r.img * A.objects + A.k) * 12, A.origins + (int64_t)r.img * 3,
                                 A.dirs + ((int64_t)r.img * A.rays + ray) * 3, r.in_scene);
    const float u = (A.perturb && !A.t_in) ? A.rand[r.gs] : 0.f;
    const float t = pe_sample_t_or(A.t_in, r.gs, pr, p, P, A.perturb != 0, u);
    pe_position(pr, t, r.x);
    r.active = pe_in_box(ob, r.x);
}

__global__ void __launch_bounds__(CHAIN_THREADS, 1) pe_bwd_bfwd_kernel(const PeBwdTcArgs B, const int64_t tile0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + BB_A_BYTES;
    unsigned char* ring = smem + 2 * BB_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BB_SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;
    uint64_t* a_ready = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    unsigned char* ones = smem + BB_SMEM_ONES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    const int64_t total = B.tile_begin[A.images];
    const int64_t tile_end = pe_min64(total, tile0 + B.tile_capacity);
    const int64_t tiles = tile_end > tile0 ? tile_end - tile0 : 0;

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
    if (warp == 2) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const unsigned char* src = blob + L.tcb_base;
                for (int l = 0; l < 7; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    bb_layer_spec(l, n, slabs, chunk0, has_bias);
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
        if (elect_one()) {
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(a_hi); R.a_addr[1] = smem_u32(a_lo);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = 2; R.x3 = 2;
            const uint64_t ones_desc = umma_smem_desc(smem_u32(ones), 128, 0);
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int l = 0; l < 7; ++l) {
                    int n, slabs, chunk0; bool has_bias;
                    bb_layer_spec(l, n, slabs, chunk0, has_bias);
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
        const int wq = warp & 3;
        const int m = (wq << 5) | lane;
        const uint32_t taddr = tmem_base + (((uint32_t)wq * 32u) << 16);
        Sync1 sync{acc_full, a_ready, 0u, lane, nullptr, nullptr, 0u};
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        int img_cursor = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            BRow r;
            load_row_prebend(B, tile0 + t, tile_end, m, img_cursor, r);
            unsigned char* st = B.bstash + t * BS_BYTES;
            uint32_t* mask = reinterpret_cast<uint32_t*>(st + (int64_t)BS_MASK * CHUNK_BYTES);
            const bool use = r.listed && r.active;
            {
                // the bender's input: annealed Fourier features of x / size (positional_ray_bender_model.py:96-100) | deformation code
                const float xn[3] = {__fdiv_rn(r.x[0], size[0]), __fdiv_rn(r.x[1], size[1]), __fdiv_rn(r.x[2], size[2])};
                const float* dfm = A.deformation + (int64_t)r.img * 32;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = 8 * c + i;
                        v[i] = !use ? 0.f : (e < 39 ? pe_encoding_value(xn, 3, e, ob.b_anneal) : (e < 71 ? __ldg(dfm + (e - 39)) : 0.f));
                    }
                    const int off = c * CHUNK_BYTES + m * 16;
                    split_store8(a_hi + BB_ENC_CHUNK0 * CHUNK_BYTES + off, a_lo + BB_ENC_CHUNK0 * CHUNK_BYTES + off, v, false);
                    if (r.store) split_store8(st + BS_ENC * CHUNK_BYTES + off, nullptr, v, false);
                }
            }
            sync.arrive_ready();
#pragma unroll 1
            for (int l = 0; l < 6; ++l) {
                sync.wait_acc();
                uint32_t v[2][32];
                tmem_ld32(taddr, v[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_wait_ld_regs(v[c & 1]);
                    if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                    float y[32];
                    uint32_t bits = 0;
#pragma unroll
                    for (int q = 0; q < 32; ++q) { y[q] = __uint_as_float(v[c & 1][q]); bits |= (y[q] > 0.f) ? (1u << q) : 0u; }
                    if (r.store) mask[(l * 4 + c) * PE_BWD_TILE + m] = bits;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int off = (c * 4 + cc) * CHUNK_BYTES + m * 16;
                        split_store8(a_hi + off, a_lo + off, y + 8 * cc, true);
                        if (r.store) split_store8(st + BS_H(l) * CHUNK_BYTES + off, nullptr, y + 8 * cc, true);
                    }
                }
                sync.arrive_ready();
            }
            sync.wait_acc();
            uint32_t v[16];
            tmem_ld16(taddr, v);
            tmem_wait_ld_regs16(v);
            tc_fence_before();
            // displacement = clamp(out * size) into the box (clamp_output :116-140): remember which axes the network output still moves
            int cf = 0;
            float dsp[3] = {0.f, 0.f, 0.f};
            if (use) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float raw = __fmul_rn(__uint_as_float(v[a]), size[a]);
                    const float lo = __fsub_rn(ob.bbox[2 * a], r.x[a]), hi = __fsub_rn(ob.bbox[2 * a + 1], r.x[a]);
                    float d = fmaxf(raw, lo);
                    const bool pass = !(raw < lo) && !(d > hi);
                    d = fminf(d, hi);
                    if (ob.canonical_pose) d = __fmul_rn(d, 0.f);
                    if (pass) cf |= 1 << a;
                    dsp[a] = d;
                }
            }
            if (r.store)
                *reinterpret_cast<float4*>(st + (int64_t)BS_AUX * CHUNK_BYTES + m * 16) = make_float4(__int_as_float(cf), dsp[0], dsp[1], dsp[2]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

__global__ void __launch_bounds__(CHAIN_THREADS, 1) pe_bwd_bchain_kernel(const PeBwdTcArgs B, const int64_t tile0, const int lo_on) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem + 2 * BB_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BB_SMEM_BAR);
    uint64_t* empty_bar = full_bar + NUM_STAGES;
    uint64_t* acc_full = empty_bar + NUM_STAGES;
    uint64_t* a_ready = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    const int64_t total = B.tile_begin[A.images];
    const int64_t tile_end = pe_min64(total, tile0 + B.tile_capacity);
    const int64_t tiles = tile_end > tile0 ? tile_end - tile0 : 0;
    constexpr int STEPS = 8;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(acc_full, 1); mbar_init(a_ready, 4);
        mbar_fence_init();
    }
    fence_proxy_async();
    if (warp == 2) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const unsigned char* src = blob + L.tcbT_base;
                for (int s = 0; s < STEPS; ++s) {
                    const StepSpec st = bb_chain_step(s);
                    const uint32_t bytes = (uint32_t)st.n * PE_TC_SLAB_K * 2;
                    for (int k = 0; k < st.slabs; ++k) {
                        for (int pass = 0; pass < ((lo_on & 2) ? 2 : 1); ++pass) {
                            mbar_wait(empty_bar + stage, ph ^ 1);
                            mbar_arrive_expect_tx(full_bar + stage, bytes);
                            bulk_copy_g2s(ring + stage * STAGE_BYTES, src + (int64_t)pass * L.tcbT_bytes_per_pass, bytes, full_bar + stage);
                            if (++stage == NUM_STAGES) { stage = 0; ph ^= 1; }
                        }
                        src += bytes;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t ready_phase = 0;
            MmaRing R;
            R.full_bar = full_bar; R.empty_bar = empty_bar; R.acc_full = acc_full;
            R.a_addr[0] = smem_u32(smem); R.a_addr[1] = smem_u32(smem + BB_A_BYTES);
            R.ring_addr = smem_u32(ring); R.tmem_base = tmem_base;
            R.stage = 0; R.phase = 0; R.num_passes = (lo_on & 2) ? 2 : 1; R.x3 = 1;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int s = 0; s < STEPS; ++s) {
                    const StepSpec st = bb_chain_step(s);
                    const uint32_t idesc = umma_idesc_f16(TILE_M, st.n);
                    const uint32_t lbo_b = (uint32_t)st.n * 16;
                    mbar_wait(a_ready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    mma_layer<false, 1, 1>(R, s, st.n, st.slabs, 0, false, idesc, lbo_b);
                }
            }
        }
    } else if (warp >= 4) {
        ChainCtx C;
        C.lane = lane; C.lo_on = lo_on & 1;
        C.m = ((warp & 3) << 5) | lane;
        C.a_hi = smem; C.a_lo = smem + BB_A_BYTES;
        C.cst = nullptr;
        C.taddr = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
        const float S = B.scale[4];                      // the bender's own call-wide scale (pe_launch_bwd_scale_bender)
        {
            int e;
            frexpf(S, &e);
            C.kS = e - 1;
        }
        const int m = C.m;
        const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
        Sync1 sync{acc_full, a_ready, 0u, lane, nullptr, nullptr, 0u};
        int img_cursor = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            BRow r;
            load_row_prebend(B, tile0 + t, tile_end, m, img_cursor, r);
            unsigned char* st = B.bstash + t * BS_BYTES;
            const uint32_t* mask = reinterpret_cast<const uint32_t*>(st + (int64_t)BS_MASK * CHUNK_BYTES);
            const bool use = r.listed && r.active;
            // ---- dL/d displacement -> dL/d network output (inside the clamp) or -> dL/d position (clamped to a box face) ----
            float gpos[3] = {0.f, 0.f, 0.f}, gout[3] = {0.f, 0.f, 0.f};
            if (r.listed) {
                const float* gb = B.g_bent + r.gs * 3;
                gpos[0] = gb[0]; gpos[1] = gb[1]; gpos[2] = gb[2];
            }
            if (use) {
                const float4 aux = *reinterpret_cast<const float4*>(st + (int64_t)BS_AUX * CHUNK_BYTES + m * 16);
                const int cf = __float_as_int(aux.x);
                const float d[3] = {aux.y, aux.z, aux.w};
                const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const float gdm = B.g_dm[r.gs];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float gd = 0.f;
                    if (!ob.canonical_pose) {
                        gd = gpos[a];
                        if (nrm > 0.f) gd = fmaf(gdm, d[a] / nrm, gd);       // torch.norm backward, 0 at the origin
                    }
                    if ((cf >> a) & 1) gout[a] = gd * size[a];
                    else gpos[a] -= gd;
                }
            }
            {
                const float mx = fmaxf(fabsf(gout[0]), fmaxf(fabsf(gout[1]), fabsf(gout[2])));
                C.km = renorm_exp(mx, 0);
                C.mop = ldexpf(mx, C.km);
                const float s_row = ldexpf(1.f, C.km);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (c == 0) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) { v[a] = gout[a] * s_row; z[a] = gout[a] * S; }
                    }
                    store_g8_hilo(C.a_hi, C.a_lo, c, m, v, C.lo_on);
                    if (r.store) {
                        const int off = (BS_GOUT + c) * CHUNK_BYTES + m * 16;
                        split_store8(st + off, nullptr, z, false);
                    }
                }
            }
            sync.arrive_ready();
            // ---- steps 0-6 share one copy of the plain epilogue ----
            //   step 0: output head; steps 1, 2: layers 5, 4 -> gradients of the pre-activations of layers 5, 4, 3
            //   step 3: the input half of the skip layer's input gradient, parked (hi + lo) behind the activations
            //   steps 4-6: layers 3 (hidden half), 2, 1 -> gradients of the pre-activations of layers 2, 1, 0
            int km_parked = 0;
#pragma unroll 1
            for (int step = 0; step <= 6; ++step) {
                sync.wait_acc();
                if (step == 3) {
                    uint32_t v[3][32];
                    tmem_ld32(C.taddr, v[0]);
                    tmem_ld32(C.taddr + 32, v[1]);
                    tmem_ld32(C.taddr + 64, v[2]);
                    tmem_wait_ld_regs(v[0]);
                    tmem_wait_ld_regs(v[1]);
                    tmem_wait_ld_regs(v[2]);
                    const int fexp = renorm_exp(C.mop, C.km);
                    km_parked = C.km + fexp;
                    const float f = ldexpf(1.f, fexp - TCT_WEXP);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            float y[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v[c][8 * cc + i]) * f;
                            store_g8_hilo(C.a_hi, C.a_lo, BB_ENC_CHUNK0 + c * 4 + cc, m, y);
                        }
                    sync.arrive_ready();
                    continue;
                }
                const int l = step < 3 ? 5 - step : 6 - step;
                chain_epilogue_plain<4>(C, mask + l * 4 * PE_BWD_TILE, st + BS_GP(l) * CHUNK_BYTES, false, 0.f, r.store);
                sync.arrive_ready();
            }
            // ---- step 7: gradient of the bender's input = layer 0's input gradient + the parked half ----
            sync.wait_acc();
            {
                uint32_t v[3][32];
                tmem_ld32(C.taddr, v[0]);
                tmem_ld32(C.taddr + 32, v[1]);
                tmem_ld32(C.taddr + 64, v[2]);
                tmem_wait_ld_regs(v[0]);
                tmem_wait_ld_regs(v[1]);
                tmem_wait_ld_regs(v[2]);
                const float fa = ldexpf(1.f, -C.km - TCT_WEXP), fp = ldexpf(1.f, -km_parked);
                float ge[3][32];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        float ph[8], pl[8];
                        unpack8(*reinterpret_cast<const uint4*>(C.a_hi + (BB_ENC_CHUNK0 + c * 4 + cc) * CHUNK_BYTES + m * 16), ph);
                        unpack8(*reinterpret_cast<const uint4*>(C.a_lo + (BB_ENC_CHUNK0 + c * 4 + cc) * CHUNK_BYTES + m * 16), pl);
#pragma unroll
                        for (int i = 0; i < 8; ++i) ge[c][cc * 8 + i] = use ? fmaf(__uint_as_float(v[c][8 * cc + i]), fa, (ph[i] + pl[i]) * fp) : 0.f;
                    }
                if (use) {
                    // annealed positional encoding backward (annealable_positional_encoder.py:54-76): columns [x | sin, cos per octave] * weight
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float xn = __fdiv_rn(r.x[a], size[a]);
                        float gsum = ge[0][a];
#pragma unroll
                        for (int o = 0; o < 6; ++o) {
                            const float f = (float)(1 << o);
                            float s, c;
                            sincosf(__fmul_rn(f, xn), &s, &c);
                            const int is = 3 + 6 * o + a, ic = is + 3;
                            gsum = fmaf(ob.b_anneal[o] * f, c * ge[is >> 5][is & 31] - s * ge[ic >> 5][ic & 31], gsum);
                        }
                        gpos[a] += gsum / size[a];
                    }
                }
                if (B.g_deformation) {
                    // the deformation code is replicated over the samples of its image: column sums of input columns 39 .. 70
                    float* gd = B.g_deformation + (int64_t)r.img * 32;
#pragma unroll
                    for (int c = 1; c < 3; ++c) {
                        const float sum = warp_transpose_sum(ge[c], lane);
                        const int col = c * 32 + lane;
                        if (r.store && col >= 39 && col < 71 && sum != 0.f) atomicAdd(gd + (col - 39), sum);
                    }
                }
                if (r.listed) {
                    float* dst = B.g_pos + r.gs * 3;
                    dst[0] = gpos[0]; dst[1] = gpos[1]; dst[2] = gpos[2];
                }
            }
            tc_fence_before();
            named_bar_sync(CHAIN_BAR, 128);        // the parked gradient is dead before the next tile's operand overwrites it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// =====================================================================================================================
// 3. dW / db
// =====================================================================================================================
struct DwItem {
    int32_t g_chunk;       // stash chunk of the 128-column block of the M operand (the gradient; the trunk output for the alpha head)
    int32_t a_chunk;       // stash chunk of the N operand
    int32_t n;             // N operand columns (multiple of 16, <= 256)
    int32_t rows, cols;    // valid rows / columns of the 128 x n product
    int32_t ld;            // row stride of the output
    int32_t fold;          // add the valid columns of a row into one output element (alpha head: the N operand is [graw hi | graw lo])
    float* out;            // fp32 gradient, accumulated
    float* bias;           // fp32 bias gradient of the block's rows (column sums of the M operand) or NULL
};
struct DwArgs {
    const unsigned char* stash;
    const int32_t* tile_begin;
    int32_t images, tile_capacity;
    int64_t tile0, tile_bytes;
    const float* scale;            // [S, 1 / S] of the stash's gradients
    int32_t splits, bender;
    PeObjectParamGrads gw;
};
constexpr int DW_ITEMS = 25;

// Work item `base`:
//   base 0-15: trunk layer l = base / 2, output rows 128 (base % 2) ..;  16-17: the encoding columns of the skip layer;  18-19: head layer 0;
//   20: head layer 3;  21-22: head layer 6 (192 rows);  23-24: alpha head (M operand = trunk output, N operand = [graw hi | graw lo])
__device__ __forceinline__ bool dw_item(int base, const PeObjectParamGrads& gw, DwItem& it) {
    it.rows = 128; it.fold = 0; it.bias = nullptr; it.out = nullptr;
    float* bias = nullptr;
    if (base < 16) {
        const int l = base >> 1, mb = base & 1;
        it.ld = l == 0 ? 63 : (l == 4 ? 319 : 256);
        it.g_chunk = FS_GP(l) + 16 * mb;
        it.a_chunk = l == 0 ? FS_ENC : FS_H(l - 1);
        it.n = l == 0 ? 64 : 256; it.cols = l == 0 ? 63 : 256;
        if (gw.backbone_w[l]) it.out = gw.backbone_w[l] + (int64_t)mb * 128 * it.ld;
        if (gw.backbone_b[l]) bias = gw.backbone_b[l] + mb * 128;
    } else if (base < 18) {
        const int mb = base - 16;
        it.ld = 319; it.g_chunk = FS_GP(4) + 16 * mb; it.a_chunk = FS_ENC; it.n = 64; it.cols = 63;
        if (gw.backbone_w[4]) it.out = gw.backbone_w[4] + (int64_t)mb * 128 * 319 + 256;
    } else if (base < 20) {
        const int mb = base - 18;
        it.ld = 256; it.g_chunk = FS_GX1 + 16 * mb; it.a_chunk = FS_H(7); it.n = 256; it.cols = 256;
        if (gw.head0_w) it.out = gw.head0_w + (int64_t)mb * 128 * 256;
    } else if (base == 20) {
        it.ld = 256; it.g_chunk = FS_GX2; it.a_chunk = FS_Y1; it.n = 256; it.cols = 256;
        it.out = gw.head3_w;
    } else if (base < 23) {
        const int mb = base - 21;
        it.ld = 128; it.g_chunk = FS_GF + 16 * mb; it.a_chunk = FS_Y2; it.n = 128; it.cols = 128; it.rows = mb ? 64 : 128;
        if (gw.head6_w) it.out = gw.head6_w + (int64_t)mb * 128 * 128;
        if (gw.head6_b) bias = gw.head6_b + mb * 128;
    } else {
        const int mb = base - 23;
        it.ld = 1; it.g_chunk = FS_H(7) + 16 * mb; it.a_chunk = FS_GRAW; it.n = 16; it.cols = 2; it.fold = 1;
        if (gw.alpha_w) it.out = gw.alpha_w + mb * 128;
        return it.out != nullptr;
    }
    it.bias = bias;
    if (!it.out) it.cols = 0;
    return it.out != nullptr || it.bias != nullptr;
}
// the ray bender's products: base 0-5: layer l (128 rows; the skip layer's N operand is [bh2 | input], contiguous in the stash); 6: output head
constexpr int DW_BENDER_ITEMS = 7;
__device__ __forceinline__ bool dw_item_bender(int base, const PeObjectParamGrads& gw, DwItem& it) {
    it.rows = 128; it.fold = 0; it.bias = nullptr; it.out = nullptr;
    float* bias = nullptr;
    if (base < 6) {
        const int l = base;
        it.g_chunk = BS_GP(l);
        it.a_chunk = l == 0 ? BS_ENC : BS_H(l - 1);
        it.n = l == 0 ? 96 : (l == 3 ? 224 : 128);
        it.cols = l == 0 ? 71 : (l == 3 ? 199 : 128);
        it.ld = it.cols;
        it.out = gw.bender_w[l];
        bias = gw.bender_b[l];
    } else {
        it.g_chunk = BS_GOUT; it.a_chunk = BS_H(5); it.n = 128; it.cols = 128; it.ld = 128; it.rows = 3;
        it.out = gw.bender_out_w;
    }
    it.bias = bias;
    if (!it.out) it.cols = 0;
    return it.out != nullptr || it.bias != nullptr;
}

constexpr int DW_THREADS = 192;                                // producer, MMA + TMEM, 4 epilogue warps
constexpr int DW_G_BYTES = 16 * CHUNK_BYTES, DW_A_BYTES = 32 * CHUNK_BYTES, DW_STAGE = DW_G_BYTES + DW_A_BYTES;
constexpr int DW_SMEM_ONES = 2 * DW_STAGE, DW_SMEM_BAR = DW_SMEM_ONES + 2 * CHUNK_BYTES, DW_SMEM_TOTAL = DW_SMEM_BAR + 128;

__global__ void __launch_bounds__(DW_THREADS, 1) pe_bwd_dw_kernel(const DwArgs D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + DW_SMEM_BAR);    // [2]
    uint64_t* empty_bar = full_bar + 2;                                      // [2]
    uint64_t* acc_full = empty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    unsigned char* ones = smem + DW_SMEM_ONES;                               // 128 rows x 16 columns, column 0 = 1
    const int warp = threadIdx.x >> 5;
    DwItem it;
    if (!(D.bender ? dw_item_bender(blockIdx.x, D.gw, it) : dw_item(blockIdx.x, D.gw, it))) return;
    const int64_t total = D.tile_begin[D.images];
    const int64_t tile_end = pe_min64(total, D.tile0 + D.tile_capacity);
    const int64_t count = tile_end > D.tile0 ? tile_end - D.tile0 : 0;
    const int64_t per = (count + D.splits - 1) / D.splits;
    const int64_t t0 = pe_min64((int64_t)blockIdx.y * per, count), t1 = pe_min64(t0 + per, count);
    if (t0 >= t1) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 128 * 16; i += DW_THREADS) {
        const int row = i >> 4, col = i & 15;
        reinterpret_cast<__half*>(ones + (col >> 3) * CHUNK_BYTES + row * 16)[col & 7] = __float2half_rn(col == 0 ? 1.f : 0.f);
    }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_bytes = (uint32_t)(it.n / 8) * CHUNK_BYTES;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = t0; t < t1; ++t) {
                const unsigned char* st = D.stash + t * D.tile_bytes;
                mbar_wait(empty_bar + stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + stage, DW_G_BYTES + a_bytes);
                bulk_copy_g2s(smem + stage * DW_STAGE, st + (int64_t)it.g_chunk * CHUNK_BYTES, DW_G_BYTES, full_bar + stage);
                bulk_copy_g2s(smem + stage * DW_STAGE + DW_G_BYTES, st + (int64_t)it.a_chunk * CHUNK_BYTES, a_bytes, full_bar + stage);
                if (++stage == 2) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t idesc = umma_idesc_f16_major(128, it.n, 1, 1);
            const uint32_t idesc_b = umma_idesc_f16_major(128, 16, 1, 1);
            for (int64_t t = t0; t < t1; ++t) {
                mbar_wait(full_bar + stage, phase);
                tc_fence_after();
                const uint32_t g_addr = smem_u32(smem + stage * DW_STAGE), a_addr = g_addr + DW_G_BYTES;
#pragma unroll
                for (int j = 0; j < 8; ++j) {                 // K = the tile's 128 samples: 16 rows (2 row groups of 128 B) per MMA
                    const uint64_t dg = umma_smem_desc(g_addr + j * 256, 128, CHUNK_BYTES);
                    const uint64_t da = umma_smem_desc(a_addr + j * 256, 128, CHUNK_BYTES);
                    const uint32_t accum = (t != t0 || j != 0) ? 1u : 0u;
                    umma_f16_ss(tmem_base, dg, da, idesc, accum);
                    if (it.bias) umma_f16_ss(tmem_base + 256, dg, umma_smem_desc(smem_u32(ones) + j * 256, 128, CHUNK_BYTES), idesc_b, accum);
                }
                umma_commit(empty_bar + stage);
                if (++stage == 2) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else {
        const int wq = warp & 3;
        const int lane = threadIdx.x & 31;
        const int row = (wq << 5) | lane;
        const uint32_t taddr = tmem_base + (((uint32_t)wq * 32u) << 16);
        const float invS = D.scale[1];
        mbar_wait(acc_full, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < it.n; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_wait_ld_regs16(v);
            if (row < it.rows) {
                if (it.fold) {
                    float val = 0.f;
#pragma unroll
                    for (int q = 0; q < 16; ++q) if (c0 + q < it.cols) val += __uint_as_float(v[q]);
                    val *= invS;
                    if (c0 == 0 && val != 0.f) atomicAdd(it.out + (int64_t)row * it.ld, val);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float val = __uint_as_float(v[q]) * invS;
                        if (c0 + q < it.cols && val != 0.f) atomicAdd(it.out + (int64_t)row * it.ld + c0 + q, val);
                    }
                }
            }
        }
        if (it.bias) {
            uint32_t v[16];
            tmem_ld16(taddr + 256, v);
            tmem_wait_ld_regs16(v);
            const float val = __uint_as_float(v[0]) * invS;
            if (row < it.rows && val != 0.f) atomicAdd(it.bias + row, val);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// gradient scale: S = 2^k with S * (largest upstream gradient magnitude) ~ 256 (fp16 operands: 256 x headroom below overflow,
// 2^-32 of it above the smallest subnormal)
// =====================================================================================================================
struct AbsmaxSegments {
    const float* x[4];
    int64_t n[4];
};

// one launch for every tensor whose largest magnitude sets a scale: blockIdx.y selects the segment
__global__ void pe_bwd_absmax_kernel(const __grid_constant__ AbsmaxSegments S, unsigned int* __restrict__ out) {
    const float* __restrict__ x = S.x[blockIdx.y];
    const int64_t n = S.n[blockIdx.y];
    if (!x || n == 0) return;
    float mx = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v == v && v < INFINITY) mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out + blockIdx.y, __float_as_uint(mx));
}

static int launch_absmax(const AbsmaxSegments& S, int segments, unsigned int* out, cudaStream_t stream) {
    int64_t most = 0;
    for (int i = 0; i < segments; ++i)
        if (S.x[i]) most = S.n[i] > most ? S.n[i] : most;
    if (most == 0) return PE_OK;
    pe_bwd_absmax_kernel<<<dim3((unsigned)pe_min64((most + 255) / 256, 592), segments), 256, 0, stream>>>(S, out);
    PE_LAUNCH_CHECK("pe_bwd_absmax_kernel");
    return PE_OK;
}

__global__ void pe_bwd_scale_kernel(const unsigned int* __restrict__ mx, float* __restrict__ scale) {
    // mx: [0] |dL/dF_obj|, [1] |dL/dF_glob|, [2] |dL/d raw alpha|, [3] |alpha_head.weight|
    const float m = fmaxf(__uint_as_float(mx[0]) + __uint_as_float(mx[1]), __uint_as_float(mx[2]) * __uint_as_float(mx[3]));
    float S = 1.f;
    if (m > 0.f) {
        int e = (int)floorf(log2f(256.f / m));
        e = max(-60, min(60, e));
        S = exp2f((float)e);
    }
    scale[0] = S;
    scale[1] = 1.f / S;
    scale[2] = __uint_as_float(mx[3]);
}

__global__ void pe_bwd_scale_bender_kernel(const unsigned int* __restrict__ mx, float msize, float* __restrict__ scale) {
    const float m = (__uint_as_float(mx[0]) + __uint_as_float(mx[1])) * msize;
    float S = 1.f;
    if (m > 0.f) {
        int e = (int)floorf(log2f(256.f / m));
        e = max(-60, min(60, e));
        S = exp2f((float)e);
    }
    scale[4] = S;
    scale[5] = 1.f / S;
}

// transposed operand of one chain step: element (n, k) = w[k * ld + col0 + n] (nn.Linear weight [out][in], n = input, k = output)
// every step of the transposed streams of a model in ONE launch (blockIdx.y = step): packing runs every training step
constexpr int TCT_TABLE = 24;
struct TcTTable {
    const float* w[TCT_TABLE];
    unsigned char* hi[TCT_TABLE];
    unsigned char* lo[TCT_TABLE];
    int32_t ld[TCT_TABLE], col0[TCT_TABLE], n_real[TCT_TABLE], N[TCT_TABLE], K[TCT_TABLE], k_real[TCT_TABLE];
};
__global__ void pe_tcT_pack_kernel(const __grid_constant__ TcTTable T) {
    const int it = blockIdx.y;
    const float* __restrict__ w = T.w[it];
    unsigned char* __restrict__ hi = T.hi[it];
    unsigned char* __restrict__ lo = T.lo[it];
    const int ld = T.ld[it], col0 = T.col0[it], n_real = T.n_real[it], N = T.N[it], K = T.K[it], k_real = T.k_real[it];
    const int total = N * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / N, n = i - k * N;
        // scaled by 2^TCT_WEXP: the lo half of a weight of typical size 0.03 would otherwise be a SUBNORMAL fp16 (quantum 6e-8 = 2e-6 of the
        // weight -- a systematic error the cancelling gradient sums amplify); the chain's epilogues fold the factor into the row exponent
        const float v = (n < n_real && k < k_real) ? ldexpf(w[(int64_t)k * ld + col0 + n], TCT_WEXP) : 0.f;
        const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        const int slab = k / PE_TC_SLAB_K, kk = k - slab * PE_TC_SLAB_K;
        const int64_t off = (int64_t)slab * N * PE_TC_SLAB_K * 2 + (int64_t)(kk >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
    }
}

}  // namespace

bool pe_bwd_tc_object_ok(const PeObjectDesc& ob) {
    const char* env = getenv("PE_BWD_TC");
    if (env && atoi(env) == 0) return false;
    return pe_tc_field_ok(ob);
}

int pe_tcT_pack(const PeObjectDesc& d, const PeLayout& L, const PeObjectParams& p, void* packed, cudaStream_t stream) {
    (void)d;
    if (!L.tcT_base) return PE_OK;
    TcTTable T = {};
    int n = 0, max_total = 0;
    auto add = [&](const float* w, int ld, int col0, int n_real, int N, int K, int k_real, unsigned char* hi, unsigned char* lo) {
        T.w[n] = w; T.hi[n] = hi; T.lo[n] = lo; T.ld[n] = ld; T.col0[n] = col0; T.n_real[n] = n_real; T.N[n] = N; T.K[n] = K; T.k_real[n] = k_real;
        ++n;
        max_total = N * K > max_total ? N * K : max_total;
    };
    unsigned char* hi = (unsigned char*)packed + L.tcT_base;
    unsigned char* lo = hi + L.tcT_bytes_per_pass;
    struct Item { const float* w; int ld, col0, n_real, N, K; };
    const Item items[12] = {
        {p.head6_w, 128, 0, 128, 128, 192},          {p.head3_w, 256, 0, 256, 256, 128},          {p.head0_w, 256, 0, 256, 256, 256},
        {p.backbone_w[7], 256, 0, 256, 256, 256},    {p.backbone_w[6], 256, 0, 256, 256, 256},    {p.backbone_w[5], 256, 0, 256, 256, 256},
        {p.backbone_w[4], 319, 256, 63, 64, 256},    {p.backbone_w[4], 319, 0, 256, 256, 256},    {p.backbone_w[3], 256, 0, 256, 256, 256},
        {p.backbone_w[2], 256, 0, 256, 256, 256},    {p.backbone_w[1], 256, 0, 256, 256, 256},    {p.backbone_w[0], 63, 0, 63, 64, 256}};
    int64_t off = 0;
    for (int s = 0; s < 12; ++s) {
        const Item& it = items[s];
        if (!it.w) { pe_set_error("missing parameter tensor for the transposed weight stream (step %d)", s); return PE_ERR_INVALID; }
        add(it.w, it.ld, it.col0, it.n_real, it.N, it.K, it.K, hi + off, lo + off);
        off += (int64_t)it.N * it.K * 2;
    }
    if (off != L.tcT_bytes_per_pass) { pe_set_error("internal: transposed weight stream size mismatch"); return PE_ERR_INVALID; }
    if (L.tcbT_base) {
        // ray bender: OUTT (N'128,K'32) L5T L4T (128,128) L3encT (96,128) L3T L2T L1T (128,128) L0T (96,128)
        unsigned char* bhi = (unsigned char*)packed + L.tcbT_base;
        unsigned char* blo = bhi + L.tcbT_bytes_per_pass;
        struct BItem { const float* w; int ld, col0, n_real, N, K, k_real; };
        const BItem bitems[8] = {
            {p.bender_out_w, 128, 0, 128, 128, 32, 3},   {p.bender_w[5], 128, 0, 128, 128, 128, 128}, {p.bender_w[4], 128, 0, 128, 128, 128, 128},
            {p.bender_w[3], 199, 128, 71, 96, 128, 128}, {p.bender_w[3], 199, 0, 128, 128, 128, 128}, {p.bender_w[2], 128, 0, 128, 128, 128, 128},
            {p.bender_w[1], 128, 0, 128, 128, 128, 128}, {p.bender_w[0], 71, 0, 71, 96, 128, 128}};
        int64_t boff = 0;
        for (int s = 0; s < 8; ++s) {
            const BItem& it = bitems[s];
            if (!it.w) { pe_set_error("missing ray-bender parameter tensor for the transposed weight stream (step %d)", s); return PE_ERR_INVALID; }
            add(it.w, it.ld, it.col0, it.n_real, it.N, it.K, it.k_real, bhi + boff, blo + boff);
            boff += (int64_t)it.N * it.K * 2;
        }
        if (boff != L.tcbT_bytes_per_pass) { pe_set_error("internal: transposed ray-bender weight stream size mismatch"); return PE_ERR_INVALID; }
    }
    static_assert(12 + 8 <= TCT_TABLE, "transposed-stream table too small");
    pe_tcT_pack_kernel<<<dim3((max_total + 255) / 256 > 64 ? 64 : (max_total + 255) / 256, n), 256, 0, stream>>>(T);
    PE_LAUNCH_CHECK("pe_tcT_pack_kernel");
    return PE_OK;
}

// bit 0: the dX chains carry hi + lo gradient operands (PE_BWD_CHAIN_LO=0, diagnostic: plain fp16, lo halves zeroed);
// bit 1: the transposed weight stream runs hi + lo (PE_BWD_CHAIN_WLO=1).  Default: hi only -- G_hi W + G_lo W, two MMAs per k-step
// and half the weight traffic out of L2 of these one-tile-per-CTA kernels: measured on the golden scenes (tests/gpu_chain_wlo.py) the
// gradients do not move (worst deviation from the exact fp32 backward 2.7e-3 / 4.7e-2 / 5.5e-3 / 3.1e-3 with the lo pass, 2.8e-3 /
// 4.8e-2 / 5.5e-3 / 3.0e-3 without: the recompute's ReLU masks set the error, not the weights' 11 bits), the dense cfg3 step drops
// from 38.4 to 36.7 ms.
static int chain_lo_on() {
    const char* env = getenv("PE_BWD_CHAIN_LO");
    const char* wenv = getenv("PE_BWD_CHAIN_WLO");
    return ((env && atoi(env) == 0) ? 0 : 1) | ((wenv && atoi(wenv) != 0) ? 2 : 0);
}

int pe_launch_bwd_fwd(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    const int grid = (int)pe_min64((int64_t)args.tile_capacity, sm_count);
    if (grid <= 0) return PE_OK;
    pe_bwd_fwd_kernel<<<grid, CHAIN_THREADS, SMEM_TOTAL, stream>>>(args, tile0);
    PE_LAUNCH_CHECK("pe_bwd_fwd_kernel");
    return PE_OK;
}

int pe_launch_bwd_chain(const PeBwdTcArgs& args, int64_t tile0, int phase, int sm_count, cudaStream_t stream) {
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FCHAIN_SMEM));
    const int grid = (int)pe_min64((int64_t)args.tile_capacity, sm_count);
    if (grid <= 0) return PE_OK;
    pe_bwd_chain_kernel<<<grid, FCHAIN_THREADS, FCHAIN_SMEM, stream>>>(args, tile0, phase, chain_lo_on());
    PE_LAUNCH_CHECK("pe_bwd_chain_kernel");
    return PE_OK;
}

int pe_launch_bwd_scale(const PeBwdTcArgs& args, float* scale, unsigned int* scratch, cudaStream_t stream) {
    const PeFieldArgs& A = args.f;
    PE_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(unsigned int), stream));
    const int64_t nf = (int64_t)A.images * A.rays * A.ob.features, ns = (int64_t)A.images * A.rays * A.ob.positions;
    AbsmaxSegments S{};
    S.x[0] = args.g_feat_obj, S.n[0] = nf;
    S.x[1] = args.g_feat_glob, S.n[1] = nf;
    S.x[2] = args.g_raw, S.n[2] = ns;
    S.x[3] = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(A.ob.packed) + A.L.alpha_w), S.n[3] = A.ob.width;
    if (int rc = launch_absmax(S, 4, scratch, stream)) return rc;
    pe_bwd_scale_kernel<<<1, 1, 0, stream>>>(scratch, scale);
    PE_LAUNCH_CHECK("pe_bwd_scale_kernel");
    return PE_OK;
}

int pe_launch_bwd_dw(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream) {
    DwArgs D = {};
    D.stash = args.stash; D.tile_begin = args.tile_begin; D.images = args.f.images; D.tile_capacity = args.tile_capacity;
    D.tile0 = tile0; D.scale = args.scale; D.gw = args.gw; D.tile_bytes = FS_BYTES;
    D.splits = (int)pe_min64(pe_min64(args.tile_capacity, 65535), (3 * sm_count + DW_ITEMS - 1) / DW_ITEMS);
    if (D.splits < 1) D.splits = 1;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_TOTAL));
    pe_bwd_dw_kernel<<<dim3(DW_ITEMS, D.splits), DW_THREADS, DW_SMEM_TOTAL, stream>>>(D);
    PE_LAUNCH_CHECK("pe_bwd_dw_kernel");
    return PE_OK;
}

// ---- ray bender ----
bool pe_bwd_tc_bender_ok(const PeObjectDesc& ob, const PeLayout& L) {
    const char* env = getenv("PE_BWD_TC_BENDER");
    if (env && atoi(env) == 0) return false;
    return pe_tc_bender_ok(ob) && L.tcb_base != 0 && L.tcbT_base != 0;
}

// call-wide scale of the bender's gradient stash from |dL/d bent position| (left by the field's chain) and |dL/d |displacement||
int pe_launch_bwd_scale_bender(const PeBwdTcArgs& args, float* scale, unsigned int* scratch, cudaStream_t stream) {
    const PeFieldArgs& A = args.f;
    PE_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(unsigned int), stream));
    const int64_t ns = (int64_t)A.images * A.rays * A.ob.positions;
    if (ns == 0) return PE_OK;
    AbsmaxSegments S{};
    S.x[0] = args.g_bent, S.n[0] = 3 * ns;
    S.x[1] = args.g_dm, S.n[1] = ns;
    if (int rc = launch_absmax(S, 2, scratch, stream)) return rc;
    const PeObjectDesc& ob = A.ob;
    const float msize = fmaxf(ob.bbox[1] - ob.bbox[0], fmaxf(ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]));
    pe_bwd_scale_bender_kernel<<<1, 1, 0, stream>>>(scratch, msize, scale);
    PE_LAUNCH_CHECK("pe_bwd_scale_bender_kernel");
    return PE_OK;
}

int pe_launch_bwd_bender(const PeBwdTcArgs& args, int64_t tile0, int sm_count, cudaStream_t stream) {
    const int grid = (int)pe_min64((int64_t)args.tile_capacity, sm_count);
    if (grid <= 0) return PE_OK;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_bfwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_SMEM_TOTAL));
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_bchain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BB_SMEM_TOTAL));
    pe_bwd_bfwd_kernel<<<grid, CHAIN_THREADS, BB_SMEM_TOTAL, stream>>>(args, tile0);
    PE_LAUNCH_CHECK("pe_bwd_bfwd_kernel");
    pe_bwd_bchain_kernel<<<grid, CHAIN_THREADS, BB_SMEM_TOTAL, stream>>>(args, tile0, chain_lo_on());
    PE_LAUNCH_CHECK("pe_bwd_bchain_kernel");
    DwArgs D = {};
    D.stash = args.bstash; D.tile_begin = args.tile_begin; D.images = args.f.images; D.tile_capacity = args.tile_capacity;
    D.tile0 = tile0; D.scale = args.scale + 4; D.gw = args.gw; D.tile_bytes = BS_BYTES; D.bender = 1;
    D.splits = (int)pe_min64(pe_min64(args.tile_capacity, 65535), (2 * sm_count + DW_BENDER_ITEMS - 1) / DW_BENDER_ITEMS);
    if (D.splits < 1) D.splits = 1;
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_TOTAL));
    pe_bwd_dw_kernel<<<dim3(DW_BENDER_ITEMS, D.splits), DW_THREADS, DW_SMEM_TOTAL, stream>>>(D);
    PE_LAUNCH_CHECK("pe_bwd_dw_kernel");
    return PE_OK;
}

size_t pe_bwd_tc_stash_bytes(int64_t tiles) { return (size_t)tiles * FS_BYTES; }

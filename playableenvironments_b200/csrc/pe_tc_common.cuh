// Device code shared by the two tensor-core field kernels (pe_field_tc.cu: one CTA per SM, lockstep tile pair;
// pe_field_tc2.cu: CTA pair with cta_group::2, epilogues overlapped with MMAs): operand layout, layer table, and the
// complete per-tile work of an epilogue group (sampling, Fourier features, layer epilogues, volume rendering).
#pragma once

#include "pe_kernels.cuh"
#include "pe_umma.cuh"

namespace pe_tc {
using namespace pe;

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 2048;                 // 8 K-columns of a 128-row operand: 16 row groups x 128 B
constexpr int A_CHUNKS = 40;                      // K columns 0..255: activations, 256..319: positional encoding
constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES;   // 80 KB per tile
constexpr int PE_CHUNK0 = 32;
constexpr int NUM_LAYERS = 11;                    // L0..L7, H0, H3, H6
constexpr int SCRATCH_STRIDE = 97;                // floats per row of the compositing scratch (bank-conflict free)
constexpr int SCR_T = 50 * 1024, SCR_SH = SCR_T + 512, SCR_W = SCR_SH + 512;   // byte offsets inside the A buffer
// per-tile constants staged in the (dead after L4) positional-encoding columns of the A buffer
constexpr int CST_BASE = PE_CHUNK0 * CHUNK_BYTES; // byte offset inside the A buffer
constexpr int CST_SC1 = 0, CST_SH1 = 256, CST_SC2 = 512, CST_SH2 = 640, CST_AW = 768;
// folded-head mode (see epilogue_tile): scratch in the same dead columns, float offsets from CST_BASE
constexpr int FOLD_T = 1280, FOLD_W = 1408, FOLD_WF = 1536, FOLD_SH = 1664, FOLD_PART = 1680, FOLD_PART_S = 2704;
constexpr int FOLD_K = 128;                       // width of the last hidden layer (input of head layer 6)

// layer l: N outputs, `slabs` K=32 weight slabs, first A chunk, bias added by the rank-1 MMA
__host__ __device__ __forceinline__ void layer_spec(int l, int& n, int& slabs, int& chunk0, bool& has_bias) {
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
__device__ __forceinline__ void store_a8_relu(unsigned char* a_base, int chunk, int m, const float* v) {
    uint4 q;
    q.x = relu_pack_half2(v[0], v[1]); q.y = relu_pack_half2(v[2], v[3]); q.z = relu_pack_half2(v[4], v[5]); q.w = relu_pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(a_base + chunk * CHUNK_BYTES + m * 16) = q;
}

// fp16x3 mode: activations are kept as hi + lo fp16 pairs (hi = fp16(max(v,0)), lo = fp16(max(v,0) - hi)) in two operand buffers.
// Packed conversions: one cvt(.relu).f16x2 per pair for hi, unpack, subtract, one cvt.f16x2 for lo.
__device__ __forceinline__ void store_a8_hilo(unsigned char* a_hi, unsigned char* a_lo, int chunk, int m, const float* v, bool relu) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = v[2 * i], b = v[2 * i + 1];
        h[i] = relu ? relu_pack_half2(a, b) : pack_half2(a, b);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        l[i] = pack_half2(a - hf.x, b - hf.y);
    }
    *reinterpret_cast<uint4*>(a_hi + chunk * CHUNK_BYTES + m * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_lo + chunk * CHUNK_BYTES + m * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

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
template <int MODE, int N, bool kHiLo = false>
__device__ __forceinline__ float hidden_epilogue(uint32_t tcol, unsigned char* abuf, int chunk0, int m, const float* __restrict__ c0s,
                                                 const float* __restrict__ c1s, unsigned char* abuf_lo = nullptr, float* h_out = nullptr) {
    // h_out (MODE 1, train-mode recompute for the backward): this row's post-ReLU trunk output in fp32, first of its N columns
    // tcol: TMEM address of the first of the N columns handled here; chunk0: A-operand chunk (= column / 8) they are stored to;
    // c0s / c1s: per-column constants, already offset to the first column
    uint32_t v[2][32];
    float alpha = 0.f;
    tmem_ld32(tcol, v[0]);
#pragma unroll
    for (int c = 0; c < N / 32; ++c) {
        tmem_wait_ld_regs(v[c & 1]);
        if (c + 1 < N / 32) tmem_ld32(tcol + (c + 1) * 32, v[(c + 1) & 1]);
        float y[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = __uint_as_float(v[c & 1][q]);
        if (MODE == 1 && h_out) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(h_out + c * 32 + 4 * q) =
                    make_float4(fmaxf(y[4 * q], 0.f), fmaxf(y[4 * q + 1], 0.f), fmaxf(y[4 * q + 2], 0.f), fmaxf(y[4 * q + 3], 0.f));
        }
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
        for (int cc = 0; cc < 4; ++cc) {
            if (kHiLo) store_a8_hilo(abuf, abuf_lo, chunk0 + c * 4 + cc, m, y + 8 * cc, true);
            else store_a8_relu(abuf, chunk0 + c * 4 + cc, m, y + 8 * cc);
        }
    }
    return alpha;
}

// torch.sigmoid of the rarely used colour-output path, kept out of line so that its expf / division bodies do not sit 96 times in the
// epilogue's instruction stream
static __device__ __noinline__ float sigmoid_out_of_line(float f) { return 1.f / (1.f + expf(-f)); }

// Train mode, BatchNorm phases 2 and 0: the trunk output of this row comes from the cache phase 1 wrote (fp32, post-ReLU: the very
// values hidden_epilogue<1> saw in the accumulators), so the alpha head and the next A operand are bit-identical to a recompute.
template <int N, bool kHiLo>
__device__ __forceinline__ float trunk_from_cache(const float* __restrict__ h_in, unsigned char* abuf, int chunk0, int m,
                                                  const float* __restrict__ aw, unsigned char* abuf_lo) {
    float alpha = 0.f;
#pragma unroll
    for (int c = 0; c < N / 32; ++c) {
        float y[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = h_in ? __ldg(reinterpret_cast<const float4*>(h_in + c * 32 + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w4 = *reinterpret_cast<const float4*>(aw + c * 32 + 4 * q);
            alpha = fmaf(fmaxf(y[4 * q + 0], 0.f), w4.x, alpha);
            alpha = fmaf(fmaxf(y[4 * q + 1], 0.f), w4.y, alpha);
            alpha = fmaf(fmaxf(y[4 * q + 2], 0.f), w4.z, alpha);
            alpha = fmaf(fmaxf(y[4 * q + 3], 0.f), w4.w, alpha);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            if (kHiLo) store_a8_hilo(abuf, abuf_lo, chunk0 + c * 4 + cc, m, y + 8 * cc, true);
            else store_a8_relu(abuf, chunk0 + c * 4 + cc, m, y + 8 * cc);
        }
    }
    return alpha;
}

// 16-column variants of the TMEM load / wait
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld_regs16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// Everything an epilogue thread needs that does not change between tiles.
struct TileCtx {
    const PeFieldArgs* A;
    const PeIntegrated* G2;      // outputs of the composed scene when this object IS the scene (else all NULL)
    unsigned char* abuf;         // this group's A operand (shared memory)
    unsigned char* abuf_lo;      // fp16x3 mode: operand buffer of the low halves
    uint32_t taddr;              // TMEM address of this thread's lane quadrant and this group's accumulator columns
    uint32_t bar_id;             // named barrier of the group
    int m, lane, wq;             // row of the tile (= TMEM lane), lane, lane quadrant (warp % 4)
    int half;                    // kSplit == 2: which half of the columns of row m this thread handles
    int gw;                      // warp index inside the group (0 .. 4*kSplit-1)
    int P, rpt, rows_used, tiles_per_image;
    int64_t total_tiles;
    float size[3];
    float alpha_bias;
    const float* alpha_w;
    bool single;
    int stat_phase;              // train-mode BatchNorm: 1 / 2 = this launch accumulates the statistics of head layer 0 / 3 and stops there
    bool resume;                 // train mode, phases 2 and 0 with a trunk cache (PeFieldArgs.h7_out): the layers start at head layer 0
    bool fold;                   // folded-head mode: composite the 128-wide input of head layer 6, skip that layer's MMAs
    int dbg;                     // PE_TC_TIMELINE=1: thread 0 of epilogue group 0 of CTA 0 prints clock64 stamps of its third tile
};

// 32 of the 64 positional-encoding columns of one sample (columns 32*H .. 32*H+31); layout of positional_encoder.py:59-64:
// [x y z | sin(2^0 xyz) cos(2^0 xyz) | sin(2^1 xyz) ... ] (+ one zero pad column).  tp/tl: x/(2 pi) as a two-float value.
template <int H>
__device__ __forceinline__ void encode_half(const float (&xn)[3], const float (&tp)[3], const float (&tl)[3], float (&enc)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int idx = 32 * H + i;
        float v;
        if (idx < 3) v = xn[idx];
        else if (idx == 63) v = 0.f;
        else {
            const int o = (idx - 3) / 6, r = (idx - 3) % 6, fn = r / 3, a = r % 3;
            const float f = (float)(1 << o);
            const float turns = tp[a] * f;                           // exact (power of two)
            const float fr = (turns - rintf(turns)) + tl[a] * f;     // fractional turns in [-0.5, 0.5]
            const float ang = fr * 6.2831855f;
            v = fn ? __cosf(ang) : __sinf(ang);
        }
        enc[i] = v;
    }
}

// Sampling state of one tile row (transform_rays, z bounds, create_ray_positions, in-box mask).
struct RowSample {
    bool tile_valid, valid, inbox, in_scene;
    int img, ray0, p;
    int64_t ray;
    float t, dnorm;
    float x[3];
};

// world-space direction (and stratified-sampling number) of a tile row, loaded ahead of use
struct RayPrefetch { float d[3]; float u; };

template <bool kNoPrepass = false>
__device__ __forceinline__ void prefetch_row(const TileCtx& X, int64_t tile, RayPrefetch& pf) {
    const PeFieldArgs& A = *X.A;
    pf.d[0] = pf.d[1] = pf.d[2] = 0.f; pf.u = 0.f;
    if (tile >= X.total_tiles || X.m >= X.rows_used) return;
    if (!kNoPrepass && A.tile_list) tile = A.tile_list[tile];
    const int img = (int)(tile / X.tiles_per_image);
    const int r = (int)(tile - (int64_t)img * X.tiles_per_image) * X.rpt + X.m / X.P;
    if (r >= A.rays) return;
    const int64_t ray = (int64_t)img * A.rays + r;
    const float* dw = A.dirs + ray * 3;
    pf.d[0] = dw[0]; pf.d[1] = dw[1]; pf.d[2] = dw[2];
    if (A.perturb && !A.t_in) pf.u = A.rand[ray * X.P + (X.m % X.P)];
}

template <bool kNoPrepass = false>
__device__ __forceinline__ void sample_row(const TileCtx& X, int64_t tile, RowSample& s, const RayPrefetch* pf = nullptr) {
    const PeFieldArgs& A = *X.A;
    const PeObjectDesc& ob = A.ob;
    const int m = X.m, P = X.P, rpt = X.rpt;
    s.tile_valid = tile < X.total_tiles;
    if (!kNoPrepass && s.tile_valid && A.tile_list) tile = A.tile_list[tile];            // pre-pass mode: only the non-empty tiles are listed
    s.img = s.tile_valid ? (int)(tile / X.tiles_per_image) : 0;
    s.ray0 = s.tile_valid ? (int)(tile - (int64_t)s.img * X.tiles_per_image) * rpt : 0;
    s.valid = false; s.inbox = false;
    s.t = 0.f; s.dnorm = 0.f; s.ray = -1; s.p = 0;
    s.in_scene = A.ois ? A.ois[(int64_t)s.img * A.objects + A.k] != 0 : true;
    s.x[0] = s.x[1] = s.x[2] = 0.f;
    if (s.tile_valid && m < X.rows_used) {
        const int rl = m / P;
        const int r = s.ray0 + rl;
        if (r < A.rays) {
            s.valid = true;
            s.p = m - rl * P;
            s.ray = (int64_t)s.img * A.rays + r;
            float dw[3];
            if (pf) { dw[0] = pf->d[0]; dw[1] = pf->d[1]; dw[2] = pf->d[2]; }
            else { const float* dg = A.dirs + s.ray * 3; dw[0] = dg[0]; dw[1] = dg[1]; dw[2] = dg[2]; }
            if (!kNoPrepass && A.bent) {
                // sampled (and bent) by the pre-pass: position, parameter t and masks come from memory
                const int64_t gs = s.ray * P + s.p;
                s.inbox = (A.flags[gs] & 2) != 0;
                s.t = A.t_out[gs];
                if (s.inbox) { s.x[0] = A.bent[gs * 3]; s.x[1] = A.bent[gs * 3 + 1]; s.x[2] = A.bent[gs * 3 + 2]; }
            } else {
                const PeRay pr = pe_make_ray(ob, A.w2o + ((int64_t)s.img * A.objects + A.k) * 12, A.origins + (int64_t)s.img * 3, dw, s.in_scene);
                const float u = (A.perturb && !A.t_in) ? (pf ? pf->u : A.rand[s.ray * P + s.p]) : 0.f;
                s.t = pe_sample_t_or(A.t_in, s.ray * P + s.p, pr, s.p, P, A.perturb != 0, u);
                pe_position(pr, s.t, s.x);
                s.inbox = pe_in_box(ob, s.x);
            }
            s.dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dw[0], dw[0]), __fmul_rn(dw[1], dw[1])), __fmul_rn(dw[2], dw[2])));
        }
    }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 q;
    q.x = pack_half2(v[0], v[1]); q.y = pack_half2(v[2], v[3]); q.z = pack_half2(v[4], v[5]); q.w = pack_half2(v[6], v[7]);
    return q;
}

// Fourier features sin/cos(2^o * x) of this thread's share of the 64 encoding columns, as packed fp16 operand chunks
// (kHiLo: hi chunks then lo chunks).  The argument is reduced EXACTLY (x/(2 pi) as a two-float value, scaled by the power
// of two, integer part dropped), then evaluated with the SFU on [-pi, pi] (abs error < 5e-7, far below the fp16 rounding of
// the operand).  Same values as positional_encoder.py:59-64 up to that error.
template <int kSplit, bool kHiLo>
__device__ __forceinline__ void encode_row(const TileCtx& X, const float (&x)[3], uint4 (&q)[8]) {
    const float xn[3] = {__fdiv_rn(x[0], X.size[0]), __fdiv_rn(x[1], X.size[1]), __fdiv_rn(x[2], X.size[2])};
    float tp[3], tl[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float c_hi = 0.15915494f, c_lo = 6.4206382e-9f;      // 1/(2 pi) = c_hi + c_lo
        tp[a] = xn[a] * c_hi;
        tl[a] = fmaf(xn[a], c_lo, fmaf(xn[a], c_hi, -tp[a]));
    }
    float enc[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (kSplit == 2 && h != X.half) continue;
        if (h == 0) encode_half<0>(xn, tp, tl, enc); else encode_half<1>(xn, tp, tl, enc);
        const int base = kSplit == 2 ? 0 : 4 * h;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (kHiLo) {
                float hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { hi[i] = __half2float(__float2half_rn(enc[8 * c + i])); lo[i] = enc[8 * c + i] - hi[i]; }
                q[c] = pack8(hi);
                q[4 + c] = pack8(lo);
            } else {
                q[base + c] = pack8(enc + 8 * c);
            }
        }
    }
}

template <int kSplit, bool kHiLo>
__device__ __forceinline__ void store_enc(const TileCtx& X, const uint4 (&q)[8]) {
    const int m = X.m;
    if (kHiLo) {
        const int c0 = PE_CHUNK0 + (kSplit == 2 ? 4 * X.half : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            *reinterpret_cast<uint4*>(X.abuf + (c0 + c) * CHUNK_BYTES + m * 16) = q[c];
            *reinterpret_cast<uint4*>(X.abuf_lo + (c0 + c) * CHUNK_BYTES + m * 16) = q[4 + c];
        }
    } else if (kSplit == 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(X.abuf + (PE_CHUNK0 + 4 * X.half + c) * CHUNK_BYTES + m * 16) = q[c];
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(X.abuf + (PE_CHUNK0 + c) * CHUNK_BYTES + m * 16) = q[c];
    }
}

// Look-ahead state carried from one tile to the next (folded-head mode): the next tile's rows are sampled and encoded
// while the tensor core runs the last two layers of the current tile.
struct TileAhead {
    bool have = false;
    RayPrefetch pf;
    RowSample rs;
    uint4 enc[8];
};

// Per-tile work of one epilogue thread.  `Sync` provides wait_acc() (accumulators of the next layer are complete) and
// arrive_ready() (this thread's part of the next A operand is written and its TMEM reads are done).
// kSplit threads (in different warps of the same lane quadrant) share one row: each handles 1/kSplit of the columns.
// kFoldOnly: instantiation for the folded-head mode without pre-pass (the headline path): everything else compiles away
template <int kSplit, bool kHiLo, bool kStats = false, bool kFoldOnly = false, class Sync>
__device__ __forceinline__ void epilogue_tile(const TileCtx& X, int64_t tile, Sync& sync, TileAhead& ahead, int64_t next_tile) {
    static_assert(!(kHiLo && kSplit != 2) , "fp16x3 mode runs with two threads per row");
    constexpr int GROUP = TILE_M * kSplit;           // threads of the group
    const PeFieldArgs& A = *X.A;
    const PeIntegrated& G2 = *X.G2;
    const PeObjectDesc& ob = A.ob;
    const int m = X.m, lane = X.lane, wq = X.wq, P = X.P, rpt = X.rpt, hf = X.half;
    const int tid = m + TILE_M * hf;
    unsigned char* abuf = X.abuf;
    const uint32_t taddr = X.taddr, bar_id = X.bar_id;
    float* scr = reinterpret_cast<float*>(abuf);
    float* t_s = reinterpret_cast<float*>(abuf + SCR_T);
    float* sh_s = reinterpret_cast<float*>(abuf + SCR_SH);
    float* w_s = reinterpret_cast<float*>(abuf + SCR_W);
    float* cst = reinterpret_cast<float*>(abuf + CST_BASE);
    float* alpha_part = cst + 1024;                  // [kSplit][128] partial alpha-head dot products

#ifdef PE_TC_TIMELINE      // diagnostic build (-DPE_TC_TIMELINE, run with PE_TC_TIMELINE=1): in-kernel timeline of one tile
    long long ts[36];
    const bool rec = X.dbg != 0 && blockIdx.x == 0 && m == 0 && hf == 0 && bar_id == 1 && tile / (2 * gridDim.x) == 2;
#define PE_STAMP(i) do { if (rec) ts[i] = clock64(); } while (0)
#else
#define PE_STAMP(i) do { } while (0)
#endif
    PE_STAMP(0);
    // ---- sampling + Fourier features (unless the previous tile's epilogue already did them) ----
    const bool resume = kStats && X.resume;
    if (resume) {
        sample_row<kFoldOnly>(X, tile, ahead.rs);       // masks and ray parameters; the encoding and the trunk are not evaluated again
    } else {
        if (!ahead.have) {
            sample_row<kFoldOnly>(X, tile, ahead.rs);
            encode_row<kSplit, kHiLo>(X, ahead.rs.x, ahead.enc);
        }
        store_enc<kSplit, kHiLo>(X, ahead.enc);
        sync.arrive_ready();
    }
    PE_STAMP(1);
    ahead.have = false;
    const bool tile_valid = ahead.rs.tile_valid, valid = ahead.rs.valid, inbox = ahead.rs.inbox, in_scene = ahead.rs.in_scene;
    const int img = ahead.rs.img, ray0 = ahead.rs.ray0, p = ahead.rs.p;
    const int64_t ray = ahead.rs.ray;
    const float t = ahead.rs.t, dnorm = ahead.rs.dnorm;
    float raw_alpha = ob.empty_space_alpha;

    // ---- alpha compositing of the tile's rays (compute_position_distances / compute_alphas / compute_weights, :153-214) and
    // the per-ray scalars of ObjectComposer.integrate (:758-772); scratch pointers differ between the two modes ----
    const int64_t gs = valid ? ray * P + p : 0;
    auto row_weight = [&](float* t_s, float* sh_s, float* w_s, float raw_alpha_v) -> float {
        t_s[m] = t;
        named_bar_sync(bar_id, GROUP);
        if (kSplit == 2) raw_alpha_v = alpha_part[m] + alpha_part[TILE_M + m];
        raw_alpha_v += X.alpha_bias;
        float raw = (inbox && in_scene) ? raw_alpha_v : ob.empty_space_alpha;
        if (valid && hf == 0) {
            if (A.raw_out) A.raw_out[gs] = raw;
            if (A.t_out && !A.bent) A.t_out[gs] = t;
            if (A.inbox_out) A.inbox_out[gs] = inbox ? 1 : 0;
            if (A.dispmag_out && !A.bent) A.dispmag_out[gs] = 0.f;
        }
        float alpha = 0.f;
        if (valid) {
            const float delta = __fmul_rn(p == P - 1 ? 1e10f : __fsub_rn(t_s[m + 1], t), dnorm);     // :153-178
            if (A.noise) raw = __fadd_rn(raw, A.noise[gs]);                                            // :193-195
            alpha = __fsub_rn(1.f, expf(__fmul_rn(-fmaxf(raw, 0.f), delta)));                          // :197
        }
        // exclusive cumprod of (1 - alpha + 1e-10) along the samples of each ray (compute_weights :199-214):
        // segmented warp scan + carry across the warps a ray spans
        const float shifted = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
        const bool head = p == 0;
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
        if (lane == 31 && hf == 0) { sh_s[wq] = incl; sh_s[4 + wq] = closed ? 1.f : 0.f; }
        named_bar_sync(bar_id, GROUP);
        float T = excl;
        if (!closed) {                                   // the ray started in an earlier warp of the tile
            for (int v = wq - 1; v >= 0; --v) {
                T *= sh_s[v];
                if (sh_s[4 + v] != 0.f) break;
            }
        }
        const float w = valid ? alpha * T : 0.f;
        if (hf == 0) {
            w_s[m] = w;
            if (valid) {
                if (A.integ.weights) A.integ.weights[gs] = w;
                if (X.single && G2.weights) G2.weights[gs] = w;
            }
        }
        return w;
    };
    // per-ray scalars (:758-772): one warp per ray, lanes stride the samples
    auto ray_scalars = [&](const float* t_s, const float* w_s) {
        for (int rl = X.gw; rl < rpt; rl += 4 * kSplit) {
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
                for (int oi = 0; oi < (X.single ? 2 : 1); ++oi) {
                    const PeIntegrated& O = *outs[oi];
                    if (O.opacity) O.opacity[gr] = opacity;
                    if (O.depth) O.depth[gr] = depth;
                    if (O.disparity) O.disparity[gr] = disparity;
                    if (O.integrated_displacements_magnitude) O.integrated_displacements_magnitude[gr] = 0.f;
                    if (O.integrated_divergence) O.integrated_divergence[gr] = 0.f;
                }
            }
        }
    };

    // ---- the 10 hidden tensor-core layers ----
    const bool fold = kFoldOnly ? true : X.fold;
    float4 pre[2 / kSplit];
    const uint32_t tcol = taddr + hf * (256 / kSplit);        // first accumulator column of this thread for 256-wide layers
#pragma unroll 1
    for (int l = resume ? 7 : 0; l < 10; ++l) {
        if (fold && l == 8) {
            // folded-head mode: the raw alphas are known after L7, so the compositing weights are computed here, while the
            // tensor core runs head layer 0
            const float w = row_weight(cst + FOLD_T, cst + FOLD_SH, cst + FOLD_W, raw_alpha);
            const float wf = inbox ? w : 0.f;
            float sw = wf;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
            if (hf == 0) {
                cst[FOLD_WF + m] = wf;
                if (lane == 0) cst[FOLD_PART_S + wq] = sw;
            }
            named_bar_sync(bar_id, GROUP);
            ray_scalars(cst + FOLD_T, cst + FOLD_W);
#ifdef PE_TC_SAMPLE_AT_L8
            sample_row<kFoldOnly>(X, next_tile, ahead.rs);
#else
            prefetch_row<kFoldOnly>(X, next_tile, ahead.pf);   // look-ahead: the next tile's ray data (latency hidden behind head layer 0)
#endif
        }
        if (fold && l == 9) {
#ifndef PE_TC_SAMPLE_AT_L8
            sample_row<kFoldOnly>(X, next_tile, ahead.rs, &ahead.pf);   // ... its rows and their encoding (behind head layer 3)
#endif
            encode_row<kSplit, kHiLo>(X, ahead.rs.x, ahead.enc);
            ahead.have = true;
        }
        PE_STAMP(2 + 3 * l);
        const bool cached = resume && l == 7;           // this "layer" is the trunk cache: no accumulators to wait for
        if (!cached) sync.wait_acc();
        PE_STAMP(3 + 3 * l);
        if (l == 4 || cached) {
            // the encoding columns are dead once L4 has run: reuse them for the constants of the later epilogues
            // (AdaIn scale/shift of this image and the alpha-head weights); the loads overlap this layer's epilogue
            const float* a1 = A.aff1 + (int64_t)img * 512;
            const float* a2 = A.aff2 + (int64_t)img * 256;
#pragma unroll
            for (int j = 0; j < 2 / kSplit; ++j) {
                const int i0 = (j * GROUP + tid) * 4;                      // 1024 floats: sc1|sh1 (512), sc2|sh2 (256), alpha_w (256)
                const float* src = i0 < 512 ? a1 + i0 : (i0 < 768 ? a2 + (i0 - 512) : X.alpha_w + (i0 - 768));
                pre[j] = __ldg(reinterpret_cast<const float4*>(src));
            }
        }
        if (cached) {
#pragma unroll
            for (int j = 0; j < 2 / kSplit; ++j) *reinterpret_cast<float4*>(cst + (j * GROUP + tid) * 4) = pre[j];
        }
        if (l == 7) named_bar_sync(bar_id, GROUP);                         // constants written by the whole group at l == 4
        constexpr int W = 256 / kSplit, W2 = 128 / kSplit;
        if (kStats && X.stat_phase != 0 && l == 7 + X.stat_phase) {
            // Statistics phase of train-mode BatchNorm (adain.py:47; BatchNorm1d over all evaluated samples of this object): this layer was
            // issued transposed (TMEM lane = feature, column = sample; 256-wide head layer 0 as two blocks of 128 columns), so each thread
            // sums its feature over the tile's evaluated samples; double-precision atomics like the fp32 kernel (accumulate_stats)
            float* rowmask = cst + FOLD_WF;
            if (hf == 0) rowmask[m] = (valid && inbox) ? 1.f : 0.f;
            named_bar_sync(bar_id, GROUP);
            const int C = X.stat_phase == 1 ? 256 : 128;
            double* stats = A.stats + (X.stat_phase == 1 ? 0 : 2 * 256 + 2);
            constexpr int CW = 128 / kSplit;                // sample columns per thread
            for (int blk = 0; blk < C / 128; ++blk) {
                float sum = 0.f, sq = 0.f;
#pragma unroll
                for (int q = 0; q < CW / 32; ++q) {
                    const int col0 = hf * CW + q * 32;
                    uint32_t v[32];
                    tmem_ld32(taddr + blk * 128 + col0, v);
                    tmem_wait_ld_regs(v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float y = __uint_as_float(v[i]) * rowmask[col0 + i];
                        sum += y;
                        sq = fmaf(y, y, sq);
                    }
                }
                atomicAdd(stats + blk * 128 + m, (double)sum);
                atomicAdd(stats + C + blk * 128 + m, (double)sq);
            }
            if (tid < 32) {                                  // number of samples the statistics run over (stats[2C])
                float cnt = rowmask[tid] + rowmask[tid + 32] + rowmask[tid + 64] + rowmask[tid + 96];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                if (tid == 0 && cnt != 0.f) atomicAdd(stats + 2 * C, (double)cnt);
            }
            tc_fence_before();
            named_bar_sync(bar_id, GROUP);        // scratch is dead before the next tile's encoding overwrites it
            return;
        }
        if (fold && l == 9) {
            // Folded head: integrated_features = sum_p w_p (W6 h_p + b6) = W6 (sum_p w_p h_p) + b6 sum_p w_p, so the
            // volume-rendering sum (:749) is taken over the 128-wide h (fp32, straight from the accumulators) and head
            // layer 6 runs once per ray afterwards (by an otherwise idle warp of the CTA) instead of once per sample.
            // Head layer 3 was issued TRANSPOSED for this (weights as the M operand, samples as N): TMEM lane = feature,
            // column = sample, so the sum over a ray's samples is a plain in-thread loop.
            const int c = m;                                // feature handled by this thread
            const float sc = cst[CST_SC2 + c], sh = cst[CST_SH2 + c];
            const float* wfs = cst + FOLD_WF;               // in-box compositing weight of each sample (row) of the tile
            float* part = cst + FOLD_PART;                  // kSplit == 2: [half][ray][128] partial sums
            constexpr int CW = 128 / kSplit;                // sample columns per thread
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < CW / 32; ++q) {
                const int col0 = hf * CW + q * 32;
                uint32_t v[32];
                tmem_ld32(taddr + col0, v);
                tmem_wait_ld_regs(v);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wfs + col0 + 4 * i);      // smem broadcast
                    a0 = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * i + 0]), sc, sh), 0.f), w4.x, a0);
                    a1 = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * i + 1]), sc, sh), 0.f), w4.y, a1);
                    a0 = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * i + 2]), sc, sh), 0.f), w4.z, a0);
                    a1 = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * i + 3]), sc, sh), 0.f), w4.w, a1);
                }
                acc += a0 + a1;
                if ((col0 + 32) % P == 0 || q == CW / 32 - 1) {          // last chunk of a ray (or of this thread's share of it)
                    const int rl = col0 / P;                // rl == rpt: the unused tail rows of the tile (128 % P != 0)
                    const int r = ray0 + rl;
                    if (kSplit == 1) {
                        if (tile_valid && rl < rpt && r < A.rays) A.fold_v[((int64_t)img * A.rays + r) * FOLD_K + c] = acc;
                    } else if (rl < rpt) {
                        part[(hf * 4 + rl) * FOLD_K + c] = acc;
                    }
                    acc = 0.f;
                }
            }
            tc_fence_before();
            named_bar_sync(bar_id, GROUP);
            const int wpr = P >> 5;                          // warps (row quadrants) per ray
            if (kSplit == 2 && hf == 0) {
                for (int rl = 0; rl < rpt; ++rl) {
                    const int r = ray0 + rl;
                    if (!tile_valid || r >= A.rays) continue;
                    // samples of ray rl: columns [rl*P, (rl+1)*P); half h owns columns [64h, 64h+64)
                    float sum = 0.f;
                    if (rl * P < 64) sum += part[(0 * 4 + rl) * FOLD_K + c];
                    if ((rl + 1) * P > 64) sum += part[(1 * 4 + rl) * FOLD_K + c];
                    A.fold_v[((int64_t)img * A.rays + r) * FOLD_K + c] = sum;
                }
            }
            if (tid < rpt && tile_valid && ray0 + tid < A.rays) {
                float sum = 0.f;
                for (int j = 0; j < wpr; ++j) sum += cst[FOLD_PART_S + tid * wpr + j];
                A.fold_s[(int64_t)img * A.rays + ray0 + tid] = sum;
            }
            sync.arrive_fold();                   // the head-6 warp of the CTA turns the sums into integrated_features
            named_bar_sync(bar_id, GROUP);        // scratch is dead before the next tile's encoding overwrites it
            PE_STAMP(4 + 3 * l);
#ifdef PE_TC_TIMELINE
            if (rec) {
                printf("PE_TC timeline (cycles since tile start; per layer: wait begin, acc ready, epilogue done): enc %lld |", ts[1] - ts[0]);
                for (int q = 0; q < 10; ++q) printf(" L%d %lld %lld %lld |", q, ts[2 + 3 * q] - ts[0], ts[3 + 3 * q] - ts[0], ts[4 + 3 * q] - ts[0]);
                printf("\n");
            }
#endif
            return;
        }
        if (l < 7) hidden_epilogue<0, W, kHiLo>(tcol, abuf, hf * (W / 8), m, nullptr, nullptr, X.abuf_lo);
        else if (l == 7) {
            // train-mode recompute for the backward: the trunk output goes to memory so that the BatchNorm-reduction passes of
            // pe_field_bwd_kernel start from it instead of recomputing bender, encoding and trunk
            // train mode with a trunk cache: phase 1 writes this row's trunk output, phases 2 and 0 (and the fp32 field backward's
            // BatchNorm passes) start from it
            float* h_out = (kStats && A.h7_out && X.stat_phase == 1 && valid) ? A.h7_out + gs * 256 + hf * W : nullptr;
            if (cached) raw_alpha = trunk_from_cache<W, kHiLo>(valid ? A.h7_out + gs * 256 + hf * W : nullptr, abuf, hf * (W / 8), m,
                                                               cst + CST_AW + hf * W, X.abuf_lo);
            else raw_alpha = hidden_epilogue<1, W, kHiLo>(tcol, abuf, hf * (W / 8), m, cst + CST_AW + hf * W, nullptr, X.abuf_lo, h_out);
        }
        else if (l == 8) hidden_epilogue<2, W, kHiLo>(tcol, abuf, hf * (W / 8), m, cst + CST_SC1 + hf * W, cst + CST_SH1 + hf * W, X.abuf_lo);
        else hidden_epilogue<2, W2, kHiLo>(taddr + hf * W2, abuf, hf * (W2 / 8), m, cst + CST_SC2 + hf * W2, cst + CST_SH2 + hf * W2, X.abuf_lo);
        if (l == 4) {
#pragma unroll
            for (int j = 0; j < 2 / kSplit; ++j) *reinterpret_cast<float4*>(cst + (j * GROUP + tid) * 4) = pre[j];
        }
        if (l == 7 && kSplit == 2) alpha_part[hf * TILE_M + m] = raw_alpha;
        sync.arrive_ready();
        PE_STAMP(4 + 3 * l);
    }

    if (kFoldOnly) return;       // (unreachable: the folded-head path returned inside the loop; drops the code below from that instantiation)
    // ---- last layer: features in TMEM -> volume rendering of the tile's rays (ObjectComposer.integrate :724-784) ----
    // row scalars are computed redundantly by the kSplit threads of a row (identical values), stores by half 0 only
    sync.wait_acc();
    const float w = row_weight(t_s, sh_s, w_s, raw_alpha);
    const float wf = inbox ? w : 0.f;
    constexpr int FC = 96 / kSplit;                  // feature columns per thread per pass (96 or 48)
    for (int pass = 0; pass < 2; ++pass) {
        const int c_first = pass * 96 + hf * FC;     // first feature column of this thread in this pass
#pragma unroll
        for (int c = 0; c < FC / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(taddr + c_first + c * 16, v);
            tmem_wait_ld_regs16(v);
            if (A.apply_activation) {      // sigmoid on colour outputs (object_composer.py:548-549; no shipped config): out of line
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(sigmoid_out_of_line(__uint_as_float(v[q])));
            }
            // per-sample features an output (multi-object scenes: input of the compositor): the scratch holds the masked, UNWEIGHTED
            // feature so that it can be written out coalesced and the reduction below applies the weights; else the weighted feature
            const float scale = A.feat_out ? (inbox ? 1.f : 0.f) : wf;       // (head-6 bias already added by the rank-1 MMA)
            const bool zero = A.feat_out && !inbox;
#pragma unroll
            for (int q = 0; q < 16; ++q) scr[m * SCRATCH_STRIDE + hf * FC + c * 16 + q] = zero ? 0.f : scale * __uint_as_float(v[q]);
        }
        named_bar_sync(bar_id, GROUP);
        if (A.feat_out && tile_valid) {
            // the tile's rows are consecutive sample slots: 96 consecutive floats per row, consecutive threads -> consecutive addresses
            const int rows_valid = min(rpt, A.rays - ray0) * P;
            float* dst = A.feat_out + ((int64_t)img * A.rays + ray0) * P * 192 + pass * 96;
            for (int idx = tid; idx < rows_valid * 96; idx += GROUP) {
                const int row = idx / 96, c = idx - row * 96;
                dst[(int64_t)row * 192 + c] = scr[row * SCRATCH_STRIDE + c];
            }
        }
        if (A.integ.integrated_features || (X.single && G2.integrated_features)) {
            for (int item = tid; item < rpt * 96; item += GROUP) {
                const int rl = item / 96, c = item - rl * 96;
                const int r = ray0 + rl;
                if (tile_valid && r < A.rays) {
                    const float* col = scr + rl * P * SCRATCH_STRIDE + c;
                    const float* wr = w_s + rl * P;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    int j = 0;
                    if (A.feat_out) {
                        for (; j + 4 <= P; j += 4) {
                            s0 = fmaf(wr[j + 0], col[(j + 0) * SCRATCH_STRIDE], s0); s1 = fmaf(wr[j + 1], col[(j + 1) * SCRATCH_STRIDE], s1);
                            s2 = fmaf(wr[j + 2], col[(j + 2) * SCRATCH_STRIDE], s2); s3 = fmaf(wr[j + 3], col[(j + 3) * SCRATCH_STRIDE], s3);
                        }
                        for (; j < P; ++j) s0 = fmaf(wr[j], col[j * SCRATCH_STRIDE], s0);
                    } else {
                        for (; j + 4 <= P; j += 4) {
                            s0 += col[(j + 0) * SCRATCH_STRIDE]; s1 += col[(j + 1) * SCRATCH_STRIDE];
                            s2 += col[(j + 2) * SCRATCH_STRIDE]; s3 += col[(j + 3) * SCRATCH_STRIDE];
                        }
                        for (; j < P; ++j) s0 += col[j * SCRATCH_STRIDE];
                    }
                    const float sum = (s0 + s1) + (s2 + s3);
                    const int64_t o = ((int64_t)img * A.rays + r) * 192 + pass * 96 + c;
                    if (A.integ.integrated_features) A.integ.integrated_features[o] = sum;
                    if (X.single && G2.integrated_features) G2.integrated_features[o] = sum;
                }
            }
        }
        named_bar_sync(bar_id, GROUP);
    }
    ray_scalars(t_s, w_s);
    tc_fence_before();
    named_bar_sync(bar_id, GROUP);        // scratch is dead before the next tile's encoding overwrites it
}

// ---- weight ring + MMA issue, shared by the forward kernel (pe_field_tc.cu) and the backward kernels (pe_bwd_tc.cu) ----
constexpr int STAGE_BYTES = 16384;                // largest slab: 256 rows x 32 k x 2 B
constexpr int NUM_STAGES = 4;

struct Sync1 {        // epilogue <-> MMA handshakes inside one CTA
    uint64_t* acc_full;
    uint64_t* a_ready;
    uint32_t phase;
    int lane;
    uint64_t* h6_full;
    uint64_t* h6_done;
    uint32_t h6_phase;
    __device__ __forceinline__ void wait_acc() { mbar_wait(acc_full, phase); phase ^= 1; tc_fence_after(); }
    // every thread publishes its operand writes to the async proxy and orders its TMEM reads; ONE arrival per warp
    // (128 serialized arrivals on one mbarrier cost several hundred cycles per layer)
    __device__ __forceinline__ void arrive_ready() {
        fence_proxy_async(); tc_fence_before(); __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
    }
    // folded-head mode: this warp's per-ray sums are in global memory (release: visible to the head-6 warp of the CTA)
    // (two-way handshake: the head-6 warp must have consumed the previous tile's phase before it can complete again)
    __device__ __forceinline__ void arrive_fold() {
        __syncwarp();
        if (lane == 0) { mbar_wait(h6_done, h6_phase ^ 1); mbar_arrive(h6_full); }
        h6_phase ^= 1;
    }
};

// State of the MMA-issuing thread that persists across layers: position in the weight ring, operand addresses.
struct MmaRing {
    uint64_t *full_bar, *empty_bar, *acc_full;
    uint32_t a_addr[2], ring_addr, tmem_base;
    int stage; uint32_t phase;
    int num_passes, x3;
    // hi/lo-split modes: TMEM column offset of a second accumulator for the SMALL partial products (A_lo W_hi, A_hi W_lo, A_lo W_lo), 0 =
    // everything into one accumulator.  tcgen05.mma truncates its fp32 accumulation toward zero (measured: tests/gpu_tc_accumulate.py,
    // -0.5 ulp per instruction on same-sign data, against +-0.3 ulp of a rounded add): every small product added to an accumulator that
    // already holds the large sum costs a truncation at the LARGE sum's ulp.  Kept apart, the small terms truncate at their own
    // magnitude (2^-11 of the large one) and the large accumulator sees one add per k-step instead of three or four; the epilogue adds
    // the two once, rounded.
    uint32_t small_off = 0;
};

// Issues the MMAs of one layer for both tiles (slab by slab as the weights land).  kSwap: operand roles exchanged
// (D^T = W * A^T, folded-head mode, head layer 3) -- a separate instantiation so that the common loop stays branch-free.
// kX3Mode: 0 = one A buffer per tile (fp16 / fp16x2), 1 = fp16x3 (hi and lo A buffers of one tile), 2 = all four partial products
template <bool kSwap, int kMBlocks, int kX3Mode>
__device__ __forceinline__ void mma_layer(MmaRing& R, int l, int n, int slabs, int chunk0, bool has_bias, uint32_t idesc, uint32_t lbo_b) {
    (void)l; (void)n;
    // kSwap: D^T = W * A^T -- the weight slab is the M operand (blocks of 128 of its n rows: block b starts 2048 B into every K chunk and
    // lands in accumulator columns 128 b ..), the tile's 128 samples are N; idesc must describe M = 128, N = 128
    constexpr int mblocks = kMBlocks;                // 2: the 256-row slab of head layer 0 in the statistics phase
    const int num_passes = R.num_passes;
    for (int s = 0; s < slabs; ++s) {
        for (int pass = 0; pass < num_passes; ++pass) {
            mbar_wait(R.full_bar + R.stage, R.phase);
            tc_fence_after();
            const bool last = !has_bias && (s == slabs - 1) && (pass == num_passes - 1);
            const uint32_t b_addr = R.ring_addr + R.stage * STAGE_BYTES;
            if (kX3Mode != 0) {
                // pass 0 (W_hi): A_hi and A_lo; pass 1 (W_lo): A_hi only
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t a_chunk = chunk0 + 4 * s + 2 * j;
                    const uint64_t da_hi = umma_smem_desc(R.a_addr[0] + a_chunk * CHUNK_BYTES, CHUNK_BYTES, 128);
                    const uint64_t da_lo = umma_smem_desc(R.a_addr[1] + a_chunk * CHUNK_BYTES, CHUNK_BYTES, 128);
                    const uint32_t accum = (s | pass | j) != 0 ? 1u : 0u;
#pragma unroll
                    for (int b = 0; b < mblocks; ++b) {
                        const uint64_t db = umma_smem_desc(b_addr + 2 * j * lbo_b + b * 2048, lbo_b, 128);
                        const uint32_t td = R.tmem_base + b * 128;
                        if (R.small_off) {
                            // pass 0: A_hi W_hi -> large accumulator, A_lo W_hi -> small one; pass 1: A_hi W_lo (and A_lo W_lo) -> small one
                            const uint32_t ts = td + R.small_off;
                            umma_f16_ss(pass == 0 ? td : ts, kSwap ? db : da_hi, kSwap ? da_hi : db, idesc, pass == 0 ? ((s | j) != 0 ? 1u : 0u) : 1u);
                            if (pass == 0 || kX3Mode == 2) umma_f16_ss(ts, kSwap ? db : da_lo, kSwap ? da_lo : db, idesc, accum);
                            continue;
                        }
                        umma_f16_ss(td, kSwap ? db : da_hi, kSwap ? da_hi : db, idesc, accum);
                        // x3 == 2 (ray bender): the lo x lo term too -- its output feeds 2^9-octave Fourier features downstream
                        if (pass == 0 || kX3Mode == 2) umma_f16_ss(td, kSwap ? db : da_lo, kSwap ? da_lo : db, idesc, 1u);
                    }
                }
                if (last) umma_commit(R.acc_full + 0);
            } else if (kMBlocks == 1) {
                // the common case, kept minimal (the issue loop shares its scheduler with two epilogue warps): descriptor words
                // advanced by 32-bit adds
                constexpr uint32_t hi = umma_desc_hi(128);
                const uint32_t b_lo = umma_desc_lo(b_addr, lbo_b);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t a_lo = umma_desc_lo(R.a_addr[g] + (chunk0 + 4 * s) * CHUNK_BYTES, CHUNK_BYTES);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t al = a_lo + j * (2 * CHUNK_BYTES >> 4), bl = b_lo + j * (2 * lbo_b >> 4);
                        const uint32_t accum = (s | pass | j) != 0 ? 1u : 0u;
                        if (kSwap) umma_f16_ss_words(R.tmem_base + g * 256, bl, hi, al, hi, idesc, accum);
                        else umma_f16_ss_words(R.tmem_base + g * 256, al, hi, bl, hi, idesc, accum);
                    }
                    if (last) umma_commit(R.acc_full + g);
                }
            } else {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t a_chunk = chunk0 + 4 * s + 2 * j;
                        const uint64_t da = umma_smem_desc(R.a_addr[g] + a_chunk * CHUNK_BYTES, CHUNK_BYTES, 128);
#pragma unroll
                        for (int b = 0; b < mblocks; ++b) {
                            const uint64_t db = umma_smem_desc(b_addr + 2 * j * lbo_b + b * 2048, lbo_b, 128);
                            umma_f16_ss(R.tmem_base + g * 256 + b * 128, kSwap ? db : da, kSwap ? da : db, idesc, (s | pass | j) != 0 ? 1u : 0u);
                        }
                    }
                    if (last) umma_commit(R.acc_full + g);
                }
            }
            umma_commit(R.empty_bar + R.stage);
            if (++R.stage == NUM_STAGES) { R.stage = 0; R.phase ^= 1; }
        }
    }
}

}  // namespace pe_tc

// CUDA-core fp32 evaluation of one object field over tiles of 64 sample slots:
// sampling -> in-box mask -> (ray bender) -> positional encoding -> style-modulated MLP -> raw outputs.
// Exact-precision path for arbitrary architectures; the tcgen05 path (pe_field_tc.cu) covers the
// shipped shape.  Replaces, per sample, RayBendingStyleNerfModel.forward
// (model/nerf_models/ray_bending_style_nerf_model.py:137-219) and everything it calls.
#include "pe_kernels.cuh"

namespace {

constexpr int TM = 64;          // sample slots per tile
constexpr int NT = 256;         // threads per block
constexpr int KC = 16;          // weight rows staged per step

struct Smem {
    float* buf0;   // [Wmax][TM]
    float* buf1;   // [Wmax][TM]
    float* enc;    // [Emax][TM]
    float* wS;     // [KC][Nmax8]
    float* pos;    // [3][TM] sample position (object space)
    float* bent;   // [3][TM] bent position
    float* aux;    // [9][TM]: object-space origin(3) + direction(3) (skybox field input), displacement(3)
    int* flags;    // [TM] bit0: in-box (outer mask), bit1: inner mask, bit2: slot valid
};

// out[n][m] = act( sum_k in[k][m] * WT[k][n] + bias[n] ), in = seg0 (K0 rows) followed by seg1 (K1 rows).
// mode 0: linear, 1: ReLU, 2: ReLU(x*sc[n]+sh[n]) (AdaIn with folded BatchNorm)
__device__ void dense_layer(const float* __restrict__ seg0, int K0, const float* __restrict__ seg1, int K1,
                            const float* __restrict__ WT, const float* __restrict__ bias, int N,
                            float* __restrict__ out, float* __restrict__ wS, int mode,
                            const float* __restrict__ sc, const float* __restrict__ sh) {
    const int tid = threadIdx.x;
    const int mg = tid & 7, ng = tid >> 3;
    const int N8 = (N + 7) & ~7;
    const int K = K0 + K1;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const bool active = ng * 8 < N8;
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        for (int idx = tid; idx < KC * N8; idx += NT) {
            const int kk = idx / N8, n = idx - kk * N8;
            const int k = k0 + kk;
            wS[idx] = (k < K && n < N) ? __ldg(WT + (int64_t)k * N + n) : 0.f;
        }
        __syncthreads();
        if (active) {
            const int kend = min(KC, K - k0);
            for (int kk = 0; kk < kend; ++kk) {
                const int k = k0 + kk;
                const float* src = k < K0 ? seg0 + k * TM : seg1 + (k - K0) * TM;
                const float4 a0 = *reinterpret_cast<const float4*>(src + mg * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(src + mg * 8 + 4);
                const float4 w0 = *reinterpret_cast<const float4*>(wS + kk * N8 + ng * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(wS + kk * N8 + ng * 8 + 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
    }
    // `out` may alias neither input segment (ping-pong buffers), so no barrier is needed before writing
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = ng * 8 + j;
            if (n < N) {
                const float b = bias ? __ldg(bias + n) : 0.f;
                float scn = 1.f, shn = 0.f;
                if (mode == 2) { scn = sc[n]; shn = sh[n]; }
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float x = acc[i][j] + b;
                    if (mode == 1) x = fmaxf(x, 0.f);
                    if (mode == 2) x = fmaxf(fmaf(x, scn, shn), 0.f);
                    v[i] = x;
                }
                *reinterpret_cast<float4*>(out + n * TM + mg * 8) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(out + n * TM + mg * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    __syncthreads();
}

// per-channel sum / sum of squares over the in-box samples of the tile (train-mode BatchNorm, adain.py:47)
__device__ void accumulate_stats(const float* __restrict__ x, int C, const int* __restrict__ flags, double* __restrict__ stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool u0 = (flags[lane] & 2) != 0, u1 = (flags[lane + 32] & 2) != 0;
    for (int c = warp; c < C; c += NT / 32) {
        const float a = u0 ? x[c * TM + lane] : 0.f;
        const float b = u1 ? x[c * TM + lane + 32] : 0.f;
        float s = a + b, q = a * a + b * b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
            atomicAdd(stats + c, (double)s);
            atomicAdd(stats + C + c, (double)q);
        }
    }
    if (warp == 0) {        // number of samples the statistics run over (stats[2C])
        int n = (u0 ? 1 : 0) + (u1 ? 1 : 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) atomicAdd(stats + 2 * C, (double)n);
    }
}

__global__ void __launch_bounds__(NT, 1) pe_field_fp32_kernel(const PeFieldArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const int W = ob.width;
    const int Wmax = max(max(W, ob.features), ob.bender_kind == PE_BENDER_POSITIONAL ? ob.b_width : 0);
    const int Emax = max(L.enc, L.b_enc);
    const int Nmax8 = (Wmax + 7) & ~7;
    Smem S;
    {
        float* p = reinterpret_cast<float*>(smem_raw);
        S.buf0 = p; p += Wmax * TM;
        S.buf1 = p; p += Wmax * TM;
        S.enc = p; p += Emax * TM;
        S.wS = p; p += KC * Nmax8;
        S.pos = p; p += 3 * TM;
        S.bent = p; p += 3 * TM;
        S.aux = p; p += 9 * TM;
        S.flags = reinterpret_cast<int*>(p);
    }
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    auto P32 = [&](int64_t off) { return reinterpret_cast<const float*>(blob + off); };

    const int tid = threadIdx.x;
    const int P = A.explicit_positions ? 1 : ob.positions;
    const int64_t slots_per_image = (int64_t)A.rays * P;
    const int tiles_per_image = (int)((slots_per_image + TM - 1) / TM);
    const int64_t total_tiles = (int64_t)tiles_per_image * A.images;
    const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};

    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int img = (int)(tile / tiles_per_image);
        const int64_t slot0 = (tile - (int64_t)img * tiles_per_image) * TM;
        const bool in_scene = A.ois ? A.ois[(int64_t)img * A.objects + A.k] != 0 : true;
        __syncthreads();
        // ---- 1. sampling -------------------------------------------------------------------
        if (tid < TM) {
            const int64_t s = slot0 + tid;
            int flag = 0;
            float x[3] = {0.f, 0.f, 0.f};
            if (s < slots_per_image) {
                flag = 4;
                const int r = (int)(s / P), p = (int)(s - (int64_t)r * P);
                const int64_t gs = (int64_t)img * slots_per_image + s;
                if (A.explicit_positions) {
                    const float* xp = A.positions + gs * 3;
                    x[0] = xp[0]; x[1] = xp[1]; x[2] = xp[2];
                    if (ob.nerf_kind == PE_NERF_SKYBOX_V3) {
                        const float* o = A.origins + (int64_t)img * 3;
                        const float* d = A.dirs + ((int64_t)img * A.rays + r) * 3;
                        for (int a = 0; a < 3; ++a) { S.aux[a * TM + tid] = o[a]; S.aux[(3 + a) * TM + tid] = d[a]; }
                    }
                } else {
                    const PeRay ray = pe_make_ray(ob, A.w2o + ((int64_t)img * A.objects + A.k) * 12, A.origins + (int64_t)img * 3,
                                                  A.dirs + ((int64_t)img * A.rays + r) * 3, in_scene);
                    const float u = (A.perturb && !A.t_in) ? A.rand[gs] : 0.f;
                    const float t = pe_sample_t_or(A.t_in, gs, ray, p, P, A.perturb != 0, u);
                    pe_position(ray, t, x);
                    if (A.t_out) A.t_out[gs] = t;
                    for (int a = 0; a < 3; ++a) { S.aux[a * TM + tid] = ray.o[a]; S.aux[(3 + a) * TM + tid] = ray.d[a]; }
                }
                if (pe_in_box(ob, x)) flag |= 1;
            }
            S.flags[tid] = flag;
            for (int a = 0; a < 3; ++a) { S.pos[a * TM + tid] = x[a]; S.bent[a * TM + tid] = x[a]; S.aux[(6 + a) * TM + tid] = 0.f; }
        }
        const int any_inbox = __syncthreads_or(tid < TM ? (S.flags[tid] & 1) : 0);
        if (!any_inbox) {   // whole tile is empty space: features 0 (never read), alpha = empty_space_alpha
            if (tid < TM && (S.flags[tid] & 4) && (A.phase == 0 || A.phase == PE_PHASE_PREPASS)) {
                const int64_t gs = (int64_t)img * slots_per_image + slot0 + tid;
                A.raw_out[gs] = ob.empty_space_alpha;
                A.inbox_out[gs] = 0;
                if (A.phase == PE_PHASE_PREPASS) A.flags[gs] = 0;
                if (A.dispmag_out) A.dispmag_out[gs] = 0.f;
                if (A.disp_out) { A.disp_out[gs * 3] = 0.f; A.disp_out[gs * 3 + 1] = 0.f; A.disp_out[gs * 3 + 2] = 0.f; }
            }
            continue;
        }

        // ---- 2. ray bender (positional_ray_bender_model.py:81-163) ------------------------------
        if (ob.bender_kind == PE_BENDER_POSITIONAL) {
            const int Eb = 3 * (1 + 2 * ob.b_octaves);
            const float* dfm = A.deformation + (int64_t)img * ob.deformation_features;
            for (int idx = tid; idx < L.b_enc * TM; idx += NT) {
                const int e = idx / TM, m = idx - e * TM;
                float v;
                if (e < Eb) {
                    const float xn[3] = {__fdiv_rn(S.pos[m], size[0]), __fdiv_rn(S.pos[TM + m], size[1]), __fdiv_rn(S.pos[2 * TM + m], size[2])};
                    v = pe_encoding_value(xn, 3, e, ob.b_anneal);
                } else {
                    v = __ldg(dfm + (e - Eb));
                }
                S.enc[idx] = v;
            }
            __syncthreads();
            const float* cur = S.enc;
            int curK = L.b_enc;
            float* nxt = S.buf0;
            for (int l = 0; l < ob.b_layers; ++l) {
                const bool skip = l == ob.b_skip;
                dense_layer(cur, curK, S.enc, skip ? L.b_enc : 0, P32(L.bd_w[l]), P32(L.bd_b[l]), ob.b_width, nxt, S.wS, 1, nullptr, nullptr);
                cur = nxt; curK = ob.b_width;
                nxt = (nxt == S.buf0) ? S.buf1 : S.buf0;
            }
            dense_layer(cur, curK, nullptr, 0, P32(L.bd_out_w), nullptr, 3, nxt, S.wS, 0, nullptr, nullptr);
            if (tid < TM) {
                for (int a = 0; a < 3; ++a) {
                    const float x = S.pos[a * TM + tid];
                    float dsp = __fmul_rn(nxt[a * TM + tid], size[a]);
                    dsp = fmaxf(dsp, __fsub_rn(ob.bbox[2 * a], x));          // clamp_output :116-140
                    dsp = fminf(dsp, __fsub_rn(ob.bbox[2 * a + 1], x));
                    if (ob.canonical_pose) dsp = __fmul_rn(dsp, 0.f);
                    if (!(S.flags[tid] & 1)) dsp = 0.f;
                    S.aux[(6 + a) * TM + tid] = dsp;
                    S.bent[a * TM + tid] = __fadd_rn(x, dsp);
                }
            }
            __syncthreads();
        }
        // inner mask of the field on the bent position (adain_style_nerf_model.py:171-184)
        if (tid < TM) {
            const float xb[3] = {S.bent[tid], S.bent[TM + tid], S.bent[2 * TM + tid]};
            int f = S.flags[tid];
            if ((f & 1) && (ob.nerf_kind == PE_NERF_SKYBOX_V3 || pe_in_box(ob, xb))) f |= 2;
            S.flags[tid] = f;
        }
        __syncthreads();
        if (A.phase == PE_PHASE_PREPASS) {
            // sampling + ray bender only: the field itself runs on the tensor cores (pe_field_tc.cu) over the tiles that hold
            // at least one sample whose bent position is in the box; everything else keeps the empty-space values set here
            if (tid < TM && (S.flags[tid] & 4)) {
                const int64_t gs = (int64_t)img * slots_per_image + slot0 + tid;
                const int f = S.flags[tid];
                A.raw_out[gs] = ob.empty_space_alpha;
                A.inbox_out[gs] = 0;
                A.flags[gs] = (uint8_t)(f & 3);
                float d2 = 0.f;
                for (int c = 0; c < 3; ++c) {
                    const float dsp = (f & 1) ? S.aux[(6 + c) * TM + tid] : 0.f;
                    if (A.disp_out) A.disp_out[gs * 3 + c] = dsp;
                    d2 += dsp * dsp;
                    A.bent[gs * 3 + c] = S.bent[c * TM + tid];
                }
                if (A.dispmag_out) A.dispmag_out[gs] = sqrtf(d2);
            }
            continue;
        }

        // ---- 3. positional encoding (positional_encoder.py:41-65) ------------------------------
        for (int idx = tid; idx < L.enc * TM; idx += NT) {
            const int e = idx / TM, m = idx - e * TM;
            float xin[6];
            if (ob.nerf_kind == PE_NERF_SKYBOX_V3) {      // skybox_adain_style_nerf_model_v3.py:88-95
                const float d0 = S.aux[3 * TM + m], d1 = S.aux[4 * TM + m], d2 = S.aux[5 * TM + m];
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
                xin[0] = __fdiv_rn(S.aux[m], size[0]); xin[1] = __fdiv_rn(S.aux[TM + m], size[1]); xin[2] = __fdiv_rn(S.aux[2 * TM + m], size[2]);
                xin[3] = __fdiv_rn(d0, nrm); xin[4] = __fdiv_rn(d1, nrm); xin[5] = __fdiv_rn(d2, nrm);
            } else {                                        // adain_style_nerf_model.py:119-123
                xin[0] = __fdiv_rn(S.bent[m], size[0]); xin[1] = __fdiv_rn(S.bent[TM + m], size[1]); xin[2] = __fdiv_rn(S.bent[2 * TM + m], size[2]);
                xin[3] = xin[4] = xin[5] = 0.f;
            }
            S.enc[idx] = pe_encoding_value(xin, L.in_dims, e, nullptr);
        }
        __syncthreads();

        // ---- 4. backbone (adain_style_nerf_model.py:126-135) ------------------------------------
        const float* cur = S.enc;
        int curK = L.enc;
        float* nxt = S.buf0;
        for (int l = 0; l < ob.layers; ++l) {
            const bool skip = l == ob.skip;
            dense_layer(cur, curK, S.enc, skip ? L.enc : 0, P32(L.bb_w[l]), P32(L.bb_b[l]), W, nxt, S.wS, 1, nullptr, nullptr);
            cur = nxt; curK = W;
            nxt = (nxt == S.buf0) ? S.buf1 : S.buf0;
        }
        // alpha head (:138): one dot product per sample, kept in a register of thread m
        float raw_alpha = ob.empty_space_alpha;
        if (tid < TM) {
            if (ob.nerf_kind == PE_NERF_SKYBOX_V3) {
                raw_alpha = 10.0f;                           // occupied_space_alpha, skybox v3 :34,109
            } else {
                const float* aw = P32(L.alpha_w);
                float s = 0.f;
                for (int c = 0; c < W; ++c) s = fmaf(cur[c * TM + tid], __ldg(aw + c), s);
                raw_alpha = s + __ldg(P32(L.alpha_b));
            }
        }
        // ---- 5. feature head with AdaIn (adain_sequential.py:14-28, adain.py:21-61) ------------
        const float* h = cur;
        float* x1 = nxt;                                  // the other ping-pong buffer
        const bool stats1 = A.training && A.phase == 1;
        const bool stats2 = A.training && A.phase == 2;
        const float* aff1 = A.aff1 + (int64_t)img * 2 * W;
        const float* aff2 = A.aff2 + (int64_t)img * W;    // 2 * (W/2)
        dense_layer(h, W, nullptr, 0, P32(L.head0_w), nullptr, W, x1, S.wS, stats1 ? 0 : 2, aff1, aff1 + W);
        if (stats1) { accumulate_stats(x1, W, S.flags, A.stats); continue; }
        float* x2 = const_cast<float*>(h);                // h is dead now
        dense_layer(x1, W, nullptr, 0, P32(L.head3_w), nullptr, W / 2, x2, S.wS, stats2 ? 0 : 2, aff2, aff2 + W / 2);
        if (stats2) { accumulate_stats(x2, W / 2, S.flags, A.stats + 2 * W + 2); continue; }
        float* fo = x1;
        dense_layer(x2, W / 2, nullptr, 0, P32(L.head6_w), P32(L.head6_b), ob.features, fo, S.wS, 0, nullptr, nullptr);

        // ---- 6. outputs (ray_bending_style_nerf_model.py:200-217, object_composer.py:547-549) -----
        const int F = ob.features;
        if (tid < TM && (S.flags[tid] & 4)) {
            const int64_t gs = (int64_t)img * slots_per_image + slot0 + tid;
            const int f = S.flags[tid];
            float a = (f & 2) ? raw_alpha : ob.empty_space_alpha;
            if (!in_scene) a = ob.empty_space_alpha;
            A.raw_out[gs] = a;
            A.inbox_out[gs] = (f & 2) ? 1 : 0;
            float d2 = 0.f;
            for (int c = 0; c < 3; ++c) {
                const float dsp = (f & 1) ? S.aux[(6 + c) * TM + tid] : 0.f;
                if (A.disp_out) A.disp_out[gs * 3 + c] = dsp;
                d2 += dsp * dsp;
            }
            if (A.dispmag_out) A.dispmag_out[gs] = sqrtf(d2);
        }
        for (int idx = tid; idx < F * TM; idx += NT) {
            const int m = idx / F, c = idx - m * F;
            const int f = S.flags[m];
            if (f & 4) {
                const int64_t gs = (int64_t)img * slots_per_image + slot0 + m;
                float v = (f & 2) ? fo[c * TM + m] : 0.f;
                if (A.apply_activation) v = 1.f / (1.f + expf(-v));
                A.feat_out[gs * F + c] = v;
            }
        }
    }
}

}  // namespace

size_t pe_field_fp32_smem_bytes(const PeObjectDesc& ob, const PeLayout& L) {
    const int Wmax = max(max(ob.width, ob.features), ob.bender_kind == PE_BENDER_POSITIONAL ? ob.b_width : 0);
    const int Emax = max(L.enc, L.b_enc);
    const int Nmax8 = (Wmax + 7) & ~7;
    return sizeof(float) * ((size_t)2 * Wmax * TM + (size_t)Emax * TM + (size_t)KC * Nmax8 + 15 * TM) + sizeof(int) * TM;
}

int pe_launch_field_fp32(const PeFieldArgs& args, int sm_count, cudaStream_t stream) {
    const PeObjectDesc& ob = args.ob;
    if (ob.width > 256 || ob.width % 8 || ob.features > 256 || (ob.bender_kind == PE_BENDER_POSITIONAL && ob.b_width > 256)) {
        pe_set_error("fp32 field kernel supports widths/features up to 256 (width multiple of 8)");
        return PE_ERR_UNSUPPORTED;
    }
    const size_t smem = pe_field_fp32_smem_bytes(ob, args.L);
    if (smem > 227 * 1024) {
        pe_set_error("fp32 field kernel needs %zu bytes of shared memory (> 227 KB)", smem);
        return PE_ERR_UNSUPPORTED;
    }
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_field_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int P = args.explicit_positions ? 1 : ob.positions;
    const int64_t tiles = (((int64_t)args.rays * P + TM - 1) / TM) * args.images;
    if (tiles == 0) return PE_OK;
    const int grid = (int)pe_min64(tiles, (int64_t)sm_count * 4);
    pe_field_fp32_kernel<<<grid, NT, smem, stream>>>(args);
    PE_LAUNCH_CHECK("pe_field_fp32_kernel");
    return PE_OK;
}

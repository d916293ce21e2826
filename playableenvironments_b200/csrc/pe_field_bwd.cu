// Backward of one object field, exact fp32 on the CUDA cores, any architecture the forward fp32 kernel handles.
// Per tile of 32 sample slots the block recomputes the forward pass (sampling -> ray bender -> positional encoding ->
// style-modulated MLP), keeping every layer's activations in a block-private stash, then walks the layers in reverse:
//   dX = dY * W            (same register-tiled product as the forward, reading the nn.Linear [out][in] tensor)
//   dW += dY^T * X, db += sum dY   (warp-tiled outer product, accumulated into the parameter gradients with RED.ADD)
// down to the gradient of the sample position.  What autograd does for the reference by replaying
// model/nerf_models/ray_bending_style_nerf_model.py:137-219, adain_style_nerf_model.py:106-199, positional_ray_bender_model.py:81-163,
// model/layers/adain.py:21-61 and model/positional_encoder.py:41-65 op by op.
#include "pe_kernels.cuh"

namespace {

constexpr int TB = 32;            // sample slots per tile
constexpr int TS = 36;            // row stride in floats: float4 accesses of 8 consecutive rows hit 32 distinct banks
constexpr int NT = 256;
constexpr int KC = 16;            // reduction rows staged per step
constexpr int NPASS = 256;        // output columns per pass of the tile product
constexpr int CMAX = 384;         // rows of an activation / gradient buffer (W + E at the skip layer)
constexpr int EMAX = 128;         // rows of the encoding buffer

struct Stash {                    // row offsets (rows of TS floats) inside the block's stash
    int benc, bh[PE_MAX_LAYERS], enc, h[PE_MAX_LAYERS], x1pre, y1, x2pre, y2, rows;
};

__host__ __device__ inline Stash stash_layout(const PeObjectDesc& ob, const PeLayout& L) {
    Stash s = {};
    int r = 0;
    if (ob.bender_kind == PE_BENDER_POSITIONAL) {
        s.benc = r; r += L.b_enc;
        for (int l = 0; l < ob.b_layers; ++l) { s.bh[l] = r; r += ob.b_width; }
    }
    s.enc = r; r += L.enc;
    for (int l = 0; l < ob.layers; ++l) { s.h[l] = r; r += ob.width; }
    s.x1pre = r; r += ob.width;
    s.y1 = r; r += ob.width;
    s.x2pre = r; r += ob.width / 2;
    s.y2 = r; r += ob.width / 2;
    s.rows = r;
    return s;
}

struct Smem {
    float *bufA, *bufB, *bufX;    // [CMAX][TS]
    float* enc;                   // [EMAX][TS] forward: encoding (skip connection); backward: its gradient
    float* wS;                    // [KC][NPASS]
    float *pos, *bent, *aux;      // [3][TB], [3][TB], [9][TB] (object-space origin, direction, displacement)
    float *graw, *gdm;            // [TB] upstream gradients of the raw alpha / displacement magnitude of the slot
    int *flags, *clampf, *slot;   // [TB]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out(n, m) = sum_k in[k][m] * Wm[k * ldw + n],  in = seg0 rows [0,K0) followed by seg1 rows [0,K1);  epi(n, m0, acc[4]) per thread
template <class Epi>
__device__ __forceinline__ void gemm_tile(const float* seg0, int K0, const float* seg1, int K1, const float* __restrict__ Wm, int ldw, int N,
                                          float* wS, Epi epi) {
    const int tid = threadIdx.x, mg = tid & 7, ng = tid >> 3;
    const int K = K0 + K1;
    for (int nb = 0; nb < N; nb += NPASS) {
        const int Np = min(NPASS, N - nb);
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
        const bool active = ng * 8 < Np;
        for (int k0 = 0; k0 < K; k0 += KC) {
            __syncthreads();
            for (int idx = tid; idx < KC * NPASS; idx += NT) {
                const int kk = idx / NPASS, n = idx - kk * NPASS;
                const int k = k0 + kk;
                wS[idx] = (k < K && n < Np) ? __ldg(Wm + (int64_t)k * ldw + nb + n) : 0.f;
            }
            __syncthreads();
            if (active) {
                const int kend = min(KC, K - k0);
                for (int kk = 0; kk < kend; ++kk) {
                    const int k = k0 + kk;
                    const float* src = k < K0 ? seg0 + k * TS : seg1 + (k - K0) * TS;
                    const float4 a = *reinterpret_cast<const float4*>(src + mg * 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(wS + kk * NPASS + ng * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(wS + kk * NPASS + ng * 8 + 4);
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(av[i], wv[j], acc[j][i]);
                }
            }
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = nb + ng * 8 + j;
                if (n < N) epi(n, mg * 4, acc[j]);
            }
        }
    }
    __syncthreads();
}

// dW[n * ldw + k] += sum_m G[n][m] * X[k][m]  (n < N, k < K);  db[n] += sum_m G[n][m]
__device__ void outer_acc(const float* G, int N, const float* X, int K, float* __restrict__ dW, int ldw, float* __restrict__ db) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (dW) {
        for (int nb = 0; nb < N; nb += 32) {
            const int n0 = nb + warp * 4;
            if (n0 >= N) continue;
            for (int kc = 0; kc < K; kc += 128) {
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
                for (int m4 = 0; m4 < TB; m4 += 4) {
                    float4 a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a[i] = (n0 + i < N) ? *reinterpret_cast<const float4*>(G + (n0 + i) * TS + m4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = kc + lane + 32 * j;
                        b[j] = k < K ? *reinterpret_cast<const float4*>(X + k * TS + m4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            acc[i][j] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, acc[i][j]))));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = kc + lane + 32 * j;
                        if (n0 + i < N && k < K && acc[i][j] != 0.f) atomicAdd(dW + (int64_t)(n0 + i) * ldw + k, acc[i][j]);
                    }
            }
        }
    }
    if (db) {
        for (int n = warp; n < N; n += NT / 32) {
            const float v = warp_sum(G[n * TS + lane]);
            if (lane == 0 && v != 0.f) atomicAdd(db + n, v);
        }
    }
}

// rows [0, rows) of a stash block -> shared buffer (identical layout)
__device__ __forceinline__ void load_rows(float* dst, const float* src, int rows) {      // src was written by this block: no __restrict__ / ld.nc
    const int n4 = rows * TS / 4;
    for (int i = threadIdx.x; i < n4; i += NT) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}

__device__ __forceinline__ float enc_value(const float* x, int dims, int e, const float* anneal) {
    if (e < dims) return x[e];
    const int q = e - dims;
    const int oct = q / (2 * dims);
    const int rem = q - oct * 2 * dims;
    const int fn = rem / dims, dim = rem - fn * dims;
    float v = fn ? cosf(__fmul_rn(exp2f((float)oct), x[dim])) : sinf(__fmul_rn(exp2f((float)oct), x[dim]));
    if (anneal) v = __fmul_rn(v, anneal[oct]);
    return v;
}

// gradient of the encoding input `dim` given the gradient column genc[e * TS] of the encoding rows
__device__ __forceinline__ float enc_backward(const float* genc, const float* x, int dims, int octaves, int dim, const float* anneal) {
    float g = genc[dim * TS];
    for (int o = 0; o < octaves; ++o) {
        const float f = exp2f((float)o);
        float s, c;
        sincosf(__fmul_rn(f, x[dim]), &s, &c);
        const float gs = genc[(dims + o * 2 * dims + dim) * TS];
        const float gc = genc[(dims + o * 2 * dims + dims + dim) * TS];
        const float w = anneal ? anneal[o] : 1.f;
        g = fmaf(w * f, c * gs - s * gc, g);
    }
    return g;
}

// AdaIn backward for one layer of C channels.  buf[c][m] holds g = dL/d(x*sc+sh) (ReLU mask already applied).
//   per image:    A[c] += sum_m g,  B[c] += sum_m g*x                   (gradients of the style-predicted bias / scale)
//   train mode:   S1[c] += sum_m g*sc, S2[c] += sum_m g*sc*x            (cross-sample terms of the BatchNorm backward)
//   in place:     buf = mask ? g*sc - k1 - x*k2 : 0                       (dL/dx)
__device__ void adain_backward(float* buf, int C, const float* xpre, const float* __restrict__ sc, const float* __restrict__ k1,
                               const float* __restrict__ k2, const int* flags, float* __restrict__ sumA, float* __restrict__ sumB,
                               double* __restrict__ S1, double* __restrict__ S2, bool do_sums, bool do_bn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool use = (flags[lane] & 2) != 0;
    for (int c = warp; c < C; c += NT / 32) {
        const float g = use ? buf[c * TS + lane] : 0.f;
        const float x = xpre[c * TS + lane];
        const float s = sc[c];
        if (do_sums) {
            const float a = warp_sum(g), b = warp_sum(g * x);
            if (lane == 0) { if (a != 0.f) atomicAdd(sumA + c, a); if (b != 0.f) atomicAdd(sumB + c, b); }
        }
        if (do_bn) {
            const float a = warp_sum(g * s), b = warp_sum(g * s * x);
            if (lane == 0) { atomicAdd(S1 + c, (double)a); atomicAdd(S2 + c, (double)b); }
        }
        buf[c * TS + lane] = use ? g * s - k1[c] - x * k2[c] : 0.f;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) pe_field_bwd_kernel(const PeFieldBwdArgs B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PeFieldArgs& A = B.f;
    const PeObjectDesc& ob = A.ob;
    const PeLayout& L = A.L;
    const int W = ob.width, F = ob.features;
    const bool positional = ob.bender_kind == PE_BENDER_POSITIONAL;
    const bool skybox = ob.nerf_kind == PE_NERF_SKYBOX_V3;
    Smem S;
    {
        float* p = reinterpret_cast<float*>(smem_raw);
        S.bufA = p; p += CMAX * TS;
        S.bufB = p; p += CMAX * TS;
        S.bufX = p; p += CMAX * TS;
        S.enc = p; p += EMAX * TS;
        S.wS = p; p += KC * NPASS;
        S.pos = p; p += 3 * TB;
        S.bent = p; p += 3 * TB;
        S.aux = p; p += 9 * TB;
        S.graw = p; p += TB;
        S.gdm = p; p += TB;
        S.flags = reinterpret_cast<int*>(p); p += TB;
        S.clampf = reinterpret_cast<int*>(p); p += TB;
        S.slot = reinterpret_cast<int*>(p);
    }
    const Stash ST = stash_layout(ob, L);
    float* stash = B.stash + (int64_t)blockIdx.x * B.stash_floats;
    auto st = [&](int row) { return stash + (int64_t)row * TS; };
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(ob.packed);
    auto P32 = [&](int64_t off) { return reinterpret_cast<const float*>(blob + off); };

    const int tid = threadIdx.x;
    const int P = ob.positions;
    const int64_t slots_per_image = (int64_t)A.rays * P;
    const int tiles_per_image = (int)((slots_per_image + TB - 1) / TB);
    const int64_t total_tiles = (int64_t)tiles_per_image * A.images;
    const float size[3] = {ob.bbox[1] - ob.bbox[0], ob.bbox[3] - ob.bbox[2], ob.bbox[5] - ob.bbox[4]};
    const int phase = B.bwd_phase;
    const bool full = phase == 0;
    const float* k1_1 = B.bn_fix;            // BatchNorm 1: k1[W], k2[W]; BatchNorm 2: k1[W/2], k2[W/2]
    const float* k2_1 = B.bn_fix + W;
    const float* k1_2 = B.bn_fix + 2 * W;
    const float* k2_2 = B.bn_fix + 2 * W + W / 2;

    // ray-bender-only mode: the field's backward ran on the tensor cores (pe_bwd_tc.cu) and left dL/d bent position in g_bent_in
    const bool bender_only = B.g_bent_in != nullptr;
    const bool compact = B.slot_list != nullptr;
    const int64_t num_tiles = compact ? (int64_t)B.tile_begin[A.images] : total_tiles;
    int img_c = 0;                                   // compacted numbering: image of the current tile (tiles are visited in order)
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int img;
        if (compact) {
            while (tile >= B.tile_begin[img_c + 1]) ++img_c;
            img = img_c;
        } else {
            img = (int)(tile / tiles_per_image);
        }
        const int64_t gsi = (int64_t)img * slots_per_image;        // first slot of the image
        __syncthreads();
        if (tid < TB) {
            // slot (inside the image) handled by row tid of this tile, -1: none
            int64_t s;
            if (compact) {
                const int64_t e = (tile - B.tile_begin[img]) * TB + tid;
                s = e < B.slot_count[img] ? (int64_t)B.slot_list[gsi + e] : -1;
            } else {
                s = (tile - (int64_t)img * tiles_per_image) * TB + tid;
                if (s >= slots_per_image) s = -1;
            }
            S.slot[tid] = (int)s;
        }
        const bool in_scene = A.ois ? A.ois[(int64_t)img * A.objects + A.k] != 0 : true;
        __syncthreads();
        float* bufs[2] = {S.bufA, S.bufB};
        const float* cur = nullptr;
        int curK = 0, which = 0;
        // BatchNorm-reduction passes with the trunk output of the forward recompute at hand: only the head is evaluated
        const bool from_cache = !full && B.h7_cache != nullptr && B.inbox_in != nullptr;
        if (from_cache) {
            if (tid < TB) {
                const int s = S.slot[tid];
                int flag = 0;
                if (s >= 0) flag = 4 | (B.inbox_in[gsi + s] ? 3 : 0);
                S.flags[tid] = flag;
            }
            const int any = __syncthreads_or(tid < TB ? (S.flags[tid] & 2) : 0);
            if (!any) continue;
            for (int idx = tid; idx < W * TB; idx += NT) {
                const int m = idx / W, n = idx - m * W;                 // consecutive threads: consecutive features of one sample
                bufs[0][n * TS + m] = (S.flags[m] & 2) ? B.h7_cache[(gsi + S.slot[m]) * W + n] : 0.f;
            }
            cur = bufs[0]; curK = W; which = 1;
            __syncthreads();
        } else {
        // ================================ forward recompute ================================
        if (tid < TB) {
            const int64_t s = S.slot[tid];
            int flag = 0;
            float x[3] = {0.f, 0.f, 0.f};
            float graw = 0.f, gdm = 0.f;
            for (int a = 0; a < 9; ++a) S.aux[a * TB + tid] = 0.f;
            if (s >= 0) {
                flag = 4;
                const int r = (int)(s / P), p = (int)(s - (int64_t)r * P);
                const int64_t gs = gsi + s;
                const PeRay ray = pe_make_ray(ob, A.w2o + ((int64_t)img * A.objects + A.k) * 12, A.origins + (int64_t)img * 3,
                                              A.dirs + ((int64_t)img * A.rays + r) * 3, in_scene);
                const float u = (A.perturb && !A.t_in) ? A.rand[gs] : 0.f;
                const float t = pe_sample_t_or(A.t_in, gs, ray, p, P, A.perturb != 0, u);
                pe_position(ray, t, x);
                for (int a = 0; a < 3; ++a) { S.aux[a * TB + tid] = ray.o[a]; S.aux[(3 + a) * TB + tid] = ray.d[a]; }
                if (pe_in_box(ob, x)) flag |= 1;
                graw = B.g_raw ? B.g_raw[gs] : 0.f;
                gdm = B.g_dm ? B.g_dm[gs] : 0.f;
            }
            S.flags[tid] = flag;
            S.clampf[tid] = 0;
            S.graw[tid] = graw;
            S.gdm[tid] = (flag & 1) ? gdm : 0.f;
            for (int a = 0; a < 3; ++a) { S.pos[a * TB + tid] = x[a]; S.bent[a * TB + tid] = x[a]; }
        }
        const int any_inbox = __syncthreads_or(tid < TB ? (S.flags[tid] & 1) : 0);
        if (!any_inbox) {
            if (full && tid < TB && (S.flags[tid] & 4)) {
                if (B.div_out) B.div_out[gsi + S.slot[tid]] = 0.f;
                for (int a = 0; a < 3; ++a) B.g_pos[(gsi + S.slot[tid]) * 3 + a] = 0.f;
                if (B.g_od) for (int a = 0; a < 6; ++a) B.g_od[(gsi + S.slot[tid]) * 6 + a] = 0.f;
            }
            continue;
        }
        // ---- ray bender forward (positional_ray_bender_model.py:81-163) ----
        if (positional) {
            const int Eb = 3 * (1 + 2 * ob.b_octaves);
            const float* dfm = A.deformation + (int64_t)img * ob.deformation_features;
            for (int idx = tid; idx < L.b_enc * TB; idx += NT) {
                const int e = idx / TB, m = idx - e * TB;
                float v;
                if (e < Eb) {
                    const float xn[3] = {__fdiv_rn(S.pos[m], size[0]), __fdiv_rn(S.pos[TB + m], size[1]), __fdiv_rn(S.pos[2 * TB + m], size[2])};
                    v = enc_value(xn, 3, e, ob.b_anneal);
                } else {
                    v = __ldg(dfm + (e - Eb));
                }
                S.enc[e * TS + m] = v;
                st(ST.benc)[e * TS + m] = v;
            }
            __syncthreads();
            const float* cur = S.enc;
            int curK = L.b_enc, which = 0;          // (the bender's own chain; shadows the field's)
            for (int l = 0; l < ob.b_layers; ++l) {
                float* nxt = bufs[which];
                float* sto = st(ST.bh[l]);
                const float* bias = P32(L.bd_b[l]);
                gemm_tile(cur, curK, S.enc, l == ob.b_skip ? L.b_enc : 0, P32(L.bd_w[l]), ob.b_width, ob.b_width, S.wS,
                          [&](int n, int m0, const float* acc) {
                              const float b = __ldg(bias + n);
                              float4 v = make_float4(fmaxf(acc[0] + b, 0.f), fmaxf(acc[1] + b, 0.f), fmaxf(acc[2] + b, 0.f), fmaxf(acc[3] + b, 0.f));
                              *reinterpret_cast<float4*>(nxt + n * TS + m0) = v;
                              *reinterpret_cast<float4*>(sto + n * TS + m0) = v;
                          });
                cur = nxt; curK = ob.b_width; which ^= 1;
            }
            float* out3 = bufs[which];
            gemm_tile(cur, curK, nullptr, 0, P32(L.bd_out_w), 3, 3, S.wS, [&](int n, int m0, const float* acc) {
                *reinterpret_cast<float4*>(out3 + n * TS + m0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            });
            if (tid < TB) {
                int cf = 0;
                for (int a = 0; a < 3; ++a) {
                    const float x = S.pos[a * TB + tid];
                    const float raw = __fmul_rn(out3[a * TS + tid], size[a]);
                    const float lo = __fsub_rn(ob.bbox[2 * a], x), hi = __fsub_rn(ob.bbox[2 * a + 1], x);
                    float dsp = fmaxf(raw, lo);                      // clamp_output :116-140
                    const bool pass = !(raw < lo) && !(dsp > hi);    // gradient reaches the network output
                    dsp = fminf(dsp, hi);
                    if (ob.canonical_pose) dsp = __fmul_rn(dsp, 0.f);
                    if (!(S.flags[tid] & 1)) dsp = 0.f;
                    if (pass) cf |= 1 << a;
                    S.aux[(6 + a) * TB + tid] = dsp;
                    S.bent[a * TB + tid] = __fadd_rn(x, dsp);
                }
                S.clampf[tid] = cf;
            }
            __syncthreads();
        }
        if (tid < TB) {                 // inner mask of the field on the bent position (adain_style_nerf_model.py:171-184)
            const float xb[3] = {S.bent[tid], S.bent[TB + tid], S.bent[2 * TB + tid]};
            int f = S.flags[tid];
            if ((f & 1) && (skybox || pe_in_box(ob, xb))) f |= 2;
            S.flags[tid] = f;
            if (!(f & 2) || !in_scene || skybox) S.graw[tid] = 0.f;
        }
        __syncthreads();
        if (!bender_only) {
        // ---- positional encoding ----
        for (int idx = tid; idx < L.enc * TB; idx += NT) {
            const int e = idx / TB, m = idx - e * TB;
            float xin[6];
            if (skybox) {
                const float d0 = S.aux[3 * TB + m], d1 = S.aux[4 * TB + m], d2 = S.aux[5 * TB + m];
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
                xin[0] = __fdiv_rn(S.aux[m], size[0]); xin[1] = __fdiv_rn(S.aux[TB + m], size[1]); xin[2] = __fdiv_rn(S.aux[2 * TB + m], size[2]);
                xin[3] = __fdiv_rn(d0, nrm); xin[4] = __fdiv_rn(d1, nrm); xin[5] = __fdiv_rn(d2, nrm);
            } else {
                xin[0] = __fdiv_rn(S.bent[m], size[0]); xin[1] = __fdiv_rn(S.bent[TB + m], size[1]); xin[2] = __fdiv_rn(S.bent[2 * TB + m], size[2]);
                xin[3] = xin[4] = xin[5] = 0.f;
            }
            const float v = enc_value(xin, L.in_dims, e, nullptr);
            S.enc[e * TS + m] = v;
            st(ST.enc)[e * TS + m] = v;
        }
        __syncthreads();
        // ---- backbone ----
        cur = S.enc; curK = L.enc; which = 0;
        for (int l = 0; l < ob.layers; ++l) {
            float* nxt = bufs[which];
            float* sto = st(ST.h[l]);
            const float* bias = P32(L.bb_b[l]);
            gemm_tile(cur, curK, S.enc, l == ob.skip ? L.enc : 0, P32(L.bb_w[l]), W, W, S.wS, [&](int n, int m0, const float* acc) {
                const float b = __ldg(bias + n);
                float4 v = make_float4(fmaxf(acc[0] + b, 0.f), fmaxf(acc[1] + b, 0.f), fmaxf(acc[2] + b, 0.f), fmaxf(acc[3] + b, 0.f));
                *reinterpret_cast<float4*>(nxt + n * TS + m0) = v;
                *reinterpret_cast<float4*>(sto + n * TS + m0) = v;
            });
            cur = nxt; curK = W; which ^= 1;
        }
        }   // !bender_only
        }   // !from_cache
        float* gA = S.bufA;
        float* gB = S.bufB;
        float* gcur = gB;
        float* gnext = gA;
        if (!bender_only) {
        // ---- feature head with AdaIn ----
        const float* sc1 = A.aff1 + (int64_t)img * 2 * W;
        const float* sh1 = sc1 + W;
        const float* sc2 = A.aff2 + (int64_t)img * W;
        const float* sh2 = sc2 + W / 2;
        {
            float* nxt = bufs[which];
            float* spre = st(ST.x1pre);
            float* sy = st(ST.y1);
            gemm_tile(cur, W, nullptr, 0, P32(L.head0_w), W, W, S.wS, [&](int n, int m0, const float* acc) {
                const float s = sc1[n], h = sh1[n];
                *reinterpret_cast<float4*>(spre + n * TS + m0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                float4 v = make_float4(fmaxf(fmaf(acc[0], s, h), 0.f), fmaxf(fmaf(acc[1], s, h), 0.f), fmaxf(fmaf(acc[2], s, h), 0.f), fmaxf(fmaf(acc[3], s, h), 0.f));
                *reinterpret_cast<float4*>(nxt + n * TS + m0) = v;
                *reinterpret_cast<float4*>(sy + n * TS + m0) = v;
            });
            cur = nxt; which ^= 1;
        }
        {
            float* nxt = bufs[which];
            float* spre = st(ST.x2pre);
            float* sy = st(ST.y2);
            gemm_tile(cur, W, nullptr, 0, P32(L.head3_w), W / 2, W / 2, S.wS, [&](int n, int m0, const float* acc) {
                const float s = sc2[n], h = sh2[n];
                *reinterpret_cast<float4*>(spre + n * TS + m0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                float4 v = make_float4(fmaxf(fmaf(acc[0], s, h), 0.f), fmaxf(fmaf(acc[1], s, h), 0.f), fmaxf(fmaf(acc[2], s, h), 0.f), fmaxf(fmaf(acc[3], s, h), 0.f));
                *reinterpret_cast<float4*>(nxt + n * TS + m0) = v;
                *reinterpret_cast<float4*>(sy + n * TS + m0) = v;
            });
            cur = nxt; which ^= 1;
        }
        if (A.apply_activation) {       // sigmoid on the features (object_composer.py:548-549): its derivative needs the head output
            float* fo = S.bufX;
            const float* bias = P32(L.head6_b);
            gemm_tile(cur, W / 2, nullptr, 0, P32(L.head6_w), F, F, S.wS, [&](int n, int m0, const float* acc) {
                const float b = __ldg(bias + n);
                *reinterpret_cast<float4*>(fo + n * TS + m0) = make_float4(acc[0] + b, acc[1] + b, acc[2] + b, acc[3] + b);
            });
        }

        // ================================ backward ================================
        // upstream gradient of the per-sample features: w_obj * dL/dF_obj[ray] + w_glob * dL/dF_glob[ray]
        for (int idx = tid; idx < F * TB; idx += NT) {
            const int m = idx / F, c = idx - m * F;
            float g = 0.f;
            if (S.flags[m] & 2) {
                const int64_t gs = gsi + S.slot[m];
                const int64_t ray = (int64_t)img * A.rays + S.slot[m] / P;
                if (B.g_feat_obj) g = B.cw_obj[gs] * __ldg(B.g_feat_obj + ray * F + c);
                if (B.g_feat_glob) g = fmaf(B.cw_glob[gs], __ldg(B.g_feat_glob + ray * F + c), g);
                if (A.apply_activation) {
                    const float f = 1.f / (1.f + expf(-S.bufX[c * TS + m]));
                    g *= f * (1.f - f);
                }
            }
            gA[c * TS + m] = g;
        }
        __syncthreads();
        // ---- features_head.6 ----
        load_rows(S.bufX, st(ST.y2), W / 2);
        __syncthreads();
        if (full) outer_acc(gA, F, S.bufX, W / 2, B.gw.head6_w, W / 2, B.gw.head6_b);
        gemm_tile(gA, F, nullptr, 0, B.w.head6_w, W / 2, W / 2, S.wS, [&](int n, int m0, const float* acc) {
            const float4 y = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
            *reinterpret_cast<float4*>(gB + n * TS + m0) =
                make_float4(y.x > 0.f ? acc[0] : 0.f, y.y > 0.f ? acc[1] : 0.f, y.z > 0.f ? acc[2] : 0.f, y.w > 0.f ? acc[3] : 0.f);
        });
        float* asum = B.adain_sums + (int64_t)img * 3 * W;
        adain_backward(gB, W / 2, st(ST.x2pre), sc2, k1_2, k2_2, S.flags, asum + 2 * W, asum + 2 * W + W / 2,
                       B.bn_sums ? B.bn_sums + 2 * W : nullptr, B.bn_sums ? B.bn_sums + 2 * W + W / 2 : nullptr, full, phase == 1);
        if (phase == 1) continue;
        // ---- features_head.3 ----
        load_rows(S.bufX, st(ST.y1), W);
        __syncthreads();
        if (full) outer_acc(gB, W / 2, S.bufX, W, B.gw.head3_w, W, nullptr);
        gemm_tile(gB, W / 2, nullptr, 0, B.w.head3_w, W, W, S.wS, [&](int n, int m0, const float* acc) {
            const float4 y = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
            *reinterpret_cast<float4*>(gA + n * TS + m0) =
                make_float4(y.x > 0.f ? acc[0] : 0.f, y.y > 0.f ? acc[1] : 0.f, y.z > 0.f ? acc[2] : 0.f, y.w > 0.f ? acc[3] : 0.f);
        });
        adain_backward(gA, W, st(ST.x1pre), sc1, k1_1, k2_1, S.flags, asum, asum + W, B.bn_sums, B.bn_sums ? B.bn_sums + W : nullptr, full, phase == 2);
        if (phase == 2) continue;
        // ---- features_head.0 and the alpha head: both read the trunk output ----
        load_rows(S.bufX, st(ST.h[ob.layers - 1]), W);
        __syncthreads();
        outer_acc(gA, W, S.bufX, W, B.gw.head0_w, W, nullptr);
        if (!skybox) {
            const int warp = tid >> 5, lane = tid & 31;
            const float gr = S.graw[lane];
            if (B.gw.alpha_w) {
                for (int c = warp; c < W; c += NT / 32) {
                    const float v = warp_sum(gr * S.bufX[c * TS + lane]);
                    if (lane == 0 && v != 0.f) atomicAdd(B.gw.alpha_w + c, v);
                }
            }
            if (warp == 0 && B.gw.alpha_b) {
                const float v = warp_sum(gr);
                if (lane == 0 && v != 0.f) atomicAdd(B.gw.alpha_b, v);
            }
        }
        {
            const float* aw = skybox ? nullptr : B.w.alpha_w;
            gemm_tile(gA, W, nullptr, 0, B.w.head0_w, W, W, S.wS, [&](int n, int m0, const float* acc) {
                const float4 h = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
                const float a = aw ? __ldg(aw + n) : 0.f;
                *reinterpret_cast<float4*>(gB + n * TS + m0) =
                    make_float4(h.x > 0.f ? fmaf(a, S.graw[m0], acc[0]) : 0.f, h.y > 0.f ? fmaf(a, S.graw[m0 + 1], acc[1]) : 0.f,
                                h.z > 0.f ? fmaf(a, S.graw[m0 + 2], acc[2]) : 0.f, h.w > 0.f ? fmaf(a, S.graw[m0 + 3], acc[3]) : 0.f);
            });
        }
        // ---- trunk, last layer first; gcur = dL/d(pre-activation of layer l) ----
        for (int idx = tid; idx < L.enc * TS; idx += NT) S.enc[idx] = 0.f;       // becomes the gradient of the encoding
        gcur = gB;
        gnext = gA;
        for (int l = ob.layers - 1; l >= 0; --l) {
            const int first = l == 0 ? L.enc : W;                   // rows of the first input segment
            const int Kl = L.k_in[l];
            load_rows(S.bufX, st(l == 0 ? ST.enc : ST.h[l - 1]), first);
            if (l == ob.skip) load_rows(S.bufX + first * TS, st(ST.enc), L.enc);
            __syncthreads();
            outer_acc(gcur, W, S.bufX, Kl, B.gw.backbone_w[l], Kl, B.gw.backbone_b[l]);
            gemm_tile(gcur, W, nullptr, 0, B.w.backbone_w[l], Kl, Kl, S.wS, [&](int n, int m0, const float* acc) {
                if (l > 0 && n < first) {
                    const float4 h = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
                    *reinterpret_cast<float4*>(gnext + n * TS + m0) =
                        make_float4(h.x > 0.f ? acc[0] : 0.f, h.y > 0.f ? acc[1] : 0.f, h.z > 0.f ? acc[2] : 0.f, h.w > 0.f ? acc[3] : 0.f);
                } else {
                    float* ge = S.enc + (l > 0 ? n - first : (n < L.enc ? n : n - L.enc)) * TS + m0;
                    atomicAdd(ge, acc[0]); atomicAdd(ge + 1, acc[1]); atomicAdd(ge + 2, acc[2]); atomicAdd(ge + 3, acc[3]);
                }
            });
            float* tsw = gcur; gcur = gnext; gnext = tsw;
        }
        }   // !bender_only
        // ---- positional encoding backward -> gradient of the (bent) position, or of origin / direction for the skybox ----
        float gbent[3] = {0.f, 0.f, 0.f};
        if (bender_only) {
            if (tid < TB && (S.flags[tid] & (B.g_bent_flag ? B.g_bent_flag : 2))) {
                const float* g = B.g_bent_in + (gsi + S.slot[tid]) * 3;
                gbent[0] = g[0]; gbent[1] = g[1]; gbent[2] = g[2];
            }
        } else if (tid < TB) {
            const int m = tid;
            if (skybox) {
                if (B.g_od && (S.flags[m] & 4)) {
                    const float o[3] = {S.aux[m], S.aux[TB + m], S.aux[2 * TB + m]};
                    const float d[3] = {S.aux[3 * TB + m], S.aux[4 * TB + m], S.aux[5 * TB + m]};
                    const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    const float xin[6] = {__fdiv_rn(o[0], size[0]), __fdiv_rn(o[1], size[1]), __fdiv_rn(o[2], size[2]), d[0] / nrm, d[1] / nrm, d[2] / nrm};
                    float gx[6];
                    for (int a = 0; a < 6; ++a) gx[a] = (S.flags[m] & 2) ? enc_backward(S.enc + m, xin, 6, ob.octaves, a, nullptr) : 0.f;
                    const float dot = gx[3] * xin[3] + gx[4] * xin[4] + gx[5] * xin[5];
                    for (int a = 0; a < 3; ++a) {
                        B.g_od[(gsi + S.slot[m]) * 6 + a] = gx[a] / size[a];
                        B.g_od[(gsi + S.slot[m]) * 6 + 3 + a] = (gx[3 + a] - xin[3 + a] * dot) / nrm;      // d(d/|d|)
                    }
                }
            } else if (S.flags[m] & 2) {
                const float xn[3] = {__fdiv_rn(S.bent[m], size[0]), __fdiv_rn(S.bent[TB + m], size[1]), __fdiv_rn(S.bent[2 * TB + m], size[2])};
                for (int a = 0; a < 3; ++a) gbent[a] = enc_backward(S.enc + m, xn, 3, ob.octaves, a, nullptr) / size[a];
            }
        }
        float gpos[3] = {gbent[0], gbent[1], gbent[2]};
        // ---- ray bender backward ----
        if (positional) {
            __syncthreads();
            float* g3 = gcur;                    // rows 0..2: gradient of the bender output
            if (tid < TB) {
                const int m = tid;
                const float d[3] = {S.aux[6 * TB + m], S.aux[7 * TB + m], S.aux[8 * TB + m]};
                const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int a = 0; a < 3; ++a) {
                    float gd = 0.f;
                    if ((S.flags[m] & 1) && !ob.canonical_pose) {
                        gd = gbent[a];
                        if (nrm > 0.f) gd = fmaf(S.gdm[m], d[a] / nrm, gd);      // torch.norm backward, 0 at the origin
                    }
                    float graw = 0.f;
                    if ((S.clampf[m] >> a) & 1) graw = gd;           // inside the clamp: the network output moves the sample
                    else gpos[a] -= gd;                               // clamped to the box face: displacement = face - x
                    g3[a * TS + m] = graw * size[a];
                }
            }
            __syncthreads();
            load_rows(S.bufX, st(ST.bh[ob.b_layers - 1]), ob.b_width);
            __syncthreads();
            outer_acc(g3, 3, S.bufX, ob.b_width, B.gw.bender_out_w, ob.b_width, nullptr);
            float* gb = gnext;
            gemm_tile(g3, 3, nullptr, 0, B.w.bender_out_w, ob.b_width, ob.b_width, S.wS, [&](int n, int m0, const float* acc) {
                const float4 h = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
                *reinterpret_cast<float4*>(gb + n * TS + m0) =
                    make_float4(h.x > 0.f ? acc[0] : 0.f, h.y > 0.f ? acc[1] : 0.f, h.z > 0.f ? acc[2] : 0.f, h.w > 0.f ? acc[3] : 0.f);
            });
            for (int idx = tid; idx < L.b_enc * TS; idx += NT) S.enc[idx] = 0.f;
            float* bcur = gb;
            float* bnext = g3;
            for (int l = ob.b_layers - 1; l >= 0; --l) {
                const int first = l == 0 ? L.b_enc : ob.b_width;
                const int Kl = L.b_k_in[l];
                load_rows(S.bufX, st(l == 0 ? ST.benc : ST.bh[l - 1]), first);
                if (l == ob.b_skip) load_rows(S.bufX + first * TS, st(ST.benc), L.b_enc);
                __syncthreads();
                outer_acc(bcur, ob.b_width, S.bufX, Kl, B.gw.bender_w[l], Kl, B.gw.bender_b[l]);
                gemm_tile(bcur, ob.b_width, nullptr, 0, B.w.bender_w[l], Kl, Kl, S.wS, [&](int n, int m0, const float* acc) {
                    if (l > 0 && n < first) {
                        const float4 h = *reinterpret_cast<const float4*>(S.bufX + n * TS + m0);
                        *reinterpret_cast<float4*>(bnext + n * TS + m0) =
                            make_float4(h.x > 0.f ? acc[0] : 0.f, h.y > 0.f ? acc[1] : 0.f, h.z > 0.f ? acc[2] : 0.f, h.w > 0.f ? acc[3] : 0.f);
                    } else {
                        float* ge = S.enc + (l > 0 ? n - first : (n < L.b_enc ? n : n - L.b_enc)) * TS + m0;
                        atomicAdd(ge, acc[0]); atomicAdd(ge + 1, acc[1]); atomicAdd(ge + 2, acc[2]); atomicAdd(ge + 3, acc[3]);
                    }
                });
                float* tsw = bcur; bcur = bnext; bnext = tsw;
            }
            const int Eb = 3 * (1 + 2 * ob.b_octaves);
            if (tid < TB && (S.flags[tid] & 1)) {
                const int m = tid;
                const float xn[3] = {__fdiv_rn(S.pos[m], size[0]), __fdiv_rn(S.pos[TB + m], size[1]), __fdiv_rn(S.pos[2 * TB + m], size[2])};
                for (int a = 0; a < 3; ++a) gpos[a] += enc_backward(S.enc + m, xn, 3, ob.b_octaves, a, ob.b_anneal) / size[a];
            }
            if (B.g_deformation) {               // the deformation code is replicated over the samples of its image
                const int warp = tid >> 5, lane = tid & 31;
                for (int j = warp; j < ob.deformation_features; j += NT / 32) {
                    const float v = warp_sum((S.flags[lane] & 1) ? S.enc[(Eb + j) * TS + lane] : 0.f);
                    if (lane == 0 && v != 0.f) atomicAdd(B.g_deformation + (int64_t)img * ob.deformation_features + j, v);
                }
            }
        }
        if (tid < TB && (S.flags[tid] & 4)) {
            const bool use = (S.flags[tid] & 1) != 0;
            for (int a = 0; a < 3; ++a) B.g_pos[(gsi + S.slot[tid]) * 3 + a] = use ? gpos[a] : 0.f;
            // e . (J e): gpos = e + J^T e (identity path of bent = x + displacement, plus the path through the bender and its clamp)
            if (B.div_out) {
                float dv = 0.f;
                for (int a = 0; a < 3; ++a) dv = fmaf(gbent[a], gpos[a] - gbent[a], dv);
                B.div_out[gsi + S.slot[tid]] = use ? dv : 0.f;
            }
        }
    }
}

// Per image: the slots whose flag has a bit of `mask`, in slot order (one block per image, four slots per thread and iteration,
// warp-shuffle scan + per-warp totals).
__global__ void __launch_bounds__(1024) pe_compact_slots_kernel(const uint8_t* __restrict__ flags, int mask, int64_t slots_per_image,
                                                                 int32_t* __restrict__ list, int32_t* __restrict__ count) {
    // gridDim.x blocks share an image: each takes a contiguous range of slots (a multiple of the 4096 slots of one scan step), counts
    // its listed slots, reserves that many entries of the image's list with ONE atomic (count[] is zeroed by the launcher) and fills them
    // in slot order.  The ranges of different blocks land in reservation order: the list is ordered inside every range, which is all the
    // consumers need (tiles are independent; neighbouring slots stay neighbours).
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* f = flags + (int64_t)img * slots_per_image;
    int32_t* out = list + (int64_t)img * slots_per_image;
    const int64_t step = 4 * (int64_t)blockDim.x;
    const int64_t steps = (slots_per_image + step - 1) / step;
    const int64_t per_block = (steps + gridDim.x - 1) / gridDim.x;
    const int64_t begin = (int64_t)blockIdx.x * per_block * step;
    const int64_t end = begin + per_block * step < slots_per_image ? begin + per_block * step : slots_per_image;
    if (begin >= end) return;
    int mine = 0;
    for (int64_t s0 = begin; s0 < end; s0 += step) {
        const int64_t s = s0 + 4 * (int64_t)threadIdx.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) mine += (s + i < end && (f[s + i] & mask) != 0) ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if (lane == 0) warp_tot[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
        base = t ? atomicAdd(count + img, t) : 0;
    }
    __syncthreads();
    for (int64_t s0 = begin; s0 < end; s0 += step) {
        const int64_t s = s0 + 4 * (int64_t)threadIdx.x;
        bool keep[4];
        int c = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { keep[i] = s + i < end && (f[s + i] & mask) != 0; c += keep[i] ? 1 : 0; }
        int incl = c;                                   // inclusive scan of the per-thread counts inside the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
        __syncthreads();                                // (warp_tot of the previous step / of the counting pass has been read)
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int off = base + incl - c;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (keep[i]) out[off++] = (int32_t)(s + i);
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w]; base += t; }
    }
}

__global__ void pe_tile_prefix_kernel(const int32_t* __restrict__ count, int images, int32_t* __restrict__ tile_begin, int tile_rows) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < images; ++i) { tile_begin[i] = t; t += (count[i] + tile_rows - 1) / tile_rows; }
        tile_begin[images] = t;
    }
}

// Per image: ceil(#flagged slots / tile_rows), summed over the images into *out -- the tile count pe_compact_slots_kernel +
// pe_tile_prefix_kernel will arrive at for the same flags (pe_forward_tile_counts).
__global__ void __launch_bounds__(1024) pe_count_tiles_kernel(const uint8_t* __restrict__ flags, int mask, int64_t slots_per_image, int tile_rows,
                                                               unsigned long long* __restrict__ out) {
    __shared__ int warp_tot[32];
    const uint8_t* f = flags + (int64_t)blockIdx.x * slots_per_image;
    int mine = 0;
    for (int64_t s = threadIdx.x; s < slots_per_image; s += blockDim.x) mine += (f[s] & mask) != 0 ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
        if (t) atomicAdd(out, (unsigned long long)((t + tile_rows - 1) / tile_rows));
    }
}

}  // namespace

int pe_launch_count_tiles(const uint8_t* flags, int flag_mask, int images, int64_t slots_per_image, int tile_rows, int64_t* out,
                          cudaStream_t stream) {
    if (images == 0 || slots_per_image == 0) return PE_OK;
    pe_count_tiles_kernel<<<images, 1024, 0, stream>>>(flags, flag_mask, slots_per_image, tile_rows, reinterpret_cast<unsigned long long*>(out));
    PE_LAUNCH_CHECK("pe_count_tiles_kernel");
    return PE_OK;
}

int pe_launch_compact_slots(const uint8_t* flags, int flag_mask, int images, int64_t slots_per_image, int32_t* slot_list, int32_t* slot_count,
                            int32_t* tile_begin, cudaStream_t stream, int tile_rows) {
    if (images == 0 || slots_per_image == 0) return PE_OK;
    PE_CUDA_CHECK(cudaMemsetAsync(slot_count, 0, (size_t)images * sizeof(int32_t), stream));
    const int64_t steps = (slots_per_image + 4095) / 4096;
    const int per_image = (int)pe_min64(steps, images >= 64 ? 4 : 256 / images);
    pe_compact_slots_kernel<<<dim3(per_image, images), 1024, 0, stream>>>(flags, flag_mask, slots_per_image, slot_list, slot_count);
    PE_LAUNCH_CHECK("pe_compact_slots_kernel");
    pe_tile_prefix_kernel<<<1, 32, 0, stream>>>(slot_count, images, tile_begin, tile_rows);
    PE_LAUNCH_CHECK("pe_tile_prefix_kernel");
    return PE_OK;
}

size_t pe_field_bwd_smem_bytes() {
    return sizeof(float) * ((size_t)3 * CMAX * TS + (size_t)EMAX * TS + (size_t)KC * NPASS + 17 * TB) + sizeof(int) * 3 * TB;
}

int64_t pe_field_bwd_stash_floats(const PeObjectDesc& ob, const PeLayout& L) { return (int64_t)stash_layout(ob, L).rows * TS; }

int pe_field_bwd_grid(int sm_count) { return sm_count; }

int pe_launch_field_bwd(const PeFieldBwdArgs& args, int sm_count, cudaStream_t stream) {
    const PeObjectDesc& ob = args.f.ob;
    const PeLayout& L = args.f.L;
    const bool positional = ob.bender_kind == PE_BENDER_POSITIONAL;
    if (ob.width > 256 || ob.width % 8 || ob.features > 256 || L.enc > EMAX || ob.width + L.enc > CMAX ||
        (positional && (ob.b_width > 256 || L.b_enc > EMAX || ob.b_width + L.b_enc > CMAX))) {
        pe_set_error("field backward supports widths/features up to 256 and encodings up to %d values", EMAX);
        return PE_ERR_UNSUPPORTED;
    }
    if (args.f.explicit_positions) { pe_set_error("field backward on explicit positions is not supported"); return PE_ERR_UNSUPPORTED; }
    const size_t smem = pe_field_bwd_smem_bytes();
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_field_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (((int64_t)args.f.rays * ob.positions + TB - 1) / TB) * args.f.images;
    if (tiles == 0) return PE_OK;
    const int grid = (int)pe_min64(tiles, pe_field_bwd_grid(sm_count));
    pe_field_bwd_kernel<<<grid, NT, smem, stream>>>(args);
    PE_LAUNCH_CHECK("pe_field_bwd_kernel");
    return PE_OK;
}

// Backward kernels around the field backward (pe_field_bwd.cu):
//   pe_composite_bwd_kernel  gradients of ObjectComposer.integrate / compose (model/object_composer.py:153-214, 399-447, 724-784)
//                            w.r.t. the per-sample raw alphas, sample distances t, displacement magnitudes and features
//   pe_style_bwd_kernel      AdaIn affine + BatchNorm fold (model/layers/adain.py:21-61) -> affine_transform parameters and style code
//   pe_bn_fix_kernel         cross-sample terms of the train-mode BatchNorm backward
//   pe_geometry_bwd_kernel   sample positions / distances -> object-space ray -> slab test -> world rays and the w2o matrix
//                            (utils/lib_3d/ray_helper.py:1180-1282, model/object_composer.py:104-151, 522-523)
#include "pe_kernels.cuh"

namespace {

constexpr int WARPS = 4;
constexpr float BN_EPS = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int DOT_BATCH = 4;        // feature rows in flight per warp in the dL/dw dot products

struct Lists {                 // per-warp shared memory, TP entries each
    float *t, *raw, *dm;       // the ordered sample list
    int *id, *fl;              // (object << 16) | sample;  bit0: features present (in-box), bit1: masked by fix_object_overlaps
    float *al, *T, *w, *gw, *gd;   // alpha, exclusive transmittance, weight, dL/dw, dL/d delta
    float *ut, *uraw, *udm;    // unsorted concatenation (composition pass)
    int *uid, *ufl;
};

// Backward of ObjectComposer.integrate over the ordered list in S (n entries).  Writes per source sample (is_global: adds).
__device__ void backward_list(const PeCompositeBwdArgs& B, const Lists& S, int n, float dnorm, const float* __restrict__ noise,
                              const PeIntegratedGrads& G, int64_t ray, int lane, bool is_global, float& g_dnorm) {
    const PeCompositeArgs& A = B.f;
    const int F = A.features;
    // ---- forward quantities ----
    float carry = 1.f, opacity = 0.f, depth = 0.f;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int j = c0 + lane;
        float alpha = 0.f;
        if (j < n) {
            const float delta = __fmul_rn(j == n - 1 ? 1e10f : __fsub_rn(S.t[j + 1], S.t[j]), dnorm);
            float raw = S.raw[j];
            if (noise) raw = __fadd_rn(raw, noise[j]);
            alpha = __fsub_rn(1.f, expf(__fmul_rn(-fmaxf(raw, 0.f), delta)));
        }
        const float shifted = j < n ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
        float incl = shifted;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= v;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        const float T = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (j < n) {
            S.al[j] = alpha; S.T[j] = T; S.w[j] = alpha * T;
            opacity += alpha * T;
            depth += alpha * T * S.t[j];
        }
    }
    opacity = warp_sum(opacity);
    depth = warp_sum(depth);
    __syncwarp();
    // ---- upstream gradients of the per-ray outputs ----
    float d_op = G.opacity ? G.opacity[ray] : 0.f;
    float d_depth = G.depth ? G.depth[ray] : 0.f;
    const float d_disp = G.disparity ? G.disparity[ray] : 0.f;
    if (d_disp != 0.f) {                          // disparity = 1 / clamp(depth / opacity, 1e-10)   (:765)
        const float q = depth / opacity;
        if (q >= 1e-10f) {
            const float gq = -d_disp / (q * q);
            d_depth += gq / opacity;
            d_op -= gq * depth / (opacity * opacity);
        }
    }
    const float d_idm = G.integrated_displacements_magnitude ? G.integrated_displacements_magnitude[ray] / (float)n : 0.f;
    float dF[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        dF[i] = (G.integrated_features && c < F) ? G.integrated_features[ray * F + c] : 0.f;
    }
    // dL/dw_j = <dF, f_j> + d_opacity + d_depth * t_j + d_weights_j.  The feature term exists only for the samples inside a box: the
    // (sequential, warp-wide) dot products visit just those; the other terms are added for all samples in parallel, in the same order
    for (int j = lane; j < n; j += 32) S.gw[j] = 0.f;
    __syncwarp();
    if (G.integrated_features) {
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int j = c0 + lane;
            unsigned todo = __ballot_sync(0xffffffffu, j < n && (S.fl[j] & 1));
            // DOT_BATCH feature rows in flight per warp (latency bound otherwise); every dot product is evaluated as before
            while (todo) {
                int js[DOT_BATCH];
                const float* fp[DOT_BATCH];
                int cnt = 0;
#pragma unroll
                for (int u = 0; u < DOT_BATCH; ++u) {
                    js[u] = 0;
                    fp[u] = nullptr;
                    if (todo) {                               // warp-uniform
                        js[u] = c0 + __ffs(todo) - 1;
                        todo &= todo - 1;
                        const int id = S.id[js[u]];
                        const int k = id >> 16, p = id & 0xffff;
                        fp[u] = A.feat[k] + (ray * A.positions[k] + p) * (int64_t)F;
                        cnt = u + 1;
                    }
                }
                float v[DOT_BATCH];
#pragma unroll
                for (int u = 0; u < DOT_BATCH; ++u) {
                    float fv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = lane + 32 * i;
                        fv[i] = (u < cnt && c < F) ? __ldg(fp[u] + c) : 0.f;
                    }
                    v[u] = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[u] = fmaf(dF[i], fv[i], v[u]);
                }
#pragma unroll
                for (int u = 0; u < DOT_BATCH; ++u) {
                    if (u < cnt) {
                        const float r = warp_sum(v[u]);
                        if (lane == 0) S.gw[js[u]] = r;
                    }
                }
            }
        }
        __syncwarp();
    }
    for (int j = lane; j < n; j += 32) S.gw[j] = S.gw[j] + d_op + d_depth * S.t[j] + (G.weights ? G.weights[ray * n + j] : 0.f);
    __syncwarp();
    // dL/d alpha_j = gw_j * T_j - (sum_{i>j} gw_i w_i) / (1 - alpha_j + 1e-10)        (compute_weights :199-214)
    float suffix = 0.f;
    for (int c0 = ((n - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        const int j = c0 + lane;
        const float val = j < n ? S.gw[j] * S.w[j] : 0.f;
        float incl = val;                         // inclusive suffix sum inside the chunk
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += v;
        }
        const float after = incl - val + suffix;  // sum over i > j
        suffix += __shfl_sync(0xffffffffu, incl, 0);
        if (j < n) {
            const float alpha = S.al[j];
            const float g_alpha = S.gw[j] * S.T[j] - after / __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            const float span = j == n - 1 ? 1e10f : __fsub_rn(S.t[j + 1], S.t[j]);
            const float delta = __fmul_rn(span, dnorm);
            float raw = S.raw[j];
            if (noise) raw = __fadd_rn(raw, noise[j]);
            const float r = fmaxf(raw, 0.f);
            const float e = expf(__fmul_rn(-r, delta));
            S.gw[j] = raw > 0.f ? g_alpha * delta * e : 0.f;      // now dL/d raw alpha
            S.gd[j] = g_alpha * r * e;                             // dL/d delta
        }
    }
    __syncwarp();
    // ---- scatter to the source samples ----
    float gdn = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float span = j == n - 1 ? 1e10f : __fsub_rn(S.t[j + 1], S.t[j]);
        gdn = fmaf(S.gd[j], span, gdn);                            // delta = span * |d|
        float g_t = d_depth * S.w[j] - S.gd[j] * dnorm;
        if (j > 0) g_t = fmaf(S.gd[j - 1], dnorm, g_t);
        const int id = S.id[j], fl = S.fl[j];
        const int k = id >> 16, p = id & 0xffff;
        const int64_t dst = ray * A.positions[k] + p;
        const float cw = (fl & 1) ? S.w[j] : 0.f;
        const float g_dm = d_idm * S.w[j];
        if (!is_global) {
            B.cw_obj[k][dst] = cw;
            B.g_raw[k][dst] = S.gw[j];
            B.g_t[k][dst] = g_t;
            B.g_dm[k][dst] = g_dm;
        } else {
            B.cw_glob[k][dst] = cw;
            if (!(fl & 2)) {                                       // overlap-masked samples are constants (:354-358)
                B.g_raw[k][dst] += S.gw[j];
                B.g_t[k][dst] += g_t;
                B.g_dm[k][dst] += g_dm;
            }
        }
    }
    g_dnorm += warp_sum(gdn);
    __syncwarp();
}

__device__ __forceinline__ bool any_grad(const PeIntegratedGrads& g) {
    return g.integrated_features || g.opacity || g.weights || g.depth || g.disparity || g.integrated_displacements_magnitude;
}

__global__ void __launch_bounds__(WARPS * 32) pe_composite_bwd_kernel(const PeCompositeBwdArgs B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PeCompositeArgs& A = B.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TP = (A.total_positions + 31) & ~31;
    float* base = reinterpret_cast<float*>(smem_raw) + (size_t)warp * TP * 15;
    Lists S;
    S.t = base; S.raw = base + TP; S.dm = base + 2 * TP; S.id = reinterpret_cast<int*>(base + 3 * TP); S.fl = reinterpret_cast<int*>(base + 4 * TP);
    S.al = base + 5 * TP; S.T = base + 6 * TP; S.w = base + 7 * TP; S.gw = base + 8 * TP; S.gd = base + 9 * TP;
    S.ut = base + 10 * TP; S.uraw = base + 11 * TP; S.udm = base + 12 * TP;
    S.uid = reinterpret_cast<int*>(base + 13 * TP); S.ufl = reinterpret_cast<int*>(base + 14 * TP);
    const int64_t n_rays = (int64_t)A.images * A.rays;
    const bool global_grads = any_grad(B.g_global);
    for (int64_t ray = (int64_t)blockIdx.x * WARPS + warp; ray < n_rays; ray += (int64_t)gridDim.x * WARPS) {
        const float* d = A.dirs + ray * 3;
        const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        float g_dnorm = 0.f;
        // ---- per-object integration ----
        for (int k = 0; k < A.objects; ++k) {
            const int P = A.positions[k];
            const int64_t b = ray * P;
            __syncwarp();
            if (!any_grad(B.g_object[k])) {
                for (int p = lane; p < P; p += 32) { B.cw_obj[k][b + p] = 0.f; B.g_raw[k][b + p] = 0.f; B.g_t[k][b + p] = 0.f; B.g_dm[k][b + p] = 0.f; }
                continue;
            }
            for (int p = lane; p < P; p += 32) {
                S.t[p] = A.t[k][b + p];
                S.raw[p] = A.raw[k][b + p];
                S.dm[p] = A.dispmag[k] ? A.dispmag[k][b + p] : 0.f;
                S.id[p] = (k << 16) | p;
                S.fl[p] = A.inbox[k][b + p] ? 1 : 0;
            }
            __syncwarp();
            backward_list(B, S, P, dnorm, (A.perturb && A.noise[k]) ? A.noise[k] + b : nullptr, B.g_object[k], ray, lane, false, g_dnorm);
        }
        // ---- composition of all objects ----
        if (!global_grads) {
            for (int k = 0; k < A.objects; ++k)
                for (int p = lane; p < A.positions[k]; p += 32) B.cw_glob[k][ray * A.positions[k] + p] = 0.f;
        } else {
            __syncwarp();
            int off = 0;
            for (int k = 0; k < A.objects; ++k) {
                const int P = A.positions[k];
                const int64_t b = ray * P;
                for (int p0 = 0; p0 < P; p0 += 32) {
                    const int p = p0 + lane;
                    float t = 0.f, raw = 0.f, dm = 0.f;
                    int fl = 0;
                    if (p < P) {
                        t = A.t[k][b + p];
                        raw = A.raw[k][b + p];
                        dm = A.dispmag[k] ? A.dispmag[k][b + p] : 0.f;
                        fl = A.inbox[k][b + p] ? 1 : 0;
                    }
                    if (A.fix_overlaps && k < A.static_objects) {      // same interval test as the forward compositor
                        bool masked = false;
                        for (int dk = A.static_objects; dk < A.objects; ++dk) {
                            const float* td = A.t[dk] + ray * A.positions[dk];
                            const float v0 = td[0], v1 = td[P - 1];
                            int lo = 0, hi = 0;
                            for (int q0 = 0; q0 < P; q0 += 32) {
                                const int q = q0 + lane;
                                const float tq = q < P ? A.t[k][b + q] : INFINITY;
                                lo += __popc(__ballot_sync(0xffffffffu, tq < v0));
                                hi += __popc(__ballot_sync(0xffffffffu, tq < v1));
                            }
                            masked = masked || (p >= lo && p < hi);
                        }
                        if (masked) { raw = raw * 0.f - 10.f; t = 0.f; dm = 0.f; fl |= 2; }
                    }
                    if (p < P) { S.ut[off + p] = t; S.uraw[off + p] = raw; S.udm[off + p] = dm; S.uid[off + p] = (k << 16) | p; S.ufl[off + p] = fl; }
                }
                off += P;
            }
            __syncwarp();
            const int n = A.total_positions;
            const bool ordered = pe_lists_ordered(S.ut, n, A.positions, A.objects, lane);
            for (int j = lane; j < n; j += 32) {               // stable sort by t, same tie order as the forward
                const float tj = S.ut[j];
                const int rank = pe_compose_rank(S.ut, n, j, A.positions, A.objects, ordered);
                S.t[rank] = tj; S.raw[rank] = S.uraw[j]; S.dm[rank] = S.udm[j]; S.id[rank] = S.uid[j]; S.fl[rank] = S.ufl[j];
            }
            __syncwarp();
            backward_list(B, S, n, dnorm, (A.perturb && A.noise_global) ? A.noise_global + ray * n : nullptr, B.g_global, ray, lane, true, g_dnorm);
        }
        if (B.g_dirs && lane < 3 && g_dnorm != 0.f) B.g_dirs[ray * 3 + lane] += g_dnorm * d[lane] / dnorm;      // d|d|/dd
    }
}

// AdaIn affine (adain.py:30-32, 58-59) with the BatchNorm fold of pe_style_kernel: sc = scale/sigma, sh = bias - mean*sc.
// Given per image A[c] = sum g, Bx[c] = sum g*x over the samples:  d bias = A,  d scale = (Bx - mean*A)/sigma.
__global__ void pe_style_bwd_kernel(const PeStyleBwdArgs A) {
    extern __shared__ float denc[];           // [2C]
    const int img = blockIdx.x;
    const int C = A.channels, S = A.style_features;
    const float* sums = A.adain_sums + (int64_t)img * A.adain_stride;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean = A.run_mean[c], var = A.run_var[c];
        if (A.training) {
            const double n = A.stats[2 * C];
            if (n > 0.0) {
                const double m = A.stats[c] / n;
                mean = (float)m;
                var = (float)fmax(A.stats[C + c] / n - m * m, 0.0);
            }
        }
        const float inv = 1.f / sqrtf(var + BN_EPS);
        denc[c] = (sums[C + c] - mean * sums[c]) * inv;
        denc[C + c] = sums[c];
    }
    __syncthreads();
    // gridDim.y blocks share an image: each recomputes the 2C coefficients above (cheap) and takes a slice of the outer product and of
    // the style gradient (one warp per style feature, lanes stride the 2C rows)
    const float* style = A.style + (int64_t)img * S;
    const int part = blockIdx.y, parts = gridDim.y;
    if (A.g_aff_b && part == 0)
        for (int j = threadIdx.x; j < 2 * C; j += blockDim.x)
            if (denc[j] != 0.f) atomicAdd(A.g_aff_b + j, denc[j]);
    if (A.g_aff_w)
        for (int i = part * blockDim.x + threadIdx.x; i < 2 * C * S; i += parts * blockDim.x) {
            const int j = i / S, s = i - j * S;
            const float v = denc[j] * style[s];
            if (v != 0.f) atomicAdd(A.g_aff_w + i, v);
        }
    if (A.g_style) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
        for (int s = part * warps + warp; s < S; s += parts * warps) {
            float v = 0.f;
            for (int j = lane; j < 2 * C; j += 32) v = fmaf(A.aff_w[(int64_t)j * S + s], denc[j], v);
            v = warp_sum(v);
            if (lane == 0) A.g_style[(int64_t)img * S + s] += v;
        }
    }
}

// fwd: sum x [C], sum x^2 [C], count; sums: S1 = sum g*sc [C], S2 = sum g*sc*x [C]  ->  fix: k1 [C], k2 [C] with
// dL/dx = g*sc - k1 - x*k2 (standard BatchNorm backward, biased variance, written in terms of the raw x)
__global__ void pe_bn_fix_kernel(const double* __restrict__ fwd, const double* __restrict__ sums, int C, float* __restrict__ fix) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double n = fwd[2 * C];
    if (!(n > 0.0)) { fix[c] = 0.f; fix[C + c] = 0.f; return; }
    const double mean = fwd[c] / n;
    const double var = fmax(fwd[C + c] / n - mean * mean, 0.0) + (double)BN_EPS;
    const double k2 = (sums[C + c] - mean * sums[c]) / (n * var);
    fix[c] = (float)(sums[c] / n - mean * k2);
    fix[C + c] = (float)k2;
}

__global__ void __launch_bounds__(128) pe_geometry_bwd_kernel(const PeGeometryBwdArgs G) {
    __shared__ float red[4][15];
    const PeObjectDesc& ob = G.ob;
    const int img = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = ob.positions;
    const float* m34 = G.w2o + ((int64_t)img * G.objects + G.k) * 12;
    float acc[15];                      // d w2o (12), d world origin (3)
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = 0.f;
    if (r < G.rays) {
        const int64_t ray = (int64_t)img * G.rays + r;
        const bool in_scene = G.ois ? G.ois[(int64_t)img * G.objects + G.k] != 0 : true;
        const float ow[3] = {G.origins[img * 3], G.origins[img * 3 + 1], G.origins[img * 3 + 2]};
        const float dw[3] = {G.dirs[ray * 3], G.dirs[ray * 3 + 1], G.dirs[ray * 3 + 2]};
        PeRay pr;
        pe_transform(m34, ow, true, pr.o);
        pe_transform(m34, dw, false, pr.d);
        // slab test again, remembering which face bounds the ray (object_composer.py:104-151)
        const float eps = 1e-6f;
        float z_near = -INFINITY, z_far = INFINITY, zn_z = 0.f, zf_z = 0.f;
        int near_axis = 0, far_axis = 0;
        for (int a = 0; a < 3; ++a) {
            const float den = __fadd_rn(pr.d[a], eps);
            const float z0 = __fdiv_rn(__fsub_rn(ob.bbox[2 * a], pr.o[a]), den);
            const float z1 = __fdiv_rn(__fsub_rn(ob.bbox[2 * a + 1], pr.o[a]), den);
            const float lo = fminf(z0, z1), hi = fmaxf(z0, z1);
            if (lo > z_near) { z_near = lo; near_axis = a; zn_z = lo; }
            if (hi < z_far) { z_far = hi; far_axis = a; zf_z = hi; }
        }
        const bool missed = z_far <= z_near || !in_scene;
        if (missed) { z_near = 0.f; z_far = 0.f; }
        const bool near_pass = !missed && z_near >= ob.z_near_min && z_near <= ob.z_far_max;     // clamp :522-523
        const bool far_pass = !missed && z_far >= ob.z_near_min && z_far <= ob.z_far_max;
        pr.z_near = fminf(fmaxf(z_near, ob.z_near_min), ob.z_far_max);
        pr.z_far = fminf(fmaxf(z_far, ob.z_near_min), ob.z_far_max);
        PeRay unit_near = pr, unit_far = pr;                   // t is linear in (near, far): coefficients from unit rays
        unit_near.z_near = 1.f; unit_near.z_far = 0.f;
        unit_far.z_near = 0.f; unit_far.z_far = 1.f;
        float g_o[3] = {0.f, 0.f, 0.f}, g_d[3] = {0.f, 0.f, 0.f}, g_near = 0.f, g_far = 0.f;
        for (int p = 0; p < P; ++p) {
            const int64_t gs = ray * P + p;
            const float u = (G.perturb && !G.t_in) ? G.rand[gs] : 0.f;
            const float t = pe_sample_t_or(G.t_in, gs, pr, p, P, G.perturb != 0, u);
            const float gx[3] = {G.g_pos[gs * 3], G.g_pos[gs * 3 + 1], G.g_pos[gs * 3 + 2]};
            float gt = G.g_t ? G.g_t[gs] : 0.f;
            for (int a = 0; a < 3; ++a) {
                g_o[a] += gx[a];
                g_d[a] = fmaf(gx[a], t, g_d[a]);
                gt = fmaf(gx[a], pr.d[a], gt);
            }
            if (G.g_od) {
                for (int a = 0; a < 3; ++a) { g_o[a] += G.g_od[gs * 6 + a]; g_d[a] += G.g_od[gs * 6 + 3 + a]; }
            }
            if (G.t_in) {                 // explicit ray parameters: dL/dt goes back to the caller (the coarse members depend on the ray)
                if (G.g_t_in) G.g_t_in[gs] = gt;
                continue;
            }
            g_near = fmaf(gt, pe_sample_t(unit_near, p, P, G.perturb != 0, u), g_near);
            g_far = fmaf(gt, pe_sample_t(unit_far, p, P, G.perturb != 0, u), g_far);
        }
        if (near_pass) {                 // z = (face - o) / (d + eps)
            const float den = __fadd_rn(pr.d[near_axis], eps);
            g_o[near_axis] -= g_near / den;
            g_d[near_axis] -= g_near * zn_z / den;
        }
        if (far_pass) {
            const float den = __fadd_rn(pr.d[far_axis], eps);
            g_o[far_axis] -= g_far / den;
            g_d[far_axis] -= g_far * zf_z / den;
        }
        // o = M ow + T, d = M dw
        float g_dw[3] = {0.f, 0.f, 0.f};
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) {
                acc[a * 4 + b] = g_o[a] * ow[b] + g_d[a] * dw[b];
                acc[12 + b] = fmaf(m34[a * 4 + b], g_o[a], acc[12 + b]);
                g_dw[b] = fmaf(m34[a * 4 + b], g_d[a], g_dw[b]);
            }
            acc[a * 4 + 3] = g_o[a];
        }
        if (G.g_dirs) for (int b = 0; b < 3; ++b) G.g_dirs[ray * 3 + b] += g_dw[b];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 15; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 15) {
        const float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
        if (v != 0.f) {
            if (threadIdx.x < 12) { if (G.g_w2o) atomicAdd(G.g_w2o + ((int64_t)img * G.objects + G.k) * 12 + threadIdx.x, v); }
            else if (G.g_origins) atomicAdd(G.g_origins + img * 3 + (threadIdx.x - 12), v);
        }
    }
}

}  // namespace

int pe_launch_composite_bwd(const PeCompositeBwdArgs& args, cudaStream_t stream) {
    const PeCompositeArgs& A = args.f;
    if (A.total_positions > PE_MAX_TOTAL_POSITIONS) { pe_set_error("sum of positions_count over objects (%d) exceeds %d", A.total_positions, PE_MAX_TOTAL_POSITIONS); return PE_ERR_UNSUPPORTED; }
    if (A.features > 256) { pe_set_error("compositor supports up to 256 features"); return PE_ERR_UNSUPPORTED; }
    const int64_t n_rays = (int64_t)A.images * A.rays;
    if (n_rays == 0) return PE_OK;
    const int TP = (A.total_positions + 31) & ~31;
    const size_t smem = (size_t)WARPS * TP * 15 * sizeof(float);
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)pe_min64((n_rays + WARPS - 1) / WARPS, 148 * 16);
    pe_composite_bwd_kernel<<<grid, WARPS * 32, smem, stream>>>(args);
    PE_LAUNCH_CHECK("pe_composite_bwd_kernel");
    return PE_OK;
}

int pe_launch_style_bwd(const PeStyleBwdArgs& args, cudaStream_t stream) {
    if (args.images == 0) return PE_OK;
    pe_style_bwd_kernel<<<dim3(args.images, 8), 256, (size_t)2 * args.channels * sizeof(float), stream>>>(args);
    PE_LAUNCH_CHECK("pe_style_bwd_kernel");
    return PE_OK;
}

int pe_launch_bn_fix(const double* fwd_stats, const double* bn_sums, int channels, float* bn_fix, cudaStream_t stream) {
    pe_bn_fix_kernel<<<(channels + 127) / 128, 128, 0, stream>>>(fwd_stats, bn_sums, channels, bn_fix);
    PE_LAUNCH_CHECK("pe_bn_fix_kernel");
    return PE_OK;
}

int pe_launch_geometry_bwd(const PeGeometryBwdArgs& args, cudaStream_t stream) {
    if (args.images == 0 || args.rays == 0) return PE_OK;
    dim3 grid((args.rays + 127) / 128, args.images);
    pe_geometry_bwd_kernel<<<grid, 128, 0, stream>>>(args);
    PE_LAUNCH_CHECK("pe_geometry_bwd_kernel");
    return PE_OK;
}

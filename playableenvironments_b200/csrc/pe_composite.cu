// Volume-rendering compositor: one warp per ray.  Replaces ObjectComposer.integrate (per object and for
// the composed scene), compose (concat + sort by t + gathers) and fix_all_object_overlaps
// (model/object_composer.py:153-214, 220-397, 399-447, 724-784).
#include "pe_kernels.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int FEAT_BATCH = 1;      // feature rows in flight per warp in the weighted sum

struct RayLists {       // per-warp shared-memory sample list
    float* t;
    float* raw;
    float* dm;          // |displacement|
    float* dv;          // divergence of the displacement field
    int* src;           // (object << 16) | sample, -1 for "features are zero"
};

// ObjectComposer.integrate (model/object_composer.py:724-784) over a list already ordered by t.
// `hand`: the decoder hand-off of the composed scene (PeHandoff) or NULL; `seg` = the segment this ray belongs to (-1: none)
__device__ void integrate_list(const PeCompositeArgs& A, const RayLists& S, int n, float dnorm, const float* __restrict__ noise,
                               const PeIntegrated& out, int64_t ray, int lane, const PeHandoff* hand = nullptr, int seg = -1) {
    const int F = A.features;
    // channel groups of 32 whose weighted sum anybody reads: all of them, or -- hand-off only -- the segment's channel range
    unsigned groups = 0xffu;
    if (hand && !out.integrated_features) {
        groups = 0u;
        if (seg >= 0) for (int i = 0; i < 8; ++i) if (32 * i >= hand->channel_begin[seg] && 32 * i < hand->channel_begin[seg] + hand->channel_count[seg]) groups |= 1u << i;
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float carry = 1.f, opacity = 0.f, depth = 0.f, dsum = 0.f, vsum = 0.f;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int j = c0 + lane;
        float alpha = 0.f, t = 0.f, dm = 0.f, dv = 0.f;
        if (j < n) {
            t = S.t[j];
            dm = S.dm[j];
            dv = S.dv[j];
            // compute_position_distances :153-178 — last interval 1e10, scaled by |d|
            const float delta = __fmul_rn(j == n - 1 ? 1e10f : __fsub_rn(S.t[j + 1], t), dnorm);
            float raw = S.raw[j];
            if (noise) raw = __fadd_rn(raw, noise[j]);                       // compute_alphas :193-195
            alpha = __fsub_rn(1.f, expf(__fmul_rn(-fmaxf(raw, 0.f), delta)));  // :197
        }
        // compute_weights :199-214 — exclusive cumprod of (1 - alpha + 1e-10)
        const float shifted = j < n ? __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.f;
        float incl = shifted;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= v;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        const float w = alpha * (carry * excl);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (j < n && out.weights) out.weights[ray * n + j] = w;
        opacity += w;
        depth += w * t;
        dsum += w * dm;
        vsum += alpha * fabsf(dv);                 // mean(alphas * |divergence|) :777-778
        const int src = j < n ? S.src[j] : -1;
        // only the samples that carry features and weight (in sample order: the sum is evaluated like the reference's); in a sparse view
        // most of a ray's samples lie outside every box
        unsigned todo = groups ? __ballot_sync(0xffffffffu, j < n && w != 0.f && src >= 0) : 0u;
        // FEAT_BATCH samples' feature rows are requested before the first of them is used: the sums stay in sample order, but a ray's
        // in-box samples cost one memory round trip per batch instead of one each (this kernel is latency bound: profiles/r2_compositor.md)
        while (todo) {
            float wv[FEAT_BATCH];
            const float* fp[FEAT_BATCH];
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < FEAT_BATCH; ++u) {
                wv[u] = 0.f;
                fp[u] = nullptr;
                if (todo) {                                   // warp-uniform
                    const int jj = __ffs(todo) - 1;
                    todo &= todo - 1;
                    wv[u] = __shfl_sync(0xffffffffu, w, jj);
                    const int sj = __shfl_sync(0xffffffffu, src, jj);
                    const int k = sj >> 16, p = sj & 0xffff;
                    fp[u] = A.feat[k] + (ray * A.positions[k] + p) * (int64_t)F;
                    cnt = u + 1;
                }
            }
            float fv[FEAT_BATCH][8];
#pragma unroll
            for (int u = 0; u < FEAT_BATCH; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = lane + 32 * i;
                    fv[u][i] = (u < cnt && c < F && ((groups >> i) & 1u)) ? __ldg(fp[u] + c) : 0.f;
                }
#pragma unroll
            for (int u = 0; u < FEAT_BATCH; ++u)
                if (u < cnt) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(wv[u], fv[u][i], acc[i]);
                }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        opacity += __shfl_xor_sync(0xffffffffu, opacity, o);
        depth += __shfl_xor_sync(0xffffffffu, depth, o);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    }
    if (out.integrated_features) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < F) out.integrated_features[ray * F + c] = acc[i];
        }
    }
    if (hand && seg >= 0) {
        // channels-first store into the segment's grid: [image][channel - channel_begin][ray - ray_begin]
        const int img = (int)(ray / A.rays), r = (int)(ray - (int64_t)img * A.rays);
        const int cb = hand->channel_begin[seg], cc = hand->channel_count[seg], rc = hand->ray_count[seg];
        float* g = hand->grid[seg] + (int64_t)img * cc * rc + (r - hand->ray_begin[seg]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i - cb;
            if (c >= 0 && c < cc) g[(int64_t)c * rc] = acc[i];
        }
    }
    if (lane == 0) {
        if (out.opacity) out.opacity[ray] = opacity;
        if (out.depth) out.depth[ray] = depth;
        if (out.disparity) {                                   // :765; 0/0 stays NaN as torch.clamp keeps it
            const float q = depth / opacity;
            out.disparity[ray] = 1.f / (q != q ? q : fmaxf(q, 1e-10f));
        }
        if (out.integrated_displacements_magnitude) out.integrated_displacements_magnitude[ray] = dsum / (float)n;  // mean :772
        if (out.integrated_divergence) out.integrated_divergence[ray] = vsum / (float)n;   // zero unless the caller asked for the Hutchinson term
    }
}

__global__ void __launch_bounds__(WARPS * 32) pe_composite_kernel(const PeCompositeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TP = (A.total_positions + 31) & ~31;
    float* base = reinterpret_cast<float*>(smem_raw) + (size_t)warp * TP * 10;
    RayLists S{base, base + TP, base + 2 * TP, base + 3 * TP, reinterpret_cast<int*>(base + 4 * TP)};
    RayLists U{base + 5 * TP, base + 6 * TP, base + 7 * TP, base + 8 * TP, reinterpret_cast<int*>(base + 9 * TP)};
    const int64_t n_rays = (int64_t)A.images * A.rays;
    for (int64_t ray = (int64_t)blockIdx.x * WARPS + warp; ray < n_rays; ray += (int64_t)gridDim.x * WARPS) {
        const float* d = A.dirs + ray * 3;
        const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        // ---- per-object integration (object_composer.py:880) ----
        if (A.do_objects) {
            for (int k = 0; k < A.objects; ++k) {
                const PeIntegrated& ok = A.object[k];
                // objects that integrated themselves in the fused field kernel (or whose outputs nobody asked for) are not walked again
                if (!(ok.integrated_features || ok.opacity || ok.weights || ok.depth || ok.disparity || ok.integrated_displacements_magnitude ||
                      ok.integrated_divergence)) continue;
                const int P = A.positions[k];
                const int64_t b = ray * P;
                __syncwarp();
                for (int p = lane; p < P; p += 32) {
                    S.t[p] = A.t[k][b + p];
                    S.raw[p] = A.raw[k][b + p];
                    S.dm[p] = A.dispmag[k] ? A.dispmag[k][b + p] : 0.f;
                    S.dv[p] = A.div[k] ? A.div[k][b + p] : 0.f;
                    S.src[p] = A.inbox[k][b + p] ? ((k << 16) | p) : -1;
                }
                __syncwarp();
                integrate_list(A, S, P, dnorm, (A.perturb && A.noise[k]) ? A.noise[k] + b : nullptr, A.object[k], ray, lane);
            }
        }
        // ---- composition of all objects (object_composer.py:885-886) ----
        if (A.do_global) {
            __syncwarp();
            int off = 0;
            for (int k = 0; k < A.objects; ++k) {
                const int P = A.positions[k];
                const int64_t b = ray * P;
                for (int p0 = 0; p0 < P; p0 += 32) {
                    const int p = p0 + lane;
                    float t = 0.f, raw = 0.f, dm = 0.f, dv = 0.f;
                    int src = -1;
                    if (p < P) {
                        t = A.t[k][b + p];
                        raw = A.raw[k][b + p];
                        dm = A.dispmag[k] ? A.dispmag[k][b + p] : 0.f;
                        dv = A.div[k] ? A.div[k][b + p] : 0.f;
                        src = A.inbox[k][b + p] ? ((k << 16) | p) : -1;
                    }
                    // fix_object_overlap :295-397: static samples between the first and the last sample of a
                    // dynamic object (indices via searchsorted on the ORIGINAL static t) get alpha -10, t 0.
                    if (A.fix_overlaps && k < A.static_objects) {
                        bool masked = false;
                        for (int dk = A.static_objects; dk < A.objects; ++dk) {
                            const float* td = A.t[dk] + ray * A.positions[dk];
                            const float v0 = td[0], v1 = td[P - 1];
                            int lo = 0, hi = 0;
                            for (int q0 = 0; q0 < P; q0 += 32) {      // lower_bound = #elements < v (t is non-decreasing)
                                const int q = q0 + lane;
                                const float tq = q < P ? A.t[k][b + q] : INFINITY;
                                lo += __popc(__ballot_sync(0xffffffffu, tq < v0));
                                hi += __popc(__ballot_sync(0xffffffffu, tq < v1));
                            }
                            masked = masked || (p >= lo && p < hi);
                        }
                        if (masked) { raw = raw * 0.f - 10.f; t = 0.f; dm = 0.f; dv = 0.f; }
                    }
                    if (p < P) { U.t[off + p] = t; U.raw[off + p] = raw; U.dm[off + p] = dm; U.dv[off + p] = dv; U.src[off + p] = src; }
                }
                off += P;
            }
            __syncwarp();
            const int n = A.total_positions;
            // stable sort by t (torch.sort :435 with a deterministic tie order: concatenation index)
            const bool ordered = pe_lists_ordered(U.t, n, A.positions, A.objects, lane);
            for (int j = lane; j < n; j += 32) {
                const float tj = U.t[j];
                const int rank = pe_compose_rank(U.t, n, j, A.positions, A.objects, ordered);
                S.t[rank] = tj; S.raw[rank] = U.raw[j]; S.dm[rank] = U.dm[j]; S.dv[rank] = U.dv[j]; S.src[rank] = U.src[j];
            }
            __syncwarp();
            int seg = -1;
            if (A.handoff.segments) {
                const int r = (int)(ray % A.rays);
                for (int q = 0; q < A.handoff.segments; ++q)
                    if (r >= A.handoff.ray_begin[q] && r < A.handoff.ray_begin[q] + A.handoff.ray_count[q]) seg = q;
            }
            integrate_list(A, S, n, dnorm, (A.perturb && A.noise_global) ? A.noise_global + ray * n : nullptr, A.global, ray, lane,
                           A.handoff.segments ? &A.handoff : nullptr, seg);
        }
    }
}

}  // namespace

int pe_launch_composite(const PeCompositeArgs& args, cudaStream_t stream) {
    if (args.total_positions > PE_MAX_TOTAL_POSITIONS) {
        pe_set_error("sum of positions_count over objects (%d) exceeds %d", args.total_positions, PE_MAX_TOTAL_POSITIONS);
        return PE_ERR_UNSUPPORTED;
    }
    if (args.features > 256) { pe_set_error("compositor supports up to 256 features"); return PE_ERR_UNSUPPORTED; }
    if (args.handoff.segments < 0 || args.handoff.segments > PE_MAX_HANDOFF) { pe_set_error("hand-off: segments must be in [0,%d]", PE_MAX_HANDOFF); return PE_ERR_INVALID; }
    for (int q = 0; q < args.handoff.segments; ++q) {
        const PeHandoff& h = args.handoff;
        if (!h.grid[q] || h.ray_begin[q] < 0 || h.ray_count[q] < 0 || h.ray_begin[q] + h.ray_count[q] > args.rays || h.channel_begin[q] % 32 ||
            h.channel_count[q] % 32 || h.channel_begin[q] < 0 || h.channel_begin[q] + h.channel_count[q] > args.features) {
            pe_set_error("hand-off segment %d: ray range inside the frame, channel range a multiple of 32 inside the features, a grid", q);
            return PE_ERR_INVALID;
        }
    }
    if (args.fix_overlaps) {
        for (int s = 0; s < args.static_objects; ++s)
            for (int d = args.static_objects; d < args.objects; ++d)
                if (args.positions[s] > args.positions[d]) {
                    // the reference indexes the dynamic t at positions_count(static)-1 (object_composer.py:322) and would raise
                    pe_set_error("fix_object_overlaps needs positions_count(static %d)=%d <= positions_count(dynamic %d)=%d",
                                 s, args.positions[s], d, args.positions[d]);
                    return PE_ERR_INVALID;
                }
    }
    const int64_t n_rays = (int64_t)args.images * args.rays;
    if (n_rays == 0) return PE_OK;
    const int TP = (args.total_positions + 31) & ~31;
    const size_t smem = (size_t)WARPS * TP * 10 * sizeof(float);
    PE_CUDA_CHECK(cudaFuncSetAttribute(pe_composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)pe_min64((n_rays + WARPS - 1) / WARPS, 148 * 16);
    pe_composite_kernel<<<grid, WARPS * 32, smem, stream>>>(args);
    PE_LAUNCH_CHECK("pe_composite_kernel");
    return PE_OK;
}
